#!/bin/bash
# interior fast path of the tcgen05 update's epilogue: parity, phase accounting, A/B against the previous kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python profiles/prof_ozaki.py 256 > gpurun_out/prof_ozaki_k256_fast.txt 2>&1; cat gpurun_out/prof_ozaki_k256_fast.txt
cp draupnir_asr_b200/libdraupnir_b200.so /tmp/keep.so
for v in base new base new; do
  cp draupnir_asr_b200/ab_$v.so draupnir_asr_b200/libdraupnir_b200.so; touch draupnir_asr_b200/libdraupnir_b200.so
  echo "== $v"
  timeout 120 python profiles/trace_ozaki.py 256 2>&1 | head -1
  timeout 120 python profiles/trace_ozaki.py 128 2>&1 | head -1
done
cp /tmp/keep.so draupnir_asr_b200/libdraupnir_b200.so
