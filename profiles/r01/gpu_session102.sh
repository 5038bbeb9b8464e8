#!/bin/bash
# A/B: MMA issue order of the tcgen05 update (k-step outer / A slice inner  vs  A slice outer / k-step inner)
set -x
mkdir -p gpurun_out
cp draupnir_asr_b200/libdraupnir_b200.so /tmp/keep.so
for v in base new base new; do
  cp draupnir_asr_b200/ab_$v.so draupnir_asr_b200/libdraupnir_b200.so; touch draupnir_asr_b200/libdraupnir_b200.so
  echo "== $v"
  timeout 120 python profiles/trace_ozaki.py 256 2>&1 | head -1
  timeout 120 python profiles/trace_ozaki.py 128 2>&1 | head -1
done
cp draupnir_asr_b200/ab_new.so draupnir_asr_b200/libdraupnir_b200.so; touch draupnir_asr_b200/libdraupnir_b200.so
timeout 600 python -m pytest tests/test_gpu_ozaki.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python profiles/prof_ozaki.py 256 2>&1 | head -8
cp /tmp/keep.so draupnir_asr_b200/libdraupnir_b200.so
