#!/bin/bash
# A/B on one box: tcgen05 update with the fp64 recombination (base) vs the paired int64 recombination (new)
set -x
mkdir -p gpurun_out
for v in base new base new; do
  cp draupnir_asr_b200/ab_$v.so draupnir_asr_b200/libdraupnir_b200.so
  touch draupnir_asr_b200/libdraupnir_b200.so
  echo "== $v"
  timeout 120 python profiles/trace_ozaki.py 256 2>&1 | head -1
  timeout 120 python profiles/trace_ozaki.py 128 2>&1 | head -1
done
for v in base new; do
  cp draupnir_asr_b200/ab_$v.so draupnir_asr_b200/libdraupnir_b200.so
  touch draupnir_asr_b200/libdraupnir_b200.so
  DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_91_${v}_nostream.json 2> gpurun_out/bench_91_${v}_nostream.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_91_${v}_nostream.json')); print('$v nostream', d['ms_per_step'], 'syrk_tcgen05', d['kernel_classes']['syrk_tcgen05']['ms_per_step'])"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_91_${v}.json 2> gpurun_out/bench_91_${v}.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_91_${v}.json')); print('$v stream', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
