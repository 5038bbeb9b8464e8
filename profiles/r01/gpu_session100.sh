#!/bin/bash
# bench line with per-class achieved rates (stream overlap off: unshared per-class numbers)
set -x
mkdir -p gpurun_out
DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_final5_nostream.json 2> gpurun_out/bench_default_final5_nostream.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_default_final5_nostream.json')); print(d['ms_per_step'])
for k,v in d['kernel_classes'].items(): print(f\"{k:14s} {v['ms_per_step']:7.3f} ms  {v['achieved']:9.1f} {v['unit']:8s} {v['frac_of_peak']:.3f}\")"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_final5.json 2> gpurun_out/bench_default_final5.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_default_final5.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
