#!/bin/bash
# rotating TMEM accumulator slots in the tcgen05 update: parity, trace of one update, C4 step
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -m gpu -x -q 2>&1 | tail -4
timeout 120 python profiles/trace_ozaki.py > gpurun_out/trace_ozaki_v7.txt 2>&1; head -3 gpurun_out/trace_ozaki_v7.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_88.json 2> gpurun_out/bench_88.err; echo "bench rc=$?"
grep "per-step\|ELBO" gpurun_out/bench_88.err | head -5
python -c "
import json; d=json.load(open('gpurun_out/bench_88.json')); print(d['ms_per_step'], d['e2e']['ms_per_step']); print({k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})"
DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_88_nostream.json 2> gpurun_out/bench_88_nostream.err
python -c "
import json; d=json.load(open('gpurun_out/bench_88_nostream.json')); print(d['ms_per_step']); print({k:round(v['ms_per_step'],3) for k,v in d['kernel_classes'].items()})"
