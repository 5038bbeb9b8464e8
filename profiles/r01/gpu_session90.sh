#!/bin/bash
# tcgen05 update epilogue: anti-diagonals paired exactly in int64 before the conversion (half the fp64 instructions)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python profiles/trace_ozaki.py 256 2>&1 | head -1
timeout 120 python profiles/trace_ozaki.py 128 2>&1 | head -1
DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_90_nostream.json 2> gpurun_out/bench_90_nostream.err
python -c "
import json; d=json.load(open('gpurun_out/bench_90_nostream.json')); print('nostream', d['ms_per_step']); print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernel_classes'].items()})"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_90.json 2> gpurun_out/bench_90.err
python -c "
import json; d=json.load(open('gpurun_out/bench_90.json')); print('stream', d['ms_per_step'], d['e2e']['ms_per_step'])"
