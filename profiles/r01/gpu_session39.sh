#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
grep "per-step" gpurun_out/bench_a.err gpurun_out/bench_b.err gpurun_out/bench_c.err
