#!/bin/bash
set -x
mkdir -p gpurun_out
./profiles/microbench/dmma_shapes > gpurun_out/dmma_shapes.txt 2>&1
tail -4 gpurun_out/dmma_shapes.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
