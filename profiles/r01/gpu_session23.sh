#!/bin/bash
# GPU session 23: does DMMA share the fp64 pipe with DFMA? + re-check after reverting the GRU loops; traffic capture.
set -x
mkdir -p gpurun_out
./profiles/microbench/dmma_shapes > gpurun_out/dmma_shapes.txt 2>&1
cat gpurun_out/dmma_shapes.txt
timeout 600 python -m pytest tests/test_gpu_gru.py -x -q > gpurun_out/t_gru.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gru.log
tail -3 gpurun_out/t_gru.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
ls -la gpurun_out
