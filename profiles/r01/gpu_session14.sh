#!/bin/bash
# GPU session 14: hoisted integer row scales in the tcgen05 epilogue; fast gate non-linearities.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_gru.py -x -q > gpurun_out/t_ozaki.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_ozaki.log
tail -5 gpurun_out/t_ozaki.log
grep -q "rc=0" gpurun_out/t_ozaki.log || exit 0
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
ls -la gpurun_out
