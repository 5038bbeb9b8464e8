#!/bin/bash
# GPU session 16: interleaved-accumulator MMA issue order; parity, trace, bench.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_kernels.py -x -q > gpurun_out/t_ozaki.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_ozaki.log
tail -5 gpurun_out/t_ozaki.log
timeout 300 python profiles/trace_ozaki.py > gpurun_out/trace_ozaki.txt 2>&1
grep -q "rc=0" gpurun_out/t_ozaki.log || exit 0
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
ls -la gpurun_out
