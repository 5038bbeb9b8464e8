#!/bin/bash
# GPU session 25: table-based LCA kernel, shared fast exp in OU / categorical kernels: parity, C5 sweep, bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
grep -q "rc=0" gpurun_out/t_all.log || exit 0
timeout 900 python tests/sweep_c5.py > gpurun_out/sweep_c5.json 2> gpurun_out/sweep_c5.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
ls -la gpurun_out
