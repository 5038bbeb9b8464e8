#!/bin/bash
# GPU session 18: UV-mode decoder GRU parity + bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
grep -q "rc=0" gpurun_out/t_all.log || exit 0
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
ls -la gpurun_out
