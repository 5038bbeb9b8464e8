#!/bin/bash
# GPU session 8: register-blocked potf2/trsm, K=64 tcgen05 panel updates, collapsed decoder projection.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -15 gpurun_out/t_all.log
grep -q "rc=0" gpurun_out/t_all.log || exit 0
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 600 python profiles/step_breakdown.py c4 > gpurun_out/step_breakdown_c4.txt 2>&1
ls -la gpurun_out
