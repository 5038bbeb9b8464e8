#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_66.json 2> gpurun_out/bench_66.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_66.err
DRP_STEP_PRIORITY=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per-step"
for c in c2 c3; do timeout 600 python bench.py --workload $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per-step"; done
