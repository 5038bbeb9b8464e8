#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_55.json 2> gpurun_out/bench_55.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_55.err
timeout 600 python profiles/step_breakdown.py c4 > gpurun_out/step_breakdown_c4_v12.txt 2>&1; echo "breakdown rc=$?"
