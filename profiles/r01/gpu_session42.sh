#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tree_io.py tests/test_gpu_model.py tests/test_gpu_gru.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_42.json 2> gpurun_out/bench_42.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_42.err
timeout 600 python profiles/step_breakdown.py c4 > gpurun_out/step_breakdown_c4_v10.txt 2>&1; echo "breakdown rc=$?"
