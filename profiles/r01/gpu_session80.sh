#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_80.json 2> gpurun_out/bench_80.err; echo "bench rc=$?"
grep "per-step\|ELBO" gpurun_out/bench_80.err
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep per-step
