#!/bin/bash
# GPU session 7: wgrad kernel + parallel tree_prepare parity, full suite, bench, C5 sweep.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gru.py tests/test_gpu_kernels.py -x -q > gpurun_out/t_gru.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gru.log
tail -15 gpurun_out/t_gru.log
grep -q "rc=0" gpurun_out/t_gru.log || exit 0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python profiles/step_breakdown.py c4 > gpurun_out/step_breakdown_c4.txt 2>&1
timeout 900 python tests/sweep_c5.py > gpurun_out/sweep_c5.json 2> gpurun_out/sweep_c5.err
ls -la gpurun_out
