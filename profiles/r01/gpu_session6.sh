#!/bin/bash
# GPU session 6: GRU kernels parity, full suite, bench c4/c2/c1, step breakdown.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gru.py -x -q > gpurun_out/t_gru.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gru.log
tail -15 gpurun_out/t_gru.log
grep -q "rc=0" gpurun_out/t_gru.log || exit 0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
timeout 600 python profiles/step_breakdown.py c4 > gpurun_out/step_breakdown_c4.txt 2>&1
ls -la gpurun_out
