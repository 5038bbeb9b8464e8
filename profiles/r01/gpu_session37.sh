#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gru.py -x -q > gpurun_out/t_gru.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gru.log
tail -5 gpurun_out/t_gru.log
grep -q "rc=0" gpurun_out/t_gru.log || exit 0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
for w in c4 c3; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 600 python profiles/step_breakdown.py c4 > gpurun_out/step_breakdown_c4.txt 2>&1
