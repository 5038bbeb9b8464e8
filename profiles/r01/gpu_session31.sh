#!/bin/bash
# GPU session 31: full suite after the MLP unrolling (LCA, categorical, OU reductions), C5 sweep, benches c1-c4.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
grep -q "rc=0" gpurun_out/t_all.log || exit 0
timeout 900 python tests/sweep_c5.py > gpurun_out/sweep_c5.json 2> gpurun_out/sweep_c5.err
for w in c4 c3 c2 c1; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
