#!/bin/bash
# f4 (marginal samples from one factorisation), encoder direction-0 forward + deferred dictionary entries, bench hygiene
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_41.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_41.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_41.json 2> gpurun_out/bench_41.err; echo "bench rc=$?"
cat gpurun_out/bench_41.json
grep "per-step" gpurun_out/bench_41.err
for c in c1 c2 c3; do
  timeout 600 python bench.py --workload $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_41.json 2> gpurun_out/bench_${c}_41.err
  grep "per-step" gpurun_out/bench_${c}_41.err
done
