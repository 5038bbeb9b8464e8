#!/bin/bash
# final consolidation of round 1 (second pass, after the GRU generation 2 / stream overlap work): gpu tests, smoke, default bench (with cpu baseline), reference arm, C1-C3 lines
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final3.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_final3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final3.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke_final3.log
timeout 900 python bench.py > gpurun_out/bench_default_final3.json 2> gpurun_out/bench_default_final3.err; echo "bench rc=$?"
cat gpurun_out/bench_default_final3.json
grep "per-step" gpurun_out/bench_default_final3.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final3.json 2> gpurun_out/bench_reference_final3.err; echo "ref rc=$?"
cat gpurun_out/bench_reference_final3.json
for c in c1 c2 c3; do
  timeout 600 python bench.py --workload $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_final3.json 2> gpurun_out/bench_${c}_final3.err
  cat gpurun_out/bench_${c}_final3.json
done
