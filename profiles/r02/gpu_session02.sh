#!/bin/bash
# round 2, second call: CUDA-graph step + fused ClippedAdam; all workloads
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest02.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_pytest02.log
for w in c1 c2 c3 c4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench02_$w.json 2> gpurun_out/r02_bench02_$w.err; echo "bench $w rc=$?"
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_bench02_$w.json"))
    print("$w", d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"])
except Exception as e: print("$w failed", e)
P
  tail -3 gpurun_out/r02_bench02_$w.err
done
