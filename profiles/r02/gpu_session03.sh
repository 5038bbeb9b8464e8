#!/bin/bash
# round 2: CUDA-graph step + fused ClippedAdam + new bench: tests and all workloads
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -W "error::UserWarning:draupnir_asr_b200" > gpurun_out/r02_pytest03.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r02_pytest03.log | tail -30
for w in c1 c2 c3 c4 c5; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench03_$w.json 2> gpurun_out/r02_bench03_$w.err; echo "bench $w rc=$?"
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_bench03_$w.json"))
    print("$w", "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "graph", d.get("cuda_graph"), "eager", d.get("eager_ms_per_step"), "launches", d["gpu_launches"])
    print("   roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "cond", d.get("conditional",{}).get("mean_only_ms"), d.get("conditional",{}).get("mean_and_covariance_ms"))
    for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:12]:
        print("   %-24s %8.3f ms %6.1f launches  %10.2f %s  frac %.3f" % (k, v["ms_per_step"], v["launches_per_step"], v.get("achieved",0), v.get("unit",""), v.get("frac",0)))
except Exception as e: print("$w failed", e)
P
  tail -4 gpurun_out/r02_bench03_$w.err
done
