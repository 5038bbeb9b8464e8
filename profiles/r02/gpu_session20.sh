#!/bin/bash
# slicing fused into the trsm kernel for the 64-column in-panel updates: parity, then A/B
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest20.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r02_pytest20.log | tail -3
for fs in 1 0; do
for w in c4 c3; do
  DRP_FUSED_SLICING=$fs timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench20_${w}_fs$fs.json 2> gpurun_out/r02_bench20_${w}_fs$fs.err; echo "bench $w fs=$fs rc=$?"
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_bench20_${w}_fs$fs.json"))
    k=d["kernels"]
    print("$w fs=$fs", "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "eager(serial)", round(d["eager_ms_per_step"],3), "cond", round(d["conditional"]["mean_only_ms"],3), round(d["conditional"]["mean_and_covariance_ms"],3))
    print("   ", {n:(round(v["ms_per_step"],3), v["launches_per_step"]) for n,v in k.items() if n in ("oz_syrk_kernel","oz_slice_kernel","trsm_kernel","potf2_kernel")})
except Exception as e: print("fail", e)
P
done
done
