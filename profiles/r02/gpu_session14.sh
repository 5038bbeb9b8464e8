#!/bin/bash
# which in-panel updates should take the tensor path?  (K = 64 updates touch one 64-column tile, K = 128 two)
set -x
mkdir -p gpurun_out
for mc in 64 65 129; do
for w in c4 c3; do
  DRP_OZ_MIN_COLS=$mc timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench14_${w}_mc$mc.json 2> gpurun_out/r02_bench14_${w}_mc$mc.err; echo "bench $w mc=$mc rc=$?"
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_bench14_${w}_mc$mc.json"))
    k=d["kernels"]
    print("$w mc=$mc", "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "eager(serial)", round(d["eager_ms_per_step"],3), "cond", round(d["conditional"]["mean_only_ms"],3), round(d["conditional"]["mean_and_covariance_ms"],3))
    print("   ", {n:(round(v["ms_per_step"],3), v["launches_per_step"]) for n,v in k.items() if n in ("oz_syrk_kernel","oz_slice_kernel","trsm_kernel","potf2_kernel","syrk_kernel")})
except Exception as e: print("fail", e)
P
done
done
