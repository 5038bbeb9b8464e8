#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest15.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r02_pytest15.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke15.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke15.log
for w in c4 c3; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench15_${w}.json 2> gpurun_out/r02_bench15_${w}.err; echo "bench $w rc=$?"
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_bench15_${w}.json"))
    k=d["kernels"]
    print("$w", "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "eager(serial)", round(d["eager_ms_per_step"],3), "cond", round(d["conditional"]["mean_only_ms"],3), round(d["conditional"]["mean_and_covariance_ms"],3))
    print("   ", {n:(round(v["ms_per_step"],3), v["launches_per_step"]) for n,v in k.items() if n in ("oz_syrk_kernel","oz_slice_kernel","trsm_kernel","potf2_kernel","syrk_kernel","gru_input_sums_kernel")})
except Exception as e: print("fail", e)
P
done
