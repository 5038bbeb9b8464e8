#!/bin/bash
# bulk accumulator read-out in the tcgen05 update's epilogue: parity, then A/B at C4
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/r02_pytest08.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest08.log
for ep in bulk prog; do
  DRP_OZ_EPILOGUE=$ep timeout 900 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench08_c4_$ep.json 2> gpurun_out/r02_bench08_c4_$ep.err; echo "bench $ep rc=$?"
  grep "per-step" gpurun_out/r02_bench08_c4_$ep.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_bench08_c4_$ep.json"))
    k=d["kernels"]
    print("$ep", "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "eager(serial)", d["eager_ms_per_step"], "cond", d["conditional"]["mean_only_ms"], d["conditional"]["mean_and_covariance_ms"])
    print({n:(round(v["ms_per_step"],3), v["launches_per_step"], round(v.get("achieved",0),1)) for n,v in k.items() if n in ("oz_syrk_kernel","oz_slice_kernel","trsm_kernel","potf2_kernel")})
except Exception as e: print("fail", e)
P
done
