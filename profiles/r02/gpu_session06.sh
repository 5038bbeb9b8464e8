#!/bin/bash
# look-ahead in the factorisation: parity tests, then C3 / C4 with and without it
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest06.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r02_pytest06.log | tail -12
for la in 1 0; do
for w in c4 c3; do
  DRP_LOOKAHEAD=$la timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench06_${w}_la$la.json 2> gpurun_out/r02_bench06_${w}_la$la.err; echo "bench $w la=$la rc=$?"
  grep "per-step" gpurun_out/r02_bench06_${w}_la$la.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_bench06_${w}_la$la.json"))
    print("$w la=$la", "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "eager(serial)", d["eager_ms_per_step"], "cond", d["conditional"]["mean_only_ms"], d["conditional"]["mean_and_covariance_ms"])
except Exception as e: print("fail", e)
P
done
done
