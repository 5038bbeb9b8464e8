#!/bin/bash
set -x
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.Stream.priority_range())"
for cfg in "1 1" "1 0" "0 0"; do
  set -- $cfg
  DRP_LOOKAHEAD=$1 DRP_LOOKAHEAD_YIELD=$2 timeout 900 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench07_c4_$1$2.json 2> gpurun_out/r02_bench07_c4_$1$2.err; echo "bench la=$1 yield=$2 rc=$?"
  grep "per-step" gpurun_out/r02_bench07_c4_$1$2.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_bench07_c4_$1$2.json"))
    print("la=$1 yield=$2", "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "eager(serial)", d["eager_ms_per_step"], "cond", d["conditional"]["mean_only_ms"], d["conditional"]["mean_and_covariance_ms"])
except Exception as e: print("fail", e)
P
done
