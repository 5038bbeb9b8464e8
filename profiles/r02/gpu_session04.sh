#!/bin/bash
set -x
mkdir -p gpurun_out
for prio in -1 0; do
  DRP_GRAPH_PRIORITY=$prio timeout 900 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench04_c4_p$prio.json 2> gpurun_out/r02_bench04_c4_p$prio.err; echo "bench rc=$?"
  grep "per-step" gpurun_out/r02_bench04_c4_p$prio.err
done
DRP_GRAPH_PRIORITY=-1 timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench04_c3.json 2> gpurun_out/r02_bench04_c3.err
grep "per-step" gpurun_out/r02_bench04_c3.err
python - <<P
import json
d=json.load(open("gpurun_out/r02_bench04_c4_p-1.json"))
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]):
    print("   %-24s %8.3f ms %6.1f launches  %10.2f %s  frac %.3f" % (k, v["ms_per_step"], v["launches_per_step"], v.get("achieved",0), v.get("unit",""), v.get("frac",0)))
print(d["conditional"])
P
