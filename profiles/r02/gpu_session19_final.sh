#!/bin/bash
# final single-GPU lines of the round-2 build: default bench (with the CPU baseline), the reference arm, C1-C3, C5
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_final_c4.json 2> gpurun_out/r02_final_c4.err; echo "default rc=$?"
DRP_REF_BUDGET_S=60 timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_final_reference_c4.json 2> gpurun_out/r02_final_reference_c4.err; echo "reference rc=$?"
for w in c1 c2 c3 c5; do
  timeout 900 python bench.py --workload $w > gpurun_out/r02_final_$w.json 2> gpurun_out/r02_final_$w.err; echo "$w rc=$?"
done
python - <<P
import json
for w in ("c4","c1","c2","c3","c5"):
    d=json.load(open(f"gpurun_out/r02_final_{w}.json"))
    print(w, round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],4), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["clocks"])
r=json.load(open("gpurun_out/r02_final_reference_c4.json")); print("reference", r["value"], r["steps"], r["config"]==json.load(open("gpurun_out/r02_final_c4.json"))["config"])
P
