#!/bin/bash
# ou_grad with the dense alpha + single reduction: full suite, smoke, bench lines C1-C5 with the CPU baseline, reference arm
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest39.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest39.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02_smoke39.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke39.log
timeout 900 python bench.py > gpurun_out/r02_final2_c4.json 2> gpurun_out/r02_final2_c4.err; echo "default rc=$?"
DRP_REF_BUDGET_S=60 timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_final2_reference_c4.json 2> gpurun_out/r02_final2_reference_c4.err; echo "reference rc=$?"
for w in c1 c2 c3 c5; do
  timeout 900 python bench.py --workload $w > gpurun_out/r02_final2_$w.json 2> gpurun_out/r02_final2_$w.err; echo "$w rc=$?"
done
python - <<P
import json
for w in ("c4","c1","c2","c3","c5"):
    d=json.load(open(f"gpurun_out/r02_final2_{w}.json"))
    print(w, round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],4), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d["clocks"])
    if w == "c4": print({n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in d['kernels'].items()}, d.get('conditional'))
r=json.load(open("gpurun_out/r02_final2_reference_c4.json")); print("reference", r["value"], r["steps"], r["config"]==json.load(open("gpurun_out/r02_final2_c4.json"))["config"])
P
