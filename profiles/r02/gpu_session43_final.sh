#!/bin/bash
# final check of the committed tree: whole GPU suite, smoke, default bench; then ncu --set full of the kernels changed last
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest43.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest43.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02_smoke43.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke43.log
timeout 900 python bench.py > gpurun_out/r02_final3_c4.json 2> gpurun_out/r02_final3_c4.err; echo "default rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/r02_final3_c4.json"))
print("c4", round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],4), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), "north", round(d["north_star_kernel"]["frac"],4), d["clocks"], d["gpu_launches"])
P
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none -k regex:"(gram_tn|ou_grad|oz_slice|potrf_diag)_kernel" -s 24 -c 10 -o gpurun_out/r02_ncu_last_c4 -f $CMD > gpurun_out/r02_ncu_last.out 2> gpurun_out/r02_ncu_last.err
echo "ncu rc=$?"
ncu -i gpurun_out/r02_ncu_last_c4.ncu-rep --page raw --csv > gpurun_out/r02_ncu_last_c4.raw.csv 2> /dev/null
python profiles/summarize_ncu_csv.py gpurun_out/r02_ncu_last_c4.raw.csv > gpurun_out/r02_ncu_last_c4.summary.txt
rm -f gpurun_out/r02_ncu_last_c4.raw.csv gpurun_out/r02_ncu_last_c4.ncu-rep
grep -E "Kernel Name|gpu__time_duration|dram__bytes|dmma|dram_throughput" gpurun_out/r02_ncu_last_c4.summary.txt | head -60
