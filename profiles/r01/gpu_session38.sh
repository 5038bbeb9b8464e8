#!/bin/bash
set -x
mkdir -p gpurun_out
for i in 1 2; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_gram$i.json 2> /dev/null
  DRP_GRAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_nogram$i.json 2> /dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_c4_*gram*.json")):
    d=json.load(open(f)); print(f, round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), round(d["kernel_classes"]["reduce"]["ms_per_step"],2))
PY
