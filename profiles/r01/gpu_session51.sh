#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_51.json 2> gpurun_out/bench_51.err; echo "bench rc=$?"
grep "per-step\|clock" gpurun_out/bench_51.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_51.json'))
print(d['clocks'])
PY
timeout 600 python profiles/step_breakdown.py c4 > gpurun_out/step_breakdown_c4_v11.txt 2>&1; echo "breakdown rc=$?"
