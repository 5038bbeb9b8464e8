#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_54.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_54.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_54.json 2> gpurun_out/bench_54.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_54.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_54.json'))
print({k:round(v['ms_per_step'],2) for k,v in d['kernel_classes'].items()})
PY
DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per-step"
for c in c1 c2 c3; do timeout 600 python bench.py --workload $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per-step"; done
