#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gru.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -4
for r in 32 16 8; do timeout 200 python profiles/trace_gru.py $r > gpurun_out/trace_gru_v3_$r.txt 2>&1; cat gpurun_out/trace_gru_v3_$r.txt; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_45.json 2> gpurun_out/bench_45.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_45.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_45.json'))
print({k:round(v['ms_per_step'],2) for k,v in d['kernel_classes'].items()})
PY
DRP_GRU_FWD=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per-step"
