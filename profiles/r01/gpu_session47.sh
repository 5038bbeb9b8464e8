#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gru.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -4
for ns in 0 800 1600 2400 3200; do echo "stagger $ns"; DRP_GRU_HANDOFF=$ns timeout 200 python profiles/trace_gru.py 32 2>&1 | tail -4; done
for ns in 0 800 1600; do echo "stagger $ns rows 16"; DRP_GRU_HANDOFF=$ns timeout 200 python profiles/trace_gru.py 16 2>&1 | tail -4; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_47.json 2> gpurun_out/bench_47.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_47.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_47.json'))
print({k:round(v['ms_per_step'],2) for k,v in d['kernel_classes'].items()})
PY
