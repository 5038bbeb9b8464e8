#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gru.py tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -4
for h in 0 1; do for r in 32 16; do echo "handoff $h rows $r"; DRP_GRU_HANDOFF=$h timeout 200 python profiles/trace_gru.py $r 2>&1 | tail -4; done; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_49.json 2> gpurun_out/bench_49.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_49.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_49.json'))
print({k:round(v['ms_per_step'],2) for k,v in d['kernel_classes'].items()})
PY
DRP_GRU_HANDOFF=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per-step"
