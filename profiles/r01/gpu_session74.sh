#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gru.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_74.json 2> gpurun_out/bench_74.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_74.err
DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_74_nostream.json 2> gpurun_out/bench_74_nostream.err
grep "per-step" gpurun_out/bench_74_nostream.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_74_nostream.json'))
print({k:round(v['ms_per_step'],2) for k,v in d['kernel_classes'].items()})
PY
