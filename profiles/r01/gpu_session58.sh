#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gru.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/trace_gru.py 32 > gpurun_out/trace_gru_v6_32.txt 2>&1; tail -5 gpurun_out/trace_gru_v6_32.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_58.json 2> gpurun_out/bench_58.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_58.err
DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_58_nostream.json 2> gpurun_out/bench_58_nostream.err
grep "per-step" gpurun_out/bench_58_nostream.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_58_nostream.json'))
print({k:round(v['ms_per_step'],2) for k,v in d['kernel_classes'].items()})
PY
