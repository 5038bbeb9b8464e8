#!/bin/bash
# pipelined host-fed step: parity test + C4 / C2 bench lines
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "pipelined or host_fed or graph_replay" > gpurun_out/r02_pytest24.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest24.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench24_c4.json 2> gpurun_out/r02_bench24_c4.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench24_c4.json')); print(d['ms_per_step'], d['e2e'])"
grep "per-step" gpurun_out/r02_bench24_c4.err
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench24_c2.json 2> gpurun_out/r02_bench24_c2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench24_c2.json')); print(d['ms_per_step'], d['e2e'])"
