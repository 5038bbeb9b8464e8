#!/bin/bash
set -x
mkdir -p gpurun_out
for r in 32 16; do timeout 200 python profiles/trace_gru.py $r > gpurun_out/trace_gru_v4_$r.txt 2>&1; cat gpurun_out/trace_gru_v4_$r.txt; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_52.json 2> gpurun_out/bench_52.err; echo "bench rc=$?"
grep "per-step" gpurun_out/bench_52.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per-step"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per-step"
