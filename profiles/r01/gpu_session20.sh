#!/bin/bash
# GPU session 20: GRU kernels templated on rows per CTA (8/16/32) - parity for every variant, benches c1-c4.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
grep -q "rc=0" gpurun_out/t_all.log || exit 0
for w in c4 c3 c2 c1; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
DRP_GRU_ROWS=16 timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_rows16.json 2> gpurun_out/bench_c4_rows16.err
ls -la gpurun_out
