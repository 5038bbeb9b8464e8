#!/bin/bash
# GPU session 4: tcgen05 parity after epilogue v3, bench c4, C5 sweep.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/t_ozaki.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_ozaki.log
tail -5 gpurun_out/t_ozaki.log
grep -q "rc=0" gpurun_out/t_ozaki.log || exit 0
for s in 5 6; do
  DRP_OZAKI_SLICES=$s timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_s$s.json 2> gpurun_out/bench_c4_s$s.err
done
timeout 900 python tests/sweep_c5.py > gpurun_out/sweep_c5.json 2> gpurun_out/sweep_c5.err
ls -la gpurun_out
