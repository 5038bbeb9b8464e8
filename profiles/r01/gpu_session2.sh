#!/bin/bash
# GPU session 2 of round 1: tcgen05 trailing-update parity, full parity suite, c4 bench with/without the tensor path.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/t_ozaki.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_ozaki.log
tail -5 gpurun_out/t_ozaki.log
if grep -q "rc=0" gpurun_out/t_ozaki.log; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
  tail -5 gpurun_out/t_all.log
  for s in 0 5 6; do
    DRP_OZAKI_SLICES=$s timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_s$s.json 2> gpurun_out/bench_c4_s$s.err
  done
  DRP_OZAKI_SLICES=0 DRP_OUTER_BLOCK=256 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_s0_ob256.json 2> gpurun_out/bench_c4_s0_ob256.err
  DRP_OZAKI_SLICES=0 DRP_OUTER_BLOCK=64 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_s0_ob64.json 2> gpurun_out/bench_c4_s0_ob64.err
fi
ls -la gpurun_out
