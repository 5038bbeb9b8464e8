#!/bin/bash
# full ncu capture of the tcgen05 update of the final build (interior fast path); the plain command ran in session 100
set -x
mkdir -p gpurun_out
DRP_PRIOR_STREAM=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'oz_syrk_kernel' -s 24 -c 4 -o gpurun_out/oz_syrk_c4_final4 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final101.log 2>&1
tail -3 gpurun_out/ncu_final101.log
ls -la gpurun_out/oz_syrk_c4_final4.ncu-rep
