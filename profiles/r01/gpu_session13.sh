#!/bin/bash
# GPU session 13: ncu full capture of oz_syrk (v5, elect-one issue).
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:oz_syrk -s 52 -c 6 -o gpurun_out/oz_syrk_c4_v5 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_oz.log 2>&1
tail -5 gpurun_out/ncu_oz.log
ls -la gpurun_out
