#!/bin/bash
# GPU session 11: DMMA shape microbenchmark; ncu full captures of the GRU kernels and potf2/trsm.
set -x
mkdir -p gpurun_out
./profiles/microbench/dmma_shapes > gpurun_out/dmma_shapes.txt 2>&1
cat gpurun_out/dmma_shapes.txt
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'gru_|potf2|trsm' -s 40 -c 8 -o gpurun_out/gru_potf2_trsm_c4 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gru.log 2>&1
tail -5 gpurun_out/ncu_gru.log
ls -la gpurun_out
