#!/bin/bash
# GPU session 36 (final captures of round 1): launch list of one step, full ncu capture of the dominant kernels.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'oz_|syrk|trsm|potf2|assemble|ou_|mvn_|reduce3|cat_|sum_kernel|alpha|gru_|lca_|tree_' -c 4000 --csv --log-file gpurun_out/launches_c4_final.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'gru_fwd|gru_bwd|gru_wgrad_kernel|oz_syrk' -s 30 -c 12 -o gpurun_out/final_gru_oz_c4 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
tail -3 gpurun_out/ncu_final.log
ls -la gpurun_out | tail -8
