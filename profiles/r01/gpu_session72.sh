#!/bin/bash
# final launch list of round 1 (one ncu pass per gpurun call): every kernel of this library over the steps of a short bench run
set -x
mkdir -p gpurun_out
DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain72.log 2>&1 &&
DRP_PRIOR_STREAM=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'oz_|syrk|trsm|potf2|assemble|ou_|mvn_|reduce3|cat_|sum_kernel|alpha|gru_|gram_|tall_|lca_|tree_' -c 6000 --csv --log-file gpurun_out/launches_c4_r01_final.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches72.log 2>&1
tail -2 gpurun_out/ncu_launches72.log
wc -l gpurun_out/launches_c4_r01_final.csv
