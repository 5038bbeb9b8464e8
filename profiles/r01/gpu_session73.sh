#!/bin/bash
# final full capture of the dominant kernels (one ncu pass per gpurun call)
set -x
mkdir -p gpurun_out
DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain73.log 2>&1 &&
DRP_PRIOR_STREAM=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'gru_fwd2|gru_bwd|gru_wgrad_kernel|oz_syrk_kernel<6, 2>|tall_gemm_kernel<23|gram_tn|gru_input_sums' -s 40 -c 16 -o gpurun_out/final_r01_c4 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final73.log 2>&1
tail -3 gpurun_out/ncu_final73.log
ls -la gpurun_out/final_r01_c4.ncu-rep
