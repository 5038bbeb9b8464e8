#!/bin/bash
# tcgen05.ld throughput microbenchmark
set -x
mkdir -p gpurun_out
timeout 20 profiles/microbench/tmem_ld_rate 1 > gpurun_out/tmem_ld_rate_1cta.txt 2>&1; echo "rc=$?"; cat gpurun_out/tmem_ld_rate_1cta.txt
timeout 20 profiles/microbench/tmem_ld_rate 148 > gpurun_out/tmem_ld_rate_148cta.txt 2>&1; echo "rc=$?"; cat gpurun_out/tmem_ld_rate_148cta.txt
