#!/bin/bash
# tcgen05 MMA rate microbenchmark, elected-lane issue
set -x
mkdir -p gpurun_out
timeout 60 profiles/microbench/imma_rate 1 > gpurun_out/imma_rate_1cta.txt 2>&1; echo "rc=$?"; cat gpurun_out/imma_rate_1cta.txt
timeout 60 profiles/microbench/imma_rate 148 > gpurun_out/imma_rate_148cta.txt 2>&1; echo "rc=$?"; cat gpurun_out/imma_rate_148cta.txt
