#!/bin/bash
# does a tcgen05.commit between MMAs cost tensor-pipe time?
set -x
mkdir -p gpurun_out
timeout 20 profiles/microbench/imma_rate 148 > gpurun_out/imma_rate_148cta_v2.txt 2>&1; echo "rc=$?"; cat gpurun_out/imma_rate_148cta_v2.txt
