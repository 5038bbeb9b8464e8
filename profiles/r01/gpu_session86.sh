#!/bin/bash
# tcgen05 MMA rate microbenchmark (round-2 groundwork) + full-size parity against the reference's own losses
set -x
mkdir -p gpurun_out
timeout 60 profiles/microbench/imma_rate 1 > gpurun_out/imma_rate_1cta.txt 2>&1; echo "rc=$?"; cat gpurun_out/imma_rate_1cta.txt
timeout 60 profiles/microbench/imma_rate 148 > gpurun_out/imma_rate_148cta.txt 2>&1; echo "rc=$?"; cat gpurun_out/imma_rate_148cta.txt
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "full_size or reference_vectors" 2>&1 | tail -5
