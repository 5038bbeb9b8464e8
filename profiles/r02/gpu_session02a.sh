#!/bin/bash
# round 2: CUDA-graph step + fused ClippedAdam: GPU tests only
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest02.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r02_pytest02.log
