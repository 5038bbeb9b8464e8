#!/bin/bash
# CUDA path against the committed vectors of the reference's own code
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -15
