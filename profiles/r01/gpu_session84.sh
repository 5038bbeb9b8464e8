#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python profiles/epoch_extras.py c1 c2 c4 > gpurun_out/epoch_extras.txt 2>&1; tail -4 gpurun_out/epoch_extras.txt
