#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gru.py -m gpu -x -q 2>&1 | tail -3
DRP_PRIOR_STREAM=0 timeout 600 python profiles/step_breakdown.py c4 > gpurun_out/step_breakdown_c4_v13_nostream.txt 2>&1; echo "breakdown rc=$?"
