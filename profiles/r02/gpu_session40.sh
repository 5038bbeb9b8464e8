#!/bin/bash
# how often does a child of the widths test fail?  (one unreproduced failure in session 33)
set -x
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 600 python -m pytest tests/test_gpu_ozaki.py -q -m gpu -k "fused_panel_widths or precision_gate" -W always 2>&1 | grep -E "passed|failed|retrying|Warning: DRP" | head -5
done > gpurun_out/r02_flake40.txt 2>&1
cat gpurun_out/r02_flake40.txt
