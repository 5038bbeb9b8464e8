#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python profiles/trace_ozaki.py 256 2>&1 | head -5 | cut -c1-300
timeout 120 python profiles/trace_ozaki.py 128 2>&1 | head -5 | cut -c1-300
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
