#!/bin/bash
set -x
mkdir -p gpurun_out
for w in 0 128; do
  DRP_PANEL_WIDTH=$w timeout 400 python profiles/r02/debug_gate_stress.py 400 2>&1 | grep -v Warning | tail -12
done > gpurun_out/r02_debug36.txt 2>&1
cat gpurun_out/r02_debug36.txt
