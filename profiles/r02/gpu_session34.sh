#!/bin/bash
set -x
mkdir -p gpurun_out
for envs in "DRP_PANEL_WIDTH=0" "DRP_PANEL_WIDTH=0 DRP_FUSED_SLICING=0" "DRP_PANEL_WIDTH=0 DRP_OZAKI_SLICES=0" "DRP_PANEL_WIDTH=128" "DRP_PANEL_WIDTH=64" "DRP_PANEL_WIDTH=256"; do
  env $envs timeout 200 python profiles/r02/debug_gate_backerr.py 2>&1 | grep -v Warning | tail -4
done > gpurun_out/r02_debug34.txt 2>&1
cat gpurun_out/r02_debug34.txt
