#!/bin/bash
set -x
mkdir -p gpurun_out
for w in 0 128; do
for args in "450,90c,200c" "450" "90c,200c" "200c" "200" ; do
  DRP_PANEL_WIDTH=$w timeout 200 python profiles/r02/debug_gate_backerr2.py $args 2>&1 | grep -v Warning | tail -3
done; done > gpurun_out/r02_debug35.txt 2>&1
cat gpurun_out/r02_debug35.txt
