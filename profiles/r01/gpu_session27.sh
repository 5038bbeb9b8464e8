#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python profiles/trace_patristic.py > gpurun_out/trace_patristic.txt 2>&1
cat gpurun_out/trace_patristic.txt
