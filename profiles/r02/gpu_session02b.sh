#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python profiles/r02/debug_graph.py > gpurun_out/r02_debug_graph.log 2>&1; echo "rc=$?"
tail -120 gpurun_out/r02_debug_graph.log
