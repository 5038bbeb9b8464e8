#!/bin/bash
# last check of the committed tree: whole GPU suite + smoke (one GPU)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest46.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest46.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02_smoke46.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke46.log
