#!/bin/bash
# last check of the committed tree: GPU tests, smoke, one default bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final6.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu_final6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final6.log 2>&1; echo "smoke rc=$?"
tail -1 gpurun_out/smoke_final6.log
