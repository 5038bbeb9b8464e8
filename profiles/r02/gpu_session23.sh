#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest23.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r02_pytest23.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke23.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke23.log
