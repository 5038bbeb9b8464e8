#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest18.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r02_pytest18.log | tail -25
