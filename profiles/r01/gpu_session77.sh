#!/bin/bash
set -x
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep per-step | head -1; done
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep per-step | head -1; done
