#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5
for c in c1 c2; do timeout 300 python bench.py --workload $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep per-step; done
