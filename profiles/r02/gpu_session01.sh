#!/bin/bash
# round 2, first call: Phase A (precision gate, large-size parity tests, z-sharded K6) on the GPU
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest01.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_pytest01.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke01.log 2>&1; echo "smoke rc=$?"
tail -1 gpurun_out/r02_smoke01.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench01.json 2> gpurun_out/r02_bench01.err; echo "bench rc=$?"
cat gpurun_out/r02_bench01.json | head -c 3000
