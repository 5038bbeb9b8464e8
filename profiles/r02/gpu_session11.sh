#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r02_pytest11.log | tail -4
timeout 300 python profiles/r02/torch_kernels.py 200 300 variational > gpurun_out/r02_torch_kernels_c2.txt 2>&1
head -70 gpurun_out/r02_torch_kernels_c2.txt
timeout 300 python profiles/r02/torch_kernels.py 19 225 delta_map > gpurun_out/r02_torch_kernels_c1.txt 2>&1
head -40 gpurun_out/r02_torch_kernels_c1.txt
timeout 900 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench11_c4.json 2> gpurun_out/r02_bench11_c4.err; echo "bench rc=$?"
grep "per-step" gpurun_out/r02_bench11_c4.err
