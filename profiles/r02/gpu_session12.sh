#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r02_pytest12.log | tail -4
timeout 300 python profiles/r02/torch_kernels.py 19 225 delta_map > gpurun_out/r02_torch_kernels_c1_v2.txt 2>&1
head -3 gpurun_out/r02_torch_kernels_c1_v2.txt | tail -1
timeout 300 python profiles/r02/torch_kernels.py 200 300 variational > gpurun_out/r02_torch_kernels_c2_v2.txt 2>&1
head -3 gpurun_out/r02_torch_kernels_c2_v2.txt | tail -1
for w in c1 c2 c3 c4; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench12_$w.json 2> gpurun_out/r02_bench12_$w.err; echo "bench $w rc=$?"
  grep "per-step" gpurun_out/r02_bench12_$w.err | head -2
done
