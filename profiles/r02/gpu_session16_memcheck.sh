#!/bin/bash
# memcheck of the tall-skinny dense-layer kernels (an ncu launch-list run of the bench stopped with LaunchFailed at tall_gemm_kernel<3,1>)
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_gru.py -m gpu -x -q -k "tall_gemm or gram or input_sums" > gpurun_out/r02_memcheck_tall.log 2>&1; echo "memcheck rc=$?"
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r02_memcheck_tall.log; grep -A12 "Invalid" gpurun_out/r02_memcheck_tall.log | head -60; tail -5 gpurun_out/r02_memcheck_tall.log
