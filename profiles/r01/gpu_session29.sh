#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q > gpurun_out/t_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_k.log
tail -3 gpurun_out/t_k.log
timeout 300 python profiles/trace_patristic.py > gpurun_out/trace_patristic.txt 2>&1
cat gpurun_out/trace_patristic.txt | tail -2
timeout 900 python tests/sweep_c5.py > gpurun_out/sweep_c5.json 2> gpurun_out/sweep_c5.err
