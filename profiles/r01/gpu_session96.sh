#!/bin/bash
# phase accounting of the tcgen05 update
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python profiles/prof_ozaki.py 256 > gpurun_out/prof_ozaki_k256.txt 2>&1; cat gpurun_out/prof_ozaki_k256.txt
timeout 120 python profiles/prof_ozaki.py 128 > gpurun_out/prof_ozaki_k128.txt 2>&1; cat gpurun_out/prof_ozaki_k128.txt
