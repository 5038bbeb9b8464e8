#!/bin/bash
# 2-GPU sanity of the final build: C4, identical ELBO, scaling line
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/scale2_c4_n2_final4.json 2> gpurun_out/scale2_c4_n2_final4.err; echo "rc=$?"
grep "rank 0.*per-step\|rank 0.*ELBO" gpurun_out/scale2_c4_n2_final4.err | head -4
cat gpurun_out/scale2_c4_n2_final4.json | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29732 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-300
