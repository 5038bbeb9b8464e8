#!/bin/bash
# 4-GPU box: c4 at N = 1, 2, 4 back to back (the driver's scaling run does the same).
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_c4_n1.json 2> gpurun_out/scale_c4_n1.err
for n in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_c4_n$n.json 2> gpurun_out/scale_c4_n$n.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --workload c3 --steps 10 --warmup 3 > gpurun_out/scale_c3_n4.json 2> gpurun_out/scale_c3_n4.err
tail -2 gpurun_out/scale_c4_n4.err
