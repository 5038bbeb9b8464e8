#!/bin/bash
# 4-GPU box: c4 at N = 2, 4 with the final build (two-phase prior from the guide, side stream, dynamic work units)
set -x
mkdir -p gpurun_out
for n in 2 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale4_c4_n$n.json 2> gpurun_out/scale4_c4_n$n.err
  grep "rank 0.*per-step\|rank 0.*ELBO" gpurun_out/scale4_c4_n$n.err | head -4
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 4 --workload c3 --steps 10 --warmup 3 > gpurun_out/scale4_c3_n4.json 2> gpurun_out/scale4_c3_n4.err
grep "rank 0.*per-step\|rank 0.*ELBO" gpurun_out/scale4_c3_n4.err | head -3
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "per-step\|ELBO"
