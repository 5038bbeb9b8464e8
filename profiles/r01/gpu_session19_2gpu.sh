#!/bin/bash
# 2-GPU session: z-sharded OU prior + leaf-sharded GRUs over NCCL (c4 and c3), and the N=1 line on the same box.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi2.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err
tail -5 gpurun_out/bench_c4_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err
cat gpurun_out/bench_c4_n2.json gpurun_out/bench_c3_n2.json
