#!/bin/bash
# N = 2: NCCL collectives inside the captured step
set -x
mkdir -p gpurun_out
for w in c3 c4; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload $w --steps 10 --warmup 3 > gpurun_out/r02_scale_${w}_n2.json 2> gpurun_out/r02_scale_${w}_n2.err; echo "rc=$?"
  grep "per-step\|warm-up\|Warning\|Error" gpurun_out/r02_scale_${w}_n2.err | head -20
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_scale_${w}_n2.json")); print("$w n2", d["ms_per_step"], d["e2e"]["ms_per_step"], d["cuda_graph"], d["conditional"]["mean_only_ms"])
except Exception as e: print("fail", e)
P
done
