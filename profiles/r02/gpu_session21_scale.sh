#!/bin/bash
# scaling check of the final build: C4 on N ranks (N = number of visible GPUs)
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for w in c4 c3; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --workload $w --steps 10 --warmup 3 > gpurun_out/r02_scale_${w}_n$N.json 2> gpurun_out/r02_scale_${w}_n$N.err; echo "rc=$?"
grep "rank 0\] per-step\|Error" gpurun_out/r02_scale_${w}_n$N.err | head -6
python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_scale_${w}_n$N.json")); print("$w n$N", round(d["value"],2), "steps/s", round(d["ms_per_step"],3), "ms; e2e", round(d["e2e"]["ms_per_step"],3), d["cuda_graph"], "eager", round(d["eager_ms_per_step"],3), "cond", round(d["conditional"]["mean_only_ms"],3), round(d["conditional"]["mean_and_covariance_ms"],3), "h2d", d["e2e"]["h2d_bytes_per_step"])
except Exception as e: print("fail", e)
P
done
