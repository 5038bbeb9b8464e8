#!/bin/bash
# N = 8, C4: scaling check with the captured step (NCCL all-reduce + all-gather inside the graph)
set -x
mkdir -p gpurun_out
w=c4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --workload $w --steps 10 --warmup 3 > gpurun_out/r02_scale_${w}_n8.json 2> gpurun_out/r02_scale_${w}_n8.err; echo "rc=$?"
grep "rank 0\] per-step\|rank 0\] warm-up\|Error" gpurun_out/r02_scale_${w}_n8.err | head -20
python - <<P
import json
try:
    d=json.load(open("gpurun_out/r02_scale_${w}_n8.json")); print("$w n8", d["ms_per_step"], d["e2e"]["ms_per_step"], d["cuda_graph"], d["eager_ms_per_step"], d["conditional"]["mean_only_ms"], d["conditional"]["mean_and_covariance_ms"])
    for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"])[:14]:
        print("   %-24s %8.3f ms %6.1f launches" % (k, v["ms_per_step"], v["launches_per_step"]))
except Exception as e: print("fail", e)
P
