#!/bin/bash
# 8-GPU box: c4 at N = 1, 2, 4, 8 back to back (the driver's scaling run does the same)
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale8_c4_n1.json 2> gpurun_out/scale8_c4_n1.err
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale8_c4_n$n.json 2> gpurun_out/scale8_c4_n$n.err
  grep "per-step\|ELBO" gpurun_out/scale8_c4_n$n.err | head -4
done
grep "ELBO" gpurun_out/scale8_c4_n1.err
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/scale8_c4_n{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],2), round(d['value'],2), d['e2e']['ms_per_step'])
    except Exception as e: print(n,'ERR',e)
PY
