#!/bin/bash
# conflict-free panel layout in potf2, 8-chain row solve, two-operation quantisation + PRMT byte picks in the slicers
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest33.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest33.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench33_c4_$name.json 2> gpurun_out/r02_bench33_c4_$name.err
  echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench33_c4_$name.json'))
    k=d['kernels']
    print('$name step %.3f e2e %.3f cond %s / %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('conditional',{}).get('mean_only_ms'), d.get('conditional',{}).get('mean_and_covariance_ms')), {n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in k.items() if any(s in n for s in ('potf2','trsm','potrf','panel','oz_','syrk'))})
except Exception as e:
    print('$name failed', e)
PY
}
run pw128 DRP_PANEL_WIDTH=128
run pw0 DRP_PANEL_WIDTH=0
run pw128b DRP_PANEL_WIDTH=128
