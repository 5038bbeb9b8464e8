#!/bin/bash
# fused panel kernels v2: slices written by panel_solve, rcp-based potf2 column chain, triangular block products
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest31.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest31.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench31_c4_$name.json 2> gpurun_out/r02_bench31_c4_$name.err
  echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench31_c4_$name.json'))
    k=d['kernels']
    print('$name step %.3f e2e %.3f cond %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('conditional',{}).get('mean_only_ms')), {n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in k.items() if any(s in n for s in ('potf2','trsm','potrf','panel','oz_','syrk'))})
except Exception as e:
    print('$name failed', e)
PY
}
run pw128 DRP_PANEL_WIDTH=128
run pw128_nofs DRP_PANEL_WIDTH=128 DRP_FUSED_SLICING=0
run pw256 DRP_PANEL_WIDTH=256
run pw0 DRP_PANEL_WIDTH=0
for w in c1 c2 c3; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench31_$w.json 2> gpurun_out/r02_bench31_$w.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench31_$w.json'))
print('$w step %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']))
PY
done
