#!/bin/bash
# with the cheap slicing arithmetic: is writing the slices from panel_solve_kernel a gain now?  (+ stream priority check)
set -x
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench47_c4_$name.json 2> gpurun_out/r02_bench47_c4_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench47_c4_$name.json'))
    k=d['kernels']
    print('$name step %.3f e2e %.3f cond %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['conditional']['mean_only_ms']), {n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in k.items() if any(s in n for s in ('potrf','panel','oz_'))})
except Exception as e:
    print('$name failed', e)
PY
}
run base DRP_PANEL_SLICING=0
run slicing DRP_PANEL_SLICING=1
run base2 DRP_PANEL_SLICING=0
run slicing2 DRP_PANEL_SLICING=1
run w256s DRP_PANEL_WIDTH=256 DRP_PANEL_SLICING=1
run prio0 DRP_GRAPH_PRIORITY=0
