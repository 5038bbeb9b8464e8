#!/bin/bash
# persistent panel solve (panel_solve2_kernel): parity, A/B against the classic kernel at C4
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r02_pytest41.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest41.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench41_c4_$name.json 2> gpurun_out/r02_bench41_c4_$name.err
  echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench41_c4_$name.json'))
    k=d['kernels']
    print('$name step %.3f e2e %.3f cond %s / %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('conditional',{}).get('mean_only_ms'), d.get('conditional',{}).get('mean_and_covariance_ms')), {n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in k.items() if any(s in n for s in ('potrf','panel','oz_','syrk'))})
except Exception as e:
    print('$name failed', e)
PY
}
run persist DRP_PANEL_PERSIST=1
run classic DRP_PANEL_PERSIST=0
run persist2 DRP_PANEL_PERSIST=1
