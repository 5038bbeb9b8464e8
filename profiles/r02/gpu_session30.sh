#!/bin/bash
# fused panel kernels (potrf_diag + panel_solve): parity, then A/B of the panel width at C4 (0 = old 64-column chain)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r02_pytest30.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest30.log
for pw in 128 0 256; do
  DRP_PANEL_WIDTH=$pw timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench30_c4_pw$pw.json 2> gpurun_out/r02_bench30_c4_pw$pw.err
  echo "bench pw=$pw rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_bench30_c4_pw$pw.json'))
    k=d['kernels']
    print('pw=$pw step %.3f e2e %.3f cond %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('conditional',{}).get('mean_only_ms')), {n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in k.items() if any(s in n for s in ('potf2','trsm','potrf','panel','oz_','syrk'))})
except Exception as e:
    print('pw=$pw failed', e)
PY
  tail -3 gpurun_out/r02_bench30_c4_pw$pw.err
done
