#!/bin/bash
# panel_solve_kernel: column tiles {0,1,6,7} / {2,3,4,5} per warp half (balanced triangular products)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r02_pytest42.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest42.log
for i in 1 2; do
timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench42_c4_$i.json 2> gpurun_out/r02_bench42_c4_$i.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench42_c4_$i.json'))
k=d['kernels']
print('step %.3f e2e %.3f cond %s / %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('conditional',{}).get('mean_only_ms'), d.get('conditional',{}).get('mean_and_covariance_ms')), {n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in k.items() if any(s in n for s in ('potrf','panel','oz_','syrk'))})
PY
done
