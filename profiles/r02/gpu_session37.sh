#!/bin/bash
# gram_tn with rectangular warp blocks + adaptive stage rows: parity, bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest37.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest37.log
for i in 1 2; do
timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench37_c4_$i.json 2> gpurun_out/r02_bench37_c4_$i.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench37_c4_$i.json'))
k=d['kernels']
print('step %.3f e2e %.3f cond %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('conditional',{}).get('mean_only_ms')), {n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in k.items()})
PY
done
