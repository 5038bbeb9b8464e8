#!/bin/bash
# gram_tn templated on the warp block shape: parity of the dense-layer kernels, bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gru.py -x -q -m gpu > gpurun_out/r02_pytest38.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest38.log
timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench38_c4.json 2> gpurun_out/r02_bench38_c4.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench38_c4.json'))
k=d['kernels']
print('step %.3f e2e %.3f cond %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('conditional',{}).get('mean_only_ms')), {n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in k.items() if 'gram' in n or 'tall' in n})
PY
