#!/bin/bash
# ou_grad: one warp per row, eight rows per CTA, unrolled row walk: parity of the prior paths + bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_ozaki.py tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r02_pytest45.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest45.log
timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench45_c4.json 2> gpurun_out/r02_bench45_c4.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench45_c4.json'))
k=d['kernels']
print('step %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), {n: (round(v['ms_per_step'],3), v['launches_per_step']) for n,v in k.items() if any(s in n for s in ('ou_','cat','small'))})
PY
