#!/bin/bash
# after the CTA-pair GRU kernels + pipelined host-fed step: whole GPU suite, smoke, bench lines C1-C4, phase trace
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest29.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest29.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02_smoke29.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke29.log
timeout 150 python profiles/trace_gru.py cluster 250 > gpurun_out/r02_trace_gru_cluster_v3b.txt 2>&1
for w in c1 c2 c3 c4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02_bench29_$w.json 2> gpurun_out/r02_bench29_$w.err
  echo "bench $w rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench29_$w.json'))
k=d['kernels']
print('$w step %.3f e2e %.3f (copy inside %.3f) cpu %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_copy_inside_step'], d['cpu_baseline'] and d['cpu_baseline']['value']), {n: round(v['ms_per_step'],3) for n,v in k.items() if 'gru' in n})
PY
done
