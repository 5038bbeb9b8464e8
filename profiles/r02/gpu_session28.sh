#!/bin/bash
# cluster-pair GRU kernels v2 (st.async + mbarrier instead of barrier.cluster): parity, phase trace, A/B at C1 / C2
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gru.py -x -q -m gpu > gpurun_out/r02_pytest28.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest28.log
timeout 150 python profiles/trace_gru.py cluster 250 > gpurun_out/r02_trace_gru_cluster_v4.txt 2>&1; grep -v "^warp 8\|^warp 7\|issue |" gpurun_out/r02_trace_gru_cluster_v4.txt | head -20
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -m gpu > gpurun_out/r02_pytest28b.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest28b.log
for w in c1 c2; do
  for c in 0 1; do
    DRP_GRU_CLUSTER=$c timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench28_${w}_cl$c.json 2> gpurun_out/r02_bench28_${w}_cl$c.err
    python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench28_${w}_cl$c.json'))
k=d['kernels']
print('$w cluster=$c step %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), {n: round(v['ms_per_step'],3) for n,v in k.items() if 'gru' in n})
PY
  done
done
