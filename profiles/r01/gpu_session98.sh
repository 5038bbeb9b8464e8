#!/bin/bash
# final consolidation of round 1 (fourth pass: reference-pinned fixtures, full-size parity targets, epilogue fast path)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final4.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_final4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final4.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke_final4.log
timeout 900 python bench.py > gpurun_out/bench_default_final4.json 2> gpurun_out/bench_default_final4.err; echo "bench rc=$?"
cat gpurun_out/bench_default_final4.json
grep "per-step" gpurun_out/bench_default_final4.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final4.json 2> gpurun_out/bench_reference_final4.err; echo "ref rc=$?"
cat gpurun_out/bench_reference_final4.json
for c in c1 c2 c3; do
  timeout 600 python bench.py --workload $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_final4.json 2> gpurun_out/bench_${c}_final4.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${c}_final4.json')); print('$c', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default_final4_nostream.json 2> gpurun_out/bench_default_final4_nostream.err
python -c "
import json; d=json.load(open('gpurun_out/bench_default_final4_nostream.json')); print('nostream', d['ms_per_step']); print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernel_classes'].items()})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/ncu_launches_c4_r01_final4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_final4.log 2>&1; echo "ncu rc=$?"
