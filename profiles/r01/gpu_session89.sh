#!/bin/bash
# B ring of up to 16 stages (K = 128 updates): parity, one update at K = 256 / 128, C4 step with outer blocks of 256 / 128
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python profiles/trace_ozaki.py 256 2>&1 | head -1
timeout 120 python profiles/trace_ozaki.py 128 > gpurun_out/trace_ozaki_v8_k128.txt 2>&1; head -1 gpurun_out/trace_ozaki_v8_k128.txt
for ob in 256 128; do
DRP_OUTER_BLOCK=$ob DRP_PRIOR_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_89_ob${ob}_nostream.json 2> gpurun_out/bench_89_ob${ob}_nostream.err
python -c "
import json; d=json.load(open('gpurun_out/bench_89_ob${ob}_nostream.json')); print('outer', $ob, 'nostream', d['ms_per_step']); print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['kernel_classes'].items()})"
done
DRP_OUTER_BLOCK=128 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_89_ob128.json 2> gpurun_out/bench_89_ob128.err
python -c "
import json; d=json.load(open('gpurun_out/bench_89_ob128.json')); print('outer 128 stream', d['ms_per_step'], d['e2e']['ms_per_step'])"
