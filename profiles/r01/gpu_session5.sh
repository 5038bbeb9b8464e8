#!/bin/bash
# GPU session 5: full parity suite with the three-level blocking, bench, step breakdown (torch.profiler), ncu of oz_syrk.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
grep -q "rc=0" gpurun_out/t_all.log || exit 0
DRP_OZAKI_SLICES=6 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_s6.json 2> gpurun_out/bench_c4_s6.err
timeout 600 python profiles/step_breakdown.py c4 > gpurun_out/step_breakdown_c4.txt 2>&1
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:oz_syrk -s 20 -c 2 -o gpurun_out/oz_syrk_c4_v3 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_oz.log 2>&1
ls -la gpurun_out
