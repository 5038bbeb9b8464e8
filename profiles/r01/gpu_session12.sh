#!/bin/bash
# GPU session 12: elect-one MMA issue; parity + bench; ncu full capture of GRU kernels and oz_syrk.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/t_ozaki.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_ozaki.log
tail -5 gpurun_out/t_ozaki.log
grep -q "rc=0" gpurun_out/t_ozaki.log || exit 0
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'gru_fwd|gru_bwd|gru_wgrad_kernel|oz_syrk_kernel<6, 2>' -s 22 -c 8 -o gpurun_out/gru_oz_c4 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gru.log 2>&1
tail -5 gpurun_out/ncu_gru.log
ls -la gpurun_out
