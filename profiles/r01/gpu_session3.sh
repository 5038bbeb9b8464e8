#!/bin/bash
# GPU session 3: re-check tcgen05 parity after the epilogue rewrite, bench c4, launch list + full ncu capture of oz_syrk.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ozaki.py -x -q > gpurun_out/t_ozaki.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_ozaki.log
tail -5 gpurun_out/t_ozaki.log
grep -q "rc=0" gpurun_out/t_ozaki.log || exit 0
for s in 5 6; do
  DRP_OZAKI_SLICES=$s timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_s$s.json 2> gpurun_out/bench_c4_s$s.err
done
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'oz_|syrk|trsm|potf2|assemble|ou_|mvn_|reduce3|catlik|alpha' -c 3000 --csv --log-file gpurun_out/launches_c4_ours.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:oz_syrk -s 12 -c 2 -o gpurun_out/oz_syrk_c4 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_oz.log 2>&1
ls -la gpurun_out
