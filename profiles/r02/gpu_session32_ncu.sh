#!/bin/bash
# ncu evidence for the fused panel kernels + a fresh launch list of the bench command (one GPU)
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline"
summ() {   # $1 = report stem
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2> /dev/null
  python profiles/summarize_ncu_csv.py gpurun_out/$1.raw.csv > gpurun_out/$1.summary.txt
  rm -f gpurun_out/$1.raw.csv
}
$CMD > gpurun_out/r02_ncu_plain.json 2> gpurun_out/r02_ncu_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"(panel_solve|potrf_diag)_kernel" -s 40 -c 8 -o gpurun_out/r02_ncu_panel_c4 -f $CMD > gpurun_out/r02_ncu_panel.out 2> gpurun_out/r02_ncu_panel.err
echo "panel rc=$?"; summ r02_ncu_panel_c4; cat gpurun_out/r02_ncu_panel_c4.summary.txt | head -60
ncu -i gpurun_out/r02_ncu_panel_c4.ncu-rep --page source --csv > gpurun_out/r02_ncu_panel_c4.source.csv 2>/dev/null; wc -l gpurun_out/r02_ncu_panel_c4.source.csv
$CMD > gpurun_out/r02_ncu_plain.json 2> gpurun_out/r02_ncu_plain.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!.*tall_gemm).*$' -c 7000 --csv --log-file gpurun_out/r02_ncu_launches_c4_v2.csv $CMD > gpurun_out/r02_ncu_launches.out 2> gpurun_out/r02_ncu_launches.err
echo "launch list rc=$?"; tail -3 gpurun_out/r02_ncu_launches.err; wc -l gpurun_out/r02_ncu_launches_c4_v2.csv
du -sh gpurun_out
