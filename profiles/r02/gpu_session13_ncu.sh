#!/bin/bash
# ncu evidence for round 2: launch list of the bench command, then --set full captures of the top kernels (one GPU).
# Only CSV summaries travel back (gpurun_out is limited to 64 MiB); the tcgen05 kernel's report is kept as well.
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline"
KEYS='Kernel Name|launch__grid_size|launch__block_size|launch__registers_per_thread|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active|sm__pipe_tensor_subpipe_imma_cycles_active|sm__inst_executed_pipe_tensor_subpipe_dmma|sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum|sm__cycles_active.avg|l1tex__throughput'
summ() {   # $1 = report stem
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2> /dev/null
  python profiles/summarize_ncu_csv.py gpurun_out/$1.raw.csv > gpurun_out/$1.summary.txt
  rm -f gpurun_out/$1.raw.csv
}
$CMD > gpurun_out/r02_ncu_plain.json 2> gpurun_out/r02_ncu_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_ncu_launches_c4.csv $CMD > gpurun_out/r02_ncu_launches.out 2> gpurun_out/r02_ncu_launches.err
echo "launch list rc=$?"; tail -5 gpurun_out/r02_ncu_launches.err; tail -3 gpurun_out/r02_ncu_launches.out
$CMD > gpurun_out/r02_ncu_plain.json 2> gpurun_out/r02_ncu_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:oz_syrk_kernel -s 62 -c 8 -o gpurun_out/r02_ncu_oz_syrk_c4 -f $CMD > gpurun_out/r02_ncu_oz.out 2> gpurun_out/r02_ncu_oz.err
echo "oz rc=$?"; summ r02_ncu_oz_syrk_c4
$CMD > gpurun_out/r02_ncu_plain.json 2> gpurun_out/r02_ncu_plain.err &&
ncu --set full --clock-control none -k regex:"gru_(fwd2|bwd|wgrad)_kernel" -s 8 -c 8 -o gpurun_out/r02_ncu_gru_c4 -f $CMD > gpurun_out/r02_ncu_gru.out 2> gpurun_out/r02_ncu_gru.err
echo "gru rc=$?"; summ r02_ncu_gru_c4; rm -f gpurun_out/r02_ncu_gru_c4.ncu-rep
$CMD > gpurun_out/r02_ncu_plain.json 2> gpurun_out/r02_ncu_plain.err &&
ncu --set full --clock-control none -k regex:"(oz_slice|trsm|potf2|tall_gemm|gram_tn|gru_input_sums|ou_assemble|ou_grad)_kernel" -s 120 -c 24 -o gpurun_out/r02_ncu_misc_c4 -f $CMD > gpurun_out/r02_ncu_misc.out 2> gpurun_out/r02_ncu_misc.err
echo "misc rc=$?"; summ r02_ncu_misc_c4; rm -f gpurun_out/r02_ncu_misc_c4.ncu-rep
du -sh gpurun_out; ls -la gpurun_out
