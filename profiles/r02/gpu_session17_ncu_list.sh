#!/bin/bash
# ncu launch list of the bench command (kernel shares of a step).  The first attempt (every kernel profiled) stopped with
# LaunchFailed when tall_gemm_kernel<3,1> was profiled; this one profiles everything except tall_gemm_kernel.
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r02_ncu_plain.json 2> gpurun_out/r02_ncu_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(?!.*tall_gemm).*$' -c 9000 --csv --log-file gpurun_out/r02_ncu_launches_c4.csv $CMD > gpurun_out/r02_ncu_launches.out 2> gpurun_out/r02_ncu_launches.err
echo "launch list rc=$?"; tail -3 gpurun_out/r02_ncu_launches.err; wc -l gpurun_out/r02_ncu_launches_c4.csv; grep -c ERROR gpurun_out/r02_ncu_launches_c4.csv
