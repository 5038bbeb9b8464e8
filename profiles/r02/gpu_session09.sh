#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k patristic > gpurun_out/r02_pytest09.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest09.log
for rpc in 4 2; do
  DRP_PATRISTIC_RPC=$rpc timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench09_c5_rpc$rpc.json 2> gpurun_out/r02_bench09_c5_rpc$rpc.err; echo "rc=$?"
  python - <<P
import json
d=json.load(open("gpurun_out/r02_bench09_c5_rpc$rpc.json"))
print("rpc=$rpc value", d["value"], "ms", d["ms_per_step"])
for r in d["sweep"]: print(r["n_leaves"], round(r["patristic_ms"],4), round(r["patristic_gbs"],1), round(r["patristic_frac_of_hbm_peak"],3), "lca", round(r["patristic_lca_gbs"],1), r["lca_bit_exact"], "ou", round(r["ou_cov_gbs"],1))
P
done
