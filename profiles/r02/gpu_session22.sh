#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_tree_io.py -m gpu -x -q -k "patristic or tree" > gpurun_out/r02_pytest22.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest22.log
for rpc in 4 2; do
  DRP_PATRISTIC_RPC=$rpc timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench22_c5_rpc$rpc.json 2> gpurun_out/r02_bench22_c5_rpc$rpc.err; echo "rc=$?"
  python - <<P
import json
d=json.load(open("gpurun_out/r02_bench22_c5_rpc$rpc.json"))
print("rpc=$rpc value", round(d["value"],1), "ms", round(d["ms_per_step"],3), d["roofline"]["kernel"], round(d["roofline"]["frac"],3))
for r in d["sweep"][-3:]: print(r["n_leaves"], round(r["patristic_ms"],4), round(r["patristic_gbs"],1), round(r["patristic_frac_of_hbm_peak"],3), "lca", round(r["patristic_lca_gbs"],1), r["lca_bit_exact"], "ou", round(r["ou_cov_gbs"],1))
P
done
