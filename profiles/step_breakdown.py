#!/usr/bin/env python
"""Where one SVI step spends its device time (torch.profiler, CUDA activities): our kernels vs the torch/cuDNN ops that
remain on the path (GRU, Linear, optimiser).  python profiles/step_breakdown.py [workload] > gpurun_out/step_breakdown.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
import bench  # noqa: E402
from draupnir_asr_b200 import runtime, synthetic  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c4"]
problem = synthetic.make_problem(w["n"], w["L"], seed=0)
init = synthetic.make_parameters(w["n"], z_dim=30, seed=0)
run = runtime.build_run(problem, w["guide"], device="cuda:0", init=init, z_dim=30)
for _ in range(3):
    run.step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        run.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
