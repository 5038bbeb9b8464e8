#!/usr/bin/env python
"""Which Python lines cause host->device copies (pageable: they stall the launch thread) inside one SVI step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
import bench  # noqa: E402
from draupnir_asr_b200 import runtime, synthetic  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
problem = synthetic.make_problem(w["n"], w["L"], seed=0)
init = synthetic.make_parameters(w["n"], z_dim=30, seed=0)
run = runtime.build_run(problem, w["guide"], device="cuda:0", init=init, z_dim=30)
for _ in range(3):
    run.step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    run.step()
    torch.cuda.synchronize()
import collections
cnt = collections.Counter()
for e in prof.events():
    if e.name in ("aten::_to_copy", "aten::copy_", "aten::scalar_tensor", "aten::lift_fresh", "aten::item", "aten::_local_scalar_dense", "aten::nonzero", "aten::is_nonzero"):
        st = [s for s in (e.stack or []) if "draupnir_asr_b200" in s or "distributions" in s]
        cnt[(e.name, tuple(st[:4]))] += 1
for (nm, st), c in cnt.most_common(40):
    print(c, nm)
    for s in st:
        print("      ", s)
