#!/usr/bin/env python
"""GPU idle gaps inside one SVI step (torch.profiler kernel intervals): how much of the step is the device waiting for
the host, and between which kernels.   DRP_PRIOR_STREAM=0 python profiles/step_gaps.py [workload]"""
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
    run.step()
    torch.cuda.synchronize()
    run.step()
    torch.cuda.synchronize()
ks = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        ks.append((e.time_range.start, e.time_range.end, e.name))
ks.sort()
# second step only: kernels after the largest gap (the synchronize between the steps)
gaps = [(ks[i + 1][0] - max(k[1] for k in ks[:i + 1]), i) for i in range(len(ks) - 1)]
split = max(gaps)[1] + 1
ks = ks[split:]
span = ks[-1][1] - ks[0][0]
busy_end, busy, gl = ks[0][0], 0.0, []
for s, e, nm in ks:
    if s > busy_end:
        gl.append((s - busy_end, prev, nm))
        busy_end = s
    if e > busy_end:
        busy += e - max(s, busy_end)
        busy_end = e
    prev = nm
print(f"step span {span / 1e3:.2f} ms, device busy {busy / 1e3:.2f} ms, idle {(span - busy) / 1e3:.2f} ms in {len(gl)} gaps; {len(ks)} kernels")
print("largest gaps (us): after -> before")
for g, a, b in sorted(gl, reverse=True)[:25]:
    print(f"{g:9.1f}  {a[:70]:70s} -> {b[:70]}")
import collections
tot = collections.Counter()
for g, a, b in gl:
    tot[b[:60]] += g
print("idle time by the kernel that ends the gap (us):")
for nm, g in tot.most_common(15):
    print(f"{g:9.1f}  {nm}")
