"""Which kernels does one eager svi.step launch, by name: library (drp) vs torch; counts and device time (torch profiler)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
from draupnir_asr_b200 import runtime, synthetic as S
from torch.profiler import profile, ProfilerActivity

n, L, guide = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
problem = S.make_problem(n, L, seed=0)
run = runtime.build_run(problem, guide, device="cuda:0", init=S.make_parameters(n, seed=0), cuda_graph=False)
for _ in range(3):
    run.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        run.step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg.setdefault(ev.name[:90], [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot_n = sum(a[0] for a in agg.values()) / 3
tot_t = sum(a[1] for a in agg.values()) / 3
print(f"# {guide} n={n} L={L}: {tot_n:.0f} kernels / memcpys per step, {tot_t / 1e3:.3f} ms of device time")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{c / 3:7.1f} x  {t / 3:9.1f} us  {name}")
