#!/usr/bin/env python
"""Does device memory grow from step to step while the cyclic garbage collector is off (as inside bench.py's timed
regions)?  Prints allocated / reserved bytes, cudaMalloc counts and step times for 40 steps."""
import gc
import os
import sys
import time

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
gc.collect()
gc.disable()
for i in range(40):
    t0 = time.perf_counter()
    run.step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    st = torch.cuda.memory_stats()
    print(f"step {i:2d} {dt:7.2f} ms  allocated {torch.cuda.memory_allocated() / 2**30:7.2f} GiB  reserved {torch.cuda.memory_reserved() / 2**30:7.2f} GiB  "
          f"cudaMalloc calls {st['num_device_alloc']}  frees {st['num_device_free']}  gc counts {gc.get_count()}")
t0 = time.perf_counter()
n = gc.collect()
print(f"gc.collect(): {n} objects, {(time.perf_counter() - t0) * 1e3:.1f} ms; allocated after {torch.cuda.memory_allocated() / 2**30:.2f} GiB")
