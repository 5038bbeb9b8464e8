"""Capture diagnostics: two sets of device-resident arguments alternate; each should get its own captured graph."""
import os, sys, time, warnings, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
from draupnir_asr_b200 import runtime, synthetic as S

warnings.filterwarnings("error", message=".*operation not permitted.*")
guide, n, L = (sys.argv[1] if len(sys.argv) > 1 else "variational"), 450, 30
problem = S.make_problem(n, L, seed=0)
init = S.make_parameters(n, seed=0)
run = runtime.build_run(problem, guide, device="cuda:0", init=init, cuda_graph=True)
a = run.step_args()
b = tuple(None if t is None else t.clone() for t in a)
sets = [a, b]
try:
    for k in range(10):
        t0 = time.perf_counter()
        l = run.svi.step(*sets[k % 2])
        print(k, f"{l:.6f}", "replay" if run.svi.last_step_was_replay else "eager", f"{(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
except Exception:
    traceback.print_exc()
