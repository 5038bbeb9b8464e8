"""Capture diagnostics: tries the CUDA-graph capture of svi.step for a few configurations and prints what happens."""
import os, sys, time, warnings, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
from draupnir_asr_b200 import runtime, synthetic as S

warnings.simplefilter("always")
cases = [("delta_map", 19, 225, False), ("variational", 60, 40, False), ("variational", 450, 24, False),
         ("variational", 450, 24, True), ("delta_map", 450, 24, True)]
if len(sys.argv) > 1:
    cases = [c for i, c in enumerate(cases) if str(i) in sys.argv[1:]]
for guide, n, L, hostfed in cases:
    print("=====", guide, n, L, "host-fed" if hostfed else "resident", flush=True)
    problem = S.make_problem(n, L, seed=0)
    init = S.make_parameters(n, seed=0)
    run = runtime.build_run(problem, guide, device="cuda:0", init=init, cuda_graph=True)
    host = run.make_host_inputs() if hostfed else None
    step = (lambda: run.step_from_host(host)) if hostfed else run.step
    try:
        for k in range(5):
            t0 = time.perf_counter()
            l = step()
            print(k, f"{l:.6f}", "replay" if run.svi.last_step_was_replay else "eager", f"{(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True)
    except Exception:
        traceback.print_exc()
