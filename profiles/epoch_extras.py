#!/usr/bin/env python
"""SURVEY §8 f2, the measurement it asks for before anything is built: what do the reference's per-epoch extras around
`svi.step` cost next to the step itself on this implementation?  S/main.py:1097 calls the guide once more, :1113-1120
decodes the leaves with argmax (one decoder forward), :1121-1122 writes two checkpoints, every epoch.
    python profiles/epoch_extras.py [workload ...] > gpurun_out/epoch_extras.txt"""
import io
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


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for name in (sys.argv[1:] or ["c1", "c2", "c4"]):
    w = bench.WORKLOADS[name]
    problem = synthetic.make_problem(w["n"], w["L"], seed=0)
    init = synthetic.make_parameters(w["n"], z_dim=30, seed=0)
    run = runtime.build_run(problem, w["guide"], device="cuda:0", init=init, z_dim=30)
    for _ in range(3):
        run.step()
    args = run.step_args()
    t_step = timed(run.step)

    def extra_guide():
        with torch.no_grad():
            return {k: v.detach() for k, v in run.guide(*args).items()}       # S/main.py:1097,1111 (reads every entry)

    t_guide = timed(extra_guide)
    est = extra_guide()

    def extra_decode():
        return run.Draupnir.sample(est, 1, None, run.patristic_matrix_full, None, use_argmax=True, use_test=False)

    t_dec = timed(extra_decode)

    def extra_save():
        for mod in (run.Draupnir, run.guide):
            if hasattr(mod, "state_dict"):
                torch.save(mod.state_dict(), io.BytesIO())

    t_save = timed(extra_save)
    print(f"{name}: svi.step {t_step:.2f} ms | extra guide call {t_guide:.2f} ms | argmax decode of the leaves {t_dec:.2f} ms | "
          f"two state_dict saves (to memory) {t_save:.2f} ms | extras / step = {(t_guide + t_dec + t_save) / t_step:.2f}")
