"""TEST INFRASTRUCTURE (build container only: needs /root/reference).  Times one SVI step of the REFERENCE'S OWN code
(tests/reference_harness.py) next to the oracle port that `bench.py --impl reference` / `cpu_baseline` time on the GPU
box, on the same host cores, to show the port is a fair stand-in for the reference's CPU cost.

    python tests/time_reference_vs_port.py [n_leaves sites guide steps [name]]

With a `name`, the losses the reference produced (one per `train()` call, the noise of the variational guide fixed to the
seed-5 draw the tests regenerate) are stored in tests/golden/ref_fullsize_losses.json: full-size parity targets for the
GPU tests (C2, C3, C4 of BASELINE.json).
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reference_harness as H  # noqa: E402
from draupnir_asr_b200 import synthetic as S  # noqa: E402
from oracle import draupnir_oracle as O  # noqa: E402


def main(n=200, L=300, guide="variational", steps=3, name=None):
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(os.cpu_count())
    pr = S.make_problem(n, L, seed=0)
    init = S.make_parameters(n, seed=0)
    nets = O.Networks(variational=(guide == "variational"), seed=0)
    st = {k: ({a: b.clone() for a, b in v.items()} if isinstance(v, dict) else v.clone())
          for k, v in H.nets_state_of(nets).items()}
    run = H.build_reference_run(pr, guide, init, st)
    port = O.OracleSVI(nets, init, guide)
    eps = torch.randn(30, n, generator=torch.Generator().manual_seed(5))
    oargs = (pr.dataset_train, pr.patristic_matrix_train, pr.dataset_train_blosum, pr.blosum_weighted, eps)
    out = {}
    for who, fn in (("reference", lambda: run.ref.train.train(run.svi, run.train_input)),
                     ("port", lambda: port.step(*oargs))):
        with H.FixedNoise(eps):
            first = fn()                       # warm-up
            losses = [first]
            t0 = time.perf_counter()
            for _ in range(steps):
                losses.append(fn())
            out[who] = ((time.perf_counter() - t0) / steps, first, losses[-1], losses)
    if name is not None:
        path = os.path.join(ROOT, "tests", "golden", "ref_fullsize_losses.json")
        store = json.load(open(path)) if os.path.exists(path) else {}
        store[name] = {"n_leaves": n, "sites": L, "guide": guide, "seed": 0, "eps_seed": 5, "threads": os.cpu_count(),
                       "losses_reference": out["reference"][3], "losses_port": out["port"][3],
                       "ms_per_step_reference": out["reference"][0] * 1e3, "ms_per_step_port": out["port"][0] * 1e3}
        json.dump(store, open(path, "w"), indent=1)
    for who, (t, first, last, _) in out.items():
        print(f"{who:9s}: {t * 1e3:9.1f} ms per SVI step   first loss {first:.6f}   loss after {steps + 1} steps {last:.6f}")
    print(f"n_leaves {n}, sites {L}, {guide}, {os.cpu_count()} threads; port / reference time = "
          f"{out['port'][0] / out['reference'][0]:.3f}")


if __name__ == "__main__":
    a = sys.argv[1:]
    main(int(a[0]) if a else 200, int(a[1]) if len(a) > 1 else 300, a[2] if len(a) > 2 else "variational",
         int(a[3]) if len(a) > 3 else 3, a[4] if len(a) > 4 else None)
