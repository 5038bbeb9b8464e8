"""Stress: the ill-conditioned three-matrix factorisation many times in one process, interleaved with other work and
allocations; reports every run whose backward errors leave the expected band."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from draupnir_asr_b200 import ops, synthetic as S
n = 1100
t = S.random_tree(n, seed=2)
Dh = S.patristic_host(t)
D = torch.from_numpy(Dh[np.ix_(t.leaves, t.leaves)])
sf = torch.tensor([1.3, 0.9, 1.1], dtype=torch.float64)
sn = torch.tensor([0.6, 2e-4, 1e-3], dtype=torch.float64)
lam = torch.tensor([0.9, 1.4, 0.7], dtype=torch.float64)
K = sf[:, None, None] ** 2 * torch.exp(-D[None] / lam[:, None, None]) + (sn[:, None, None] ** 2) * torch.eye(n, dtype=torch.float64)   # S/models_utils.py:303-311
Kg = K.cuda()
cb = ops.ou_cond_bound(sf, sn, n)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
bad = 0
ref = None
junk = []
Dg = D.cuda()
g = torch.Generator().manual_seed(1)
for rep in range(reps):
    if rep % 3 == 1:                      # other work of the library in between, different sizes -> different stale memory
        m = 300 + 37 * (rep % 11)
        x = torch.randn(3, m, generator=g, dtype=torch.float64).cuda().requires_grad_(True)
        p = [v.cuda().requires_grad_(True) for v in (sf, torch.tensor([0.6, 0.5, 0.7], dtype=torch.float64), lam)]
        ops.ou_prior_logp(Dg[:m, :m].contiguous(), *p, x).sum().backward()
    if rep % 5 == 2:
        junk.append(torch.randn(1 + (rep * 7919) % 3000000, device="cuda"))
        if len(junk) > 4:
            junk.pop(0)
    L = ops.potrf(Kg, cond_bound=cb)
    E = (L @ L.transpose(-1, -2) - Kg).abs()
    back = (E.amax((1, 2)) / Kg.abs().amax((1, 2))).tolist()
    if ref is None:
        ref = L.clone()
    same = bool((L == ref).all())
    if not (back[0] < 2e-11 and back[1] < 2e-15 and back[2] < 2e-15) or not same:
        bad += 1
        d = (L - ref).abs()
        idx = d.nonzero()
        print("rep", rep, ["%.2e" % b for b in back], "bitwise same as run 0:", same, "differing entries", len(idx),
              "first", idx[:3].tolist(), "matrices", sorted(set(idx[:, 0].tolist())),
              "rows", (int(idx[:, 1].min()), int(idx[:, 1].max())) if len(idx) else None,
              "cols", (int(idx[:, 2].min()), int(idx[:, 2].max())) if len(idx) else None, flush=True)
print({k: v for k, v in os.environ.items() if k.startswith("DRP_")}, "runs", reps, "bad", bad)
