"""As debug_gate_backerr.py, but after the other operations the widths test runs first in the same process
(argv: which of them to run, e.g. '450,90,200' or '' for none; 'c' suffix = with the conditional)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from draupnir_asr_b200 import ops, synthetic as S
todo = [a for a in (sys.argv[1] if len(sys.argv) > 1 else "").split(",") if a]
for item in todo:
    n = int(item.rstrip("c"))
    t = S.random_tree(n, seed=2)
    Dh = S.patristic_host(t)
    D = torch.from_numpy(Dh[np.ix_(t.leaves, t.leaves)]).cuda()
    g = torch.Generator().manual_seed(n)
    sf, sn, lam = (torch.rand(3, generator=g, dtype=torch.float64).add(0.5).cuda().requires_grad_(True) for _ in range(3))
    x = torch.randn(3, n, generator=g, dtype=torch.float64).cuda().requires_grad_(True)
    lp = ops.ou_prior_logp(D, sf, sn, lam, x)
    lp.sum().backward()
    if item.endswith("c"):
        Df = torch.from_numpy(Dh).cuda()
        leaf = torch.as_tensor(t.leaves); internal = torch.as_tensor([i for i in range(Dh.shape[0]) if i not in set(t.leaves)])
        mean, cov = ops.ou_conditional(Df, leaf, internal, sf.detach(), sn.detach(), lam.detach(), x.detach(), jitter=0.0)
n = 1100
t = S.random_tree(n, seed=2)
D = torch.from_numpy(S.patristic_host(t)[np.ix_(t.leaves, t.leaves)])
sf = torch.tensor([1.3, 0.9, 1.1], dtype=torch.float64)
sn = torch.tensor([0.6, 2e-4, 1e-3], dtype=torch.float64)
lam = torch.tensor([0.9, 1.4, 0.7], dtype=torch.float64)
K = sf[:, None, None] ** 2 * torch.exp(-D[None] / lam[:, None, None]) + (sn[:, None, None] ** 2) * torch.eye(n, dtype=torch.float64)   # S/models_utils.py:303-311
for rep in range(3):
    L = ops.potrf(K.cuda(), cond_bound=ops.ou_cond_bound(sf, sn, n)).cpu()
    back = ((L @ L.transpose(-1, -2) - K).abs().amax((1, 2)) / K.abs().amax((1, 2))).tolist()
    E = (L @ L.transpose(-1, -2) - K).abs()[1]
    r, c = divmod(int(E.argmax()), n)
    bad = (E > 1e-13 * float(K[1].abs().max())).nonzero()
    print({k: v for k, v in os.environ.items() if k.startswith("DRP_")}, todo, "rep", rep, ["%.2e" % b for b in back], "worst of matrix 1 at", (r, c),
          "entries above 1e-13:", len(bad), "rows", (int(bad[:, 0].min()), int(bad[:, 0].max())) if len(bad) else None,
          "cols", (int(bad[:, 1].min()), int(bad[:, 1].max())) if len(bad) else None)
