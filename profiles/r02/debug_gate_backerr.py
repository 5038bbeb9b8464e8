"""Backward error of the three-matrix ill-conditioned batch (tests/test_gpu_ozaki.py::test_precision_gate_is_per_matrix)
under the current environment switches; prints one line."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from draupnir_asr_b200 import ops, synthetic as S
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
    print({k: v for k, v in os.environ.items() if k.startswith("DRP_")}, "rep", rep, ["%.2e" % b for b in back], "worst entry of matrix 1 at", (r, c))
