#!/usr/bin/env python
"""One forward (+ backward) of the product GRU kernels on C4-sized random data; the target of ncu captures.
python profiles/gru_only.py [uv]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
from draupnir_asr_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n, L, H = 2000, 500, 60
g = torch.Generator(device=dev).manual_seed(0)
whh = (torch.randn(2, 3 * H, H, dtype=torch.float64, device=dev, generator=g) * 0.1).requires_grad_(True)
bhh = (torch.randn(2, 3 * H, dtype=torch.float64, device=dev, generator=g) * 0.1).requires_grad_(True)
h0 = torch.randn(2, n, H, dtype=torch.float64, device=dev, generator=g)
if len(sys.argv) > 1 and sys.argv[1] == "uv":
    U = torch.randn(n, 2 * 3 * H, dtype=torch.float64, device=dev, generator=g).requires_grad_(True)
    V = torch.randn(L, 2 * 3 * H, dtype=torch.float64, device=dev, generator=g).requires_grad_(True)
    run = lambda: ops.gru_bidirectional_uv(U, V, whh, bhh, h0)
else:
    gi = torch.randn(n, L, 2, 3 * H, dtype=torch.float64, device=dev, generator=g).requires_grad_(True)
    run = lambda: ops.gru_bidirectional(gi, whh, bhh, h0)
for _ in range(2):
    out = run()
    out.sum().backward()
torch.cuda.synchronize()
print("ok", float(out.sum()))
