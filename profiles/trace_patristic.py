#!/usr/bin/env python
"""Device time of the two patristic kernels separately (tree_prepare vs the all-pairs LCA kernel) at N = 15,999 nodes."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
from draupnir_asr_b200 import ops, synthetic as S
from draupnir_asr_b200._lib import lib
dev = torch.device("cuda:0")
for n in (2000, 8000):
    t = S.random_tree(n, seed=n)
    N = t.n_nodes
    parent = torch.from_numpy(t.parent).to(dev); branch = torch.from_numpy(t.branch).to(dev)
    depth = torch.empty(N, dtype=torch.int32, device=dev); pre = torch.empty_like(depth); size = torch.empty_like(depth); scr = torch.empty_like(depth)
    rd = torch.empty(N, dtype=torch.float64, device=dev)
    out = torch.empty((N, N), dtype=torch.float64, device=dev)
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for rep in range(3):
        ev[0].record()
        lib.drp_tree_prepare(p(parent), p(branch), N, p(depth), p(rd), p(pre), p(size), p(scr), st)
        ev[1].record()
        md = int(depth.max())
        lib.drp_lca_patristic_f64(p(parent), p(depth), p(rd), p(pre), p(size), N, md, None, N, None, N, None, p(out), None, None, 0, st)
        ev[2].record()
        torch.cuda.synchronize()
    print(f"n={n} N={N} max_depth={md}: prepare {ev[0].elapsed_time(ev[1]):.3f} ms, depth.max()+lca kernel {ev[1].elapsed_time(ev[2]):.3f} ms "
          f"({N*N*8/ev[1].elapsed_time(ev[2])/1e6:.0f} GB/s)")
