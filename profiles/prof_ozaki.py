#!/usr/bin/env python
"""Phase accounting of one tcgen05 trailing update (debug hook drp_ozaki_set_prof): cycles every CTA's producer, MMA
issuer and first epilogue warp ran in total and spent blocked on each of their barriers - two clock reads per wait, no
global traffic inside the loops (unlike the event trace, whose hook costs ~1 k cycles per event).
    python profiles/prof_ozaki.py [K] > gpurun_out/prof_ozaki.txt"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
from draupnir_asr_b200 import ops  # noqa: E402
from draupnir_asr_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
z, R, ld, c1 = 30, 2304 + 256, 2304 + 256 + 8, 256
M = torch.randn(z, R, ld, dtype=torch.float64, device=dev)
for _ in range(2):
    ops.syrk_update(M, 0, K, c1, R, R, slices=6)             # warm-up
torch.cuda.synchronize()
n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
prof = torch.zeros(16 * n_sm, dtype=torch.int64, device=dev)
lib.drp_ozaki_set_prof(ctypes.c_void_p(prof.data_ptr()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
ops.syrk_update(M, 0, K, c1, R, R, slices=6)
b.record()
torch.cuda.synchronize()
lib.drp_ozaki_set_prof(None)
p = prof.cpu().double().reshape(n_sm, 16)
ms = a.elapsed_time(b)
tiles = p[:, 8]
print(f"# update rows {R - c1} K {K} z {z}: {ms:.3f} ms incl. slicing; {int(tiles.sum())} tiles over {n_sm} CTAs "
      f"({tiles.min():.0f}..{tiles.max():.0f} per CTA)")
mma_total = p[:, 3]
print(f"# MMA issuer lifetime: mean {mma_total.mean():.0f} cycles (min {mma_total.min():.0f}, max {mma_total.max():.0f}) "
      f"= {mma_total.mean() / tiles.mean():.0f} cycles per tile; at {ms:.3f} ms that is >= {mma_total.max() / (ms * 1e-3) / 1e9:.2f} GHz")
rows = [("producer: waiting for the A strip to be released", p[:, 1], p[:, 0]),
        ("producer: waiting for a free B stage", p[:, 2], p[:, 0]),
        ("MMA issuer: waiting for a work unit", p[:, 4], p[:, 3]),
        ("MMA issuer: waiting for the A strip", p[:, 5], p[:, 3]),
        ("MMA issuer: waiting for drained accumulators", p[:, 6], p[:, 3]),
        ("MMA issuer: waiting for a B block", p[:, 7], p[:, 3]),
        ("MMA issuer: everything else (issue)", p[:, 3] - p[:, 4] - p[:, 5] - p[:, 6] - p[:, 7], p[:, 3]),
        ("epilogue warp 2: waiting for a work unit", p[:, 10], p[:, 9]),
        ("epilogue warp 2: waiting for completed accumulators", p[:, 11], p[:, 9]),
        ("epilogue warp 2: everything else (loads, fp64, C update)", p[:, 9] - p[:, 10] - p[:, 11], p[:, 9])]
for name, x, tot in rows:
    print(f"{name:58s} {x.mean() / tiles.mean():9.0f} cycles per tile   {100 * x.sum() / tot.sum():5.1f} % of the role's lifetime")
