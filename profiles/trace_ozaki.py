#!/usr/bin/env python
"""Event trace of CTA 0 of one tcgen05 trailing update (debug hook drp_ozaki_set_trace): where do the producer, the MMA
issuer and the epilogue spend their time?   python profiles/trace_ozaki.py > gpurun_out/trace_ozaki.txt"""
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
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256          # contraction length: 256 (outer panels) or 128
z, R, ld, c1 = 30, 2304 + 256, 2304 + 256 + 8, 256
M = torch.randn(z, R, ld, dtype=torch.float64, device=dev)
ops.syrk_update(M, 0, K, c1, R, R, slices=6)             # warm-up
torch.cuda.synchronize()
trace = torch.zeros(8002, dtype=torch.int64, device=dev)
lib.drp_ozaki_set_trace(ctypes.c_void_p(trace.data_ptr()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
ops.syrk_update(M, 0, K, c1, R, R, slices=6)
b.record()
torch.cuda.synchronize()
lib.drp_ozaki_set_trace(None)
t = trace.cpu().numpy()
n = int(min(t[0], 4000))
ev = sorted(((int(t[3 + 2 * k]), int(t[2 + 2 * k])) for k in range(n)))
print(f"# update rows {R - c1} K {K} z {z}: {a.elapsed_time(b):.3f} ms total; {n} events of CTA 0 (first 400 shown)")
print("# ids: 100 producer got a_empty; 2xx producer got b_empty for B block xx; 300 MMA got A; 310 MMA got acc_empty;")
print("#      4xx MMA got b_full for block xx; 50d epilogue(warp 2) got accumulator d; 600 epilogue released accumulators")
t0 = ev[0][0]
tiles = sum(1 for _, eid in ev if eid == 600)
span = ev[-1][0] - t0
print(f"# CTA 0: {tiles} tiles in {span} cycles between its first and last event = {span / max(tiles, 1):.0f} cycles per tile; "
      f"if it was busy for the whole update the SM clock was {span / (a.elapsed_time(b) * 1e-3) / 1e9:.2f} GHz or more "
      f"(the elapsed time also covers the slicing kernel)")
last = {}
for clk, eid in ev[:400]:
    role = eid // 100
    d = clk - last.get(role, clk)
    last[role] = clk
    print(f"{clk - t0:9d}  +{d:6d}  {eid}")
