#!/usr/bin/env python
"""Phase timing of the GRU forward recurrence (profiling build with -DDRP_GRU_TRACE): per step, for warps 0 and 7 of CTA
(0,0): cycles spent issuing the next TMA stage, in the recurrent product (DMMA), waiting for the step's gi rows, in the gate
math (+ stores) and at the CTA barrier.   python profiles/trace_gru.py [rows-per-CTA] > gpurun_out/trace_gru.txt"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from draupnir_asr_b200 import build as B  # noqa: E402

out = os.path.join(ROOT, "profiles", "microbench", "libdrp_gru_trace.so")
if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in B.sources()):
    subprocess.check_call([B._nvcc(), *B.NVCC_FLAGS, "-DDRP_GRU_TRACE", "-shared", "-o", out, *B.sources(), "-lcudart"], cwd=B.CSRC)
if not torch.cuda.is_available():
    print("built", out)
    sys.exit(0)
# argv: rows per CTA (8 / 16 / 32), or "cluster" for the CTA-pair kernels (gru_fwdc_kernel / gru_bwdc_kernel); then n
cluster = len(sys.argv) > 1 and sys.argv[1] == "cluster"
if cluster:
    os.environ["DRP_GRU_CLUSTER"] = "1"
elif len(sys.argv) > 1:
    os.environ["DRP_GRU_ROWS"] = sys.argv[1]
    os.environ["DRP_GRU_CLUSTER"] = "0"
lib = ctypes.CDLL(out)
dev = torch.device("cuda:0")
n, L, H = (int(sys.argv[2]) if len(sys.argv) > 2 else 2000), 500, 60
g = torch.Generator(device=dev).manual_seed(0)
gi = torch.randn(n, L, 2, 3 * H, dtype=torch.float64, device=dev, generator=g)
whh = torch.randn(2, 3 * H, H, dtype=torch.float64, device=dev, generator=g) * 0.1
bhh = torch.randn(2, 3 * H, dtype=torch.float64, device=dev, generator=g) * 0.1
h0 = torch.randn(2, n, H, dtype=torch.float64, device=dev, generator=g)
outb = torch.empty(n, L, 2 * H, dtype=torch.float64, device=dev)
gates = torch.empty(n, L, 2, 4, H, dtype=torch.float64, device=dev)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(2):
    rc = lib.drp_gru_fwd_f64(P(gi), P(whh), P(bhh), P(h0), n, L, H, P(outb), P(gates), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
lib.drp_gru_fwd_f64(P(gi), P(whh), P(bhh), P(h0), n, L, H, P(outb), P(gates), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
b.record()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (8 * 64 * 2))()
assert lib.drp_gru_trace_read(buf) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(2, 64, 8)
print(f"# gru_fwd n={n} L={L}: {a.elapsed_time(b):.3f} ms = {a.elapsed_time(b) * 1e6 / L:.0f} ns/step; rows/CTA = {'8 per CTA pair' if cluster else os.environ.get('DRP_GRU_ROWS', 'auto')}")
if cluster:
    order = [0, 1, 7, 2, 3, 4, 6, 5]
    print("# cycles per phase (mean over steps 200..263): wait own half of h | product, own half | wait partner's half + product | wait gi | gates + h stores (local, st.async) | mbarrier arrivals | global stores | step total")
    for w, name in ((0, "warp 0"), (1, "warp 3")):
        d = np.diff(t[w][:, order], axis=1).astype(np.float64)
        step = np.diff(t[w, :, 0]).astype(np.float64)
        print(name, " ".join(f"{v:8.0f}" for v in d.mean(0)), f"| step {step.mean():8.0f} (min {step.min():.0f} max {step.max():.0f})")
print("# cycles per phase (mean over steps 200..263): issue | product | (to wait) | wait gi | gates+stores | barrier | step total")
for w, name in ((0, "warp 0"), (1, "warp 8/7")):
    d = np.diff(t[w, :, :6], axis=1).astype(np.float64)            # [64, 5]
    step = np.diff(t[w, :, 0]).astype(np.float64)
    print(name, " ".join(f"{v:8.0f}" for v in d.mean(0)), f"| step {step.mean():8.0f} (min {step.min():.0f} max {step.max():.0f})")
print("# start of step 230: second traced warp minus warp 0 =", int(t[1, 30, 0] - t[0, 30, 0]), "cycles")
# ---- backward kernel (both directions), same buffer
dout = torch.randn(n, L, 2 * H, dtype=torch.float64, device=dev, generator=g)
dgi = torch.empty(n, L, 2, 3 * H, dtype=torch.float64, device=dev)
dh0 = torch.empty(2, n, H, dtype=torch.float64, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    rc = lib.drp_gru_bwd_f64(P(dout), P(outb), P(gates), P(whh), P(h0), n, L, H, P(dgi), P(dh0), st)
    assert rc == 0, rc
a.record()
lib.drp_gru_bwd_f64(P(dout), P(outb), P(gates), P(whh), P(h0), n, L, H, P(dgi), P(dh0), st)
b.record()
torch.cuda.synchronize()
assert lib.drp_gru_trace_read(buf) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(2, 64, 8)
print(f"# gru_bwd n={n} L={L}: {a.elapsed_time(b):.3f} ms = {a.elapsed_time(b) * 1e6 / L:.0f} ns/step")
if cluster:
    order = [0, 2, 3, 4, 1, 7, 5, 6]
    print("# cycles per phase: dout loads issued | wait ring | gates + dgh stores (local, st.async) | mbarrier arrivals | global stores | wait own columns + product | wait partner's columns + product | step total")
    for w, name in ((0, "warp 0"), (1, "warp 3")):
        d = np.diff(t[w][:, order], axis=1).astype(np.float64)
        step = np.diff(t[w, :, 0]).astype(np.float64)
        print(name, " ".join(f"{v:8.0f}" for v in d.mean(0)), f"| step {step.mean():8.0f}")
print("# cycles per phase: issue | dout loads issued | wait ring | gates+stores | barrier 1 | product | barrier 2 | step total")
for w, name in ((0, "warp 0"), (1, "warp 7")):
    d = np.diff(t[w, :, :8], axis=1).astype(np.float64)
    step = np.diff(t[w, :, 0]).astype(np.float64)
    print(name, " ".join(f"{v:8.0f}" for v in d.mean(0)), f"| step {step.mean():8.0f}")
# ---- weight-gradient kernel (both directions)
dwhh = torch.empty_like(whh)
dbhh = torch.empty(2, 3 * H, dtype=torch.float64, device=dev)
lib.drp_gru_wgrad_workspace_elems.restype = ctypes.c_int64
wse = lib.drp_gru_wgrad_workspace_elems()
ws = torch.empty(wse, dtype=torch.float64, device=dev)
for _ in range(2):
    rc = lib.drp_gru_wgrad_f64(P(dgi), P(gates), P(outb), P(h0), n, L, H, P(dwhh), P(dbhh), P(ws), ctypes.c_int64(wse), st)
    assert rc == 0, rc
a.record()
lib.drp_gru_wgrad_f64(P(dgi), P(gates), P(outb), P(h0), n, L, H, P(dwhh), P(dbhh), P(ws), ctypes.c_int64(wse), st)
b.record()
torch.cuda.synchronize()
assert lib.drp_gru_trace_read(buf) == 0
t = np.frombuffer(buf, dtype=np.int64).reshape(2, 64, 8)
print(f"# gru_wgrad n={n} L={L}: {a.elapsed_time(b):.3f} ms")
print("# cycles per 32-pair stage: issue | wait ring | product | barrier | stage total")
for w, name in ((0, "warp 0"), (1, "warp 7")):
    d = np.diff(t[w, :, :5], axis=1).astype(np.float64)
    step = np.diff(t[w, :, 0]).astype(np.float64)
    print(name, " ".join(f"{v:8.0f}" for v in d.mean(0)), f"| stage {step.mean():8.0f}")
