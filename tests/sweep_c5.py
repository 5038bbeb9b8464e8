#!/usr/bin/env python
"""C5 of BASELINE.json: patristic-matrix + OU-covariance build sweep, 100-8,000 leaves, on one B200.

Test infrastructure (it checks against and times `oracle/`, which only tests/, smoke() and bench.py's CPU legs may use);
a script, not collected by pytest.

For each size: (a) LCA indices bit-exact against the scalar C oracle (all pairs up to 2,000 leaves, a seeded random
256 x 256 rows x cols subset above), distances rtol 1e-12; (b) device time (CUDA events, median of 7 after 3 warm-ups,
L2 flushed between repetitions) of the all-node patristic build and of the [30, n, n] OU covariance assembly, reported
as algorithmic GB/s against the measured HBM peak; (c) the host-CPU time of the same builds with the oracle
(parent-walk C port / torch restatement of OUKernel_Fast) on a bounded sample.

    python tests/sweep_c5.py > gpurun_out/sweep_c5.json
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

entry.build()
from draupnir_asr_b200 import ops, synthetic as S  # noqa: E402
from oracle import draupnir_oracle as O  # noqa: E402
from oracle import patristic_oracle as P  # noqa: E402

Z = 30


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def timed(fn, flush, reps=7, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    peak, src = hbm_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"hbm_peak_gbs": peak, "peak_source": src, "z_dim": Z, "sizes": []}
    torch.set_num_threads(os.cpu_count() or 1)
    for n in (100, 200, 500, 1000, 2000, 4000, 8000):
        t = S.random_tree(n, seed=n)
        N = t.n_nodes
        parent = torch.from_numpy(t.parent).to(dev)
        branch = torch.from_numpy(t.branch).to(dev)
        row = {"n_leaves": n, "n_nodes": N}
        # ---- parity
        if n <= 2000:
            pat, _, lca = ops.tree_distances(parent, branch, want_lca=True)
            t0 = time.perf_counter()
            lca_ref, pat_ref, _ = P.patristic_c(t.parent, t.branch, t.depth)
            row["cpu_patristic_s"] = time.perf_counter() - t0
            row["cpu_pairs"] = N * N
        else:
            rng = np.random.default_rng(n)
            rows = np.sort(rng.choice(N, 256, replace=False)).astype(np.int32)
            cols = np.sort(rng.choice(N, 256, replace=False)).astype(np.int32)
            pat, _, lca = ops.tree_distances(parent, branch, rows=torch.from_numpy(rows), cols=torch.from_numpy(cols), want_lca=True)
            t0 = time.perf_counter()
            lca_ref, pat_ref, _ = P.patristic_c(t.parent, t.branch, t.depth, rows, cols)
            row["cpu_patristic_s"] = time.perf_counter() - t0
            row["cpu_pairs"] = 256 * 256
        row["lca_bit_exact"] = bool(np.array_equal(lca.cpu().numpy(), lca_ref))
        row["patristic_max_rel_err"] = float(np.max(np.abs(pat.cpu().numpy() - pat_ref) / np.maximum(np.abs(pat_ref), 1e-300)))
        del pat, lca
        # ---- patristic build, all nodes
        ms = timed(lambda: ops.tree_distances(parent, branch), flush)
        row["patristic_ms"] = ms
        row["patristic_bytes"] = N * N * 8
        row["patristic_gbs"] = N * N * 8 / (ms * 1e-3) / 1e9
        ms = timed(lambda: ops.tree_distances(parent, branch, want_lca=True), flush)
        row["patristic_lca_ms"] = ms
        row["patristic_lca_gbs"] = N * N * 12 / (ms * 1e-3) / 1e9
        # the two kernels on their own (library instrumentation: CUDA events around each launch)
        import ctypes
        from draupnir_asr_b200._lib import lib
        ops.tree_distances(parent, branch)
        lib.drp_timing_enable(1)
        for _ in range(5):
            flush.zero_()
            ops.tree_distances(parent, branch)
        NK = lib.drp_timing_classes()
        kms, kln, kwk = (ctypes.c_double * NK)(), (ctypes.c_int64 * NK)(), (ctypes.c_double * NK)()
        lib.drp_timing_read(kms, kln, kwk)
        lib.drp_timing_enable(0)
        row["patristic_kernels_ms_per_build"] = kms[6] / 5          # tree_prepare + lca kernel
        row["patristic_kernels_gbs"] = N * N * 8 / (kms[6] / 5 * 1e-3) / 1e9
        # ---- OU covariance [z, n, n] over the leaves
        leaves = torch.from_numpy(t.leaves).to(dev)
        D = ops.tree_distances(parent, branch, rows=leaves, cols=leaves)[0]
        g = torch.Generator().manual_seed(0)
        sf, sn, lam = (torch.rand(Z, generator=g, dtype=torch.float64).add(0.5).to(dev) for _ in range(3))
        K = None

        def build_cov():
            nonlocal K
            K = None
            K = ops.ou_cov(D, sf, sn, lam)

        ms = timed(build_cov, flush, reps=5, warm=2)
        row["ou_cov_ms"] = ms
        row["ou_cov_bytes"] = n * n * 8 * (Z + 1)
        row["ou_cov_gbs"] = row["ou_cov_bytes"] / (ms * 1e-3) / 1e9
        if n <= 2000:
            Dc = D.cpu()
            t0 = time.perf_counter()
            K_ref = O.ou_covariance(Dc, sf.cpu(), sn.cpu(), lam.cpu())
            row["cpu_ou_cov_s"] = time.perf_counter() - t0
            row["ou_cov_max_rel_err"] = float((K.cpu() - K_ref).abs().max() / K_ref.abs().max())
            del K_ref
        K = None
        for k in ("patristic_gbs", "patristic_lca_gbs", "ou_cov_gbs"):
            row[k.replace("_gbs", "_frac_of_hbm_peak")] = row[k] / peak
        out["sizes"].append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
    out["cpu_cores"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
