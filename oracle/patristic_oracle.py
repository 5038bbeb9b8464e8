"""TEST INFRASTRUCTURE (see oracle/__init__.py) — patristic / cladistic / LCA oracle.

`lca_parent_walk` is the pure-Python small-case restatement; `patristic_c` calls the scalar C port
(`oracle/patristic_oracle.c`, built by `make -C oracle`) for sizes Python loops cannot finish in seconds.
Both follow S/utils.py:572-575 (ete3 `get_distance` arm sums), see the C file header.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    path = os.path.join(_HERE, "libdrp_oracle.so")
    src = os.path.join(_HERE, "patristic_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libdrp_oracle.so"])
    return path


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.drp_oracle_patristic.restype = None
    return _LIB


def lca_parent_walk(parent, branch, depth, rows, cols):
    """Pure Python: (lca, patristic, cladistic) for rows x cols."""
    lca = np.zeros((len(rows), len(cols)), dtype=np.int32)
    pat = np.zeros((len(rows), len(cols)), dtype=np.float64)
    cla = np.zeros((len(rows), len(cols)), dtype=np.float64)
    for a, u0 in enumerate(rows):
        for b, v0 in enumerate(cols):
            u, v, du, dv, e = int(u0), int(v0), 0.0, 0.0, 0
            while u != v:
                if depth[u] >= depth[v]:
                    du += branch[u]
                    u = int(parent[u])
                else:
                    dv += branch[v]
                    v = int(parent[v])
                e += 1
            lca[a, b], pat[a, b], cla[a, b] = u, du + dv, e
    return lca, pat, cla


def patristic_c(parent, branch, depth, rows=None, cols=None, want_lca=True, want_cladistic=False):
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    branch = np.ascontiguousarray(branch, dtype=np.float64)
    depth = np.ascontiguousarray(depth, dtype=np.int32)
    N = parent.shape[0]
    rows = np.arange(N, dtype=np.int32) if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.arange(N, dtype=np.int32) if cols is None else np.ascontiguousarray(cols, dtype=np.int32)
    lca = np.empty((len(rows), len(cols)), dtype=np.int32) if want_lca else None
    pat = np.empty((len(rows), len(cols)), dtype=np.float64)
    cla = np.empty((len(rows), len(cols)), dtype=np.float64) if want_cladistic else None
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    _lib().drp_oracle_patristic(p(parent), p(branch), p(depth), ctypes.c_int32(N), p(rows), ctypes.c_int32(len(rows)),
                                p(cols), ctypes.c_int32(len(cols)), p(lca), p(pat), p(cla))
    return lca, pat, cla
