"""CPU restatement (numpy) of what the fused panel kernels of csrc/factor.cu compute, step for step: `potrf_diag_kernel`
(per 64-column block: factor, invert, solve the rows of the diagonal region below, update the in-region blocks) and
`panel_solve_kernel` (rows below, independent 64-row groups: X_j = (B_j - sum_{i<j} X_i L_ji^T) Y_jj^T with the explicit block
inverses, optionally one step of refinement).  Checks the blocking against LAPACK for full and ragged panel widths, and that
the explicit-inverse solve + refinement keeps the backward error of an ill-conditioned OU covariance at rounding level - the
property `tests/test_gpu_ozaki.py::test_precision_gate_is_per_matrix` asserts for the kernels themselves."""
import numpy as np
import pytest

NB = 64


def _potrf_diag(M, c0, W):
    """In place on the lower-stored matrix M; returns the inverses of the 64 x 64 diagonal blocks (identity-padded)."""
    nblk = (W + NB - 1) // NB
    Y = []
    for j in range(nblk):
        nbj = min(NB, W - NB * j)
        s = c0 + NB * j
        blk = np.tril(M[s:s + nbj, s:s + nbj])
        L = np.linalg.cholesky(blk + np.tril(blk, -1).T)
        M[s:s + nbj, s:s + nbj] = L
        Lp = np.eye(NB)
        Lp[:nbj, :nbj] = L
        Y.append(np.linalg.solve(Lp, np.eye(NB)))                     # threads 0..63: rows of the identity
        nbelow = W - NB * (j + 1)
        if nbelow > 0:
            X = np.linalg.solve(Lp, M[s + NB:s + NB + nbelow, s:s + NB].T).T   # threads 64..: substitution, one row each
            M[s + NB:s + NB + nbelow, s:s + NB] = X
            for i in range(j + 1, nblk):
                for k in range(j + 1, i + 1):
                    Ti, Tk = X[(i - j - 1) * NB:(i - j) * NB], X[(k - j - 1) * NB:(k - j) * NB]
                    C = M[c0 + NB * i:c0 + NB * i + len(Ti), c0 + NB * k:c0 + NB * k + len(Tk)]
                    C -= np.tril(Ti @ Tk.T) if i == k else Ti @ Tk.T
    return Y


def _panel_solve(M, c0, W, R, Y, refine):
    nblk = (W + NB - 1) // NB
    for r0 in range(c0 + W, R, NB):                                    # one CTA per 64 rows
        nrows = min(NB, R - r0)
        X = [np.zeros((NB, NB)) for _ in range(nblk)]
        for j in range(nblk):
            nc = min(NB, W - NB * j)
            X[j][:nrows, :nc] = M[r0:r0 + nrows, c0 + NB * j:c0 + NB * j + nc]
        for j in range(nblk):
            nbj = min(NB, W - NB * j)
            for i in range(j):
                Lt = np.zeros((NB, NB))
                Lt[:nbj] = M[c0 + NB * j:c0 + NB * j + nbj, c0 + NB * i:c0 + NB * i + NB]
                X[j] -= X[i] @ Lt.T
            X0 = X[j] @ Y[j].T
            if refine:
                Ljj = np.eye(NB)
                Ljj[:nbj, :nbj] = np.tril(M[c0 + NB * j:c0 + NB * j + nbj, c0 + NB * j:c0 + NB * j + nbj])
                X0 = X0 + (X[j] - X0 @ Ljj.T) @ Y[j].T
            X[j] = X0
            M[r0:r0 + nrows, c0 + NB * j:c0 + NB * j + nbj] = X0[:nrows, :nbj]


def _factor(K, W, refine):
    R = K.shape[0]
    M = np.tril(K).copy()
    for c0 in range(0, R, W):
        Wq = min(W, R - c0)
        Y = _potrf_diag(M, c0, Wq)
        c1 = c0 + Wq
        if c1 < R:
            _panel_solve(M, c0, Wq, R, Y, refine)
            P = M[c1:, c0:c1]
            M[c1:, c1:] -= np.tril(P @ P.T)                             # the trailing update (tcgen05 / fp64 mma.sync kernels)
    return np.tril(M)


@pytest.mark.parametrize("R,W", [(300, 128), (300, 64), (500, 256), (333, 128), (130, 128), (90, 128), (401, 256)])
@pytest.mark.parametrize("refine", [False, True])
def test_blocking_equals_lapack(R, W, refine):
    rng = np.random.default_rng(R + W)
    A = rng.standard_normal((R, R))
    K = A @ A.T + R * np.eye(R)
    L = _factor(K, W, refine)
    assert np.abs(L - np.linalg.cholesky(K)).max() < 1e-11 * np.abs(K).max() ** 0.5


def test_explicit_inverse_solve_keeps_the_backward_error():
    from draupnir_asr_b200 import synthetic as S
    n = 700
    t = S.random_tree(n, seed=2)
    D = S.patristic_host(t)[np.ix_(t.leaves, t.leaves)]
    for sf, sn, lam in ((0.9, 2e-4, 1.4), (1.1, 1e-3, 0.7), (1.3, 0.6, 0.9)):
        K = sf * sf * np.exp(-D / lam) + sn * sn * np.eye(n)
        for refine in (False, True):
            L = _factor(K, 128, refine)
            assert np.abs(L @ L.T - K).max() / np.abs(K).max() < 4e-15, (sn, refine)
