"""PyTorch custom ops over the C ABI (raw device pointers + the current CUDA stream; torch is only the allocator).

Every function requires CUDA fp64 tensors and raises otherwise — there is no CPU path in the product.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import List, Optional, Tuple

import torch

from ._lib import lib

_DEFERRED: Optional[List[Tuple[torch.Tensor, str]]] = None

# The OU prior (a chain of ~150 mostly narrow kernels: 64-wide panel factorisations, solves, slicing) and the GRU
# recurrences (one CTA per SM on 126 of 148 SMs, latency-bound) are independent until the loss is summed: the prior's
# log-density is evaluated on a side stream so that the two fill each other's idle SMs.  DRP_PRIOR_STREAM=0 disables.
PRIOR_STREAM_ENABLED = os.environ.get("DRP_PRIOR_STREAM", "1") != "0"
LOOKAHEAD_ENABLED = os.environ.get("DRP_LOOKAHEAD", "0") not in ("0", "")     # opt-in: no gain measured (csrc/factor.cu)
_SIDE_STREAMS = {}


def side_stream(device) -> "torch.cuda.Stream":
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        # priorities: the step's main chain (the graph's capture stream) -2, the prior's chain here -1, the look-ahead part of
        # the factorisation's updates (`_aux`) 0 = lowest: latency chains get SMs ahead of throughput work
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=key, priority=-1)
    return st


_AUX = {}


def _aux(device):
    """(stream, fork event, join event) the factorisation's look-ahead runs on (csrc/factor.cu: the update of an outer panel
    is split by columns and its larger part overlaps the factorisation of the next panel).  The library creates no streams
    or events of its own; these are per device AND per launching stream (the prior on the side stream and a conditional
    posterior on the main stream must not share events).  Returned as three ctypes handles."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    hit = _AUX.get(key)
    if hit is None:
        if not LOOKAHEAD_ENABLED:
            hit = (None, None, None, None)
        else:
            st = torch.cuda.Stream(device=idx, priority=0)
            evs = (torch.cuda.Event(), torch.cuda.Event())
            with torch.cuda.stream(st):
                for e in evs:
                    e.record()                      # materialises the cudaEvent_t
            hit = (ctypes.c_void_p(st.cuda_stream), ctypes.c_void_p(evs[0].cuda_event), ctypes.c_void_p(evs[1].cuda_event),
                   (st, evs))                       # the torch objects keep the handles alive
        _AUX[key] = hit
    return hit[:3]


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _rc(rc: int, name: str):
    if rc != 0:
        what = f"argument group {-rc} invalid" if rc < 0 else f"CUDA error {rc - 1000}"
        raise RuntimeError(f"{name} failed: {what}")


def _req(t: torch.Tensor, name: str, dtype=torch.float64) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (draupnir_asr_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t


def _matrix_view(D: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """Accepts strided row-major views such as `patristic_matrix_sorted[1:, 1:]` without copying."""
    _req(D, name)
    if D.dim() != 2:
        raise RuntimeError(f"{name}: expected a 2-D matrix")
    if D.stride(1) != 1:
        D = D.contiguous()
    return D, D.stride(0)


@contextlib.contextmanager
def defer_info_checks():
    """Inside the block, pivot-failure flags are queued instead of synchronising; the block's owner calls
    `flush_info_checks()` where it synchronises anyway (SVI.step reads the loss there)."""
    global _DEFERRED
    prev, _DEFERRED = _DEFERRED, []
    try:
        yield
    except BaseException:
        _DEFERRED = prev                 # unwinding from another error: do not synchronise, do not mask it
        raise
    else:
        pending, _DEFERRED = _DEFERRED, prev
        _raise_if_failed(pending)


def begin_deferred():
    """Open a deferred-check scope without the context manager (a CUDA-graph capture must not examine the flags when it
    ends — nothing has run yet).  Returns the token `end_deferred` needs."""
    global _DEFERRED
    prev, _DEFERRED = _DEFERRED, []
    return (prev,)


def end_deferred(token):
    """Close the scope opened by `begin_deferred` and hand back the queued `(info, what)` pairs unexamined."""
    global _DEFERRED
    pending, _DEFERRED = _DEFERRED, token[0]
    return pending or []


def flush_info_checks():
    """Examine (and clear) the pivot flags queued so far inside `defer_info_checks()`: raises `torch.linalg.LinAlgError`
    for the first failed factorisation.  Synchronises; call it where the host waits for the device anyway."""
    global _DEFERRED
    if _DEFERRED:
        pending, _DEFERRED = _DEFERRED, []
        _raise_if_failed(pending)


def pending_failures() -> Optional[torch.Tensor]:
    """Device scalar (int64): how many of the queued pivot-flag vectors report a failure; None if nothing is queued.  Lets a
    sharded step add the flag to its all-reduce so that every rank raises together."""
    if not _DEFERRED:
        return None
    return torch.stack([(info != 0).any() for info, _ in _DEFERRED]).sum()


def _raise_if_failed(pending):
    for info, what in pending:
        bad = torch.nonzero(info)
        if bad.numel():
            b = int(bad[0, 0])
            raise torch.linalg.LinAlgError(
                f"{what}: (Batch element {b}): The factorization could not be completed because the input is not "
                f"positive-definite (the leading minor of order {int(info[b])} is not positive-definite).")


def _info_done(info: torch.Tensor, what: str):
    if _DEFERRED is not None:
        _DEFERRED.append((info, what))
    else:
        _raise_if_failed([(info, what)])


# ------------------------------------------------------------------------------------------------------------------
# K2  OUKernel_Fast.forward
# ------------------------------------------------------------------------------------------------------------------
class _OUCov(torch.autograd.Function):
    @staticmethod
    def forward(ctx, D, sf, sn, lam):
        Dv, ldd = _matrix_view(D, "t")
        n, z = Dv.shape[0], sf.shape[0]
        sf, sn, lam = (_req(v, k).contiguous() for v, k in ((sf, "sigma_f"), (sn, "sigma_n"), (lam, "lamb")))
        K = torch.empty((z, n, n), dtype=torch.float64, device=D.device)
        _rc(lib.drp_ou_cov_fwd_f64(_p(Dv), ldd, n, _p(sf), _p(sn), _p(lam), z, _p(K), _stream()), "drp_ou_cov_fwd_f64")
        ctx.save_for_backward(Dv, sf, sn, lam)
        return K

    @staticmethod
    def backward(ctx, G):
        Dv, sf, sn, lam = ctx.saved_tensors
        n, z = Dv.shape[0], sf.shape[0]
        G = G.contiguous()
        part = torch.empty((z, n, 3), dtype=torch.float64, device=G.device)
        out = torch.empty((z, 3), dtype=torch.float64, device=G.device)
        _rc(lib.drp_ou_cov_bwd_f64(_p(Dv), Dv.stride(0), n, _p(lam), z, _p(G), _p(part), _p(out), _stream()),
            "drp_ou_cov_bwd_f64")
        return None, 2.0 * sf * out[:, 0], 2.0 * sn * out[:, 2], sf * sf * out[:, 1] / (lam * lam)


def ou_cov(t: torch.Tensor, sigma_f: torch.Tensor, sigma_n: torch.Tensor, lamb: torch.Tensor) -> torch.Tensor:
    """`[z,n,n]` OU covariance sf^2 exp(-t/lam) + sn^2 I (S/models_utils.py:303-311)."""
    return _OUCov.apply(t, sigma_f, sigma_n, lamb)


# ------------------------------------------------------------------------------------------------------------------
# K3  Cholesky, potri
# ------------------------------------------------------------------------------------------------------------------
def _cond_arg(cond_bound, z: int, device):
    """Per-matrix upper bounds of the 2-norm condition number (`[z]` CUDA fp64) for the precision gate of the
    factorisation, or None: without a bound no matrix takes the split-integer tensor-core update."""
    if cond_bound is None:
        return None
    c = _req(torch.as_tensor(cond_bound, dtype=torch.float64, device=device).reshape(-1), "cond_bound").contiguous()
    if c.numel() == 1 and z != 1:
        c = c.expand(z).contiguous()
    if c.numel() != z:
        raise RuntimeError(f"cond_bound: expected {z} entries, got {c.numel()}")
    return c


def ou_cond_bound(sigma_f: torch.Tensor, sigma_n: torch.Tensor, n_nodes: int) -> torch.Tensor:
    """`(n sf^2 + sn^2) / sn^2 >= cond_2` of an OU covariance over `n_nodes` nodes and of every Schur complement of it
    (lambda_min >= sn^2, the norm is at most the trace): the bound the OU entry points of the library use themselves."""
    f2, n2 = sigma_f.detach() ** 2, sigma_n.detach() ** 2
    return (float(n_nodes) * f2 + n2) / n2


def potrf(A: torch.Tensor, check: bool = True, cond_bound=None) -> torch.Tensor:
    """Batched lower Cholesky factor with zeros above the diagonal, like `torch.linalg.cholesky` (out of place).
    `cond_bound` (optional, per matrix): see `_cond_arg`."""
    _req(A, "A")
    n = A.shape[-1]
    L = A.reshape(-1, n, n).clone()
    info = torch.empty(L.shape[0], dtype=torch.int32, device=A.device)
    cb = _cond_arg(cond_bound, L.shape[0], A.device)
    wsb = lib.drp_factor_workspace_bytes(n, L.shape[0]) if (n >= 1024 and cb is not None) else 0
    ws = torch.empty(wsb, dtype=torch.uint8, device=A.device) if wsb else None
    _rc(lib.drp_potrf_batched_f64(_p(L), n, L.shape[0], 1, _p(info), _p(cb), _p(ws), wsb, *_aux(A.device), _stream()),
        "drp_potrf_batched_f64")
    if check:
        _info_done(info, "linalg.cholesky")
    return L.reshape(A.shape)


def potri(K: torch.Tensor, cond_bound=None) -> torch.Tensor:
    """Inverse of a batch of SPD matrices through the partial-Cholesky engine."""
    _req(K, "K")
    n = K.shape[-1]
    Kc = K.reshape(-1, n, n).contiguous()
    z = Kc.shape[0]
    out = torch.empty_like(Kc)
    M = torch.empty(lib.drp_ou_workspace_elems(n, n, z), dtype=torch.float64, device=K.device)
    info = torch.empty(z, dtype=torch.int32, device=K.device)
    _rc(lib.drp_potri_batched_f64(_p(Kc), n, z, _p(out), _p(M), _p(info), _p(_cond_arg(cond_bound, z, K.device)),
                                  *_aux(K.device), _stream()), "drp_potri_batched_f64")
    _info_done(info, "potri")
    return out.reshape(K.shape)


# ------------------------------------------------------------------------------------------------------------------
# K3+K4+K5  fused OU prior  log N(x_z; 0, K_z(D))
# ------------------------------------------------------------------------------------------------------------------
class _OUPrior(torch.autograd.Function):
    @staticmethod
    def forward(ctx, D, sf, sn, lam, x):
        Dv, ldd = _matrix_view(D, "patristic_matrix")
        sf, sn, lam, x = (_req(v, k).contiguous() for v, k in ((sf, "sigma_f"), (sn, "sigma_n"), (lam, "lambd"),
                                                                  (x, "latent_z")))
        z, n = x.shape
        if Dv.shape != (n, n) or sf.shape != (z,) or sn.shape != (z,) or lam.shape != (z,):
            raise RuntimeError(f"ou_prior_logp: inconsistent shapes D{tuple(Dv.shape)} x{tuple(x.shape)} sf{tuple(sf.shape)}")
        want = any(ctx.needs_input_grad[1:])
        dev = x.device
        logp = torch.empty(z, dtype=torch.float64, device=dev)
        info = torch.empty(z, dtype=torch.int32, device=dev)
        M = torch.empty(lib.drp_ou_workspace_elems(n, n if want else 0, z), dtype=torch.float64, device=dev)
        alpha = gsum = part = None
        if want:
            alpha = torch.empty((z, n), dtype=torch.float64, device=dev)
            gsum = torch.empty((z, 3), dtype=torch.float64, device=dev)
            part = torch.empty((z, n, 3), dtype=torch.float64, device=dev)
        _rc(lib.drp_ou_prior_f64(_p(Dv), ldd, n, _p(sf), _p(sn), _p(lam), _p(x), z, int(want), _p(logp), _p(alpha),
                                 _p(gsum), _p(M), _p(part), _p(info), *_aux(dev), _stream()), "drp_ou_prior_f64")
        _info_done(info, "linalg.cholesky")
        if want:
            ctx.save_for_backward(sf, sn, lam, alpha, gsum)
        return logp

    @staticmethod
    def backward(ctx, g):
        sf, sn, lam, alpha, gsum = ctx.saved_tensors
        d_sf = g * 2.0 * sf * gsum[:, 0]
        d_sn = g * 2.0 * sn * gsum[:, 2]
        d_lam = g * sf * sf * gsum[:, 1] / (lam * lam)
        d_x = -g[:, None] * alpha
        return None, d_sf, d_sn, d_lam, d_x


_ELBO_DEPTH = 0


@contextlib.contextmanager
def elbo_scope():
    """Marks an ELBO evaluation (guide followed by model): only there is it worth starting the prior's factorisation
    from the guide — a stand-alone `guide(...)` call (S/main.py:1097) must not launch one."""
    global _ELBO_DEPTH
    _ELBO_DEPTH += 1
    try:
        yield
    finally:
        _ELBO_DEPTH -= 1


def in_elbo_scope() -> bool:
    return _ELBO_DEPTH > 0


class PriorFactor:
    """Phase 1 of the two-phase OU prior: the factored workspace of `K(sigma_f, sigma_n, lambd)` (L and -K^-1), enqueued on
    `stream`.  `key` are the objects it was derived from: a consumer presents the same ones (`matches`) or does not get it."""
    __slots__ = ("D", "ldd", "key", "M", "info", "stream", "n", "z", "used", "inputs")

    def matches(self, D, key) -> bool:
        Dv, ldd = _matrix_view(D, "patristic_matrix")
        return (not self.used and Dv.data_ptr() == self.D.data_ptr() and ldd == self.ldd and Dv.shape[0] == self.n
                and len(key) == len(self.key) and all(a is b for a, b in zip(key, self.key)))


def ou_prior_factor(patristic: torch.Tensor, sigma_f, sigma_n, lambd, key=(), stream=None, before_launch=None) -> PriorFactor:
    """Assemble and factor the OU covariances of all latent dimensions (everything of the prior that costs time) before the
    latent vectors exist; `ou_prior_logp(..., factor=...)` completes the log-density and its gradients.  Runs on `stream`
    (default: the current one) after that stream has waited for the current one."""
    Dv, ldd = _matrix_view(patristic, "patristic_matrix")
    f = PriorFactor()
    f.D, f.ldd, f.key, f.used = Dv, ldd, tuple(key), False
    sf, sn, lam = (_req(v.detach(), k).contiguous() for v, k in ((sigma_f, "sigma_f"), (sigma_n, "sigma_n"), (lambd, "lambd")))
    f.n, f.z = Dv.shape[0], sf.shape[0]
    cur = torch.cuda.current_stream()
    f.stream = cur if stream is None else stream
    f.inputs = (sf, sn, lam)
    if f.stream != cur:
        f.stream.wait_stream(cur)
        # these temporaries were allocated on the CURRENT stream and are read by kernels on f.stream: without this the caching
        # allocator may hand their memory to the next allocation of the current stream as soon as the Python references die,
        # i.e. while (in a replayed CUDA graph: possibly long before) the other stream's kernels read them
        for t in (sf, sn, lam):
            t.record_stream(f.stream)
    with torch.cuda.stream(f.stream):
        if before_launch is not None:      # e.g. a host-fed step: THIS stream (not the caller's) waits for the matrix to arrive
            before_launch()
        f.M = torch.empty(lib.drp_ou_workspace_elems(f.n, f.n, f.z), dtype=torch.float64, device=sf.device)
        f.info = torch.empty(f.z, dtype=torch.int32, device=sf.device)
        _rc(lib.drp_ou_prior_factor_f64(_p(Dv), ldd, f.n, _p(sf), _p(sn), _p(lam), f.z, _p(f.M), _p(f.info), *_aux(sf.device),
                                        _stream()), "drp_ou_prior_factor_f64")
    return f


class _OUPriorFromFactor(torch.autograd.Function):
    """log N(x_z; 0, K_z) and its gradients from a `PriorFactor` (must run on the factor's stream)."""

    @staticmethod
    def forward(ctx, sf, sn, lam, x, factor):
        sf, sn, lam, x = (_req(v, k).contiguous() for v, k in ((sf, "sigma_f"), (sn, "sigma_n"), (lam, "lambd"), (x, "latent_z")))
        z, n = x.shape
        if (n, z) != (factor.n, factor.z):
            raise RuntimeError(f"ou_prior_logp: latent_z {tuple(x.shape)} does not fit the factor ({factor.z}, {factor.n})")
        dev = x.device
        logp = torch.empty(z, dtype=torch.float64, device=dev)
        alpha = torch.empty((z, n), dtype=torch.float64, device=dev)
        gsum = torch.empty((z, 3), dtype=torch.float64, device=dev)
        part = torch.empty((z, n, 3), dtype=torch.float64, device=dev)
        yp = torch.empty(lib.drp_ou_prior_apply_workspace_elems(n, z), dtype=torch.float64, device=dev)
        _rc(lib.drp_ou_prior_apply_f64(_p(factor.D), factor.ldd, n, _p(lam), _p(x), z, _p(factor.M), _p(logp), _p(alpha), _p(gsum),
                                       _p(part), _p(yp), _stream()), "drp_ou_prior_apply_f64")
        factor.used = True
        factor.M = None                           # the workspace goes back to the allocator (stream-ordered)
        _info_done(factor.info, "linalg.cholesky")
        ctx.save_for_backward(sf, sn, lam, alpha, gsum)
        return logp

    @staticmethod
    def backward(ctx, g):
        sf, sn, lam, alpha, gsum = ctx.saved_tensors
        d_sf = g * 2.0 * sf * gsum[:, 0]
        d_sn = g * 2.0 * sn * gsum[:, 2]
        d_lam = g * sf * sf * gsum[:, 1] / (lam * lam)
        d_x = -g[:, None] * alpha
        return d_sf, d_sn, d_lam, d_x, None


def ou_prior_logp(patristic: torch.Tensor, sigma_f, sigma_n, lambd, latent_z, factor: Optional[PriorFactor] = None) -> torch.Tensor:
    """`[z]` log-densities of the zero-mean OU process prior at `latent_z [z,n]` (S/models.py:100-118), with
    analytic gradients w.r.t. sigma_f, sigma_n, lambd and latent_z.  The covariance is never materialised."""
    if factor is not None:
        return _OUPriorFromFactor.apply(sigma_f, sigma_n, lambd, latent_z, factor)
    return _OUPrior.apply(patristic, sigma_f, sigma_n, lambd, latent_z)


# ------------------------------------------------------------------------------------------------------------------
# K4/K5  general MVN log-prob from an explicit covariance
# ------------------------------------------------------------------------------------------------------------------
class _MVNLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, K, x):
        K, x = _req(K, "covariance_matrix").contiguous(), _req(x, "value").contiguous()
        z, n = x.shape
        want = any(ctx.needs_input_grad)
        dev = x.device
        logp = torch.empty(z, dtype=torch.float64, device=dev)
        info = torch.empty(z, dtype=torch.int32, device=dev)
        M = torch.empty(lib.drp_ou_workspace_elems(n, n if want else 0, z), dtype=torch.float64, device=dev)
        alpha = torch.empty((z, n), dtype=torch.float64, device=dev) if want else None
        _rc(lib.drp_mvn_logprob_f64(_p(K), _p(x), n, z, int(want), _p(logp), _p(alpha), _p(M), _p(info), None, *_aux(dev),
                                    _stream()), "drp_mvn_logprob_f64")
        _info_done(info, "linalg.cholesky")
        if want:
            ctx.save_for_backward(alpha, M)
            ctx.n = n
        return logp

    @staticmethod
    def backward(ctx, g):
        alpha, M = ctx.saved_tensors
        z, n = alpha.shape
        g = g.contiguous()
        dK = None
        if ctx.needs_input_grad[0]:
            dK = torch.empty((z, n, n), dtype=torch.float64, device=g.device)
            _rc(lib.drp_mvn_grad_cov_f64(_p(M), n, z, _p(g), _p(dK), _stream()), "drp_mvn_grad_cov_f64")
        return dK, -g[:, None] * alpha


def mvn_logprob(covariance_matrix: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
    """`[z]` zero-mean MVN log-densities for `covariance_matrix [z,n,n]`, `value [z,n]`."""
    return _MVNLogProb.apply(covariance_matrix, value)


# ------------------------------------------------------------------------------------------------------------------
# K6  conditional posterior
# ------------------------------------------------------------------------------------------------------------------
def ou_conditional(patristic_full: torch.Tensor, leaf_pos: torch.Tensor, internal_pos: torch.Tensor, sigma_f, sigma_n,
                   lambd, x_leaf, jitter: float = 0.0, want_cov: bool = True):
    """Mean `[z,ni]` (and covariance `[z,ni,ni]` + jitter) of the OU process at `internal_pos` given `x_leaf [z,n]`
    at `leaf_pos` (positions = rows of `patristic_full`).  Schur-complement form of S/models.py:167-193."""
    Dv, ldd = _matrix_view(patristic_full, "patristic_matrix")
    sf, sn, lam, x = (_req(v, k).contiguous() for v, k in ((sigma_f, "sigma_f"), (sigma_n, "sigma_n"), (lambd, "lambd"),
                                                              (x_leaf, "latent_z")))
    z, n = x.shape
    ni = int(internal_pos.numel())
    dev = x.device
    idx = torch.cat((leaf_pos.reshape(-1), internal_pos.reshape(-1))).to(device=dev, dtype=torch.int32).contiguous()
    assert idx.numel() == n + ni
    mean = torch.empty((z, ni), dtype=torch.float64, device=dev)
    cov = torch.empty((z, ni, ni), dtype=torch.float64, device=dev) if want_cov else None
    info = torch.empty(z, dtype=torch.int32, device=dev)
    M = torch.empty(lib.drp_ou_workspace_elems(n, ni, z), dtype=torch.float64, device=dev)
    _rc(lib.drp_ou_conditional_f64(_p(Dv), ldd, _p(idx), n, ni, _p(sf), _p(sn), _p(lam), _p(x), z, float(jitter), _p(mean),
                                   _p(cov), _p(M), _p(info), *_aux(dev), _stream()), "drp_ou_conditional_f64")
    _info_done(info, "linalg.inv")
    return mean, cov


# ------------------------------------------------------------------------------------------------------------------
# K10  categorical log-likelihood
# ------------------------------------------------------------------------------------------------------------------
class _CatLogLik(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, obs):
        scores = _req(scores, "logits").contiguous()
        if obs.dtype not in (torch.float64, torch.int64) or not obs.is_cuda:
            raise RuntimeError("cat_loglik: obs must be a CUDA fp64 or int64 tensor")
        obs = obs.contiguous()
        A = scores.shape[-1]
        rows = scores.numel() // A
        if obs.numel() != rows:
            raise RuntimeError(f"cat_loglik: {obs.numel()} observations for {rows} sites")
        dev = scores.device
        out = torch.empty((), dtype=torch.float64, device=dev)
        lse = torch.empty(rows, dtype=torch.float64, device=dev)
        part = torch.empty(lib.drp_cat_loglik_partials(rows), dtype=torch.float64, device=dev)
        f64 = int(obs.dtype == torch.float64)
        _rc(lib.drp_cat_loglik_fwd_f64(_p(scores), _p(obs), f64, rows, A, _p(out), _p(lse), _p(part), _stream()),
            "drp_cat_loglik_fwd_f64")
        ctx.save_for_backward(scores, obs, lse)
        return out

    @staticmethod
    def backward(ctx, g):
        scores, obs, lse = ctx.saved_tensors
        A = scores.shape[-1]
        rows = scores.numel() // A
        g = g.contiguous().reshape(1)
        ds = torch.empty_like(scores)
        _rc(lib.drp_cat_loglik_bwd_f64(_p(scores), _p(obs), int(obs.dtype == torch.float64), _p(lse), _p(g), rows, A, _p(ds),
                                       _stream()), "drp_cat_loglik_bwd_f64")
        return ds, None


def cat_loglik(scores: torch.Tensor, obs: torch.Tensor) -> torch.Tensor:
    """Scalar sum over sites of log softmax(scores)[obs] (S/models.py:352)."""
    return _CatLogLik.apply(scores, obs)


# ------------------------------------------------------------------------------------------------------------------
# the small sample sites: summed HalfNormal / Normal log-densities, one launch each way
# ------------------------------------------------------------------------------------------------------------------
def _strides2(t: torch.Tensor):
    if t.dim() == 0:
        return 0, 0
    if t.dim() == 1:
        return 0, t.stride(0)
    return t.stride(0), t.stride(1)


class _SiteLogP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, loc, scale, mode):
        rows, cols = (1, max(value.numel(), 1)) if value.dim() < 2 else tuple(value.shape)
        st = (ctypes.c_int64 * 6)(*_strides2(value), *(_strides2(loc) if loc is not None else (0, 0)), *_strides2(scale))
        out = torch.empty((), dtype=torch.float64, device=value.device)
        part = torch.empty(lib.drp_site_logp_workspace_elems(), dtype=torch.float64, device=value.device)
        _rc(lib.drp_site_logp_f64(_p(value), _p(loc), _p(scale), st, rows, cols, mode, _p(out), _p(part), _stream()),
            "drp_site_logp_f64")
        ctx.save_for_backward(value, loc, scale)
        ctx.geom = (rows, cols, mode, st)
        return out

    @staticmethod
    def backward(ctx, g):
        value, loc, scale = ctx.saved_tensors
        rows, cols, mode, st = ctx.geom
        need_v, need_m, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and loc is not None, ctx.needs_input_grad[2]
        mk = lambda need: torch.empty(value.shape, dtype=torch.float64, device=value.device) if need else None
        dv, dm, ds = mk(need_v), mk(need_m), mk(need_s)
        g = g.contiguous()
        _rc(lib.drp_site_logp_bwd_f64(_p(value), _p(loc), _p(scale), st, rows, cols, mode, _p(g), _p(dv), _p(dm), _p(ds), _stream()),
            "drp_site_logp_bwd_f64")
        return dv, dm, ds, None


def site_logp_sum(value: torch.Tensor, loc: Optional[torch.Tensor], scale: torch.Tensor) -> torch.Tensor:
    """Scalar `HalfNormal(scale).log_prob(value).sum()` (loc None) or `Normal(loc, scale).log_prob(value).sum()` for CUDA fp64
    tensors of at most two dimensions; `loc` / `scale` broadcast against `value` (expanded scalars and transposed views are
    read in place).  S/models.py:89-92, S/guides.py:118-121."""
    _req(value, "value"), _req(scale, "scale")
    # torch's broadcasting, exactly: the variational guide's [k,1]-shaped values meet [k]-shaped scales in the model and the
    # reference's log_prob is the [k,k] outer broadcast of the two (S/guides.py:33-36 vs S/models.py:89-92) — so is this one
    if loc is None:
        value, scale = torch.broadcast_tensors(value, scale)
    else:
        value, loc, scale = torch.broadcast_tensors(value, _req(loc, "loc"), scale)
    if value.dim() > 2:
        raise RuntimeError("site_logp_sum: at most two dimensions after broadcasting")
    return _SiteLogP.apply(value, loc, scale, 0 if loc is None else 1)


# ------------------------------------------------------------------------------------------------------------------
# K1  patristic / cladistic / LCA
# ------------------------------------------------------------------------------------------------------------------
def tree_distances(parent: torch.Tensor, branch: torch.Tensor, rows: Optional[torch.Tensor] = None,
                   cols: Optional[torch.Tensor] = None, want_lca: bool = False, want_cladistic: bool = False,
                   want_patristic: bool = True):
    """All-pairs (or rows x cols) patristic / cladistic distances and LCA ids of a tree given as level-order
    `parent [N] int32` (root: -1) and `branch [N] fp64` device tensors.  Returns (patristic, cladistic, lca)."""
    if not parent.is_cuda or parent.dtype != torch.int32:
        raise RuntimeError("tree_distances: parent must be a CUDA int32 tensor")
    _req(branch, "branch")
    dev = parent.device
    N = int(parent.numel())
    parent, branch = parent.contiguous(), branch.contiguous()
    depth = torch.empty(N, dtype=torch.int32, device=dev)
    pre, size, scratch = torch.empty_like(depth), torch.empty_like(depth), torch.empty_like(depth)
    rdist = torch.empty(N, dtype=torch.float64, device=dev)
    _rc(lib.drp_tree_prepare(_p(parent), _p(branch), N, _p(depth), _p(rdist), _p(pre), _p(size), _p(scratch), _stream()),
        "drp_tree_prepare")
    max_depth = int(depth.max())
    rows_t = None if rows is None else rows.to(device=dev, dtype=torch.int32).contiguous()
    cols_t = None if cols is None else cols.to(device=dev, dtype=torch.int32).contiguous()
    nr = N if rows_t is None else int(rows_t.numel())
    nc = N if cols_t is None else int(cols_t.numel())
    wsb = lib.drp_lca_workspace_bytes(nr, max_depth)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb else None
    pat = torch.empty((nr, nc), dtype=torch.float64, device=dev) if want_patristic else None
    cla = torch.empty((nr, nc), dtype=torch.float64, device=dev) if want_cladistic else None
    lca = torch.empty((nr, nc), dtype=torch.int32, device=dev) if want_lca else None
    _rc(lib.drp_lca_patristic_f64(_p(parent), _p(depth), _p(rdist), _p(pre), _p(size), N, max_depth, _p(rows_t), nr,
                                  _p(cols_t), nc, _p(lca), _p(pat), _p(cla), _p(ws), wsb, _stream()),
        "drp_lca_patristic_f64")
    return pat, cla, lca


# ------------------------------------------------------------------------------------------------------------------
# bidirectional GRU recurrence (decoder / encoder)
# ------------------------------------------------------------------------------------------------------------------
GRU_HIDDEN = lib.drp_gru_hidden_size()


def require_gru_hidden(H: int):
    """The recurrence kernels are built for the reference's gru_hidden_dim (60, S/main.py:2777).  Another size is refused
    rather than sent to cuDNN's `nn.GRU` behind the caller's back: this package has no library fallback."""
    if H != GRU_HIDDEN:
        raise RuntimeError(f"gru_hidden_dim = {H}: the sm_100a GRU kernels of draupnir_asr_b200 are built for {GRU_HIDDEN} "
                           "(rebuild csrc/gru.cu with another GH); there is no cuDNN fallback")


class _GRUBidir(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gi, whh, bhh, h0):
        gi, whh, bhh, h0 = (_req(v, k).contiguous() for v, k in ((gi, "gi"), (whh, "weight_hh"), (bhh, "bias_hh"), (h0, "h0")))
        n, L = gi.shape[0], gi.shape[1]
        H = whh.shape[-1]
        if gi.numel() != n * L * 2 * 3 * H or whh.shape != (2, 3 * H, H) or bhh.shape != (2, 3 * H) or h0.shape != (2, n, H):
            raise RuntimeError(f"gru_bidirectional: inconsistent shapes gi{tuple(gi.shape)} whh{tuple(whh.shape)} "
                               f"bhh{tuple(bhh.shape)} h0{tuple(h0.shape)}")
        want = any(ctx.needs_input_grad)
        out = torch.empty((n, L, 2 * H), dtype=torch.float64, device=gi.device)
        gates = torch.empty((n, L, 2, 4, H), dtype=torch.float64, device=gi.device) if want else None
        _rc(lib.drp_gru_fwd_f64(_p(gi), _p(whh), _p(bhh), _p(h0), n, L, H, _p(out), _p(gates), _stream()), "drp_gru_fwd_f64")
        if want:
            ctx.save_for_backward(out, gates, whh, h0)
        return out

    @staticmethod
    def backward(ctx, dout):
        dgi, dwhh, dbhh, dh0 = _gru_backward_common(ctx.saved_tensors, dout, ctx.needs_input_grad[1] or ctx.needs_input_grad[2],
                                                    ctx.needs_input_grad[3])
        return dgi if ctx.needs_input_grad[0] else None, dwhh, dbhh, dh0


def _gru_backward_common(ctx_saved, dout, need_w, need_h0):
    """BPTT + weight gradients shared by the two input modes: returns (dgi, dwhh, dbhh, dh0)."""
    out, gates, whh, h0 = ctx_saved
    n, L, H2 = out.shape
    H = H2 // 2
    dout = dout.contiguous()
    dgi = torch.empty((n, L, 2, 3 * H), dtype=torch.float64, device=out.device)
    dh0 = torch.empty_like(h0) if need_h0 else None
    _rc(lib.drp_gru_bwd_f64(_p(dout), _p(out), _p(gates), _p(whh), _p(h0), n, L, H, _p(dgi), _p(dh0), _stream()),
        "drp_gru_bwd_f64")
    dwhh = dbhh = None
    if need_w:
        dwhh = torch.empty_like(whh)
        dbhh = torch.empty((2, 3 * H), dtype=torch.float64, device=out.device)
        wse = lib.drp_gru_wgrad_workspace_elems()
        ws = torch.empty(wse, dtype=torch.float64, device=out.device)
        _rc(lib.drp_gru_wgrad_f64(_p(dgi), _p(gates), _p(out), _p(h0), n, L, H, _p(dwhh), _p(dbhh), _p(ws), wse, _stream()),
            "drp_gru_wgrad_f64")
    return dgi, dwhh, dbhh, dh0


class _GRUBidirUV(torch.autograd.Function):
    """Same recurrence with `gi[i, l] = U[i] + V[l]` formed inside the kernel (never materialised)."""

    @staticmethod
    def forward(ctx, U, V, whh, bhh, h0):
        U, V, whh, bhh, h0 = (_req(v, k).contiguous() for v, k in ((U, "U"), (V, "V"), (whh, "weight_hh"), (bhh, "bias_hh"),
                                                                    (h0, "h0")))
        n, L, H = U.shape[0], V.shape[0], whh.shape[-1]
        if U.numel() != n * 6 * H or V.numel() != L * 6 * H or whh.shape != (2, 3 * H, H) or bhh.shape != (2, 3 * H) \
                or h0.shape != (2, n, H):
            raise RuntimeError(f"gru_bidirectional_uv: inconsistent shapes U{tuple(U.shape)} V{tuple(V.shape)} "
                               f"whh{tuple(whh.shape)} h0{tuple(h0.shape)}")
        want = any(ctx.needs_input_grad)
        out = torch.empty((n, L, 2 * H), dtype=torch.float64, device=U.device)
        gates = torch.empty((n, L, 2, 4, H), dtype=torch.float64, device=U.device) if want else None
        _rc(lib.drp_gru_fwd_uv_f64(_p(U), _p(V), _p(whh), _p(bhh), _p(h0), n, L, H, _p(out), _p(gates), _stream()),
            "drp_gru_fwd_uv_f64")
        if want:
            ctx.save_for_backward(out, gates, whh, h0)
            ctx.shapes = (U.shape, V.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        dgi, dwhh, dbhh, dh0 = _gru_backward_common(ctx.saved_tensors, dout, ctx.needs_input_grad[2] or ctx.needs_input_grad[3],
                                                    ctx.needs_input_grad[4])
        dU = dV = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            # both sums in one pass over dgi (2.9 GB at 2,000 leaves x 500 sites)
            n, L = dgi.shape[0], dgi.shape[1]
            C = dgi.numel() // (n * L)
            nblk, nch = lib.drp_gru_input_sums_blocks(n), lib.drp_gru_input_sums_chunks(n, L)
            dUp = torch.empty((nch, n, C), dtype=torch.float64, device=dgi.device)
            dVp = torch.empty((nblk, L, C), dtype=torch.float64, device=dgi.device)
            _rc(lib.drp_gru_input_sums_f64(_p(dgi), n, L, C, _p(dUp), _p(dVp), _stream()), "drp_gru_input_sums_f64")
            dU = (dUp[0] if nch == 1 else dUp.sum(0)).reshape(ctx.shapes[0])
            dV = dVp.sum(0).reshape(ctx.shapes[1])
        return dU, dV, dwhh, dbhh, dh0


def gru_bidirectional_uv(U: torch.Tensor, V: torch.Tensor, weight_hh, bias_hh, h0) -> torch.Tensor:
    """`gru_bidirectional(U[:, None] + V[None], ...)` for `U [n, 2, 3H]` (or `[n, 6H]`) and `V [L, 2, 3H]` without the
    `[n, L, 2, 3H]` intermediate: the decoder's input is `[latent_i ; blosum_l]` (S/models.py:339-342)."""
    return _GRUBidirUV.apply(U, V, weight_hh, bias_hh, h0)


_GRAM_OFF = os.environ.get("DRP_GRAM", "1") == "0"      # A/B switch for measurements


def _aligned16(t: torch.Tensor) -> torch.Tensor:
    """The TMA bulk copies of `tall_gemm` / `gram_tn` need 16-byte-aligned bases.  A contiguous row slice such as
    `data_blosum[leaf_lo:leaf_hi]` (the sharded encoder's input) starts at `leaf_lo * L * 21` doubles, which is odd when
    `leaf_lo` and `L` are: such a view is copied to a fresh (512-byte-aligned) allocation once, instead of failing."""
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def gram_tn(A: torch.Tensor, B: torch.Tensor):
    """`(A.t() @ B, A.sum(0))` for tall-skinny fp64 `A [K,M]`, `B [K,N]` (K huge, M, N <= 192): the weight / bias gradient
    shapes of the dense layers around the recurrences, as one split-K kernel instead of a cuBLAS "tn" GEMM + a reduction."""
    _req(A, "A"), _req(B, "B")
    K, M = A.shape
    N = B.shape[1]
    if B.shape[0] != K:
        raise RuntimeError(f"gram_tn: A has {K} rows, B has {B.shape[0]}")
    if _GRAM_OFF or not lib.drp_gram_tn_supported(M, N):
        return A.t() @ B, A.sum(0)                  # shapes outside the kernel's range (never on the DRAUPNIR path): cuBLAS
    A, B = _aligned16(A.contiguous()), _aligned16(B.contiguous())
    C = torch.empty((M, N), dtype=torch.float64, device=A.device)
    cs = torch.empty(M, dtype=torch.float64, device=A.device)
    wse = lib.drp_gram_workspace_elems(M, N)
    ws = torch.empty(wse, dtype=torch.float64, device=A.device)
    _rc(lib.drp_gram_tn_f64(_p(A), _p(B), K, M, N, _p(C), _p(cs), _p(ws), wse, _stream()), "drp_gram_tn_f64")
    return C, cs


_TALL_OFF = os.environ.get("DRP_TALL_GEMM", "1") == "0"      # A/B switch for measurements


def tall_gemm(x: torch.Tensor, Wt: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`x @ Wt + bias` for row-major `x [M,K]`, `Wt [K,N]` (M huge, K <= 192, N <= 184) as one HBM-bound kernel; shapes
    outside that range (or DRP_TALL_GEMM=0) go to cuBLAS."""
    _req(x, "x"), _req(Wt, "Wt")
    M, K = x.shape
    N = Wt.shape[1]
    if _TALL_OFF or K > 192 or N > 184 or M == 0:
        return torch.addmm(bias, x, Wt) if bias is not None else x @ Wt
    x, Wt = _aligned16(x.contiguous()), Wt.contiguous()
    b = None if bias is None else _req(bias, "bias").contiguous()
    y = torch.empty((M, N), dtype=torch.float64, device=x.device)
    rc = lib.drp_tall_gemm_f64(_p(x), _p(Wt), _p(b), M, K, N, _p(y), _stream())
    if rc == 1:                                   # declined (W^T + ring exceed shared memory): library GEMM
        return torch.addmm(bias, x, Wt) if bias is not None else x @ Wt
    _rc(rc, "drp_tall_gemm_f64")
    return y


class _AffineTall(torch.autograd.Function):
    """`x @ W.t() + b` for `x [M,K]` with M in the millions and small K, N: forward and input gradient through
    `tall_gemm` (one read, one write), weight and bias gradients through `gram_tn` (split-K) — cuBLAS runs these fp64
    shapes far below what their memory traffic allows."""

    @staticmethod
    def forward(ctx, x, W, b):
        ctx.save_for_backward(x, W)
        return tall_gemm(x, W.t().contiguous(), b)

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = dy.contiguous()
        dW = db = dx = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dW, db = gram_tn(dy, x)
        if ctx.needs_input_grad[0]:
            dx = tall_gemm(dy, W)
        return dx, dW, db


def affine_tall(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return _AffineTall.apply(x, W, b)


class _GRUEncoderLast(torch.autograd.Function):
    """Hidden state at the LAST position, `output[:, -1]` `[n, 2H]`, of a bidirectional GRU over `x [n,L,I]` — all that
    `RNNEncoder.forward` feeds to `fc1` (S/models_utils.py:57).  The forward direction runs the recurrence kernel for
    direction 0 alone (16 instead of 32 sequences per CTA at 2,000 leaves, half the input projection); the reverse
    direction reaches position L-1 after ONE cell step from `h0[1]`, done with dense ops.  Back-propagation likewise."""

    @staticmethod
    def forward(ctx, x, wih, bih, whh, bhh, h0):
        x, wih, bih, whh, bhh, h0 = (_req(v, k).contiguous() for v, k in ((x, "input"), (wih, "weight_ih"), (bih, "bias_ih"),
                                                                          (whh, "weight_hh"), (bhh, "bias_hh"), (h0, "h0")))
        n, L, I = x.shape
        H = whh.shape[-1]
        gi0 = tall_gemm(x.reshape(n * L, I), wih[0].t().contiguous(), bih[0])                      # [n*L, 3H]
        want = any(ctx.needs_input_grad)
        out = torch.empty((n, L, 2 * H), dtype=torch.float64, device=x.device)       # direction-0 half is written
        gates = torch.empty((n, L, 2, 4, H), dtype=torch.float64, device=x.device) if want else None
        _rc(lib.drp_gru_fwd_dir0_f64(_p(gi0), _p(whh), _p(bhh), _p(h0), n, L, H, _p(out), _p(gates), _stream()),
            "drp_gru_fwd_dir0_f64")
        del gi0
        # reverse direction: position L-1 is its first step (h_prev = h0[1])
        gi1 = torch.addmm(bih[1], x[:, L - 1], wih[1].t())
        gh1 = torch.addmm(bhh[1], h0[1], whh[1].t())
        r = torch.sigmoid(gi1[:, :H] + gh1[:, :H])
        z = torch.sigmoid(gi1[:, H:2 * H] + gh1[:, H:2 * H])
        q = gh1[:, 2 * H:]
        nn = torch.tanh(gi1[:, 2 * H:] + r * q)
        h1 = (1.0 - z) * nn + z * h0[1]
        if want:
            ctx.save_for_backward(x, wih, whh, h0, out, gates, torch.stack((r, z, nn, q), dim=1))
        return torch.cat((out[:, L - 1, :H], h1), dim=1)

    @staticmethod
    def backward(ctx, dlast):
        x, wih, whh, h0, out, gates, g1 = ctx.saved_tensors
        n, L, I = x.shape
        H = whh.shape[-1]
        dev = x.device
        dlast = dlast.contiguous()
        # ---- forward direction: full BPTT, gradient enters at the last position only
        dgi0 = torch.empty((n, L, 3 * H), dtype=torch.float64, device=dev)
        dh0 = torch.empty_like(h0)
        _rc(lib.drp_gru_bwd_last_dir0_f64(_p(dlast), _p(out), _p(gates), _p(whh), _p(h0), n, L, H, _p(dgi0), _p(dh0), _stream()),
            "drp_gru_bwd_last_dir0_f64")
        dwhh = torch.empty_like(whh)
        dbhh = torch.empty((2, 3 * H), dtype=torch.float64, device=dev)
        wse = lib.drp_gru_wgrad_workspace_elems()
        ws = torch.empty(wse, dtype=torch.float64, device=dev)
        _rc(lib.drp_gru_wgrad_dir0_f64(_p(dgi0), _p(gates), _p(out), _p(h0), n, L, H, _p(dwhh), _p(dbhh), _p(ws), wse, _stream()),
            "drp_gru_wgrad_dir0_f64")
        # ---- reverse direction: its single relevant cell step
        r, z, nn, q = g1[:, 0], g1[:, 1], g1[:, 2], g1[:, 3]
        d = dlast[:, H:]
        dn = d * (1.0 - z) * (1.0 - nn * nn)
        dz = d * (h0[1] - nn) * z * (1.0 - z)
        dr = dn * q * r * (1.0 - r)
        dgi1 = torch.cat((dr, dz, dn), dim=1)                    # [n, 3H] at position L-1
        dgh1 = torch.cat((dr, dz, dn * r), dim=1)
        dwhh[1] = dgh1.t() @ h0[1]
        dbhh[1] = dgh1.sum(0)
        dh0[1] = d * z + dgh1 @ whh[1]
        # ---- input projection
        dgi0f = dgi0.reshape(n * L, 3 * H)
        xf = x.reshape(n * L, I)
        dwih0, dbih0 = gram_tn(dgi0f, xf)
        dwih = torch.stack((dwih0, dgi1.t() @ x[:, L - 1]))
        dbih = torch.stack((dbih0, dgi1.sum(0)))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = tall_gemm(dgi0f, wih[0]).reshape(n, L, I)
            dx[:, L - 1] += dgi1 @ wih[1]
        return dx, dwih, dbih, dwhh, dbhh, dh0


def gru_encoder_last(rnn: torch.nn.GRU, x: torch.Tensor, h0: torch.Tensor):
    """The encoder's use of `rnn(x, h0)`: returns `(last, full)` where `last = output[:, -1]` `[n,2H]` carries the
    gradient (see `_GRUEncoderLast`) and `full` is a `utils.Deferred` whose `.value()` is `(output [n,L,2H], h_n [2,n,H])`
    without gradient — the complete bidirectional pass, run only if somebody asks, on the weights as they were at THIS
    call (`gru_stacked_params` copies them)."""
    from .utils import Deferred
    H = rnn.hidden_size
    require_gru_hidden(H)
    _req(x, "input")
    wih, whh, bih, bhh = gru_stacked_params(rnn)
    last = _GRUEncoderLast.apply(x, wih, bih, whh, bhh, h0)
    snap = tuple(t.detach() for t in (x, wih, bih, whh, bhh, h0))

    def run_full():
        xs, wih_s, bih_s, whh_s, bhh_s, h0_s = snap
        n, L, I = xs.shape
        with torch.no_grad():
            gi = torch.addmm(bih_s.reshape(-1), xs.reshape(n * L, I), wih_s.reshape(6 * H, I).t()).reshape(n, L, 2, 3 * H)
            out = gru_bidirectional(gi, whh_s, bhh_s, h0_s)
        return out, torch.stack((out[:, -1, :H], out[:, 0, H:]))

    return last, Deferred(run_full)


def gru_bidirectional(gi: torch.Tensor, weight_hh: torch.Tensor, bias_hh: torch.Tensor, h0: torch.Tensor) -> torch.Tensor:
    """Hidden states `[n, L, 2H]` (torch's bidirectional layout) of a one-layer bidirectional GRU whose input
    projections `gi = W_ih x + b_ih` are given as `[n, L, 2, 3H]`; `weight_hh [2,3H,H]`, `bias_hh [2,3H]`, `h0 [2,n,H]`
    (direction 0 forward, 1 reverse; gate order r, z, n as in `nn.GRU`).  H must equal `GRU_HIDDEN`."""
    return _GRUBidir.apply(gi, weight_hh, bias_hh, h0)


def gru_stacked_params(rnn: torch.nn.GRU):
    """(W_ih [2,3H,I], W_hh [2,3H,H], b_ih [2,3H], b_hh [2,3H]) of a one-layer bidirectional `nn.GRU`."""
    assert rnn.num_layers == 1 and rnn.bidirectional and rnn.batch_first and rnn.bias
    return (torch.stack((rnn.weight_ih_l0, rnn.weight_ih_l0_reverse)), torch.stack((rnn.weight_hh_l0, rnn.weight_hh_l0_reverse)),
            torch.stack((rnn.bias_ih_l0, rnn.bias_ih_l0_reverse)), torch.stack((rnn.bias_hh_l0, rnn.bias_hh_l0_reverse)))


def gru_module_forward(rnn: torch.nn.GRU, x: torch.Tensor, h0: torch.Tensor):
    """Drop-in for `rnn(x, h0)` of a one-layer bidirectional batch-first `nn.GRU`: returns `(output, h_n)`."""
    H = rnn.hidden_size
    require_gru_hidden(H)
    _req(x, "input")
    wih, whh, bih, bhh = gru_stacked_params(rnn)
    n, L, I = x.shape
    gi = torch.addmm(bih.reshape(-1), x.reshape(n * L, I), wih.reshape(6 * H, I).t()).reshape(n, L, 2, 3 * H)
    out = gru_bidirectional(gi, whh, bhh, h0)
    return out, torch.stack((out[:, -1, :H], out[:, 0, H:]))


# ------------------------------------------------------------------------------------------------------------------
# test hook: one trailing update of the blocked factorisation (fp64 mma.sync kernel or int8 tcgen05 kernel)
# ------------------------------------------------------------------------------------------------------------------
def syrk_update_split(M: torch.Tensor, k0: int, K: int, c1: int, R: int, col_mid: int, col_limit: int, slices: int) -> bool:
    """The same update as `syrk_update` in the two column-range parts the factorisation's look-ahead uses: columns
    [c1, col_mid) with slicing, then [col_mid, col_limit) from the same slices with the second work-unit counter."""
    _req(M, "M")
    z, rows, ld = M.shape
    wsb = lib.drp_factor_workspace_bytes(rows, z)
    ws = torch.empty(wsb, dtype=torch.uint8, device=M.device)
    for lo, hi, do_slice, cidx in ((c1, col_mid, 1, 0), (col_mid, col_limit, 0, 1)):
        rc = lib.drp_syrk_update_range_f64(_p(M), ld, rows * ld, z, k0, K, c1, R, lo, hi, slices, _p(ws), wsb, do_slice, cidx, _stream())
        if rc == 1:
            return False
        _rc(rc, "drp_syrk_update_range_f64")
    return True


def syrk_update(M: torch.Tensor, k0: int, K: int, c1: int, R: int, col_limit: int, slices: int = 0) -> bool:
    """In place on `M [z, rows, ld]`: `M[r][c] -= sum_{k0<=k<k0+K} M[r][k] M[c][k]` for `c1 <= c <= r < R`, `c < col_limit`.
    `slices = 0` runs the fp64 mma.sync kernel, `4..7` the split-integer tcgen05 kernel.  Returns False when the tcgen05
    kernel does not accept the shape (nothing is modified then)."""
    _req(M, "M")
    assert M.dim() == 3 and M.is_contiguous()
    z, rows, ld = M.shape
    wsb = lib.drp_factor_workspace_bytes(rows, z)
    ws = torch.empty(wsb, dtype=torch.uint8, device=M.device)
    rc = lib.drp_syrk_update_f64(_p(M), ld, rows * ld, z, k0, K, c1, R, col_limit, slices, _p(ws), wsb, _stream())
    if rc == 1:
        return False
    _rc(rc, "drp_syrk_update_f64")
    return True
