"""`pyro.distributions` subset: torch distributions + the Pyro mixin (`expand_by`, `to_event`), `Delta`, and the two
B200-accelerated distributions of the DRAUPNIR path.

* `MultivariateNormal` / `Categorical` given CUDA fp64 tensors score through the sm_100a kernels
  (`ops.mvn_logprob`, `ops.cat_loglik`); `OUProcessPrior` is the fused form that never materialises the covariance.
* Given CPU tensors the generic classes are plain `torch.distributions` (that is all Pyro's are); the DRAUPNIR model
  classes themselves refuse CPU tensors, so there is no CPU route through the product path.
"""
from __future__ import annotations

import math

import torch
import torch.distributions as td
from torch.distributions import constraints
from torch.distributions.utils import _sum_rightmost


def _accel(t: torch.Tensor) -> bool:
    return isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64


class TorchDistributionMixin:
    def expand_by(self, sample_shape):
        return self.expand(torch.Size(sample_shape) + self.batch_shape)

    def to_event(self, reinterpreted_batch_ndims=None):
        if reinterpreted_batch_ndims is None:
            reinterpreted_batch_ndims = len(self.batch_shape)
        if reinterpreted_batch_ndims == 0:
            return self
        return Independent(self, reinterpreted_batch_ndims)

    def mask(self, mask):
        raise NotImplementedError("mask() is not used on the DRAUPNIR path")

    def __call__(self, sample_shape=torch.Size()):
        return self.rsample(sample_shape) if self.has_rsample else self.sample(sample_shape)


class AsyncTerm:
    """A summed log-density being evaluated on a side stream.  `join()` makes the current stream wait for it and returns
    the scalar tensor; until then nothing on the current stream may touch `value`."""
    __slots__ = ("value", "stream")

    def __init__(self, value, stream):
        self.value, self.stream = value, stream

    def join(self):
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.stream)
        self.value.record_stream(cur)
        return self.value


class Independent(td.Independent, TorchDistributionMixin):
    def log_prob_sum(self, value):
        f = getattr(self.base_dist, "log_prob_sum", None)
        return f(value) if f is not None else self.log_prob(value).sum()

    def log_prob_sum_async(self, value, scale=1.0):
        f = getattr(self.base_dist, "log_prob_sum_async", None)
        return f(value, scale) if f is not None else None


class HalfNormal(td.HalfNormal, TorchDistributionMixin):
    """torch's constructor builds `Normal(0, scale)`: the Python 0 becomes a CPU tensor that is copied to the device from
    pageable memory, and such a copy stalls the launch thread until everything queued on the stream has run (four of them
    per model evaluation).  Same distribution, location created on the device."""

    def __init__(self, scale, validate_args=None):
        if isinstance(scale, torch.Tensor):
            base_dist = td.Normal(torch.zeros_like(scale), scale, validate_args=False)
            td.TransformedDistribution.__init__(self, base_dist, td.transforms.AbsTransform(), validate_args=validate_args)
        else:
            super().__init__(scale, validate_args=validate_args)

    def expand(self, batch_shape, _instance=None):
        return HalfNormal(self.base_dist.scale.expand(torch.Size(batch_shape)), validate_args=False)

    def log_prob_sum(self, value):
        """One fused kernel each way instead of ~15 element-wise launches (csrc/sites.cu) for CUDA fp64 sites."""
        s = self.base_dist.scale
        if _accel(value) and _accel(s) and value.dim() <= 2 and s.dim() <= 2 and value.numel() > 0:
            from .. import ops
            return ops.site_logp_sum(value, None, s)
        return self.log_prob(value).sum()


class Normal(td.Normal, TorchDistributionMixin):
    def log_prob_sum(self, value):
        if _accel(value) and _accel(self.loc) and _accel(self.scale) and value.dim() <= 2 and self.loc.dim() <= 2 and value.numel() > 0:
            from .. import ops
            return ops.site_logp_sum(value, self.loc, self.scale)
        return self.log_prob(value).sum()


class Delta(td.Distribution, TorchDistributionMixin):
    """Point mass (pyro.distributions.Delta): log_prob is `log_density` at the value, -inf elsewhere."""
    has_rsample = True
    arg_constraints = {"v": constraints.dependent, "log_density": constraints.real}
    support = constraints.real

    def __init__(self, v, log_density=0.0, event_dim=0, validate_args=None):
        if event_dim > v.dim():
            raise ValueError(f"Expected event_dim <= v.dim(), actual {event_dim} vs {v.dim()}")
        batch_dim = v.dim() - event_dim
        self.v = v
        self.log_density = log_density
        super().__init__(v.shape[:batch_dim], v.shape[batch_dim:], validate_args=False)

    def __repr__(self):
        return f"Delta(v: {tuple(self.v.shape)}, event_dim: {len(self.event_shape)})"

    def expand(self, batch_shape, _instance=None):
        batch_shape = torch.Size(batch_shape)
        return Delta(self.v.expand(batch_shape + self.event_shape), self.log_density, len(self.event_shape))

    def rsample(self, sample_shape=torch.Size()):
        shape = torch.Size(sample_shape) + self.v.shape
        return self.v.expand(shape)

    def log_prob(self, x):
        lp = (x == self.v).to(x.dtype).log()
        lp = _sum_rightmost(lp, len(self.event_shape))
        return lp + self.log_density

    def log_prob_sum(self, x):
        """A guide's Delta site is scored at the very tensor it handed out (`rsample` returns `v`): log 1 = 0 per element, no
        kernel needed; anything else goes through `log_prob`."""
        same = x is self.v or (isinstance(x, torch.Tensor) and x.shape == self.v.shape and x.stride() == self.v.stride()
                               and x.data_ptr() == self.v.data_ptr() and x.dtype == self.v.dtype)
        if not same:
            return self.log_prob(x).sum()
        ld = self.log_density
        if isinstance(ld, torch.Tensor):
            return ld.sum() if ld.dim() else ld
        batch = self.v.shape[:self.v.dim() - len(self.event_shape)]
        return float(ld) * float(torch.Size(batch).numel())

    @property
    def mean(self):
        return self.v


class MultivariateNormal(td.Distribution, TorchDistributionMixin):
    """`dist.MultivariateNormal(loc, covariance_matrix)` (S/models.py:118,193,239).  CUDA fp64: batched blocked
    Cholesky + solve + analytic gradients on the GPU kernels; otherwise torch's implementation."""
    has_rsample = True
    arg_constraints = {"loc": constraints.real_vector, "covariance_matrix": constraints.positive_definite}
    support = constraints.real_vector

    def __new__(cls, loc, covariance_matrix=None, precision_matrix=None, scale_tril=None, validate_args=None):
        if covariance_matrix is not None and _accel(covariance_matrix) and covariance_matrix.dim() == 3:
            return super().__new__(cls)
        return _TorchMVN(loc, covariance_matrix=covariance_matrix, precision_matrix=precision_matrix,
                         scale_tril=scale_tril, validate_args=validate_args)

    def __init__(self, loc, covariance_matrix=None, precision_matrix=None, scale_tril=None, validate_args=None):
        z, n = covariance_matrix.shape[0], covariance_matrix.shape[-1]
        self.covariance_matrix = covariance_matrix
        self.loc = loc.expand(z, n) if loc.dim() <= 2 else loc
        self._scale_tril = None
        super().__init__(torch.Size([z]), torch.Size([n]), validate_args=False)

    @property
    def scale_tril(self):
        if self._scale_tril is None:
            from .. import ops
            self._scale_tril = ops.potrf(self.covariance_matrix)
        return self._scale_tril

    @property
    def mean(self):
        return self.loc

    def log_prob(self, value):
        from .. import ops
        return ops.mvn_logprob(self.covariance_matrix, (value - self.loc).reshape(self.loc.shape))

    def rsample(self, sample_shape=torch.Size()):
        shape = torch.Size(sample_shape) + self.loc.shape
        eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        return self.loc + torch.matmul(self.scale_tril, eps.unsqueeze(-1)).squeeze(-1)      # no [S,z,n,n] intermediate


class _TorchMVN(td.MultivariateNormal, TorchDistributionMixin):
    pass


class Categorical(td.Distribution, TorchDistributionMixin):
    """`dist.Categorical(logits=...)` (S/models.py:352,389).  CUDA fp64 logits: the summed log-likelihood and its
    gradient are one fused kernel each (`log_prob_sum`), and the constructor does not touch the logits."""
    has_enumerate_support = True
    arg_constraints = {"logits": constraints.real_vector}

    def __new__(cls, probs=None, logits=None, validate_args=None):
        if logits is not None and _accel(logits):
            return super().__new__(cls)
        return _TorchCategorical(probs=probs, logits=logits, validate_args=validate_args)

    def __init__(self, probs=None, logits=None, validate_args=None):
        self.raw_logits = logits
        super().__init__(logits.shape[:-1], torch.Size(), validate_args=False)

    @property
    def logits(self):
        return torch.log_softmax(self.raw_logits, dim=-1)

    @property
    def probs(self):
        return torch.softmax(self.raw_logits, dim=-1)

    @property
    def support(self):
        return constraints.integer_interval(0, self.raw_logits.shape[-1] - 1)

    def log_prob_sum(self, value):
        from .. import ops
        return ops.cat_loglik(self.raw_logits, value)

    def log_prob(self, value):
        return self.logits.gather(-1, value.long().unsqueeze(-1)).squeeze(-1)

    def sample(self, sample_shape=torch.Size()):
        sample_shape = torch.Size(sample_shape)
        A = self.raw_logits.shape[-1]
        flat = self.probs.reshape(-1, A)
        draws = torch.multinomial(flat, max(sample_shape.numel(), 1), True).T
        return draws.reshape(sample_shape + self.batch_shape)


class _TorchCategorical(td.Categorical, TorchDistributionMixin):
    pass


class OUProcessPrior(td.Distribution, TorchDistributionMixin):
    """z independent zero-mean Ornstein-Uhlenbeck processes over the tree: N(0, sf_z^2 exp(-D/lam_z) + sn_z^2 I).

    Equivalent to `MultivariateNormal(zeros(1,n), OUKernel_Fast(sf,sn,lam).forward(D))` (S/models.py:101-118) but
    assembly, Cholesky, solve, log-det and all gradients are one fused GPU pipeline (`ops.ou_prior_logp`)."""
    has_rsample = True
    arg_constraints = {}
    support = constraints.real_vector

    def __init__(self, patristic, sigma_f, sigma_n, lambd, z_slice=None, factor=None, validate_args=None):
        self.patristic, self.sigma_f, self.sigma_n, self.lambd = patristic, sigma_f, sigma_n, lambd
        # factor: an `ops.PriorFactor` of exactly these covariances already being computed (started from the guide); the
        # log-density then only adds the part that needs the value
        self.factor = factor
        # z_slice = (lo, hi): this process owns only these latent dimensions (multi-GPU z-sharding); log_prob then
        # returns the densities of the slice alone, the other ranks contribute theirs
        self.z_slice = z_slice
        super().__init__(torch.Size([sigma_f.shape[0]]), torch.Size([patristic.shape[0]]), validate_args=False)

    @property
    def covariance_matrix(self):
        from .. import ops
        return ops.ou_cov(self.patristic, self.sigma_f, self.sigma_n, self.lambd)

    @property
    def mean(self):
        return torch.zeros(self.batch_shape + self.event_shape, dtype=self.sigma_f.dtype, device=self.sigma_f.device)

    def _log_prob_here(self, value, factor=None):
        from .. import ops
        if self.z_slice is not None:
            lo, hi = self.z_slice
            if hi <= lo:
                return value.new_zeros(0)
            return ops.ou_prior_logp(self.patristic, self.sigma_f[lo:hi], self.sigma_n[lo:hi], self.lambd[lo:hi],
                                     value[lo:hi], factor=factor)
        return ops.ou_prior_logp(self.patristic, self.sigma_f, self.sigma_n, self.lambd, value, factor=factor)

    def log_prob(self, value):
        f, self.factor = self.factor, None
        if f is None or f.used:
            return self._log_prob_here(value)
        cur = torch.cuda.current_stream()
        if f.stream == cur:
            return self._log_prob_here(value, f)
        f.stream.wait_stream(cur)                     # the value (and whatever else the caller produced) is ready
        for t in (value, self.sigma_f, self.sigma_n, self.lambd):
            t.record_stream(f.stream)
        with torch.cuda.stream(f.stream):
            lp = self._log_prob_here(value, f)
        cur.wait_stream(f.stream)
        lp.record_stream(cur)
        return lp

    def log_prob_sum_async(self, value, scale=1.0):
        """Summed log-density evaluated on the side stream (`ops.side_stream`): forked from the current stream here,
        joined by whoever needs the number (`AsyncTerm.join`).  None when the side stream is switched off."""
        from .. import ops
        if not ops.PRIOR_STREAM_ENABLED or not value.is_cuda or value.shape[-1] < 400:
            return None                      # small trees: the prior is a few short kernels, forking costs more than it hides
        side = ops.side_stream(value.device)
        side.wait_stream(torch.cuda.current_stream())
        for t in (value, self.sigma_f, self.sigma_n, self.lambd):      # allocated on this stream, read on the side stream
            t.record_stream(side)
        with torch.cuda.stream(side):
            lp = self.log_prob(value).sum()           # a factor started from the guide lives on this same stream
            if scale != 1.0:
                lp = lp * scale
        return AsyncTerm(lp, side)

    def rsample(self, sample_shape=torch.Size()):
        from .. import ops
        L = ops.potrf(self.covariance_matrix.detach(),
                      cond_bound=ops.ou_cond_bound(self.sigma_f, self.sigma_n, self.patristic.shape[0]))
        shape = torch.Size(sample_shape) + self.batch_shape + self.event_shape
        eps = torch.randn(shape, dtype=L.dtype, device=L.device)
        return torch.matmul(L, eps.unsqueeze(-1)).squeeze(-1)          # [S,z,n]; never a [S,z,n,n] intermediate


class TorchDistribution(td.Distribution, TorchDistributionMixin):
    """Base class name reference code subclasses (`pyro.distributions.torch_distribution.TorchDistribution`,
    S/models_utils.py:18,330)."""


Distribution = TorchDistribution      # `dist.Distribution` appears in type annotations (S/models_utils.py:337)

__all__ = ["HalfNormal", "Normal", "Delta", "MultivariateNormal", "Categorical", "Independent", "OUProcessPrior",
           "TorchDistributionMixin", "TorchDistribution", "constraints"]
