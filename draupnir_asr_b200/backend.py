"""Selects the probabilistic-programming runtime: the real `pyro` when it is importable, otherwise the in-tree
compatible runtime (`pyro_compat`).  The accelerated distributions are attached to whichever `dist` is active."""
from __future__ import annotations

try:  # pragma: no cover - pyro-ppl is not present in the build image
    import pyro  # type: ignore
    if getattr(pyro, "__version__", "") == "0-compat":
        raise ImportError
    import pyro.distributions as dist  # type: ignore
    from . import pyro_compat as _compat
    USING_REAL_PYRO = True

    class _OU(_compat.distributions.OUProcessPrior, dist.torch_distribution.TorchDistributionMixin):
        pass

    class _Dist:
        """real pyro.distributions + the fused GPU distributions"""
        OUProcessPrior = _OU

        def __getattr__(self, name):
            if name == "Categorical":
                return _compat.distributions.Categorical
            if name == "MultivariateNormal":
                return _compat.distributions.MultivariateNormal
            return getattr(dist, name)

    dist = _Dist()
except ImportError:
    from . import pyro_compat as pyro
    from .pyro_compat import distributions as dist
    USING_REAL_PYRO = False

__all__ = ["pyro", "dist", "USING_REAL_PYRO"]
