"""`pyro.infer` subset: `Trace_ELBO` and `SVI` (S/main.py:1021,1059; S/train.py:75)."""
from __future__ import annotations

import torch

from .. import poutine


def _unconstrained_params(*traces):
    seen, out = set(), []
    for tr in traces:
        for _, site in tr.param_sites():
            u = site.get("unconstrained", site["value"])
            if isinstance(u, torch.Tensor) and u.requires_grad and id(u) not in seen:
                seen.add(id(u))
                out.append(u)
    return out


class Trace_ELBO:
    """Single-particle ELBO for fully reparameterised guides (every site on the DRAUPNIR path is):
    loss = -( sum_model log p - sum_guide log q ), one backward pass."""

    def __init__(self, num_particles: int = 1, **kwargs):
        if num_particles != 1:
            raise NotImplementedError("num_particles > 1 is not used by the reference path")
        self.num_particles = num_particles

    def _traces(self, model, guide, args, kwargs):
        guide_trace = poutine.trace(guide).get_trace(*args, **kwargs)
        model_trace = poutine.trace(poutine.replay(model, trace=guide_trace)).get_trace(*args, **kwargs)
        for name, site in guide_trace.sample_sites():
            if not getattr(site["fn"], "has_rsample", False):
                raise NotImplementedError(f"guide site {name!r} is not reparameterised")
        return model_trace, guide_trace

    def differentiable_loss(self, model, guide, *args, **kwargs):
        model_trace, guide_trace = self._traces(model, guide, args, kwargs)
        self.last_traces = (model_trace, guide_trace)
        return -(model_trace.log_prob_sum() - guide_trace.log_prob_sum())

    def loss(self, model, guide, *args, **kwargs):
        with torch.no_grad():
            return float(self.differentiable_loss(model, guide, *args, **kwargs))

    def loss_and_grads(self, model, guide, *args, **kwargs):
        loss = self.differentiable_loss(model, guide, *args, **kwargs)
        if isinstance(loss, torch.Tensor) and loss.requires_grad:
            loss.backward()
        return loss

    def __repr__(self):
        return "Trace_ELBO()"


class SVI:
    def __init__(self, model, guide, optim, loss, shard=None, **kwargs):
        self.model, self.guide, self.optim, self.loss = model, guide, optim, loss
        self.shard = shard     # zshard.ShardContext: local losses/gradients are summed over the ranks before the update

    def step(self, *args, **kwargs) -> float:
        """One gradient step; returns the loss as a Python float (the only host<->device synchronisation)."""
        from ... import ops
        with ops.defer_info_checks():
            loss = self.loss.loss_and_grads(self.model, self.guide, *args, **kwargs)
            params = _unconstrained_params(*self.loss.last_traces)
            if self.shard is not None:
                loss = self.shard.reduce_step(params, loss)
            self.optim(params)
            for p in params:
                p.grad = None
            value = float(loss.detach())   # synchronises; the deferred pivot checks piggy-back on it
        return value

    def evaluate_loss(self, *args, **kwargs) -> float:
        return self.loss.loss(self.model, self.guide, *args, **kwargs)


from . import autoguide  # noqa: E402,F401
