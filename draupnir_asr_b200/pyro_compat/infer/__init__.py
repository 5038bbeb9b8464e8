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
        # called after the guide and the model have run and before any site is scored (the observations are first read
        # by the likelihood kernel there): lets a host-fed step keep copying the data set while the networks compute
        self.pre_score_hook = None

    def _traces(self, model, guide, args, kwargs):
        from ... import ops
        with ops.elbo_scope():
            guide_trace = poutine.trace(guide).get_trace(*args, **kwargs)
            model_trace = poutine.trace(poutine.replay(model, trace=guide_trace), async_log_prob=True).get_trace(*args, **kwargs)
        for name, site in guide_trace.sample_sites():
            if not getattr(site["fn"], "has_rsample", False):
                raise NotImplementedError(f"guide site {name!r} is not reparameterised")
        if self.pre_score_hook is not None:
            self.pre_score_hook()
        return model_trace, guide_trace

    def differentiable_loss(self, model, guide, *args, **kwargs):
        model_trace, guide_trace = self._traces(model, guide, args, kwargs)
        self.last_traces = (model_trace, guide_trace)
        return -(model_trace.log_prob_sum() - guide_trace.log_prob_sum())

    def loss(self, model, guide, *args, **kwargs):
        with torch.no_grad():
            return float(self.differentiable_loss(model, guide, *args, **kwargs))

    def loss_and_grads(self, model, guide, *args, **kwargs):
        model_trace, guide_trace = self._traces(model, guide, args, kwargs)
        self.last_traces = (model_trace, guide_trace)
        model_lp, pending = model_trace.split_log_prob()
        loss = -(model_lp - guide_trace.log_prob_sum())
        if not pending:
            if isinstance(loss, torch.Tensor) and loss.requires_grad:
                loss.backward()
            return loss
        # terms still running on a side stream are separate roots of ONE backward call: autograd runs every node on the
        # stream of its forward op, so the recurrences' back-propagation (this stream) does not wait for the OU prior
        terms = [site["log_prob_async"].value for site in pending]
        roots = [loss] + terms
        grads = [torch.ones_like(loss)] + [-torch.ones_like(loss)] * len(terms)
        keep = [(r, g) for r, g in zip(roots, grads) if isinstance(r, torch.Tensor) and r.requires_grad]
        if keep:
            torch.autograd.backward([r for r, _ in keep], [g for _, g in keep])
        for site in pending:
            site["log_prob_sum"] = site["log_prob_async"].join()
            loss = loss - site["log_prob_sum"].detach()
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
