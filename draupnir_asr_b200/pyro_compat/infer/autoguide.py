"""`pyro.infer.autoguide.AutoDelta` (S/main.py:516) with `init_to_median` / `init_to_value`."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch.distributions import biject_to, constraints

from .. import poutine
from ..distributions import Delta


def init_to_median(site=None, num_samples: int = 15):
    """Median of `num_samples` prior draws (Pyro's default AutoDelta initialiser; RNG dependent)."""
    if site is None:
        return lambda s: init_to_median(s, num_samples)
    with torch.no_grad():
        samples = site["fn"].sample(sample_shape=(num_samples,))
        return samples.median(dim=0).values


def init_to_value(site=None, values: Optional[Dict[str, torch.Tensor]] = None, fallback=init_to_median):
    if site is None:
        return lambda s: init_to_value(s, values, fallback)
    if values and site["name"] in values:
        return values[site["name"]].detach().clone()
    return fallback(site)


class AutoDelta(torch.nn.Module):
    """MAP guide: one learnable point per latent site, stored unconstrained (`<site>_unconstrained`, the same
    attribute names PyroModule uses, so `Guide_state_dict.p` keys match), scored as `Delta` (log-density 0)."""

    def __init__(self, model, init_loc_fn=init_to_median, create_plates=None):
        super().__init__()
        self.model = model
        self.init_loc_fn = init_loc_fn
        self.prototype_trace = None
        self._sites = []

    def _setup_prototype(self, *args, **kwargs):
        with poutine.block():
            self.prototype_trace = poutine.trace(self.model).get_trace(*args, **kwargs)
        for name, site in self.prototype_trace.sample_sites():
            if site["is_observed"]:
                continue
            value = self.init_loc_fn(site).detach()
            support = site["fn"].support
            transform = biject_to(support)
            u = transform.inv(value).clone()
            self.register_parameter(f"{name}_unconstrained", torch.nn.Parameter(u))
            self._sites.append((name, transform, len(site["fn"].event_shape),
                                support is constraints.real or support is constraints.real_vector))

    def forward(self, *args, **kwargs):
        from .. import apply_stack, am_i_wrapped
        if self.prototype_trace is None:
            self._setup_prototype(*args, **kwargs)
        result = {}
        import draupnir_asr_b200.pyro_compat as pyro
        for name, transform, event_dim, identity in self._sites:
            u = getattr(self, f"{name}_unconstrained")
            value = u if identity else transform(u)
            if am_i_wrapped():
                msg = {"type": "param", "name": f"AutoDelta.{name}", "fn": None, "is_observed": False,
                       "args": (None, constraints.real), "kwargs": {}, "value": value, "unconstrained": u, "infer": {},
                       "scale": 1.0, "mask": None, "cond_indep_stack": (), "done": True, "stop": False}
                apply_stack(msg)
            result[name] = pyro.sample(name, Delta(value, event_dim=event_dim))
        return result

    def median(self, *args, **kwargs):
        return {k: v.detach() for k, v in self.forward(*args, **kwargs).items()}
