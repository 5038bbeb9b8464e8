"""`pyro.infer.autoguide`: `AutoDelta`, `AutoNormal`, `AutoDiagonalNormal` (S/main.py:516-518) with `init_to_median` /
`init_to_value` / `init_to_feasible`."""
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
        # no_grad: the prototype only records names, shapes and supports.  A trace with an autograd graph would keep the
        # networks' AccumulateGrad nodes (bound to the stream of this very first call) alive for the life of the guide,
        # which breaks a later CUDA-graph capture of the step (pyro_compat/infer: Trace_ELBO.loss_and_grads).
        with torch.no_grad(), poutine.block():
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


# ----------------------------------------------------------------------------------------------------------------------
# Mean-field normal guides (S/main.py:516-518: "normal" -> AutoNormal, "diagonal_normal" -> AutoDiagonalNormal)
# ----------------------------------------------------------------------------------------------------------------------
def init_to_feasible(site=None):
    """Pyro's default for AutoNormal: the image of 0 under the site's bijection to its support."""
    if site is None:
        return init_to_feasible
    with torch.no_grad():
        proto = site["value"] if site.get("value") is not None else site["fn"].sample()
        transform = biject_to(site["fn"].support)
        return transform(torch.zeros_like(transform.inv(proto)))


def _softplus_inv(x: torch.Tensor) -> torch.Tensor:
    return x + torch.log(-torch.expm1(-x))


class _AutoMeanField(torch.nn.Module):
    """Shared machinery: every latent site gets a Normal(loc, softplus(scale_unconstrained)) in UNCONSTRAINED space (the
    auxiliary site), its value is the bijection's image and the site itself is a `Delta` carrying the log-Jacobian, as
    pyro-ppl >= 1.6 does (`scale` under the `softplus_positive` constraint, `init_scale = 0.1`)."""

    def __init__(self, model, init_loc_fn, init_scale: float = 0.1, create_plates=None):
        super().__init__()
        if not (isinstance(init_scale, float) and init_scale > 0):
            raise ValueError(f"Expected init_scale > 0. but got {init_scale}")
        self.model = model
        self.init_loc_fn = init_loc_fn
        self._init_scale = init_scale
        self.prototype_trace = None
        self._sites = []          # (name, transform, site event_dim, unconstrained shape)

    def _latent_sites(self, *args, **kwargs):
        with torch.no_grad(), poutine.block():          # see AutoDelta._setup_prototype
            self.prototype_trace = poutine.trace(self.model).get_trace(*args, **kwargs)
        out = []
        for name, site in self.prototype_trace.sample_sites():
            if site["is_observed"]:
                continue
            transform = biject_to(site["fn"].support)
            u = transform.inv(self.init_loc_fn(site).detach()).clone()
            out.append((name, transform, len(site["fn"].event_shape), u))
        return out

    @staticmethod
    def _emit_param(name, value, unconstrained):
        from .. import apply_stack, am_i_wrapped
        if am_i_wrapped():
            apply_stack({"type": "param", "name": name, "fn": None, "is_observed": False, "args": (None, constraints.real),
                         "kwargs": {}, "value": value, "unconstrained": unconstrained, "infer": {}, "scale": 1.0,
                         "mask": None, "cond_indep_stack": (), "done": True, "stop": False})

    @staticmethod
    def _site_from_unconstrained(name, transform, event_dim, u):
        """`pyro.sample(name, Delta(transform(u), log_density = -log|d value / d u|, event_dim))`."""
        import draupnir_asr_b200.pyro_compat as pyro
        from torch.distributions.utils import _sum_rightmost
        value = transform(u)
        log_density = transform.inv.log_abs_det_jacobian(value, u)
        log_density = _sum_rightmost(log_density, log_density.dim() - value.dim() + event_dim)
        return pyro.sample(name, Delta(value, log_density=log_density, event_dim=event_dim))

    def median(self, *args, **kwargs):
        if self.prototype_trace is None:
            self._setup_prototype(*args, **kwargs)
        return {name: transform(loc).detach() for name, transform, loc in self._median_items()}


class AutoNormal(_AutoMeanField):
    """`pyro.infer.autoguide.AutoNormal`: parameters `locs.<site>` and `scales.<site>` per latent site."""

    def __init__(self, model, init_loc_fn=init_to_feasible, init_scale: float = 0.1, create_plates=None):
        super().__init__(model, init_loc_fn, init_scale, create_plates)
        self.locs = torch.nn.ParameterDict()
        self.scales_unconstrained = torch.nn.ParameterDict()

    def _setup_prototype(self, *args, **kwargs):
        for name, transform, event_dim, u in self._latent_sites(*args, **kwargs):
            self.locs[name] = torch.nn.Parameter(u)
            self.scales_unconstrained[name] = torch.nn.Parameter(_softplus_inv(torch.full_like(u, self._init_scale)))
            self._sites.append((name, transform, event_dim, u.shape))

    def _median_items(self):
        return [(name, transform, self.locs[name]) for name, transform, _, _ in self._sites]

    def forward(self, *args, **kwargs):
        import draupnir_asr_b200.pyro_compat as pyro
        from ..distributions import Normal
        if self.prototype_trace is None:
            self._setup_prototype(*args, **kwargs)
        result = {}
        for name, transform, event_dim, _ in self._sites:
            loc, su = self.locs[name], self.scales_unconstrained[name]
            scale = torch.nn.functional.softplus(su)
            self._emit_param(f"AutoNormal.locs.{name}", loc, loc)
            self._emit_param(f"AutoNormal.scales.{name}", scale, su)
            u = pyro.sample(name + "_unconstrained", Normal(loc, scale).to_event(loc.dim()), infer={"is_auxiliary": True})
            result[name] = self._site_from_unconstrained(name, transform, event_dim, u)
        return result


class AutoDiagonalNormal(_AutoMeanField):
    """`pyro.infer.autoguide.AutoDiagonalNormal`: ONE auxiliary site `_AutoDiagonalNormal_latent` ~ Normal(loc, scale) over
    the concatenation of all unconstrained latents (parameters `loc`, `scale`), unpacked site by site."""

    def __init__(self, model, init_loc_fn=init_to_median, init_scale: float = 0.1, create_plates=None):
        super().__init__(model, init_loc_fn, init_scale, create_plates)

    def _setup_prototype(self, *args, **kwargs):
        chunks = []
        for name, transform, event_dim, u in self._latent_sites(*args, **kwargs):
            self._sites.append((name, transform, event_dim, u.shape))
            chunks.append(u.reshape(-1))
        loc = torch.cat(chunks)
        self.loc = torch.nn.Parameter(loc)
        self.scale_unconstrained = torch.nn.Parameter(_softplus_inv(torch.full_like(loc, self._init_scale)))
        self.latent_dim = loc.numel()

    def _unpack(self, latent):
        pos = 0
        for name, transform, event_dim, shape in self._sites:
            size = int(torch.Size(shape).numel())
            yield name, transform, event_dim, latent[..., pos:pos + size].reshape(latent.shape[:-1] + tuple(shape))
            pos += size

    def _median_items(self):
        return [(name, transform, u) for name, transform, _, u in self._unpack(self.loc)]

    def get_posterior(self):
        from ..distributions import Normal
        return Normal(self.loc, torch.nn.functional.softplus(self.scale_unconstrained)).to_event(1)

    def forward(self, *args, **kwargs):
        import draupnir_asr_b200.pyro_compat as pyro
        if self.prototype_trace is None:
            self._setup_prototype(*args, **kwargs)
        self._emit_param("AutoDiagonalNormal.loc", self.loc, self.loc)
        self._emit_param("AutoDiagonalNormal.scale", torch.nn.functional.softplus(self.scale_unconstrained), self.scale_unconstrained)
        latent = pyro.sample("_AutoDiagonalNormal_latent", self.get_posterior(), infer={"is_auxiliary": True})
        return {name: self._site_from_unconstrained(name, transform, event_dim, u)
                for name, transform, event_dim, u in self._unpack(latent)}
