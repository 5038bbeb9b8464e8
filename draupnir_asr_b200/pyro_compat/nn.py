"""`pyro.nn.PyroParam` / `PyroModule` subset (S/guides.py:33-36)."""
from __future__ import annotations

import torch
from torch.distributions import constraints, transform_to


class PyroParam:
    def __init__(self, init_value, constraint=constraints.real, event_dim=None):
        self.init_value, self.constraint, self.event_dim = init_value, constraint, event_dim


class PyroModule(torch.nn.Module):
    """nn.Module whose `PyroParam` attributes are stored as `<name>_unconstrained` parameters, returned constrained on
    access, and announced to the active trace as param sites so that SVI optimises them."""

    def __init__(self, name: str = ""):
        super().__init__()
        self.__dict__["_pyro_params"] = {}
        self.__dict__["_pyro_name"] = name

    def __setattr__(self, name, value):
        if isinstance(value, PyroParam):
            init = value.init_value() if callable(value.init_value) else value.init_value
            with torch.no_grad():
                u = transform_to(value.constraint).inv(init.detach()).clone()
            self._pyro_params[name] = value.constraint
            super().__setattr__(f"{name}_unconstrained", torch.nn.Parameter(u))
            return
        super().__setattr__(name, value)

    def __getattr__(self, name):
        params = self.__dict__.get("_pyro_params")
        if params is not None and name in params:
            from . import am_i_wrapped, apply_stack
            u = super().__getattr__(f"{name}_unconstrained")
            c = params[name]
            value = u if c is constraints.real else transform_to(c)(u)
            if am_i_wrapped():
                full = f"{self._pyro_name}.{name}" if self._pyro_name else name
                msg = {"type": "param", "name": full, "fn": None, "is_observed": False, "args": (None, c), "kwargs": {},
                       "value": value, "unconstrained": u, "infer": {}, "scale": 1.0, "mask": None,
                       "cond_indep_stack": (), "done": True, "stop": False}
                apply_stack(msg)
            return value
        return super().__getattr__(name)
