"""Minimal Pyro-compatible runtime for the DRAUPNIR model/guide path (pyro-ppl is not installed in this image).

Only what S/models.py, S/guides.py, S/train.py and S/main.py:1021-1059 use is provided (S/ =
/root/reference/draupnir/src/draupnir/): `sample`, `plate`, `module`, `param`, the param store, `poutine.trace/replay`,
`distributions`, `infer.SVI/Trace_ELBO`, `infer.autoguide.AutoDelta / AutoNormal / AutoDiagonalNormal`, `optim.ClippedAdam`, `nn.PyroParam/PyroModule`,
`contrib.easyguide.EasyGuide`.  Semantics follow pyro-ppl 1.8/1.9 for these features: effect handlers
as a messenger stack, size-less plates are no-ops, Delta guide sites score 0, positive parameters live in log space.

`install_as_pyro()` registers this package as `pyro` in `sys.modules` when the real one is not importable, so
reference-style code (`import pyro; import pyro.distributions as dist`) runs unchanged.
"""
from __future__ import annotations

import sys
from collections import OrderedDict
from typing import Optional

import torch
from torch.distributions import constraints, transform_to

_PYRO_STACK: list = []


# ------------------------------------------------------------------------------------------------------------------
# parameter store
# ------------------------------------------------------------------------------------------------------------------
class ParamStoreDict:
    """name -> unconstrained leaf tensor (+ constraint).  nn.Module parameters registered through `pyro.module` are
    stored as themselves (real constraint)."""

    def __init__(self):
        self._params = OrderedDict()
        self._constraints = {}

    def clear(self):
        self._params.clear()
        self._constraints.clear()

    def __contains__(self, name):
        return name in self._params

    def __len__(self):
        return len(self._params)

    def keys(self):
        return self._params.keys()

    def setdefault_unconstrained(self, name, unconstrained, constraint=constraints.real):
        if name not in self._params:
            self._params[name] = unconstrained
            self._constraints[name] = constraint
        return self._params[name]

    def replace_unconstrained(self, name, unconstrained, constraint=constraints.real):
        self._params[name] = unconstrained
        self._constraints[name] = constraint

    def get_unconstrained(self, name):
        return self._params[name]

    def get_param(self, name, init_tensor=None, constraint=constraints.real):
        if name not in self._params:
            if init_tensor is None:
                raise KeyError(f"param {name!r} has no initial value")
            if callable(init_tensor):
                init_tensor = init_tensor()
            with torch.no_grad():
                u = transform_to(constraint).inv(init_tensor.detach()).clone()
            u.requires_grad_(True)
            self._params[name] = u
            self._constraints[name] = constraint
        return self[name]

    def __getitem__(self, name):
        u = self._params[name]
        c = self._constraints[name]
        return u if c is constraints.real else transform_to(c)(u)

    def named_parameters(self):
        return self._params.items()

    def get_state(self):
        return {"params": {k: v.detach().clone() for k, v in self._params.items()},
                "constraints": dict(self._constraints)}

    def set_state(self, state):
        for k, v in state["params"].items():
            if k in self._params:
                with torch.no_grad():
                    self._params[k].copy_(v)
            else:
                self._params[k] = v.clone().requires_grad_(True)
                self._constraints[k] = state["constraints"].get(k, constraints.real)


_PARAM_STORE = ParamStoreDict()


def get_param_store() -> ParamStoreDict:
    return _PARAM_STORE


def clear_param_store():
    _PARAM_STORE.clear()


def enable_validation(is_validate: bool = True):
    """Draupnir_example.py:159 switches validation off; here that also stops torch.distributions from validating
    arguments (each validation is a device synchronisation on the GPU)."""
    torch.distributions.Distribution.set_default_validate_args(bool(is_validate))


def set_rng_seed(seed: int):
    torch.manual_seed(seed)


# ------------------------------------------------------------------------------------------------------------------
# messenger machinery
# ------------------------------------------------------------------------------------------------------------------
def am_i_wrapped() -> bool:
    return len(_PYRO_STACK) > 0


def _default_process_message(msg):
    if msg["done"] or msg["value"] is not None:
        msg["done"] = True
        return
    if msg["type"] == "sample":
        fn = msg["fn"]
        msg["value"] = fn.rsample() if getattr(fn, "has_rsample", False) else fn.sample()
    elif msg["type"] == "param":
        name, (init, constraint) = msg["name"], msg["args"]
        msg["value"] = _PARAM_STORE.get_param(name, init, constraint)
        msg["unconstrained"] = _PARAM_STORE.get_unconstrained(name)
    msg["done"] = True


def apply_stack(msg):
    pointer = 0
    for pointer, frame in enumerate(reversed(_PYRO_STACK)):
        frame._process_message(msg)
        if msg.get("stop"):
            break
    _default_process_message(msg)
    for frame in _PYRO_STACK[len(_PYRO_STACK) - pointer - 1:]:
        frame._postprocess_message(msg)
    return msg


def sample(name: str, fn, *args, obs=None, infer: Optional[dict] = None, **kwargs):
    """pyro.sample: draws (or scores `obs`) at a named site under the active handlers."""
    if not am_i_wrapped():
        if obs is not None:
            return obs
        return fn.rsample() if getattr(fn, "has_rsample", False) else fn.sample()
    msg = {"type": "sample", "name": name, "fn": fn, "is_observed": obs is not None, "args": args, "kwargs": kwargs,
           "value": obs, "infer": infer or {}, "scale": 1.0, "mask": None, "cond_indep_stack": (), "done": False,
           "stop": False}
    return apply_stack(msg)["value"]


def param(name: str, init_tensor=None, constraint=constraints.real, event_dim=None):
    if not am_i_wrapped():
        return _PARAM_STORE.get_param(name, init_tensor, constraint)
    msg = {"type": "param", "name": name, "fn": None, "is_observed": False, "args": (init_tensor, constraint),
           "kwargs": {}, "value": None, "infer": {}, "scale": 1.0, "mask": None, "cond_indep_stack": (), "done": False,
           "stop": False}
    return apply_stack(msg)["value"]


def module(name: str, nn_module: torch.nn.Module, update_module_params: bool = False):
    """pyro.module: registers every parameter as `name$$$<path>` so SVI optimises it (S/models.py:333-334)."""
    for pname, p in nn_module.named_parameters():
        full = f"{name}$$${pname}"
        if full not in _PARAM_STORE or _PARAM_STORE.get_unconstrained(full) is not p:
            _PARAM_STORE.replace_unconstrained(full, p)
        if am_i_wrapped():
            msg = {"type": "param", "name": full, "fn": None, "is_observed": False, "args": (None, constraints.real),
                   "kwargs": {}, "value": p, "unconstrained": p, "infer": {}, "scale": 1.0, "mask": None,
                   "cond_indep_stack": (), "done": True, "stop": False}
            apply_stack(msg)
    return nn_module


from . import poutine  # noqa: E402
from .poutine import plate  # noqa: E402
from . import distributions  # noqa: E402
from . import optim  # noqa: E402
from . import nn  # noqa: E402
from . import infer  # noqa: E402
from . import contrib  # noqa: E402

__version__ = "0-compat"


def install_as_pyro(force: bool = False) -> bool:
    """Expose this package as `pyro` (and its submodules) unless the real pyro-ppl can be imported."""
    if not force:
        try:
            import importlib
            if "pyro" in sys.modules and sys.modules["pyro"] is not sys.modules[__name__]:
                return False
            importlib.import_module("pyro")
            return False
        except ImportError:
            pass
    me = sys.modules[__name__]
    sys.modules["pyro"] = me
    for sub in ("poutine", "distributions", "optim", "nn", "infer", "infer.autoguide", "contrib", "contrib.easyguide"):
        sys.modules[f"pyro.{sub}"] = sys.modules[f"{__name__}.{sub}"]
    import types
    td_mod = types.ModuleType("pyro.distributions.torch_distribution")      # S/models_utils.py:18 imports it by path
    td_mod.TorchDistribution = distributions.TorchDistribution
    td_mod.TorchDistributionMixin = distributions.TorchDistributionMixin
    sys.modules["pyro.distributions.torch_distribution"] = td_mod
    return True
