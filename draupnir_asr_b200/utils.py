"""Host helpers the hot path shares with the reference's utils (S/utils.py)."""
from __future__ import annotations

import torch


def squeeze_tensor(required_ndims: int, tensor: torch.Tensor) -> torch.Tensor:
    """Drop size-1 dimensions until `required_ndims` remain (S/utils.py:1874-1902); used by gp_prior on the
    `[z,1]`-shaped PyroParams of the variational guide."""
    while tensor.dim() > required_ndims:
        if tensor.shape[0] == 1:
            tensor = tensor.squeeze(0)
        elif 1 in tensor.shape:
            tensor = tensor.squeeze(list(tensor.shape).index(1))
        else:
            break
    return tensor


def require_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"draupnir_asr_b200 runs on CUDA (sm_100a) only; got device {dev!r}. There is no CPU path — "
                           "use the reference implementation for CPU runs.")
    return dev


class Deferred:
    """A value produced on first use (memoised)."""
    __slots__ = ("_fn", "_value", "_done")

    def __init__(self, fn):
        self._fn, self._value, self._done = fn, None, False

    def value(self):
        if not self._done:
            self._value, self._done, self._fn = self._fn(), True, None
        return self._value


class DeferredDict(dict):
    """A `dict` some of whose values are `Deferred`: they are evaluated the first time somebody reads them (`d[k]`,
    `.get`, `.items()`, `.values()`, pickling) and behave like ordinary entries from then on.  The variational guide
    returns its dictionary (S/guides.py:123-133) in this form: `svi.step` discards the guide's return value, so the
    entries the loss does not depend on (the encoder's full hidden-state tensors) cost nothing inside a training step and
    are exact when a caller does look at them.  `raw(k)` hands the entry on without evaluating it."""

    def raw(self, key):
        return dict.__getitem__(self, key)

    def _resolve(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, Deferred):
            v = v.value()
            dict.__setitem__(self, key, v)
        return v

    def materialize(self):
        for k in list(dict.keys(self)):
            self._resolve(k)
        return self

    def __getitem__(self, key):
        return self._resolve(key)

    def __iter__(self):          # an overridden __iter__ makes dict(d), {**d} and other.update(d) go through __getitem__
        return dict.__iter__(self)

    def get(self, key, default=None):
        return self._resolve(key) if key in self else default

    def pop(self, key, *default):
        if key in self:
            self._resolve(key)
        return dict.pop(self, key, *default)

    def items(self):
        return dict.items(self.materialize())

    def values(self):
        return dict.values(self.materialize())

    def copy(self):
        return DeferredDict(dict.copy(self))

    def __eq__(self, other):
        return dict.__eq__(self.materialize(), other)

    __hash__ = None

    def __reduce__(self):                       # pickles (Map_estimates.p, S/main.py:1332) as a plain dict of values
        return (dict, (dict(dict.items(self.materialize())),))
