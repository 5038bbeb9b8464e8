"""`pyro.optim` subset: `ClippedAdam` (S/main.py:1039-1041) and `Adam`.

Pyro's optimisers are called with the set of parameters seen in the step and keep per-parameter state lazily.
ClippedAdam (pyro-ppl 1.8/1.9): `lr <- lr * lrd` at the top of every step, gradients clamped ELEMENT-WISE to
+-clip_norm, optional L2 weight decay added to the gradient, Adam moments, `denom = sqrt(v) + eps`.
The update runs as a handful of multi-tensor (`torch._foreach_*`) launches over all parameters at once.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable

import torch


class ClippedAdam:
    def __init__(self, optim_args: Dict, clip_args=None):
        a = dict(optim_args)
        self.lr = float(a.get("lr", 1e-3))
        self.betas = tuple(a.get("betas", (0.9, 0.999)))
        self.eps = float(a.get("eps", 1e-8))
        self.weight_decay = float(a.get("weight_decay", 0.0))
        self.clip_norm = float(a.get("clip_norm", 10.0))
        self.lrd = float(a.get("lrd", 1.0))
        self.state: Dict[torch.Tensor, dict] = {}
        self.step_count = 0

    def _clip(self, grads):
        torch._foreach_clamp_min_(grads, -self.clip_norm)
        torch._foreach_clamp_max_(grads, self.clip_norm)

    def __call__(self, params: Iterable[torch.Tensor], *args, **kwargs):
        params = [p for p in params if p.grad is not None]
        if not params:
            return
        self.lr *= self.lrd
        b1, b2 = self.betas
        fresh = [p for p in params if p not in self.state]
        for p in fresh:
            self.state[p] = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
        groups: Dict[int, list] = {}
        for p in params:
            st = self.state[p]
            st["step"] += 1
            groups.setdefault(st["step"], []).append(p)
        with torch.no_grad():
            for step, ps in groups.items():
                grads = [p.grad for p in ps]
                self._clip(grads)
                if self.weight_decay != 0:
                    grads = torch._foreach_add(grads, ps, alpha=self.weight_decay)
                m = [self.state[p]["exp_avg"] for p in ps]
                v = [self.state[p]["exp_avg_sq"] for p in ps]
                torch._foreach_mul_(m, b1)
                torch._foreach_add_(m, grads, alpha=1 - b1)
                torch._foreach_mul_(v, b2)
                torch._foreach_addcmul_(v, grads, grads, value=1 - b2)
                denom = torch._foreach_sqrt(v)
                torch._foreach_add_(denom, self.eps)
                step_size = self.lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
                torch._foreach_addcdiv_(ps, m, denom, value=-step_size)

    def get_state(self):
        return {"lr": self.lr, "state": {id(p): s for p, s in self.state.items()}}

    def save(self, filename):
        torch.save({"lr": self.lr, "states": [dict(s) for s in self.state.values()]}, filename)


class Adam(ClippedAdam):
    """torch Adam semantics expressed with the same machinery (no clipping, no decay of lr)."""

    def __init__(self, optim_args: Dict, clip_args=None):
        super().__init__(dict(optim_args, clip_norm=float("inf"), lrd=1.0))

    def _clip(self, grads):
        return None
