"""`pyro.optim` subset: `ClippedAdam` (S/main.py:1039-1041) and `Adam`.

Pyro's optimisers are called with the set of parameters seen in the step and keep per-parameter state lazily.
ClippedAdam (pyro-ppl 1.8/1.9): `lr <- lr * lrd` at the top of every step, gradients clamped ELEMENT-WISE to
+-clip_norm, optional L2 weight decay added to the gradient, Adam moments, `denom = sqrt(v) + eps`.
CUDA fp64 parameters (everything on the DRAUPNIR path) are updated by `drp_clipped_adam_f64` (csrc/adam.cu): two launches
for all parameters together, step counters and learning rate on the device, hence capturable in a CUDA graph with the rest
of the step.  Other tensors (the CPU tests of this runtime) take a handful of multi-tensor (`torch._foreach_*`) launches.
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
        self._dev_state = {}          # device -> (state tensor [1 + 2 * slots]: lr | (step, step size) per slot, next free slot)

    _SLOTS = 1024

    def _device_update(self, params):
        """All CUDA fp64 parameters in one call of the fused kernel (per 72 tensors)."""
        import ctypes
        from .. import ops
        lib = ops.lib
        dev = params[0].device
        ds = self._dev_state.get(dev)
        if ds is None:
            st = torch.zeros(1 + 2 * self._SLOTS, dtype=torch.float64, device=dev)
            st[0] = self.lr
            ds = self._dev_state[dev] = [st, 0]
        st = ds[0]
        for p in params:
            s = self.state[p]
            if "slot" not in s:
                if ds[1] >= self._SLOTS:
                    raise RuntimeError("ClippedAdam: more than 1024 parameter tensors")
                s["slot"], ds[1] = ds[1], ds[1] + 1
                if s["step"]:                       # state restored from a checkpoint: the device counter continues there
                    st[1 + 2 * s["slot"]] = float(s["step"])
        cap = lib.drp_clipped_adam_max_tensors()
        b1, b2 = self.betas
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for lo in range(0, len(params), cap):
            chunk = params[lo:lo + cap]
            n = len(chunk)
            ptrs = (ctypes.c_void_p * (4 * n))()
            numel = (ctypes.c_int64 * n)()
            slot = (ctypes.c_int64 * n)()
            for i, p in enumerate(chunk):
                s, g = self.state[p], p.grad
                if not (p.is_contiguous() and g.is_contiguous()):
                    raise RuntimeError("ClippedAdam: parameters and gradients must be contiguous")
                ptrs[4 * i], ptrs[4 * i + 1] = p.data_ptr(), g.data_ptr()
                ptrs[4 * i + 2], ptrs[4 * i + 3] = s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr()
                numel[i], slot[i] = p.numel(), s["slot"]
            rc = lib.drp_clipped_adam_f64(ptrs, numel, slot, n, int(lo == 0), ctypes.c_void_p(st.data_ptr()), self.lrd, b1, b2,
                                          self.eps, self.weight_decay, self.clip_norm, stream)
            if rc != 0:
                raise RuntimeError(f"drp_clipped_adam_f64 failed: {rc}")

    def _clip(self, grads):
        torch._foreach_clamp_min_(grads, -self.clip_norm)
        torch._foreach_clamp_max_(grads, self.clip_norm)

    def __call__(self, params: Iterable[torch.Tensor], *args, **kwargs):
        params = [p for p in params if p.grad is not None]
        if not params:
            return
        fresh = [p for p in params if p not in self.state]
        for p in fresh:
            self.state[p] = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
        fused = [p for p in params if p.is_cuda and p.dtype == torch.float64]
        params = [p for p in params if not (p.is_cuda and p.dtype == torch.float64)]
        if fused:
            if params:
                raise RuntimeError("ClippedAdam: CUDA fp64 and other parameters in one step")
            self._device_update(fused)
            self.lr *= self.lrd                     # host mirrors of the device-side values (checkpoints, get_state)
            for p in fused:
                self.state[p]["step"] += 1
            self.step_count += 1
            return
        self.lr *= self.lrd
        b1, b2 = self.betas
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
        torch.save({"lr": self.lr, "states": [{k: v for k, v in s.items() if k != "slot"} for s in self.state.values()]}, filename)


class Adam(ClippedAdam):
    """torch Adam semantics expressed with the same machinery (no clipping, no decay of lr)."""

    def __init__(self, optim_args: Dict, clip_args=None):
        super().__init__(dict(optim_args, clip_norm=float("inf"), lrd=1.0))

    def _clip(self, grads):
        return None
