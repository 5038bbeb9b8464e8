"""Effect handlers: trace, replay, block, plate (the subset Trace_ELBO + AutoDelta + the DRAUPNIR model need)."""
from __future__ import annotations

import functools
from collections import OrderedDict

import torch


class Messenger:
    def __init__(self, fn=None):
        self.fn = fn
        if fn is not None:
            functools.update_wrapper(self, fn, updated=[])

    def __enter__(self):
        from . import _PYRO_STACK
        _PYRO_STACK.append(self)
        return self

    def __exit__(self, exc_type, exc, tb):
        from . import _PYRO_STACK
        if _PYRO_STACK and _PYRO_STACK[-1] is self:
            _PYRO_STACK.pop()
        elif self in _PYRO_STACK:            # unwinding after an exception inside nested handlers
            del _PYRO_STACK[_PYRO_STACK.index(self):]
        return False

    def __call__(self, *args, **kwargs):
        with self:
            return self.fn(*args, **kwargs)

    def _process_message(self, msg):
        pass

    def _postprocess_message(self, msg):
        pass


class Trace:
    """Ordered record of the sample and param sites of one execution."""

    def __init__(self):
        self.nodes = OrderedDict()
        self.return_value = None

    def add_node(self, name, site):
        # sample sites are keyed by their name, param sites by "name@param": a PyroParam and a sample site may
        # share a name (S/guides.py:33 vs :108)
        key = name if site["type"] == "sample" else f"{name}@param"
        if key in self.nodes and site["type"] == "sample":
            raise RuntimeError(f"Multiple sample sites named '{name}'")
        self.nodes[key] = site

    def sample_sites(self):
        return [(k, v) for k, v in self.nodes.items() if v["type"] == "sample"]

    def param_sites(self):
        return [(k, v) for k, v in self.nodes.items() if v["type"] == "param"]

    def compute_log_prob(self, join_async: bool = True):
        for _, site in self.sample_sites():
            if "log_prob_sum" not in site and "log_prob_async" in site:
                if join_async:
                    site["log_prob_sum"] = site["log_prob_async"].join()
                continue
            if "log_prob_sum" not in site:
                fused = getattr(site["fn"], "log_prob_sum", None) if site["mask"] is None else None
                if fused is not None:
                    lp = fused(site["value"])
                else:
                    lp = site["fn"].log_prob(site["value"])
                    if site["mask"] is not None:
                        lp = torch.where(site["mask"], lp, torch.zeros_like(lp))
                    lp = lp.sum()
                if site["scale"] != 1.0:
                    lp = lp * site["scale"]
                site["log_prob_sum"] = lp

    def log_prob_sum(self):
        self.compute_log_prob()
        total = 0.0
        for _, site in self.sample_sites():
            total = total + site["log_prob_sum"]
        return total

    def split_log_prob(self):
        """(sum of the site terms evaluated on the current stream, [AsyncTerm, ...] still running on a side stream)."""
        self.compute_log_prob(join_async=False)
        total, pending = 0.0, []
        for _, site in self.sample_sites():
            if "log_prob_sum" in site:
                total = total + site["log_prob_sum"]
            else:
                pending.append(site)
        return total, pending


class trace(Messenger):
    def __init__(self, fn=None, param_only: bool = False, async_log_prob: bool = False):
        super().__init__(fn)
        self.param_only = param_only
        # async_log_prob: sites whose distribution offers `log_prob_sum_async` start scoring on a side stream the
        # moment the site is recorded (the OU prior overlaps with the decoder that way)
        self.async_log_prob = async_log_prob
        self.trace = Trace()

    def __enter__(self):
        self.trace = Trace()
        return super().__enter__()

    def _postprocess_message(self, msg):
        if self.param_only and msg["type"] != "param":
            return
        if msg["type"] in ("sample", "param"):
            site = dict(msg)
            self.trace.add_node(msg["name"], site)
            if self.async_log_prob and msg["type"] == "sample" and msg["mask"] is None:
                start = getattr(msg["fn"], "log_prob_sum_async", None)
                term = start(msg["value"], msg["scale"]) if start is not None else None
                if term is not None:
                    site["log_prob_async"] = term

    def get_trace(self, *args, **kwargs) -> Trace:
        self.trace.return_value = self(*args, **kwargs)       # Pyro records it as the `_RETURN` node
        return self.trace


class replay(Messenger):
    """Latent sample sites take the values recorded in `trace` (the guide's)."""

    def __init__(self, fn=None, trace: Trace = None):
        super().__init__(fn)
        self.guide_trace = trace

    def _process_message(self, msg):
        if msg["type"] == "sample" and not msg["is_observed"] and msg["name"] in self.guide_trace.nodes:
            site = self.guide_trace.nodes[msg["name"]]
            if site["type"] == "sample":
                msg["value"] = site["value"]
                msg["done"] = True
                msg["infer"] = dict(site.get("infer", {}), **msg["infer"])


class block(Messenger):
    def __init__(self, fn=None, hide_fn=None, expose=None, hide=None):
        super().__init__(fn)
        self.expose, self.hide = expose, hide

    def _process_message(self, msg):
        hidden = True
        if self.expose is not None:
            hidden = msg["name"] not in self.expose
        elif self.hide is not None:
            hidden = msg["name"] in self.hide
        if hidden:
            msg["stop"] = True


class scale(Messenger):
    """poutine.scale: multiplies the log-density of the enclosed sample sites by `scale`."""

    def __init__(self, fn=None, scale: float = 1.0):
        super().__init__(fn)
        self.scale = scale

    def _process_message(self, msg):
        if msg["type"] == "sample":
            msg["scale"] = msg["scale"] * self.scale


class CondIndepStackFrame(tuple):
    def __new__(cls, name, dim, size):
        return super().__new__(cls, (name, dim, size))

    name = property(lambda self: self[0])
    dim = property(lambda self: self[1])
    size = property(lambda self: self[2])


class plate(Messenger):
    """pyro.plate.  Size-less plates (S/models.py:308,321, S/guides.py:102) and full-size plates (S/models.py:348)
    carry scale 1 and do not change any shape on this path; subsampling is not part of the accelerated path."""

    def __init__(self, name, size=None, subsample_size=None, subsample=None, dim=None, use_cuda=None, device=None):
        super().__init__(None)
        if subsample is not None or (subsample_size is not None and size is not None and subsample_size < size):
            raise NotImplementedError("plate subsampling is outside the accelerated path (SURVEY.md §2 row 1, plating)")
        self.name, self.size, self.dim, self.device = name, size, dim, device
        self.indices = None if size is None else torch.arange(size, device=device)

    def __enter__(self):
        super().__enter__()
        return self.indices

    def _process_message(self, msg):
        if msg["type"] == "sample":
            msg["cond_indep_stack"] = (CondIndepStackFrame(self.name, self.dim, self.size),) + tuple(msg["cond_indep_stack"])
