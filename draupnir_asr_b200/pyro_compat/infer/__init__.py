"""`pyro.infer` subset: `Trace_ELBO` and `SVI` (S/main.py:1021,1059; S/train.py:75)."""
from __future__ import annotations

import os

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
        """Loss and gradients of one particle.  The gradients are taken with `torch.autograd.grad` w.r.t. the parameters
        the two traces saw and then stored in `.grad` — the result `loss.backward()` leaves behind, but without going
        through the parameters' AccumulateGrad nodes: those are bound to the stream they were created on and outlive a
        step whenever anything (a loss the caller kept, an earlier trace) still references an old autograd graph, which a
        CUDA-graph capture of the step cannot tolerate."""
        model_trace, guide_trace = self._traces(model, guide, args, kwargs)
        self.last_traces = (model_trace, guide_trace)
        model_lp, pending = model_trace.split_log_prob()
        loss = -(model_lp - guide_trace.log_prob_sum())
        params = _unconstrained_params(model_trace, guide_trace)
        # terms still running on a side stream are separate roots of ONE backward pass: autograd runs every node on the
        # stream of its forward op, so the recurrences' back-propagation (this stream) does not wait for the OU prior
        terms = [site["log_prob_async"].value for site in pending]
        roots = [loss] + terms
        seeds = [torch.ones_like(loss)] + [-torch.ones_like(t) for t in terms] if isinstance(loss, torch.Tensor) else []
        keep = [(r, g) for r, g in zip(roots, seeds) if isinstance(r, torch.Tensor) and r.requires_grad]
        if keep and params:
            grads = torch.autograd.grad([r for r, _ in keep], params, [g for _, g in keep], allow_unused=True)
            for p, g in zip(params, grads):
                if g is not None:
                    if not g.is_contiguous():      # e.g. the gradient of a transposed view: `.grad` has the parameter's layout
                        g = g.contiguous()
                    p.grad = g if p.grad is None else p.grad + g
        for site in pending:
            site["log_prob_sum"] = site["log_prob_async"].join()
            loss = loss - site["log_prob_sum"].detach()
        return loss

    def __repr__(self):
        return "Trace_ELBO()"


def _recover_generator(device, rng_state) -> bool:
    """A capture that dies half-way leaves the device's default generator state in capture mode, with a non-zero
    intra-graph offset (every later random draw, and every replay of a graph captured EARLIER with that state, would
    raise).  A trivial capture that succeeds runs the generator's capture prologue / epilogue and so puts the very same state
    object back in order; the seed / offset the failed capture started from are then restored.  Returns True if that worked
    (earlier graphs remain usable).  Otherwise a fresh state object is swapped in — eager draws work again, but graphs that
    registered the old one must be dropped — and the function returns False."""
    try:
        torch.cuda.synchronize(device)
    except Exception:       # noqa: BLE001 - the error of the failed capture may surface here once more
        pass
    try:
        dummy = torch.cuda.CUDAGraph()
        with torch.cuda.graph(dummy):
            torch.zeros(1, device=device)
        del dummy
        torch.cuda.set_rng_state(rng_state, device)
        torch.cuda.synchronize(device)
        return True
    except Exception:       # noqa: BLE001
        pass
    try:
        fresh = torch.Generator(device=device)
        fresh.set_state(rng_state)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        torch.cuda.default_generators[idx].graphsafe_set_state(fresh)
    except Exception:       # noqa: BLE001 - best effort
        pass
    return False


class _CapturedStep:
    """One `svi.step` recorded as a CUDA graph for one set of argument tensors."""
    __slots__ = ("graph", "out_host", "infos", "params", "grads", "traces", "stream", "static_args")


class SVI:
    """`pyro.infer.SVI` (S/main.py:1059) — `step(*args)` is one ELBO evaluation, its backward pass and one optimiser update
    (S/train.py:75), returning the loss as a Python float.

    CUDA-graph replay.  With static shapes the whole of guide -> model -> backward (and, sharded, the
    NCCL all-reduce) is a fixed sequence of a few hundred kernels on two or three streams, re-derived by the Python trace
    machinery every step: ~6.5 ms of host work, which IS the step below ~800 leaves and on 8 GPUs.  After `graph_warmup`
    eager steps with the same argument tensors (same storage, shape, dtype) the step is therefore captured once
    (`torch.cuda.graph`, the side streams forked and joined inside the capture) and replayed afterwards: one graph launch,
    then the loss and the pivot-failure flag arrive in pinned host memory (a memcpy node of the graph), the host checks the
    flag — a failed Cholesky raises `torch.linalg.LinAlgError` BEFORE any parameter is touched, as in the reference — and
    the optimiser (two launches, csrc/adam.cu, counters on the device) runs on the captured gradient buffers.  The values
    a caller reads afterwards (`loss.last_traces`, the guide's returned dictionary) are the graph's static tensors and hold
    the latest step's numbers.  `cuda_graph=False` / DRP_CUDA_GRAPH=0 keeps every step eager; a capture that fails (an
    operation that cannot be captured somewhere in user code) falls back to eager steps for good, with a warning.
    """

    def __init__(self, model, guide, optim, loss, shard=None, cuda_graph=None, graph_warmup: int = 2, **kwargs):
        self.model, self.guide, self.optim, self.loss = model, guide, optim, loss
        self.shard = shard     # zshard.ShardContext: local losses/gradients are summed over the ranks before the update
        self.cuda_graph = (os.environ.get("DRP_CUDA_GRAPH", "1") != "0") if cuda_graph is None else bool(cuda_graph)
        self.graph_warmup = int(graph_warmup)
        self.step_prologue = None          # callable run at the start of every step (host-fed steps enqueue their copies there)
        self._graphs = {}                  # key -> _CapturedStep | number of eager steps so far | False (not capturable)
        self._capture_failures = 0
        self._capture_stream = None
        self.last_step_was_replay = False

    # ---------------------------------------------------------------------------------------------------------------
    def reset_graphs(self):
        """Forget every captured step (after swapping tensors the model reads that are not arguments of `step`)."""
        self._graphs.clear()

    def _graph_key(self, args):
        """(exact key, signature key, device) of a call, or Nones if it cannot be captured.  The exact key names the argument
        STORAGE (a caller that passes the same tensors every step - `Run.step`, the host-fed staging buffers - replays with no
        copy at all); the signature key names only shapes / strides / dtypes: a caller that brings a fresh tensor every epoch -
        the reference's loop takes `dataset` from a DataLoader (S/train.py:73-75) - gets a graph captured on copies the SVI
        object owns, into which each step's arguments are copied (device to device) before the replay."""
        exact, sig, device = [], [], None
        for a in args:
            if a is None:
                exact.append(None)
                sig.append(None)
            elif isinstance(a, torch.Tensor) and a.is_cuda:
                sig.append((tuple(a.shape), tuple(a.stride()), a.dtype, a.device.index))
                exact.append((a.data_ptr(),) + sig[-1])
                device = a.device
            else:
                return None, None, None
        if device is None or not torch.is_grad_enabled():
            return None, None, None
        if self.shard is not None:
            import torch.distributed as dist
            if dist.get_backend(self.shard.group) != "nccl":      # gloo moves CUDA tensors through the host: not capturable
                return None, None, None
        pro = id(self.step_prologue)
        return (tuple(exact), pro), (tuple(sig), pro, "signature"), device

    def _evaluate(self, args, kwargs):
        """Prologue, ELBO and gradients, cross-rank reduction: everything of a step up to (not including) the optimiser.
        Returns (loss, failure flag or None, parameters)."""
        from ... import ops
        if self.step_prologue is not None:
            self.step_prologue()
        loss = self.loss.loss_and_grads(self.model, self.guide, *args, **kwargs)
        params = _unconstrained_params(*self.loss.last_traces)
        flag = ops.pending_failures()
        if self.shard is not None:
            loss, flag = self.shard.reduce_step(params, loss, flag)
        return loss, flag, params

    def _eager_step(self, args, kwargs) -> float:
        from ... import ops
        with ops.defer_info_checks():
            loss, flag, params = self._evaluate(args, kwargs)
            value = float(loss.detach())   # synchronises; the deferred pivot checks piggy-back on it
            ops.flush_info_checks()        # a failed factorisation raises here, before any parameter is updated
            if self.shard is not None and flag is not None and float(flag) != 0.0:
                raise torch.linalg.LinAlgError("linalg.cholesky: the factorisation failed on another rank of the sharded run")
        self.optim(params)
        for p in params:
            p.grad = None
        self.last_step_was_replay = False
        return value

    def _capture(self, args, device) -> "_CapturedStep":
        from ... import ops
        cs = _CapturedStep()
        cs.graph = torch.cuda.CUDAGraph()
        cs.out_host = torch.zeros(2, dtype=torch.float64).pin_memory()
        self.loss.last_traces = None       # the previous step's autograd graph is not needed any more
        rng_state = torch.cuda.get_rng_state(device)
        token = ops.begin_deferred()
        # The capture stream has HIGH priority and the kernel nodes inherit it: the step's main chain (the recurrences, one CTA
        # per SM for milliseconds) gets SMs ahead of the prior's chain on the (default-priority) side stream whenever both have
        # a kernel pending.  The other order starves the recurrences behind the persistent tcgen05 update kernels: 26-33 ms per
        # 2,000-leaf step instead of 22.4 (profiles/r02/bench_graph_v1_c4.json).
        # ONE capture stream per SVI object: the parameters' AccumulateGrad nodes are bound to the stream they were created on
        # and live as long as any captured step's traces do, so a second capture (other argument tensors) on another stream
        # would hand the engine a stream mismatch - a cross-stream wait onto a stream outside the capture, which breaks it
        # now and then (profiles/r02/gpu_session24.sh: the second of two alternating argument sets)
        if self._capture_stream is None or self._capture_stream.device != device:
            prio = int(os.environ.get("DRP_GRAPH_PRIORITY", "-2"))
            self._capture_stream = torch.cuda.Stream(device=device, priority=prio)
        stream = self._capture_stream
        try:
            with torch.cuda.graph(cs.graph, stream=stream):
                loss, flag, params = self._evaluate(args, {})
                zero = torch.zeros((), dtype=torch.float64, device=device)
                out = torch.stack((loss.detach().reshape(()).to(torch.float64), zero if flag is None else flag.to(torch.float64)))
                cs.out_host.copy_(out, non_blocking=True)
            cs.infos = ops.end_deferred(token)
        except BaseException:
            ops.end_deferred(token)
            if not _recover_generator(device, rng_state):
                self._graphs = {k: (self.graph_warmup if isinstance(v, _CapturedStep) else v) for k, v in self._graphs.items()}
            raise
        cs.static_args = args                  # the tensors the graph reads
        cs.params = params
        cs.grads = [p.grad for p in params]
        cs.traces = self.loss.last_traces
        for p in params:                   # nothing has run yet: the capture only recorded the work
            p.grad = None
        return cs

    def _replay(self, cs: "_CapturedStep", args=None) -> float:
        from ... import ops
        if args is not None:
            for a, st in zip(args, cs.static_args):
                if a is not None and a is not st and (a.data_ptr() != st.data_ptr()):
                    st.copy_(a, non_blocking=True)           # same signature, other storage: device-to-device, stream-ordered
        cs.graph.replay()
        torch.cuda.current_stream().synchronize()
        value, failed = float(cs.out_host[0]), float(cs.out_host[1])
        self.loss.last_traces = cs.traces
        if failed != 0.0:
            ops._raise_if_failed(cs.infos)
            raise torch.linalg.LinAlgError("linalg.cholesky: the factorisation failed on another rank of the sharded run")
        live = [(p, g) for p, g in zip(cs.params, cs.grads) if g is not None]
        for p, g in live:
            p.grad = g
        self.optim([p for p, _ in live])
        for p, _ in live:
            p.grad = None
        self.last_step_was_replay = True
        return value

    def _try_capture(self, key, args, device):
        try:
            entry = self._graphs[key] = self._capture(args, device)
        except Exception as e:        # noqa: BLE001 - whatever broke the capture, the eager path still works
            import warnings
            self._capture_failures += 1
            # the usual cause is transient (somebody still holds an old autograd graph whose AccumulateGrad nodes are bound to
            # the default stream): try again a few eager steps later, three times in all, before giving up for good
            self._graphs[key] = False if self._capture_failures >= 3 else -8
            torch.cuda.synchronize()
            warnings.warn(f"svi.step could not be captured in a CUDA graph ({type(e).__name__}: {e}); "
                          + ("continuing with eager steps" if self._capture_failures >= 3 else "eager steps for now, another attempt later"))
            return None
        return entry

    def step(self, *args, **kwargs) -> float:
        """One gradient step; returns the loss as a Python float (the only host<->device synchronisation)."""
        if self.cuda_graph and not kwargs:
            key, sig, device = self._graph_key(args)
            if key is not None:
                entry = self._graphs.get(key, 0)
                if isinstance(entry, _CapturedStep):
                    return self._replay(entry)
                owned = self._graphs.get(sig, 0)
                if isinstance(owned, _CapturedStep):             # same signature, other storage: copy in, replay
                    return self._replay(owned, args)
                if entry is not False and owned is not False:
                    if entry >= self.graph_warmup:               # the same tensors for the third time: capture on them
                        got = self._try_capture(key, args, device)
                        if got is not None:
                            return self._replay(got)
                    elif entry == 0 and owned >= self.graph_warmup + 1:
                        # the storage keeps changing under a constant signature: capture on copies this object owns
                        own = tuple(None if a is None else a.clone() for a in args)
                        got = self._try_capture(sig, own, device)
                        if got is not None:
                            return self._replay(got, args)
                    else:
                        if len(self._graphs) > 256:              # counters of storages that never came back
                            self._graphs = {k: v for k, v in self._graphs.items() if not isinstance(v, int)}
                        self._graphs[key] = entry + 1
                        if entry == 0:
                            self._graphs[sig] = owned + 1
        return self._eager_step(args, kwargs)

    def evaluate_loss(self, *args, **kwargs) -> float:
        return self.loss.loss(self.model, self.guide, *args, **kwargs)


from . import autoguide  # noqa: E402,F401
