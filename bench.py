#!/usr/bin/env python
"""bench.py — SVI steps/s of DRAUPNIR's tree-OU prior + decoder ELBO path on B200 (contract: see the task prompt).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl ours|reference]

A "step" is one full `svi.step` (guide forward, model log-joint, backward, ClippedAdam update) on a seeded synthetic
tree + alignment.  Default workload: c4 = 2,000 leaves / 500 sites / variational guide / z = 30 (the configuration
BASELINE.json's target is quoted on; it fits one GPU).  c5 is BASELINE.json's patristic-matrix + OU-covariance build sweep
(100-8,000 leaves; a "step" there is one build of both at the largest size).  One JSON line goes to stdout; everything
else to stderr.

What the line's numbers are: `value` — steps/s with device-resident inputs, the step replayed from its CUDA
graph; `e2e` — the same through `Run.step_from_host(host, prefetch=host)`: every step copies its inputs from pinned host
memory, one step early (double-buffered), and reads the loss back (`ms_per_step_copy_inside_step`: without the overlap, the
copies being memcpy nodes of the step's own graph); `kernels` — per-kernel device time / achieved rate from an extra, EAGER, instrumented pass after the timed regions
(CUDA events around every launch of the library on its launching stream); `roofline` — the kernel with the largest share
of that pass; `north_star_kernel` — always the tcgen05 trailing update; `conditional` — K6 at this size; `cpu_baseline` —
the oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c1": dict(n=19, L=225, guide="delta_map", desc="Randall CFP shape: 19 leaves, 225 sites, delta_map"),
    "c2": dict(n=200, L=300, guide="variational", desc="200-leaf random tree, 300 sites, variational"),
    "c3": dict(n=800, L=400, guide="delta_map", desc="800-leaf tree, 400 sites, delta_map"),
    "c4": dict(n=2000, L=500, guide="variational", desc="2,000-leaf tree, 500 sites, variational"),
    "c5": dict(n=8000, L=0, guide=None, desc="patristic-matrix + OU-covariance build sweep, 100-8,000 leaves"),
}
C5_SIZES = (100, 200, 500, 1000, 2000, 4000, 8000)
Z_DIM = 30
METRIC = "svi_steps_per_sec"
UNIT = "steps/s"
# Builder constants (NOT in MEASURED_PEAKS.json; every line that uses them says so in `peak_source`):
FP64_TENSOR_NOMINAL_TFLOPS = 40.0          # B200 data sheet, dense fp64 (tensor and CUDA-core figures coincide)
FP64_DMMA_MEASURED_TFLOPS = 37.15          # mma.sync.m8n8k4.f64 issue-bound loop on this pool (profiles/microbench/dmma_shapes.cu)
INT8_OVER_BF16 = 2.0                       # dense int8 tensor rate / dense bf16 rate (data sheet ratio; the int8 rate itself is unmeasured)


_REAL_STDOUT = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"], "tensor_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and clock-event (throttle) reasons DURING the timed region — the quantities of the nvidia-smi line in
    B200_PROFILING.md (clocks.sm, clocks.max.sm, clocks_event_reasons.*), read through NVML in a background thread every
    250 ms.  An `nvidia-smi -lms` child process was used first: its queries stalled the launches of the running context
    for 60-160 ms once in a while (single 95-190 ms steps among 30 ms ones); the in-process NVML calls do not."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.gpu, self.thread, self.stop_flag = gpu_index, None, False
        self.sm, self.mx, self.reasons, self.error = [], [], set(), None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.mx.append(float(mx))
                mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.25)
            pynvml.nvmlShutdown()
        except Exception as e:  # pragma: no cover
            self.error = e

    def start(self):
        import threading
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def wait_first_sample(self, timeout_s: float = 5.0):
        t0 = time.time()
        while not self.sm and self.error is None and time.time() - t0 < timeout_s:
            time.sleep(0.02)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml"}
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=5)
        if self.error is not None:
            log("clock sampler unavailable:", self.error)
        if self.sm:
            sm = sorted(self.sm)
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(self.mx), reasons=sorted(self.reasons), samples=len(sm))
        return out


class SmiClockSampler:
    """Fallback when NVML cannot be loaded in-process: an `nvidia-smi -lms 200` child (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)
            self.proc = None

    def wait_first_sample(self, timeout_s: float = 5.0):
        t0 = time.time()
        while self.proc is not None and time.time() - t0 < timeout_s:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                return
            time.sleep(0.05)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception as e:  # pragma: no cover
            log("clock log unreadable:", e)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def make_clock_sampler(gpu_index: int):
    try:
        import pynvml
        pynvml.nvmlInit()
        return ClockSampler(gpu_index)
    except Exception as e:
        log("NVML not usable in-process (", e, "): falling back to an nvidia-smi child")
        return SmiClockSampler(gpu_index)


def load_synthetic_standalone():
    """`draupnir_asr_b200/synthetic.py` as a stand-alone module: the reference arm needs the seeded inputs but must not
    import the package, whose `__init__` loads the CUDA library (numpy / torch only; nothing of the product runs)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_drp_synthetic", os.path.join(ROOT, "draupnir_asr_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_drp_synthetic"] = mod
    spec.loader.exec_module(mod)
    return mod


def config_dict(args, w, world):
    """Identical keys and values in both arms (the driver compares them)."""
    return {"workload": f"{args.workload}: {w['desc']}, z_dim={Z_DIM}", "n_leaves": w["n"], "sites": w["L"], "guide": w["guide"],
            "z_dim": Z_DIM, "gru_hidden_dim": 60, "l2": "flushed between timed iterations", "ranks": world}


def oracle_steps_per_sec(w, budget_s: float, max_steps: int, warmup: int, synthetic):
    """The reference's arithmetic (fp64 torch restatement = oracle port) on the host cores; wall-clock bounded."""
    import torch
    from oracle import draupnir_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    problem = synthetic.make_problem(w["n"], w["L"], seed=0)
    init = synthetic.make_parameters(w["n"], z_dim=Z_DIM, seed=0)
    nets = O.Networks(z_dim=Z_DIM, variational=(w["guide"] == "variational"), seed=0)
    svi = O.OracleSVI(nets, init, w["guide"], z_dim=Z_DIM)
    eps = torch.randn(Z_DIM, w["n"], dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    args = (problem.dataset_train, problem.patristic_matrix_train, problem.dataset_train_blosum, problem.blosum_weighted, eps)
    t_begin = time.perf_counter()
    done_w = 0
    for _ in range(warmup):
        if done_w >= 1 and time.perf_counter() - t_begin > 0.35 * budget_s:
            break
        svi.step(*args)
        done_w += 1
    times = []
    while len(times) < max_steps:
        t0 = time.perf_counter()
        svi.step(*args)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > budget_s:
            break
    sps = len(times) / sum(times)
    return sps, cores, len(times), f"{len(times)} full SVI step(s) of the same workload after {done_w} warm-up, {cores} threads, fp64"


def c5_cpu_sample(synthetic, budget_s: float):
    """C5 on the host: parent-walk LCA patristic rows (scalar C oracle) and OUKernel_Fast in torch, on a bounded sample."""
    import numpy as np
    import torch
    from oracle import draupnir_oracle as O
    from oracle import patristic_oracle as P
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = WORKLOADS["c5"]["n"]
    t = synthetic.random_tree(n, seed=n)
    N = t.n_nodes
    rows = np.arange(0, N, max(1, N // 1024), dtype=np.int32)[:1024]
    cols = np.arange(N, dtype=np.int32)
    P.patristic_c(t.parent, t.branch, t.depth, rows[:8], cols, want_lca=False)         # warm-up (loads the library)
    t0 = time.perf_counter()
    P.patristic_c(t.parent, t.branch, t.depth, rows, cols, want_lca=False)
    t_pat = time.perf_counter() - t0
    b_pat = len(rows) * N * 8
    m = 2000
    g = torch.Generator().manual_seed(0)
    D = torch.rand(m, m, generator=g, dtype=torch.float64)
    sf, sn, lam = (torch.rand(Z_DIM, generator=g, dtype=torch.float64) + 0.5 for _ in range(3))
    O.ou_covariance(D[:200, :200], sf, sn, lam)
    reps, t_ou = 0, 0.0
    while reps < 1 or (t_ou < min(10.0, budget_s) and reps < 5):
        t0 = time.perf_counter()
        O.ou_covariance(D, sf, sn, lam)
        t_ou += time.perf_counter() - t0
        reps += 1
    t_ou /= reps
    b_ou = m * m * 8 * (Z_DIM + 1)
    gbs = (b_pat + b_ou) / (t_pat + t_ou) / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": (f"{len(rows)} of {N} patristic rows at {n} leaves (scalar C parent-walk, 1 thread: {b_pat / t_pat / 1e9:.3f} GB/s) + "
                       f"OUKernel_Fast [30,{m},{m}] in torch on {cores} threads ({b_ou / t_ou / 1e9:.3f} GB/s), fp64"),
            "patristic_gbs": b_pat / t_pat / 1e9, "ou_cov_gbs": b_ou / t_ou / 1e9}


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    synthetic = load_synthetic_standalone()
    budget = float(os.environ.get("DRP_REF_BUDGET_S", "150"))
    if args.workload == "c5":
        cpu = c5_cpu_sample(synthetic, budget)
        line = {"impl": "reference", "metric": C5_METRIC, "value": cpu["value"], "unit": "GB/s", "n_gpus": args.gpus, "steps": 1,
                "warmup": 1, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_dict(args, w, world), "cpu_baseline": cpu,
                "e2e": {"value": cpu["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return
    sps, cores, timed, sample = oracle_steps_per_sec(w, budget, max(1, args.steps), max(1, args.warmup), synthetic)
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": timed,
            "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / sps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, w, world),
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
def _event_ms(fn, reps, flush=None, warm=1):
    import torch
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def read_kernel_table(lib, steps, peaks):
    """Per-kernel device time and achieved rate from the library's instrumentation (CUDA events around every launch on the
    launching stream).  Flop-counted kernels are set against the measured dense bf16 rate (the contract's tensor peak;
    fp64 arithmetic cannot reach it, so the fraction of the fp64 DMMA rate - a builder constant - rides along), byte-counted
    kernels against the measured HBM copy rate."""
    import ctypes
    NK = lib.drp_timing_classes()
    kms, kln, kwk = (ctypes.c_double * NK)(), (ctypes.c_int64 * NK)(), (ctypes.c_double * NK)()
    lib.drp_timing_read(kms, kln, kwk)
    table = {}
    for k in range(NK):
        if not kln[k]:
            continue
        name = lib.drp_timing_class_name(k).decode()
        flop = bool(lib.drp_timing_class_is_flops(k))
        row = {"ms_per_step": kms[k] / steps, "launches_per_step": kln[k] / steps, "avg_launch_ms": kms[k] / kln[k],
               "work_per_step": kwk[k] / steps, "bound": "tensor" if flop else "hbm"}
        if kms[k] > 0 and kwk[k] > 0:
            rate = kwk[k] / (kms[k] * 1e-3) / (1e12 if flop else 1e9)
            pk = peaks["tensor_sustained"] if flop else peaks["hbm"]
            row.update(achieved=rate, unit="TFLOP/s" if flop else "GB/s", peak=pk, frac=rate / pk)
            if flop:
                row["frac_of_fp64_dmma_builder_constant"] = rate / FP64_DMMA_MEASURED_TFLOPS
        table[name] = row
    return table


def roofline_entry(name, row, step_ms, peaks, traffic):
    if row is None or "achieved" not in row:
        return None
    tensor = row["bound"] == "tensor"
    out = {"bound": row["bound"], "kernel": name, "achieved": row["achieved"], "peak": row["peak"], "unit": row["unit"],
           "frac": row["frac"], "traffic": traffic.get(name),
           "peak_source": (f"{peaks['source']} MEASURED_PEAKS.json: " + ("dense bf16, sustained" if tensor else "HBM copy")),
           "avg_launch_ms": row["avg_launch_ms"], "launches_per_step": row["launches_per_step"],
           "algorithmic_work_per_launch": row["work_per_step"] / row["launches_per_step"],
           "ms_per_step": row["ms_per_step"], "share_of_step": row["ms_per_step"] / step_ms,
           "timed": "CUDA events around each launch on its launching stream; eager, single-stream instrumented pass after the timed "
                    "regions (share_of_step is relative to that serialised step)"}
    if tensor:
        out["frac_of_fp64_dmma"] = row["achieved"] / FP64_DMMA_MEASURED_TFLOPS
        out["fp64_dmma_peak_source"] = f"builder constant {FP64_DMMA_MEASURED_TFLOPS} TFLOP/s (profiles/microbench/dmma_shapes.cu), not in MEASURED_PEAKS.json"
    if name == "oz_syrk_kernel":       # split-integer update: every fp64-equivalent flop costs S(S+1)/2 int8 MACs on the tensor pipe
        S = int(os.environ.get("DRP_OZAKI_SLICES", "6") or 6)
        pairs = S * (S + 1) // 2
        int8_peak = INT8_OVER_BF16 * row["peak"]
        out["int8_tops_executed"] = row["achieved"] * pairs
        out["frac_of_int8_peak"] = row["achieved"] * pairs / int8_peak
        out["int8_peak_source"] = f"{INT8_OVER_BF16} x measured dense bf16 (data-sheet ratio, a builder assumption; the int8 rate is not in MEASURED_PEAKS.json)"
        out["note"] = (f"fp64-accurate trailing update executed as {pairs} int8 GEMMs (S = {S} byte slices) on tcgen05.mma.kind::i8; "
                       "'achieved' counts fp64-equivalent flops (2 K per updated entry)")
    elif tensor:
        out["note"] = "fp64 kernel (the reference dtype) on mma.sync.m8n8k4.f64: the bf16 peak in 'peak' is not reachable by fp64 arithmetic"
    return out


def load_traffic():
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(path)).get("kernels", {})
    except Exception:
        return {}


def finish_ranks(world):
    """Multi-rank exit.  Tearing down a NCCL communicator whose collectives are nodes of live CUDA graphs can block for
    minutes; the ranks meet at a barrier, flush, and leave without running the destructors."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    if _REAL_STDOUT is not None:
        _REAL_STDOUT.flush()
    os._exit(0)


def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    from draupnir_asr_b200 import runtime, synthetic, zshard
    from draupnir_asr_b200._lib import lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shard = zshard.ShardContext(rank, world, Z_DIM, w["n"]) if world > 1 else None

    # the patristic matrices come from the product's own builder (K1, csrc/patristic.cu)
    problem = synthetic.make_problem(w["n"], w["L"], seed=0, builder="device", device=dev)
    init = synthetic.make_parameters(w["n"], z_dim=Z_DIM, seed=0)
    run = runtime.build_run(problem, w["guide"], device=dev, init=init, z_dim=Z_DIM, shard=shard)
    host = run.make_host_inputs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        # the cyclic garbage collector is kept out of the timed region (a generation-2 pass over the autograd graphs of a
        # step costs tens of ms on the host and stalls the launch thread); the collection itself runs BEFORE the barrier that
        # opens the region (`quiesce`), so that the GPU does not sit idle between the barrier and the first timed step
        import gc
        gc.disable()
        try:
            evs = []
            for _ in range(k):
                flush.zero_()                                             # L2 flush between timed iterations (untimed)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            per = [a.elapsed_time(b) for a, b in evs]
            log(f"[rank {rank}] per-step ms: " + " ".join(f"{v:.2f}" for v in per))
            return sum(per)
        finally:
            gc.enable()

    def quiesce():
        import gc
        gc.collect()
        barrier()

    # the sampler starts BEFORE the warm-up: NVML's start-up (attaching to every GPU of the box) can stall the launches
    # of a running context for tens of ms, which must not land inside a timed step; its polling does not
    sampler = make_clock_sampler(local_rank)
    sampler.start()
    W = max(args.warmup, 3)
    for _ in range(W):                                   # two eager steps, then the step is captured and replayed
        loss = run.step()
    graphed = bool(run.svi.last_step_was_replay)
    log(f"[rank {rank}] warm-up done, -ELBO {loss:.4f}; step replayed from a CUDA graph: {graphed}")
    sampler.wait_first_sample()
    quiesce()                                                          # ranks leave the (rank-dependent) set-up together
    ms = timed(run.step, args.steps)
    for _ in range(W):                                                 # the host-fed path has its own warm-up / capture
        run.step_from_host(host)
    quiesce()
    ms_e2e_serial = timed(lambda: run.step_from_host(host), args.steps)     # copies inside the step (memcpy nodes of its graph)
    # e2e proper: the loop of a double-buffered loader — every step copies its full inputs from pinned host memory, one step
    # early (while the previous step computes), into the other of two staging sets; each set has its own captured graph
    for _ in range(max(W, 7)):
        run.step_from_host(host, prefetch=host)
    graphed_e2e = bool(run.svi.last_step_was_replay)
    quiesce()
    ms_e2e = timed(lambda: run.step_from_host(host, prefetch=host), args.steps)
    barrier()
    clocks = sampler.stop()

    # ---- instrumented pass: the same step, eager, with CUDA events around every launch of the library
    # (the prior's side stream is switched off for it, so that every kernel is timed ALONE: with the overlap on, a kernel's
    # span includes the time it waits for SMs the other chain holds)
    from draupnir_asr_b200 import ops as _ops
    P = max(2, min(args.steps, 5))
    run.svi.cuda_graph = False
    _ops.PRIOR_STREAM_ENABLED = False
    l0 = lib.drp_launch_count()
    run.step()
    launches_per_step = int(lib.drp_launch_count() - l0)
    torch.cuda.synchronize()
    lib.drp_timing_enable(int(launches_per_step * (P + 1)) + 64)
    ms_eager = timed(run.step, P)
    peaks = measured_peaks()
    kernels = read_kernel_table(lib, P, peaks)
    lib.drp_timing_enable(0)
    run.svi.cuda_graph = True
    _ops.PRIOR_STREAM_ENABLED = os.environ.get("DRP_PRIOR_STREAM", "1") != "0"

    # ---- K6: conditional posterior of the internal nodes at this size (S/models.py:149-242), z-sharded like the prior
    with torch.no_grad():
        est = run.guide(*run.step_args())
        est = {k: est[k].detach() for k in ("sigma_f", "sigma_n", "lambd", "latent_z")}
        D = run.Draupnir
        n, ni, N = problem.n_leaves, problem.n_internal, problem.n_leaves + problem.n_internal
        reps = 3 if w["n"] <= 800 else 2
        ms_map = _event_ms(lambda: D.conditional_samplingMAP(est, run.patristic_matrix_full), reps, flush)
        ms_cov = _event_ms(lambda: D._conditional(est, run.patristic_matrix_full, D.internal_nodes, D.n_internal, 0.0, 1e-6, True),
                           reps, flush)
    t2 = torch.tensor([ms, ms_e2e, ms_map, ms_cov, ms_e2e_serial], dtype=torch.float64, device=dev)
    h2d = torch.tensor([float(run.host_bytes_per_step(host))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)                    # max over ranks
        dist.all_reduce(h2d, op=dist.ReduceOp.SUM)                   # bytes copied by all ranks together
    ms, ms_e2e, ms_map, ms_cov, ms_e2e_serial = (float(v) for v in t2)
    if rank != 0:
        finish_ranks(world)
        return

    step_ms = ms / args.steps
    traffic = load_traffic()
    timed_rows = {k: v for k, v in kernels.items() if "achieved" in v and k not in ("misc", "small_reductions")}
    dom = max(timed_rows, key=lambda k: timed_rows[k]["ms_per_step"]) if timed_rows else None
    eager_step_ms = ms_eager / P
    roofline = roofline_entry(dom, kernels.get(dom), eager_step_ms, peaks, traffic) if dom else None
    north = roofline_entry("oz_syrk_kernel", kernels.get("oz_syrk_kernel"), eager_step_ms, peaks, traffic)
    ours_map = Z_DIM * (n ** 3 / 3.0 + n * n * ni)                  # elimination of the leaf block, mean only
    ours_cov = Z_DIM * (n ** 3 / 3.0 + n * n * ni + n * ni * ni)
    ref_flops = Z_DIM * (2.0 * N ** 3 + 4.0 * ni ** 3)                # the reference's three dense inverses
    conditional = {"mean_only_ms": ms_map, "mean_and_covariance_ms": ms_cov,
                   "algorithmic_tflops_mean_only": ours_map / (ms_map * 1e-3) / 1e12,
                   "algorithmic_tflops_mean_and_covariance": ours_cov / (ms_cov * 1e-3) / 1e12,
                   "reference_flops": ref_flops, "reference_equivalent_tflops": ref_flops / (ms_cov * 1e-3) / 1e12,
                   "flop_saving_vs_reference": ref_flops / ours_cov,
                   "what": "conditional_samplingMAP (mean [z,ni]) and conditional mean + covariance [z,ni,ni] of the internal nodes given "
                           "the leaves, Schur complement on the leaf block (S/models.py:149-242), CUDA events, median"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        budget = float(os.environ.get("DRP_CPU_BUDGET_S", "25"))
        sps, cores, _, sample = oracle_steps_per_sec(w, budget, 8, 1, load_synthetic_standalone())
        cpu = {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    value = args.steps / (ms * 1e-3)
    cfg = config_dict(args, w, world)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "parallelism": "single GPU" if world == 1 else f"z-sharded OU prior + leaf-sharded decoder over {world} GPUs", "roofline": roofline, "north_star_kernel": north, "cpu_baseline": cpu,
            "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d[0]),
                    "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / args.steps, "cuda_graph": graphed_e2e,
                    "pipelining": "double-buffered: Run.step_from_host(host, prefetch=next) copies step k+1's inputs from pinned "
                                  "host memory while step k computes; every timed step moves h2d_bytes_per_step",
                    "ms_per_step_copy_inside_step": ms_e2e_serial / args.steps},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_note": (f"{launches_per_step} kernels of libdraupnir_b200.so per step (counted on an eager step) x {args.steps} timed steps; "
                                  + ("in the timed region they are nodes of the replayed CUDA graph" if graphed else "launched eagerly")),
            "cuda_graph": graphed, "eager_ms_per_step": eager_step_ms, "clocks": clocks, "conditional": conditional,
            "kernels": kernels}
    emit(line)
    finish_ranks(world)


# ---------------------------------------------------------------------------------------------------------------------
C5_METRIC = "patristic_plus_ou_covariance_build_gbs"


def run_c5(args, w, rank, world, local_rank):
    """BASELINE.json config 5: all-node patristic matrix (K1) + [30, n, n] OU covariance over the leaves (K2), 100-8,000
    leaves.  Independent trees shard trivially: with N ranks every rank builds its own copy (weak scaling, no collective)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    from draupnir_asr_b200 import ops, synthetic as S
    from draupnir_asr_b200._lib import lib
    from oracle import patristic_oracle as P

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = make_clock_sampler(local_rank)
    sampler.start()
    g = torch.Generator().manual_seed(0)
    sf, sn, lam = (torch.rand(Z_DIM, generator=g, dtype=torch.float64).add(0.5).to(dev) for _ in range(3))
    sweep = []
    K = None
    for n in C5_SIZES:
        t = S.random_tree(n, seed=n)
        N = t.n_nodes
        parent, branch = torch.from_numpy(t.parent).to(dev), torch.from_numpy(t.branch).to(dev)
        row = {"n_leaves": n, "n_nodes": N}
        # parity: LCA indices bit-exact against the scalar C oracle (all pairs up to 2,000 leaves, a 256 x 256 sample above)
        if n <= 2000:
            pat, _, lca = ops.tree_distances(parent, branch, want_lca=True)
            lca_ref, pat_ref, _ = P.patristic_c(t.parent, t.branch, t.depth)
        else:
            rng = np.random.default_rng(n)
            rows = np.sort(rng.choice(N, 256, replace=False)).astype(np.int32)
            cols = np.sort(rng.choice(N, 256, replace=False)).astype(np.int32)
            pat, _, lca = ops.tree_distances(parent, branch, rows=torch.from_numpy(rows), cols=torch.from_numpy(cols), want_lca=True)
            lca_ref, pat_ref, _ = P.patristic_c(t.parent, t.branch, t.depth, rows, cols)
        row["lca_bit_exact"] = bool(np.array_equal(lca.cpu().numpy(), lca_ref))
        row["patristic_max_rel_err"] = float(np.max(np.abs(pat.cpu().numpy() - pat_ref) / np.maximum(np.abs(pat_ref), 1e-300)))
        del pat, lca
        ms = _event_ms(lambda: ops.tree_distances(parent, branch), 7, flush, warm=3)
        row.update(patristic_ms=ms, patristic_bytes=N * N * 8, patristic_gbs=N * N * 8 / (ms * 1e-3) / 1e9)
        ms = _event_ms(lambda: ops.tree_distances(parent, branch, want_lca=True), 5, flush, warm=1)
        row.update(patristic_lca_ms=ms, patristic_lca_gbs=N * N * 12 / (ms * 1e-3) / 1e9)
        leaves = torch.from_numpy(t.leaves).to(dev)
        D = ops.tree_distances(parent, branch, rows=leaves, cols=leaves)[0]

        def build_cov():
            nonlocal K
            K = None
            K = ops.ou_cov(D, sf, sn, lam)

        ms = _event_ms(build_cov, 5, flush, warm=2)
        row.update(ou_cov_ms=ms, ou_cov_bytes=n * n * 8 * (Z_DIM + 1), ou_cov_gbs=n * n * 8 * (Z_DIM + 1) / (ms * 1e-3) / 1e9)
        K = None
        for k in ("patristic_gbs", "patristic_lca_gbs", "ou_cov_gbs"):
            row[k.replace("_gbs", "_frac_of_hbm_peak")] = row[k] / peaks["hbm"]
        sweep.append(row)
        log(json.dumps(row))
    # ---- the timed "steps": one build of both at the largest size, K of them; then per-kernel events for the roofline
    n = C5_SIZES[-1]
    t = S.random_tree(n, seed=n)
    N = t.n_nodes
    parent_h = torch.from_numpy(t.parent).pin_memory()
    branch_h = torch.from_numpy(t.branch).pin_memory()
    parent, branch = parent_h.to(dev), branch_h.to(dev)
    leaves = torch.from_numpy(t.leaves).to(dev)
    out_h = torch.zeros(2, dtype=torch.float64).pin_memory()

    def build(parent, branch):
        nonlocal K
        K = None
        full = ops.tree_distances(parent, branch)[0]
        D = ops.tree_distances(parent, branch, rows=leaves, cols=leaves)[0]          # the leaves' matrix the prior consumes
        K = ops.ou_cov(D, sf, sn, lam)
        return full, K

    def step_resident():
        build(parent, branch)

    def step_host():
        p, b = parent_h.to(dev, non_blocking=True), branch_h.to(dev, non_blocking=True)
        full, Kc = build(p, b)
        out_h.copy_(torch.stack((full[N - 1].sum(), Kc[Z_DIM - 1, n - 1].sum())), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    W = max(args.warmup, 3)
    for _ in range(W):
        step_resident()
    sampler.wait_first_sample()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step_resident(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    for _ in range(W):
        step_host()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step_host(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms_e2e = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop()
    l0 = lib.drp_launch_count()
    step_resident()
    launches_per_step = int(lib.drp_launch_count() - l0)
    P_ = 5
    lib.drp_timing_enable(launches_per_step * (P_ + 1) + 64)
    for _ in range(P_):
        flush.zero_()
        step_resident()
    kernels = read_kernel_table(lib, P_, peaks)
    lib.drp_timing_enable(0)
    t2 = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t2[0]), float(t2[1])
    if rank != 0:
        finish_ranks(world)
        return
    bytes_step = N * N * 8 + n * n * 8 + n * n * 8 * (Z_DIM + 1)      # all-node matrix + leaf matrix + D read once, K written
    traffic = load_traffic()
    step_ms = ms / args.steps
    dom = max((k for k in kernels if "achieved" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = c5_cpu_sample(load_synthetic_standalone(), float(os.environ.get("DRP_CPU_BUDGET_S", "25")))
    cfg = config_dict(args, w, world)
    line = {"parallelism": "single GPU" if world == 1 else f"{world} independent builds, one per GPU (no collective)",
            "metric": C5_METRIC, "value": world * bytes_step * args.steps / (ms * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": roofline_entry(dom, kernels[dom], step_ms, peaks, traffic), "cpu_baseline": cpu,
            "e2e": {"value": world * bytes_step * args.steps / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": int(N * 12), "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks, "kernels": kernels, "sweep": sweep,
            "all_lca_bit_exact": all(r["lca_bit_exact"] for r in sweep), "hbm_peak_gbs": peaks["hbm"]}
    emit(line)
    finish_ranks(world)


def main():
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner, cuDNN notices) are sent to
    # stderr for the duration of the run and the line is written to the saved descriptor at the end
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("DRP_WORKLOAD", "c4"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    elif args.workload == "c5":
        run_c5(args, w, rank, world, local_rank)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
