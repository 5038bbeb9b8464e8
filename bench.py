#!/usr/bin/env python
"""bench.py — SVI steps/s of DRAUPNIR's tree-OU prior + decoder ELBO path on B200 (contract: see the task prompt).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4] [--impl ours|reference]

A "step" is one full `svi.step` (guide forward, model log-joint, backward, ClippedAdam update) on a seeded synthetic
tree + alignment.  Default workload: c4 = 2,000 leaves / 500 sites / variational guide / z = 30 (the configuration
BASELINE.json's target is quoted on; it fits one GPU).  One JSON line goes to stdout; everything else to stderr.
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
}
Z_DIM = 30
METRIC = "svi_steps_per_sec"
UNIT = "steps/s"
FP64_TENSOR_NOMINAL_TFLOPS = 40.0          # B200 data sheet, dense fp64 (tensor and CUDA-core figures coincide)
FP64_DMMA_MEASURED_TFLOPS = 37.15          # mma.sync.m8n8k4.f64 issue-bound loop on this pool (profiles/microbench/dmma_shapes.cu)
KCLASS = ["syrk", "trsm", "potf2", "assemble", "reduce", "catlik", "patristic", "other", "syrk_tcgen05", "slice", "gru"]
FLOP_CLASSES = (0, 1, 2, 8, 10)


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


def oracle_steps_per_sec(w, budget_s: float, max_steps: int, warmup: int):
    """The reference's arithmetic (fp64 torch restatement = oracle port) on the host cores; wall-clock bounded."""
    import torch
    from draupnir_asr_b200 import synthetic
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
        if time.perf_counter() - t_begin > 0.35 * budget_s:
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
    return sps, cores, f"{len(times)} full SVI step(s) of the same workload after {done_w} warm-up, {cores} threads, fp64"


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    budget = float(os.environ.get("DRP_REF_BUDGET_S", "150"))
    sps, cores, sample = oracle_steps_per_sec(w, budget, max(1, args.steps), args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / sps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}, z_dim={Z_DIM}", "n_leaves": w["n"], "sites": w["L"],
                       "guide": w["guide"], "z_dim": Z_DIM, "device": "host CPU"},
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    from draupnir_asr_b200 import runtime, synthetic, zshard
    from draupnir_asr_b200._lib import lib
    import ctypes

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shard = zshard.ShardContext(rank, world, Z_DIM, w["n"]) if world > 1 else None

    problem = synthetic.make_problem(w["n"], w["L"], seed=0)
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
        # step costs tens of ms on the host and stalls the launch thread); it runs between the regions instead
        import gc
        gc.collect()
        gc.disable()
        try:
            return _timed(fn, k)
        finally:
            gc.enable()

    def _timed(fn, k):
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

    # the sampler starts BEFORE the warm-up: NVML's start-up (attaching to every GPU of the box) can stall the launches
    # of a running context for tens of ms, which must not land inside a timed step; its polling does not
    sampler = make_clock_sampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        lw = lib.drp_launch_count()
        loss = run.step()
        launches_warm = lib.drp_launch_count() - lw
    log(f"[rank {rank}] warm-up done, -ELBO {loss:.4f}; {launches_warm} launches of this library per step")
    sampler.wait_first_sample()
    torch.cuda.synchronize()
    lib.drp_timing_enable(int(launches_warm * (args.steps + 1)) + 64)   # events for the whole timed region, created now
    barrier()                                                          # ranks leave the (rank-dependent) set-up together
    l0 = lib.drp_launch_count()
    ms = timed(run.step, args.steps)
    launches = lib.drp_launch_count() - l0
    NK = lib.drp_timing_classes()
    kms = (ctypes.c_double * NK)()
    kln = (ctypes.c_int64 * NK)()
    kwk = (ctypes.c_double * NK)()
    lib.drp_timing_read(kms, kln, kwk)
    lib.drp_timing_enable(0)
    for _ in range(max(args.warmup, 3)):                                # the host-fed path has its own warm-up (copy stream, events)
        run.step_from_host(host)
    barrier()
    ms_e2e = timed(lambda: run.step_from_host(host), args.steps)
    barrier()
    clocks = sampler.stop()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    h2d = torch.tensor([float(run.host_bytes_per_step(host))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                     # max over ranks
        dist.all_reduce(h2d, op=dist.ReduceOp.SUM)                   # bytes copied by all ranks together
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    per_class = {KCLASS[k]: {"ms_per_step": kms[k] / args.steps, "launches_per_step": kln[k] / args.steps,
                             "work_per_step": kwk[k] / args.steps} for k in range(NK) if kln[k]}
    # achieved rate of every class against the peak that bounds it (north star: tensor-pipe utilisation for the
    # Cholesky / GEMM kernels, HBM GB/s for assembly and likelihood); with the prior on a side stream the classes overlap,
    # so these per-class rates are lower bounds of what each kernel does alone (DRP_PRIOR_STREAM=0 gives the unshared ones)
    for k in range(NK):
        if kln[k] and kms[k] > 0:
            flop = k in FLOP_CLASSES
            rate = kwk[k] / (kms[k] * 1e-3) / (1e12 if flop else 1e9)
            pk = peaks["tensor_sustained"] if flop else peaks["hbm"]
            per_class[KCLASS[k]].update({"achieved": rate, "unit": "TFLOP/s" if flop else "GB/s", "frac_of_peak": rate / pk,
                                         "bound": "tensor" if flop else "hbm"})
    dom = max(range(NK), key=lambda k: kms[k])
    tensor_bound = dom in FLOP_CLASSES
    avg_ms = kms[dom] / max(kln[dom], 1)
    work_per_launch = kwk[dom] / max(kln[dom], 1)
    achieved = work_per_launch / (avg_ms * 1e-3) / (1e12 if tensor_bound else 1e9)
    peak = peaks["tensor_sustained"] if tensor_bound else peaks["hbm"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get(KCLASS[dom])
        except Exception:
            traffic = None
    roofline = {"bound": "tensor" if tensor_bound else "hbm", "kernel": KCLASS[dom] + "_kernel", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": f"{peaks['source']} ({'bf16 dense sustained' if tensor_bound else 'HBM copy'})",
                "avg_launch_ms": avg_ms, "algorithmic_work_per_launch": work_per_launch,
                "share_of_step": kms[dom] / ms}
    if tensor_bound:
        roofline["frac_of_fp64_dmma_measured"] = achieved / FP64_DMMA_MEASURED_TFLOPS
        if dom == 8:      # split-integer update: every fp64-equivalent flop costs S(S+1)/2 int8 MACs on the tensor pipe
            S = int(os.environ.get("DRP_OZAKI_SLICES", "6") or 6)
            pairs = S * (S + 1) // 2
            int8_peak = 2.0 * peak                               # dense int8 = 2 x dense bf16
            roofline["int8_tops_executed"] = achieved * pairs
            roofline["frac_of_int8_peak"] = achieved * pairs / int8_peak
            roofline["note"] = (f"fp64-accurate update executed as {pairs} int8 GEMMs (S = {S} slices) on tcgen05.mma.kind::i8; "
                                f"'achieved' counts fp64-equivalent flops, the tensor pipe itself runs {achieved * pairs:.0f} TOP/s = "
                                f"{achieved * pairs / int8_peak:.3f} of the int8 peak (2 x measured bf16)")
        else:
            roofline["note"] = ("fp64 kernel (reference dtype) on mma.sync.m8n8k4.f64; against the measured fp64 DMMA peak of "
                                f"{FP64_DMMA_MEASURED_TFLOPS} TFLOP/s the fraction is {achieved / FP64_DMMA_MEASURED_TFLOPS:.3f} "
                                "(the bf16 peak in 'peak' is not reachable by fp64 arithmetic)")
        roofline["frac_of_fp64_nominal"] = achieved / FP64_TENSOR_NOMINAL_TFLOPS

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        budget = float(os.environ.get("DRP_CPU_BUDGET_S", "25"))
        sps, cores, sample = oracle_steps_per_sec(w, budget, 8, 1 if w["n"] <= 800 else 0)
        cpu = {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    value = args.steps / (ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}, z_dim={Z_DIM}", "n_leaves": w["n"], "sites": w["L"],
                       "guide": w["guide"], "z_dim": Z_DIM, "gru_hidden_dim": 60, "l2": "flushed between timed iterations",
                       "parallelism": "single GPU" if world == 1 else f"z-sharded OU prior + leaf-sharded decoder over {world} GPUs"},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d[0]),
                    "d2h_bytes_per_step": 8 + 4 * Z_DIM, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "kernel_classes": per_class}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
