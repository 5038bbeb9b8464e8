"""Caller-side glue: the namedtuples `S/main.py:45-59` hands to the model classes and a builder that wires a synthetic
problem into model + guide + SVI the way `draupnir_train` does (S/main.py:991-1059).  Host logic only."""
from __future__ import annotations

from collections import namedtuple
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from .backend import pyro
from .guides import DRAUPNIRGUIDES
from .models import DRAUPNIRModel_classic
from .pyro_compat.infer import SVI, Trace_ELBO
from .pyro_compat.infer.autoguide import AutoDelta, AutoDiagonalNormal, AutoNormal, init_to_value
from .pyro_compat.optim import ClippedAdam
from .synthetic import SyntheticProblem

# S/main.py:54-59
ModelLoad = namedtuple("ModelLoad", ["z_dim", "align_seq_len", "device", "args", "build_config", "leaves_nodes",
                                     "n_tree_levels", "gru_hidden_dim", "pretrained_params", "aa_frequencies", "blosum",
                                     "blosum_max", "blosum_weighted", "dataset_train_blosum", "variable_score",
                                     "internal_nodes", "graph_coo", "nodes_representations_array", "dgl_graph",
                                     "children_dict", "closest_leaves_dict", "descendants_dict", "clades_dict_all",
                                     "leaves_testing", "plate_unordered", "one_hot_encoding"])
BuildConfig = namedtuple("BuildConfig", ["alignment_file", "use_ancestral", "n_test", "build_graph", "aa_probs", "triTSNE",
                                         "align_seq_len", "leaves_testing", "batch_size", "plate_subsample_size",
                                         "script_dir", "no_testing"])

# S/main.py:2761-2779
DEFAULT_CONFIG = {"lr": 1e-3, "beta1": 0.9, "beta2": 0.999, "eps": 1e-8, "weight_decay": 0.0, "clip_norm": 10.0,
                  "lrd": 1.0, "z_dim": 30, "gru_hidden_dim": 60}


def default_args(select_guide: str, num_epochs: int = 1, embedding_dim: int = 50, device="cuda"):
    """The argparse fields of Draupnir_example.py:62-143 that reach the model/guide constructors."""
    return SimpleNamespace(num_epochs=num_epochs, embedding_dim=embedding_dim, position_embedding_dim=30,
                           max_indel_size=5, batch_by_clade=False, kappa_addition=5, plating=False, use_cuda=True,
                           select_guide=select_guide, plate_unordered=False, device=device, use_scheduler=False)


def model_load_from_problem(problem: SyntheticProblem, select_guide: str, device, z_dim=30, gru_hidden_dim=60,
                            embedding_dim=50) -> ModelLoad:
    dev = torch.device(device)
    build = BuildConfig(alignment_file=None, use_ancestral=True, n_test=0, build_graph=False, aa_probs=problem.aa_probs,
                        triTSNE=False, align_seq_len=problem.align_seq_len, leaves_testing=False,
                        batch_size=problem.n_leaves,      # S/main.py:146: batch_size 1 => number of sequences
                        plate_subsample_size=None, script_dir=None, no_testing=False)
    to = lambda t: t.to(dev)
    return ModelLoad(z_dim=z_dim, align_seq_len=problem.align_seq_len, device=dev,
                     args=default_args(select_guide, embedding_dim=embedding_dim, device=dev), build_config=build,
                     leaves_nodes=to(problem.leaves_nodes), n_tree_levels=int(problem.tree.depth.max()) + 1,
                     gru_hidden_dim=gru_hidden_dim, pretrained_params=None, aa_frequencies=to(problem.aa_frequencies),
                     blosum=to(problem.blosum), blosum_max=to(problem.blosum_max),
                     blosum_weighted=to(problem.blosum_weighted), dataset_train_blosum=to(problem.dataset_train_blosum),
                     variable_score=to(problem.variable_score), internal_nodes=to(problem.internal_nodes), graph_coo=None,
                     nodes_representations_array=None, dgl_graph=None, children_dict=None, closest_leaves_dict=None,
                     descendants_dict=None, clades_dict_all=None, leaves_testing=False, plate_unordered=False,
                     one_hot_encoding=False)


class Run(SimpleNamespace):
    """model (`Draupnir`), `guide`, `svi`, `optim` and the device-resident step inputs of one training run."""

    def step_args(self):
        return (self.dataset_train, self.patristic_matrix_train, None, self.dataset_train_blosum, None, None)

    def step(self) -> float:
        return self.svi.step(*self.step_args())

    # ---- the same call with HOST inputs: what a caller holding CPU tensors pays per step --------------------------
    def make_host_inputs(self):
        """Pinned host copies of the tensors `svi.step` consumes (S/train.py:75): the dataset, the leaves' patristic
        matrix and, for the variational guide, the BLOSUM-encoded dataset."""
        pin = lambda t: t.detach().cpu().contiguous().pin_memory()
        host = {"dataset": pin(self.problem.dataset_train), "patristic": pin(self.problem.patristic_matrix_train)}
        if self.select_guide == "variational":
            host["blosum"] = pin(self.problem.dataset_train_blosum)
        self._staging = {k: v.to(self.device) for k, v in host.items()}       # rows outside a rank's block stay valid
        # Sharded over NCCL: every rank needs the WHOLE patristic matrix (its z-slice of the prior covers all leaves), but it
        # need not cross PCIe eight times: each rank copies its 1/N of the rows and one all-gather over NVLink rebuilds the
        # matrix on every GPU (32 MB at 2,000 leaves: 4 MB of H2D + ~0.1 ms of all-gather per rank instead of 32 MB of H2D)
        self._patristic_rows = None
        shard = getattr(self.Draupnir, "shard", None)
        if shard is not None and shard.world > 1:
            import torch.distributed as dist
            if dist.get_backend(shard.group) == "nccl":
                rows, cols = host["patristic"].shape
                per = (rows + shard.world - 1) // shard.world
                gathered = torch.zeros((per * shard.world, cols), dtype=host["patristic"].dtype, device=self.device)
                gathered[:rows].copy_(self._staging["patristic"])
                self._staging["patristic"] = gathered[:rows]
                lo = min(rows, shard.rank * per)
                self._patristic_rows = (lo, min(rows, lo + per), per, gathered,
                                        torch.zeros((per, cols), dtype=gathered.dtype, device=self.device))
                # the pipelined form (`prefetch=`) gathers on the copy stream WHILE the previous step's graph, which holds
                # all-reduces of the run's communicator, executes: it needs a communicator of its own
                self._prefetch_group = dist.new_group(ranks=list(range(shard.world)), backend="nccl")
        self._slots, self._pending = None, None
        return host

    def _leaf_rows(self):
        """Rows of the per-leaf inputs this rank reads: all of them on one GPU, its leaf block when the run is sharded
        (the model and the guide slice `family_data` / `data_blosum` to that block, so the other rows are never touched)."""
        shard = getattr(self.Draupnir, "shard", None)
        return (0, self.problem.n_leaves) if shard is None else (shard.leaf_lo, shard.leaf_hi)

    def host_bytes_per_step(self, host) -> int:
        """Bytes this rank copies host -> device per step."""
        lo, hi = self._leaf_rows()
        total = 0
        for k, v in host.items():
            rows = hi - lo
            if k == "patristic":
                pr = getattr(self, "_patristic_rows", None)
                rows = v.shape[0] if pr is None else pr[1] - pr[0]
            total += rows * (v.numel() // v.shape[0]) * v.element_size()
        return int(total)

    # ---- pipelined form: the NEXT step's inputs cross PCIe while this step computes ---------------------------------
    def _slot(self, k: int):
        """Staging set k (0 / 1) of the pipelined host-fed step: device copies of the inputs, plus — sharded — the
        buffers of the row-sharded patristic copy."""
        if self._slots is None:
            self._slots = [None, None]
            self._slot_ready = [torch.cuda.Event(), torch.cuda.Event()]
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
        if self._slots[k] is None:
            st = {name: torch.empty_like(v) for name, v in self._staging.items()}
            for name, v in self._staging.items():        # rows this rank never copies stay valid
                st[name].copy_(v)
            pr = self._patristic_rows
            if pr is not None:
                gathered = torch.zeros_like(pr[3])
                gathered[: st["patristic"].shape[0]].copy_(self._staging["patristic"])
                st["patristic"] = gathered[: st["patristic"].shape[0]]
                st["_gathered"], st["_mine"] = gathered, torch.zeros_like(pr[4])
            self._slots[k] = st
        return self._slots[k]

    def _enqueue_copy(self, host, k: int):
        """Copy the host inputs into staging set k on the copy stream; `_slot_ready[k]` fires when they have landed.  The
        set's previous reader has finished: every step ends with the host reading its loss."""
        st, cs = self._slot(k), self._copy_stream
        lo, hi = self._leaf_rows()
        cs.wait_stream(torch.cuda.current_stream())       # idle between steps; orders the set's initialisation before the copy
        with torch.cuda.stream(cs):
            pr = self._patristic_rows
            if pr is None:
                st["patristic"].copy_(host["patristic"], non_blocking=True)
            else:
                import torch.distributed as dist
                r_lo, r_hi = pr[0], pr[1]
                if r_hi > r_lo:
                    st["_mine"][: r_hi - r_lo].copy_(host["patristic"][r_lo:r_hi], non_blocking=True)
                dist.all_gather_into_tensor(st["_gathered"], st["_mine"], group=self._prefetch_group)
            if "blosum" in host:
                st["blosum"][lo:hi].copy_(host["blosum"][lo:hi], non_blocking=True)
            st["dataset"][lo:hi].copy_(host["dataset"][lo:hi], non_blocking=True)
            self._slot_ready[k].record(cs)

    def _pipelined_step(self, host, prefetch) -> float:
        pend, self._pending = self._pending, None
        if pend is not None and pend[1] is host:
            k = pend[0]                                   # already on its way (or landed): issued during the previous step
        else:
            k = 0 if pend is None else 1 - pend[0]
            self._enqueue_copy(host, k)
        if prefetch is not None:                          # enqueued BEFORE this step's launch: the two overlap
            self._enqueue_copy(prefetch, 1 - k)
            self._pending = (1 - k, prefetch)
        torch.cuda.current_stream().wait_event(self._slot_ready[k])
        st = self._slots[k]
        return self.svi.step(st["dataset"], st["patristic"], None, st.get("blosum"), None, None)

    def step_from_host(self, host, prefetch=None) -> float:
        """H2D of the step inputs (pinned, asynchronous) -> svi.step -> loss back to the host as a float.

        `prefetch`: the host inputs of the NEXT call (the next batch of the caller's loader; may be `host` itself).  Their
        copy is enqueued on the copy stream before this step is launched and lands in the other of two staging sets while
        this step computes — the loop S/train.py:73-75 with a pinned, double-buffered loader; the next call must then pass
        that same object as `host` and must not have modified it in between.  Each staging set gets its own captured graph
        (no copies inside); every step still moves its full inputs across PCIe, one step early.


        The copies run on a copy stream and every consumer waits for its own input only (event waits on the stream that reads
        it, no host synchronisation): the patristic matrix is awaited by the stream that assembles the covariances (the
        prior's side stream when the guide starts the factorisation, so the main stream never waits for it), the
        BLOSUM-encoded data set in front of the encoder, the data set itself when the likelihood is scored, after both
        networks ran.  Of the first two the SMALLER goes first: on one GPU that is the matrix (the factorisation then has
        the machine to itself while the 168 MB BLOSUM tensor crosses PCIe), on a leaf-sharded rank it is the rank's slice of
        the BLOSUM tensor (the recurrences are the critical path there and must not queue behind 32 MB of matrix)."""
        if prefetch is not None or getattr(self, "_pending", None) is not None:
            return self._pipelined_step(host, prefetch)
        if getattr(self, "_events", None) is None:
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            self._events = {k: torch.cuda.Event() for k in ("patristic", "blosum", "dataset")}
            self._armed = False
            real_model, real_guide = self.svi.model, self.svi.guide

            def wait_for(*keys):
                if self._armed:
                    for k in keys:
                        torch.cuda.current_stream().wait_event(self._events[k])

            def guide_after_inputs(*a, **k):
                wait_for(*self._guide_waits)
                return real_guide(*a, **k)

            def model_after_inputs(*a, **k):
                wait_for(*self._model_waits)
                return real_model(*a, **k)

            self._wait_for = wait_for
            self.svi.model, self.svi.guide = model_after_inputs, guide_after_inputs
            if hasattr(self.guide, "_encode"):                         # DRAUPNIRGUIDES: waits for the BLOSUM copy itself
                self.guide.__dict__["_before_encode"] = lambda: wait_for("blosum")
                self._guide_waits = ()
            elif hasattr(self.guide, "prototype_trace"):               # auto-guides read no data
                self._guide_waits = ()
            else:
                self._guide_waits = ("patristic", "blosum")
            hooked_model = hasattr(self.Draupnir, "_before_prior") and hasattr(self.svi.loss, "pre_score_hook")
            if hooked_model:
                self.Draupnir._before_prior = lambda: wait_for("patristic")
                self.svi.loss.pre_score_hook = lambda: wait_for("patristic", "blosum", "dataset")   # joins the copy stream
                self._model_waits = ()
            else:                                   # an ELBO without the hook (real Pyro): the model waits for everything
                self._model_waits = ("patristic", "blosum", "dataset")
            self._host_prologues = {}
        key = tuple(id(v) for v in host.values())
        prologue = self._host_prologues.get(key)
        if prologue is None:
            lo, hi = self._leaf_rows()
            blosum_bytes = (hi - lo) * (host["blosum"].numel() // host["blosum"].shape[0]) * 8 if "blosum" in host else 0
            pr = self._patristic_rows
            matrix_bytes = host["patristic"].numel() * 8 if pr is None else (pr[1] - pr[0]) * host["patristic"].shape[1] * 8
            matrix_first = "blosum" not in host or matrix_bytes <= blosum_bytes

            def prologue():
                # runs at the start of the step — eagerly, or ONCE inside the CUDA-graph capture of the step, where the
                # copies become memcpy nodes from the pinned buffers and the event waits become graph dependencies
                main, cs = torch.cuda.current_stream(), self._copy_stream
                cs.wait_stream(main)               # the previous step has finished reading the staging buffers
                with torch.cuda.stream(cs):
                    def matrix():
                        pr = self._patristic_rows
                        if pr is None:
                            self._staging["patristic"].copy_(host["patristic"], non_blocking=True)
                        else:                       # this rank's rows over PCIe, the others' over NVLink
                            import torch.distributed as dist
                            r_lo, r_hi, per, gathered, mine = pr
                            if r_hi > r_lo:
                                mine[: r_hi - r_lo].copy_(host["patristic"][r_lo:r_hi], non_blocking=True)
                            dist.all_gather_into_tensor(gathered, mine, group=self.Draupnir.shard.group)
                        self._events["patristic"].record(cs)

                    def blosum():
                        if "blosum" in host:
                            self._staging["blosum"][lo:hi].copy_(host["blosum"][lo:hi], non_blocking=True)
                        self._events["blosum"].record(cs)

                    for fn in ((matrix, blosum) if matrix_first else (blosum, matrix)):
                        fn()
                    self._staging["dataset"][lo:hi].copy_(host["dataset"][lo:hi], non_blocking=True)
                    self._events["dataset"].record(cs)
                self._armed = True

            self._host_prologues[key] = prologue
        st = self._staging
        self.svi.step_prologue = prologue
        try:
            loss = self.svi.step(st["dataset"], st["patristic"], None, st.get("blosum"), None, None)
        finally:
            self.svi.step_prologue = None
            self._armed = False
        return loss


def build_run(problem: SyntheticProblem, select_guide: str, device="cuda", init: Optional[Dict[str, torch.Tensor]] = None,
              z_dim=30, gru_hidden_dim=60, embedding_dim=50, config: Optional[dict] = None, seed: int = 0,
              shard=None, cuda_graph: Optional[bool] = None) -> Run:
    """Mirror of `draupnir_train`'s setup (S/main.py:991-1059): ModelLoad -> DRAUPNIRModel_classic -> guide
    (`AutoDelta` for delta_map, `DRAUPNIRGUIDES` otherwise) -> Trace_ELBO -> ClippedAdam -> SVI."""
    assert select_guide in ("delta_map", "variational", "normal", "diagonal_normal")          # S/main.py:511-524
    cfg = dict(DEFAULT_CONFIG, **(config or {}))
    pyro.clear_param_store()
    pyro.enable_validation(False)
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        ml = model_load_from_problem(problem, select_guide, device, z_dim, gru_hidden_dim, embedding_dim)
        Draupnir = DRAUPNIRModel_classic(ml)
        Draupnir.shard = shard
        if select_guide == "delta_map":
            values = None if init is None else {k: v.to(ml.device) for k, v in init.items()}
            guide = AutoDelta(Draupnir.model, init_loc_fn=init_to_value(values=values))
        elif select_guide in ("normal", "diagonal_normal"):
            cls = AutoNormal if select_guide == "normal" else AutoDiagonalNormal
            if init is None:
                guide = cls(Draupnir.model)
            else:
                guide = cls(Draupnir.model, init_loc_fn=init_to_value(values={k: v.to(ml.device) for k, v in init.items()}))
        else:
            guide = DRAUPNIRGUIDES(Draupnir.model, ml, Draupnir)
            if init is not None:
                with torch.no_grad():
                    for k in ("alpha", "sigma_f", "sigma_n", "lambd"):
                        getattr(guide, f"{k}_unconstrained").copy_(init[k].reshape(-1, 1).log().to(ml.device))
    finally:
        torch.random.set_rng_state(gen_state)
    optim = ClippedAdam({"lr": cfg["lr"], "betas": (cfg["beta1"], cfg["beta2"]), "eps": cfg["eps"],
                         "weight_decay": cfg["weight_decay"], "clip_norm": cfg["clip_norm"], "lrd": cfg["lrd"]})
    svi = SVI(Draupnir.model, guide, optim, Trace_ELBO(), shard=shard, cuda_graph=cuda_graph)
    dev = ml.device
    return Run(Draupnir=Draupnir, guide=guide, svi=svi, optim=optim, model_load=ml, device=dev,
               dataset_train=problem.dataset_train.to(dev), patristic_matrix_train=problem.patristic_matrix_train.to(dev),
               patristic_matrix_full=problem.patristic_matrix_full.to(dev),
               dataset_train_blosum=problem.dataset_train_blosum.to(dev), problem=problem, select_guide=select_guide)


def load_networks(run: Run, state: Dict[str, Dict[str, torch.Tensor]]):
    """Copy identically named network weights (decoder / embed / encoder / embeddingencoder / h_0_*) into a run —
    parity tests inject one set of weights into both the oracle and this implementation."""
    with torch.no_grad():
        run.Draupnir.decoder.load_state_dict(state["decoder"])
        run.Draupnir.embed.load_state_dict(state["embed"])
        run.Draupnir.h_0_MODEL.copy_(state["h_0_MODEL"])
        if run.select_guide == "variational":
            run.guide.encoder.load_state_dict(state["encoder"])
            run.guide.embeddingencoder.load_state_dict(state["embeddingencoder"])
            run.guide.h_0_GUIDE.copy_(state["h_0_GUIDE"])
