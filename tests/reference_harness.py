"""TEST INFRASTRUCTURE: runs the reference's OWN model / guide / sampling code (unmodified, read from /root/reference)
on the CPU, so that the oracle and the committed fixtures are pinned to the reference's arithmetic and not only to a
restatement of it.

What is real and what is substituted:
  * REAL (executed from /root/reference/draupnir/src/draupnir/, nothing copied): `models.py` (`DRAUPNIRModelClass`,
    `DRAUPNIRModel_classic`: `gp_prior`, `model_delta_map`, `model_variational`, `conditional_sampling`,
    `conditional_samplingMAP`, `map_sampling`, `sample`), `models_utils.py` (`OUKernel_Fast`, `RNNDecoder_Tiling`,
    `RNNEncoder`, `EmbedComplex`, `EmbedComplexEncoder`), `guides.py` (`DRAUPNIRGUIDES.guide_batch`), `train.py`
    (`train`), `utils.py` (`squeeze_tensor`, `create_blosum`, ...), on torch's own `nn.GRU`, `MultivariateNormal`,
    `Categorical`, `linalg.inv` — i.e. every floating-point operation of the path.
  * SUBSTITUTED: `pyro` is not installed in this image, so `import pyro` resolves to `draupnir_asr_b200.pyro_compat`
    (the messenger stack, `Trace_ELBO` sum, `AutoDelta`, `ClippedAdam` as restated in SURVEY §8c) — on CPU tensors its
    distributions are thin subclasses of `torch.distributions`.  Plot / tree / graph / structure libraries the
    reference imports at module scope but never calls on this path (`ignite`, `ete3`, `dgl`, `matplotlib`, `seaborn`,
    `Bio`) resolve to empty stand-ins.

`/root/reference` exists in the build container only: tests that use this module skip when it is absent, and the
vectors it produces are committed under tests/golden/ (`make_reference_golden.py`).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
import warnings

REFERENCE_SRC = "/root/reference/draupnir/src/draupnir"
_STUB_ROOTS = ("ignite", "ete3", "dgl", "matplotlib", "seaborn", "Bio", "umap", "prody", "gdown")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "models.py"))


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    __all__: list = []
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (_Anything,), {})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


_loaded = None


def import_reference():
    """Returns a namespace with the reference's modules: `.models`, `.models_utils`, `.guides`, `.train`, `.utils`."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise ImportError("the reference tree is not present on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import draupnir_asr_b200.pyro_compat as compat
    compat.install_as_pyro(force=True)       # always the shim: the fixtures must not depend on which pyro-ppl is installed
    for name in _STUB_ROOTS:
        try:                                     # only stand in for what is really missing
            importlib.import_module(name)
        except ImportError:
            if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
                sys.meta_path.insert(0, _StubFinder())
    pkg = types.ModuleType("draupnir")           # the package WITHOUT its __init__ (which imports the plotting stack)
    pkg.__path__ = [REFERENCE_SRC]
    sys.modules["draupnir"] = pkg
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ns = types.SimpleNamespace(
            models_utils=importlib.import_module("draupnir.models_utils"),
            utils=importlib.import_module("draupnir.utils"),
            models=importlib.import_module("draupnir.models"),
            guides=importlib.import_module("draupnir.guides"),
            train=importlib.import_module("draupnir.train"))
    _loaded = ns
    return ns


# ------------------------------------------------------------------------------------------------------------------
# one reference run wired like S/main.py:991-1059 (ModelLoad -> DRAUPNIRModel_classic -> guide -> SVI), on the CPU
# ------------------------------------------------------------------------------------------------------------------
def build_reference_run(problem, select_guide: str, init, nets_state, config=None):
    """`nets_state`: {"decoder","embed","h_0_MODEL"[, "encoder","embeddingencoder","h_0_GUIDE"]} as produced by
    `oracle.Networks`; loaded with strict state-dict matching, which also proves the parameter names are the
    reference's.  Returns a namespace (Draupnir, guide, svi, args)."""
    import torch
    import draupnir_asr_b200.pyro_compat as pyro
    from draupnir_asr_b200 import runtime as R
    from draupnir_asr_b200.pyro_compat.infer import SVI, Trace_ELBO
    from draupnir_asr_b200.pyro_compat.infer.autoguide import AutoDelta, init_to_value
    from draupnir_asr_b200.pyro_compat.optim import ClippedAdam

    ref = import_reference()
    cfg = dict(R.DEFAULT_CONFIG, **(config or {}))
    pyro.clear_param_store()
    pyro.enable_validation(False)
    ml = R.model_load_from_problem(problem, select_guide, "cpu")
    ml.args.use_cuda = False
    Draupnir = ref.models.DRAUPNIRModel_classic(ml)
    with torch.no_grad():
        Draupnir.decoder.load_state_dict(nets_state["decoder"], strict=True)
        Draupnir.embed.load_state_dict(nets_state["embed"], strict=True)
        Draupnir.h_0_MODEL.copy_(nets_state["h_0_MODEL"])
    if select_guide == "delta_map":
        guide = AutoDelta(Draupnir.model, init_loc_fn=init_to_value(values=dict(init)))       # S/main.py:516
    else:
        guide = ref.guides.DRAUPNIRGUIDES(Draupnir.model, ml, Draupnir)                       # S/main.py:1018-1020
        with torch.no_grad():
            guide.encoder.load_state_dict(nets_state["encoder"], strict=True)
            guide.embeddingencoder.load_state_dict(nets_state["embeddingencoder"], strict=True)
            guide.h_0_GUIDE.copy_(nets_state["h_0_GUIDE"])
            for k in ("alpha", "sigma_f", "sigma_n", "lambd"):
                getattr(guide, f"{k}_unconstrained").copy_(init[k].reshape(-1, 1).log())
    optim = ClippedAdam({"lr": cfg["lr"], "betas": (cfg["beta1"], cfg["beta2"]), "eps": cfg["eps"],
                         "weight_decay": cfg["weight_decay"], "clip_norm": cfg["clip_norm"], "lrd": cfg["lrd"]})
    svi = SVI(Draupnir.model, guide, optim, Trace_ELBO())                                     # S/main.py:1059
    args = (problem.dataset_train, problem.patristic_matrix_train, None, problem.dataset_train_blosum, None, None)
    train_input = {"patristic_matrix_model": problem.patristic_matrix_train, "cladistic_matrix_full": None,
                   "dataset_train_blosum": problem.dataset_train_blosum, "train_loader": [problem.dataset_train],
                   "map_estimates": None}                                                     # S/train.py:66-70
    return types.SimpleNamespace(ref=ref, Draupnir=Draupnir, guide=guide, svi=svi, args=args, train_input=train_input,
                                 model_load=ml)


def nets_state_of(nets):
    """State dicts of an `oracle.Networks` in the form `build_reference_run` and `runtime.load_networks` take."""
    st = {"decoder": nets.decoder.state_dict(), "embed": nets.embed.state_dict(), "h_0_MODEL": nets.h_0_MODEL.detach()}
    if getattr(nets, "encoder", None) is not None:
        st.update(encoder=nets.encoder.state_dict(), embeddingencoder=nets.embeddingencoder.state_dict(),
                  h_0_GUIDE=nets.h_0_GUIDE.detach())
    return st


class FixedNoise:
    """Makes `Normal.rsample` / `MultivariateNormal.rsample` return a given standard-normal tensor (both draw through
    `torch.distributions.utils._standard_normal`), so the reference and the oracle see the same draw."""

    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        import torch
        import torch.distributions.multivariate_normal as tm
        import torch.distributions.normal as tn
        self._mods = [(tn, tn._standard_normal), (tm, tm._standard_normal)]
        eps = self.eps

        def fixed(shape, dtype, device):
            if torch.Size(shape).numel() == eps.numel():
                return eps.to(device=device, dtype=dtype).reshape(shape)
            return self._mods[0][1](shape, dtype, device)

        tn._standard_normal = fixed
        tm._standard_normal = fixed
        return self

    def __exit__(self, *exc):
        for mod, orig in self._mods:
            mod._standard_normal = orig
        return False
