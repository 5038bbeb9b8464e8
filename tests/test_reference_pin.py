"""Pins the oracle to the REFERENCE'S OWN code.

tests/golden/ref_*.npz were produced by executing the reference's models.py / models_utils.py / guides.py / train.py /
utils.py (in place, from /root/reference) on the CPU — tests/golden/make_reference_golden.py, tests/reference_harness.py
say what was real and what was substituted.  Here:
  (a) the oracle must reproduce those vectors (runs everywhere, the fixtures are committed);
  (b) where the reference tree is present (the build container), the reference is executed live on further shapes and
      compared with the oracle call by call.
Tolerances: the oracle restates the same torch calls, so agreement is at rounding level (1e-12 relative; most values
are bit-identical with one thread).
"""
import os

import numpy as np
import pytest
import torch

import reference_harness as H
from draupnir_asr_b200 import synthetic as S
from oracle import draupnir_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"ref_c1_delta_map": (19, 225, "delta_map", 0), "ref_c1_variational": (19, 225, "variational", 0),
         "ref_n48_delta_map": (48, 64, "delta_map", 3), "ref_n48_variational": (48, 64, "variational", 3)}
RTOL = 1e-12
torch.set_default_dtype(torch.float64)


def _close(a, b, rtol=RTOL, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(1.0, float(np.abs(b).max())) if b.size else 1.0
    assert float(np.abs(a - b).max()) <= rtol * scale if b.size else True, (what, float(np.abs(a - b).max()), scale)


def _oracle_outputs(n, L, guide, seed, eps, cond_eps):
    """The same quantities make_reference_golden.run_case stores, computed by the oracle."""
    pr = S.make_problem(n, L, seed=seed)
    init = S.make_parameters(n, seed=seed)
    nets = O.Networks(variational=(guide == "variational"), seed=seed)
    svi = O.OracleSVI(nets, init, guide)
    args = (pr.dataset_train, pr.patristic_matrix_train, pr.dataset_train_blosum, pr.blosum_weighted, eps)
    loss, terms, _ = svi.loss(*args)
    out = {"loss0": float(loss.detach())}
    out.update({f"term_{k}": float(v) for k, v in terms.items()})
    loss.backward()
    for k, u in svi.u.items():
        out[f"grad_{k}"] = u.grad.detach().numpy().copy() if k == "latent_z" else u.grad.detach().reshape(-1).numpy().copy()
        u.grad = None
    groups = {"decoder": nets.decoder, "embed": nets.embed}
    if guide == "variational":
        groups.update(encoder=nets.encoder, embeddingencoder=nets.embeddingencoder)
    for k, mod in groups.items():
        out[f"gradnorm_{k}"] = float(torch.cat([p.grad.reshape(-1) for p in mod.parameters()]).norm())
    for p in nets.parameters():
        p.grad = None
    out["losses"] = np.array([svi.step(*args) for _ in range(3)])
    with torch.no_grad():
        vals = {k: v.detach() for k, v in svi.constrained().items()}
        if guide == "variational":
            res, _ = O.guide_batch(nets, vals, pr.dataset_train_blosum, eps)
            vals["latent_z"] = res["latent_z"]
        out["map_latent_z"] = vals["latent_z"].numpy().copy()
        mean = O.conditional_samplingMAP(vals, pr.patristic_matrix_full, pr.internal_nodes)
        out["cond_mean"] = mean.numpy().copy()
        out["cond_draw"] = O.conditional_sampling(vals, pr.patristic_matrix_full, pr.internal_nodes, cond_eps).numpy().copy()
        so = O.sample(nets, mean, pr.blosum_weighted)
        out["argmax_internal"] = so.aa_sequences.numpy().copy()
        out["logits_internal_lse0"] = float(torch.logsumexp(so.logits[..., 0].reshape(-1), 0))
        out["argmax_leaves"] = O.sample(nets, vals["latent_z"].T, pr.blosum_weighted).aa_sequences.numpy().copy()
    return out, pr


def _compare(got, want, guide):
    _close(got["loss0"], want["loss0"], what="loss0")
    for k in ("alpha", "sigma_f", "sigma_n", "lambd", "latent_z", "aa_sequences", "log_q"):
        _close(got[f"term_{k}"], want[f"term_{k}"], what=f"term_{k}")
    for k in ("alpha", "sigma_f", "sigma_n", "lambd") + (("latent_z",) if guide == "delta_map" else ()):
        g, w = np.asarray(got[f"grad_{k}"]), np.asarray(want[f"grad_{k}"])
        assert np.linalg.norm(g - w) <= 1e-10 * max(1.0, np.linalg.norm(w)), k
    for k in ("decoder", "embed") + (("encoder", "embeddingencoder") if guide == "variational" else ()):
        _close(got[f"gradnorm_{k}"], want[f"gradnorm_{k}"], rtol=1e-10, what=f"gradnorm_{k}")
    _close(got["losses"], want["losses"], what="losses")
    _close(got["map_latent_z"], want["map_latent_z"], rtol=1e-10, what="map_latent_z")
    _close(got["cond_mean"], want["cond_mean"], rtol=1e-9, what="cond_mean")          # three dense inverses: conditioning
    _close(got["cond_draw"], want["cond_draw"], rtol=1e-9, what="cond_draw")
    _close(got["logits_internal_lse0"], want["logits_internal_lse0"], rtol=1e-10, what="logits")
    assert np.array_equal(got["argmax_internal"], want["argmax_internal"])
    assert np.array_equal(got["argmax_leaves"], want["argmax_leaves"])


# ------------------------------------------------------------------- (a) oracle vs vectors of the reference's own code
@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_reproduces_reference_vectors(case):
    n, L, guide, seed = CASES[case]
    g = np.load(os.path.join(GOLDEN, f"{case}.npz"))
    got, pr = _oracle_outputs(n, L, guide, seed, torch.from_numpy(g["eps"]), torch.from_numpy(g["cond_eps"]))
    _compare(got, g, guide)
    want_keys = ["alpha", "lambd", "latent_z", "sigma_f", "sigma_n"]
    if guide == "variational":                                                          # S/guides.py:123-133
        want_keys = sorted(want_keys + ["z_loc", "z_scale", "rnn_final_bidirectional", "rnn_final_hidden_state",
                                        "rnn_hidden_states"])
    assert list(g["guide_keys"]) == want_keys
    # the synthetic generator's restatement of the input pipeline (S/utils.py:1752-1760,1820-1851)
    _close(pr.aa_frequencies.numpy(), g["aa_frequencies"], what="aa_frequencies")
    _close(pr.blosum_weighted.numpy(), g["blosum_weighted"], what="blosum_weighted")
    _close(pr.blosum_max.numpy(), g["blosum_max"], what="blosum_max")
    _close(pr.variable_score.numpy(), g["variable_score"], what="variable_score")
    _close(pr.dataset_train_blosum.sum(-1).numpy(), g["dataset_train_blosum_sum"], what="dataset_train_blosum")


def test_reference_vectors_equal_the_oracle_made_fixtures():
    """The older c1_*.npz fixtures were written by the oracle before the reference could be executed; the vectors the
    reference's code produced for the same case are identical."""
    for guide in ("delta_map", "variational"):
        a = np.load(os.path.join(GOLDEN, f"c1_{guide}.npz"))
        b = np.load(os.path.join(GOLDEN, f"ref_c1_{guide}.npz"))
        _close(a["loss0"], b["loss0"])
        _close(a["losses"], b["losses"])
        for k in ("alpha", "sigma_f", "sigma_n", "lambd", "latent_z", "aa_sequences", "log_q"):
            _close(a[f"term_{k}"], b[f"term_{k}"], what=k)
        if guide == "delta_map":          # (the older variational fixture conditions on the initial latent instead)
            _close(a["cond_mean"], b["cond_mean"], rtol=1e-9)
            assert np.array_equal(a["argmax_internal"], b["argmax_internal"])


# ------------------------------------------------------------------- (b) the reference executed live (build container)
needs_reference = pytest.mark.skipif(not H.reference_available(), reason="/root/reference is not on this machine")


@needs_reference
def test_reference_imports_and_accepts_the_same_state_dicts():
    ref = H.import_reference()
    assert ref.models.DRAUPNIRModel_classic.__module__ == "draupnir.models"
    assert os.path.realpath(ref.models.__file__).startswith("/root/reference/")
    pr = S.make_problem(7, 12, seed=2)
    nets = O.Networks(variational=True, seed=2)
    run = H.build_reference_run(pr, "variational", S.make_parameters(7, seed=2), H.nets_state_of(nets))   # strict loads
    assert sorted(run.Draupnir.decoder.state_dict()) == sorted(nets.decoder.state_dict())
    assert sorted(run.guide.encoder.state_dict()) == sorted(nets.encoder.state_dict())


@needs_reference
@pytest.mark.parametrize("n,L,guide,seed", [(12, 20, "delta_map", 7), (12, 20, "variational", 7), (31, 9, "variational", 8),
                                            (2, 5, "delta_map", 9)])
def test_reference_live_matches_oracle(n, L, guide, seed):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(GOLDEN, "make_reference_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    want = mk.run_case(n, L, guide, seed)
    got, _ = _oracle_outputs(n, L, guide, seed, torch.from_numpy(want["eps"]), torch.from_numpy(want["cond_eps"]))
    _compare(got, want, guide)


@needs_reference
def test_reference_ou_kernel_and_squeeze():
    ref = H.import_reference()
    torch.manual_seed(0)
    t = torch.rand(9, 9)
    t = t + t.T
    t.fill_diagonal_(0.0)
    sf, sn, lam = torch.rand(4) + 0.5, torch.rand(4) + 0.5, torch.rand(4) + 0.5
    want = ref.models_utils.OUKernel_Fast(sf, sn, lam).forward(t)
    assert torch.equal(O.ou_covariance(t, sf, sn, lam), want)
    # shapes the path produces ([3,1] / [z,1] PyroParams, [3] / [z] AutoDelta values) plus the docstring's example; the
    # reference's loop does not terminate for e.g. (1,3,1) -> 1 (stale index of the first unit dimension), not exercised
    for shape, nd in [((30, 1, 1, 40), 2), ((3, 1), 1), ((30, 1), 1), ((1, 1, 5), 1), ((4, 5), 2), ((3,), 1)]:
        x = torch.zeros(shape)
        assert O.squeeze_tensor(nd, x).shape == ref.utils.squeeze_tensor(nd, x).shape, shape


@needs_reference
@pytest.mark.parametrize("n", [2, 9, 40])
def test_reference_loader_reads_our_tree_files(tmp_path, n):
    """Row a1 / f3 layout: the patristic / cladistic CSVs and the level-order file written by `tree_io` go through the
    REFERENCE'S own loader steps (S/main.py:196-207 with pandas, then `symmetrize_and_clean`, `rename_axis`,
    `pandas_to_numpy` of S/utils.py:931-979 and the `matrix_sort` of S/load_utils.py:452-458 restated inline because it is
    a closure) and must give the header matrix `tree_io.load_distance_matrix` + `sort_by_node_id` produce — which is the
    matrix the oracle builds from the tree.  (The builder of the distances themselves, ete3 / R ape, cannot run here.)"""
    import pandas as pd
    from draupnir_asr_b200 import tree_io as T
    from oracle import patristic_oracle as P
    import test_tree_io as TT
    ref = H.import_reference()
    text, st = TT._random_newick(n, seed=100 + n)
    t = T.rename_tree_internal_nodes(T.read_newick(text))
    order = t.file_order()
    depth = TT._depth(t)
    _, pat, cla = P.patristic_c(t.parent, t.branch, depth, order, order, want_cladistic=True)
    base = os.path.join(tmp_path, "fam")
    labels = [t.names[i] for i in order]
    T._write_matrix_csv(base + "_patristic_distance_matrix.csv", labels, np.triu(pat))
    T._write_matrix_csv(base + "_cladistic_distance_matrix.csv", labels, cla)
    T.write_tree_levelorder_info(base + "_tree_levelorder_info.csv", t)
    # --- the reference's steps, S/main.py:196-207 ---
    patristic = pd.read_csv(base + "_patristic_distance_matrix.csv", low_memory=False, float_precision="round_trip")
    patristic = patristic.rename(columns={'Unnamed: 0': 'rows'})
    patristic.set_index('rows', inplace=True)
    cladistic = pd.read_csv(base + "_cladistic_distance_matrix.csv", index_col="rows", low_memory=False)
    info = pd.read_csv(base + "_tree_levelorder_info.csv", sep="\t", index_col=False, low_memory=False)
    info["0"] = info["0"].astype(str)
    info.drop('Unnamed: 0', inplace=True, axis=1)
    names = np.asarray(info["0"].tolist())
    assert names.tolist() == t.names
    _, full, full_cla = P.patristic_c(t.parent, t.branch, depth, want_cladistic=True)
    for frame, want, mine in ((patristic, full, base + "_patristic_distance_matrix.csv"),
                              (cladistic, full_cla, base + "_cladistic_distance_matrix.csv")):
        m = ref.utils.symmetrize_and_clean(frame, ancestral=True)
        m = ref.utils.rename_axis(m, names, name_file="fam")
        m = ref.utils.pandas_to_numpy(m)
        m = torch.from_numpy(np.asarray(m, dtype=np.float64))
        _, idx = torch.sort(m[:, 0])                                            # matrix_sort, S/load_utils.py:452-458
        m = m[idx][:, idx].numpy()
        got = T.sort_by_node_id(T.load_distance_matrix(mine, T.read_tree_levelorder_names(base + "_tree_levelorder_info.csv")))
        assert np.array_equal(m, got)
        assert np.isneginf(m[0, 0]) and np.array_equal(m[1:, 1:], want.astype(np.float64))
    # the reference reads with pandas' default float parser (not round-trip exact): same matrix to a few ulp
    default = pd.read_csv(base + "_patristic_distance_matrix.csv", low_memory=False).rename(columns={'Unnamed: 0': 'rows'})
    default.set_index('rows', inplace=True)
    assert np.allclose(default.to_numpy(dtype=np.float64), patristic.to_numpy(dtype=np.float64), rtol=1e-14, atol=0.0)
