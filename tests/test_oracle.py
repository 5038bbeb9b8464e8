"""Cross-checks of the CPU oracle that need no reference (tests/test_reference_pin.py pins it to the reference's own
code; the reference ships no golden vectors for this path).

(i) closed forms on tiny trees, (ii) the precision-partition conditional (the reference's form) against an independent
Schur-complement derivation, (iii) parent-walk LCA vs brute-force path sums vs the C port, on random / caterpillar /
balanced trees incl. hypothesis-generated ones, (iv) the committed golden fixtures, (v) torch broadcasting quirks the
restatement relies on.
"""
import math
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from draupnir_asr_b200 import synthetic as S
from oracle import draupnir_oracle as O
from oracle import patristic_oracle as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
torch.set_default_dtype(torch.float64)


# ----------------------------------------------------------- (i) closed forms ------------------------------------
def test_ou_covariance_closed_form_two_leaves():
    t = torch.tensor([[0.0, 0.7], [0.7, 0.0]])
    sf, sn, lam = torch.tensor([1.3, 0.4]), torch.tensor([0.5, 0.9]), torch.tensor([2.0, 0.3])
    K = O.ou_covariance(t, sf, sn, lam)
    for z in range(2):
        off = sf[z] ** 2 * math.exp(-0.7 / lam[z])
        diag = sf[z] ** 2 + sn[z] ** 2
        assert torch.allclose(K[z], torch.tensor([[diag, off], [off, diag]]), rtol=1e-15)


def test_mvn_logprob_closed_form_two_leaves():
    t = torch.tensor([[0.0, 0.7], [0.7, 0.0]])
    sf, sn, lam = torch.tensor([1.3]), torch.tensor([0.5]), torch.tensor([2.0])
    x = torch.tensor([[0.3, -1.1]])
    a = float(sf[0] ** 2 + sn[0] ** 2)
    b = float(sf[0] ** 2 * math.exp(-0.7 / 2.0))
    det = a * a - b * b
    maha = (a * 0.3 ** 2 - 2 * b * 0.3 * -1.1 + a * 1.1 ** 2) / det
    want = -0.5 * (2 * math.log(2 * math.pi) + math.log(det) + maha)
    vals = {"alpha": torch.ones(3), "sigma_f": sf - 1e-6, "sigma_n": sn - 1e-6, "lambd": lam - 1e-6, "latent_z": x}
    pm = torch.from_numpy(S.with_id_header(t.numpy(), np.array([1, 2])))
    _, lp = O.gp_prior(vals, pm, 1)
    assert abs(float(lp["latent_z"]) - want) < 1e-13
    hn = math.log(2) - 0.5 * math.log(2 * math.pi)        # HalfNormal(1) at x: log 2 + log N(x; 0, 1)
    assert abs(float(lp["alpha"]) - 3 * (hn - 0.5)) < 1e-13


def test_alpha_broadcast_quirk_of_the_variational_guide():
    """S/guides.py:33 keeps alpha as a [3,1] PyroParam; scored against the [3]-shaped HalfNormal of S/models.py:89
    torch broadcasts to [3,3], i.e. the alpha log-prior is counted three times.  The restatement must inherit that."""
    v = torch.tensor([0.5, 1.0, 2.0])
    d = torch.distributions.Independent(torch.distributions.HalfNormal(torch.tensor(1.0)).expand([3]), 1)
    assert torch.allclose(d.log_prob(v.reshape(3, 1)).sum(), 3 * d.log_prob(v))


# ------------------------------------------------- (ii) conditional: precision partition vs Schur ----------------
@pytest.mark.parametrize("n", [2, 7, 40])
def test_conditional_precision_partition_equals_schur(n):
    pr = S.make_problem(n, 5, seed=n)
    par = S.make_parameters(n, z_dim=4, seed=n)
    mean, cov = O.conditional_moments(par, pr.patristic_matrix_full, pr.internal_nodes)
    K = O.ou_covariance(pr.patristic_matrix_full[1:, 1:], par["sigma_f"], par["sigma_n"], par["lambd"])
    a, b = torch.from_numpy(pr.tree.internal), torch.from_numpy(pr.tree.leaves)
    Kbb, Kab, Kaa = K[:, b][:, :, b], K[:, a][:, :, b], K[:, a][:, :, a]
    sol = torch.linalg.solve(Kbb, torch.cat((par["latent_z"][:, :, None], Kab.transpose(1, 2)), dim=2))
    assert torch.allclose(mean, (Kab @ sol[:, :, :1]).squeeze(-1), rtol=1e-9, atol=1e-11)
    assert torch.allclose(cov, Kaa - Kab @ sol[:, :, 1:], rtol=1e-8, atol=1e-10)
    eps = torch.randn(4, pr.n_internal, generator=torch.Generator().manual_seed(0))
    draw = O.conditional_sampling(par, pr.patristic_matrix_full, pr.internal_nodes, eps)
    assert draw.shape == (pr.n_internal, 4) and torch.isfinite(draw).all()
    assert torch.equal(O.conditional_samplingMAP(par, pr.patristic_matrix_full, pr.internal_nodes), mean.T)


# ----------------------------------------------------------- (iii) patristic -------------------------------------
def _brute_force(t):
    """Path sums via explicit ancestor sets — independent of both walks."""
    N = t.n_nodes
    anc = []
    for v in range(N):
        chain, u = [], v
        while u >= 0:
            chain.append(u)
            u = int(t.parent[u])
        anc.append(chain)
    lca = np.zeros((N, N), dtype=np.int32)
    for i in range(N):
        si = set(anc[i])
        for j in range(N):
            lca[i, j] = next(u for u in anc[j] if u in si)
    r = t.root_dist
    return lca, r[:, None] + r[None, :] - 2 * r[lca]


@pytest.mark.parametrize("n,shape", [(2, "agglomerate"), (3, "caterpillar"), (19, "agglomerate"), (33, "caterpillar"),
                                     (32, "balanced"), (60, "agglomerate")])
def test_lca_walk_vs_brute_force_vs_c(n, shape):
    t = S.random_tree(n, seed=n, shape=shape)
    lca_bf, pat_bf = _brute_force(t)
    lca_py, pat_py, cla_py = P.lca_parent_walk(t.parent, t.branch, t.depth, range(t.n_nodes), range(t.n_nodes))
    lca_c, pat_c, cla_c = P.patristic_c(t.parent, t.branch, t.depth, want_cladistic=True)
    assert np.array_equal(lca_py, lca_bf) and np.array_equal(lca_c, lca_bf)
    assert np.array_equal(pat_py, pat_c) and np.array_equal(cla_py, cla_c)
    np.testing.assert_allclose(pat_c, pat_bf, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(pat_c, S.patristic_host(t), rtol=1e-12, atol=1e-13)
    d = t.depth.astype(np.float64)
    assert np.array_equal(cla_c, d[:, None] + d[None, :] - 2 * d[lca_bf])


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 40), st.integers(0, 10_000), st.sampled_from(["agglomerate", "caterpillar", "balanced"]))
def test_lca_properties(n, seed, shape):
    t = S.random_tree(n, seed=seed, shape=shape)
    lca, pat, _ = P.patristic_c(t.parent, t.branch, t.depth)
    N = t.n_nodes
    assert np.array_equal(lca, lca.T) and np.array_equal(np.diag(lca), np.arange(N))
    assert np.array_equal(lca[0], np.zeros(N, dtype=np.int32))                  # the root is everyone's ancestor
    assert np.all(lca <= np.minimum.outer(np.arange(N), np.arange(N)))          # level order: ancestors have smaller ids
    assert np.allclose(pat, pat.T) and np.all(pat >= 0)
    k = np.arange(1, N)
    assert np.allclose(pat[k, t.parent[k]], t.branch[k])                        # child-parent distance = branch length


# ----------------------------------------------------------- (iv) golden fixtures --------------------------------
@pytest.mark.parametrize("guide", ["delta_map", "variational"])
def test_oracle_reproduces_golden_c1(guide):
    g = np.load(os.path.join(GOLDEN, f"c1_{guide}.npz"))
    pr = S.make_problem(19, 225, seed=0)
    init = S.make_parameters(19, seed=0)
    nets = O.Networks(variational=(guide == "variational"), seed=0)
    svi = O.OracleSVI(nets, init, guide)
    args = (pr.dataset_train, pr.patristic_matrix_train, pr.dataset_train_blosum, pr.blosum_weighted,
            torch.from_numpy(g["eps"]))
    loss, terms, _ = svi.loss(*args)
    assert abs(float(loss) - float(g["loss0"])) <= 1e-9 * abs(float(g["loss0"]))
    for k, v in terms.items():
        assert abs(float(v) - float(g[f"term_{k}"])) <= 1e-9 * max(1.0, abs(float(g[f"term_{k}"]))), k
    losses = [svi.step(*args) for _ in range(3)]
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-9)
    assert losses[2] < losses[0]                                                # the optimiser makes progress


def test_golden_tree_fixture():
    g = np.load(os.path.join(GOLDEN, "tree_64.npz"))
    lca, pat, cla = P.patristic_c(g["parent"], g["branch"], g["depth"], want_cladistic=True)
    assert np.array_equal(lca, g["lca"]) and np.array_equal(pat, g["patristic"]) and np.array_equal(cla, g["cladistic"])


# ----------------------------------------------------------- (v) optimiser ---------------------------------------
def test_clipped_adam_matches_torch_adam_when_unclipped():
    torch.manual_seed(0)
    w0 = torch.randn(7)
    a, b = w0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    mine = O.ClippedAdam([a], lr=1e-2, clip_norm=1e9, lrd=1.0)
    ref = torch.optim.Adam([b], lr=1e-2, eps=1e-8)
    for _ in range(5):
        for p in (a, b):
            (p ** 3).sum().backward()
        mine.step()
        ref.step()
        ref.zero_grad()
        # torch.Adam puts eps inside the bias-corrected denominator; ClippedAdam outside: equal up to O(eps)
        assert torch.allclose(a, b, rtol=1e-6)


def test_clipped_adam_clips_elementwise_and_decays_lr():
    a = torch.tensor([0.0, 0.0], requires_grad=True)
    opt = O.ClippedAdam([a], lr=0.1, clip_norm=1.0, lrd=0.5)
    a.grad = torch.tensor([100.0, 0.5])
    opt.step()
    # first Adam step moves by lr * g/|g| regardless of magnitude; lr was halved first
    assert torch.allclose(a.detach(), torch.tensor([-0.05, -0.05]), rtol=1e-6)
    assert opt.lr == 0.05
