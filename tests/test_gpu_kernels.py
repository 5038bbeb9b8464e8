"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against the CPU oracle.

Tolerances (fp64, the reference dtype): LCA indices bit-exact; distances rtol 1e-12; covariance / Cholesky /
log-prob rtol 1e-9 or better; analytic gradients vs autograd of the oracle rtol 1e-7.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from draupnir_asr_b200 import synthetic as S  # noqa: E402
from oracle import draupnir_oracle as O  # noqa: E402
from oracle import patristic_oracle as P  # noqa: E402


def _spd_problem(n, z, seed=0):
    t = S.random_tree(n, seed=seed)
    D = torch.from_numpy(S.patristic_host(t)[np.ix_(t.leaves, t.leaves)])
    g = torch.Generator().manual_seed(seed)
    sf = torch.rand(z, generator=g, dtype=torch.float64) + 0.5
    sn = torch.rand(z, generator=g, dtype=torch.float64) + 0.5
    lam = torch.rand(z, generator=g, dtype=torch.float64) + 0.5
    x = torch.randn(z, n, generator=g, dtype=torch.float64)
    return D, sf, sn, lam, x


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


# ---------------------------------------------------------------- K1 ---------------------------------------------
@pytest.mark.parametrize("n,shape", [(2, "agglomerate"), (19, "agglomerate"), (200, "agglomerate"), (64, "caterpillar"),
                                     (700, "caterpillar"), (128, "balanced"), (1000, "agglomerate")])
def test_patristic_lca_bit_exact(cuda, n, shape):
    from draupnir_asr_b200 import ops
    t = S.random_tree(n, seed=n, shape=shape)
    pat, cla, lca = ops.tree_distances(torch.from_numpy(t.parent).to(cuda), torch.from_numpy(t.branch).to(cuda),
                                       want_lca=True, want_cladistic=True)
    lca_ref, pat_ref, cla_ref = P.patristic_c(t.parent, t.branch, t.depth, want_cladistic=True)
    assert np.array_equal(lca.cpu().numpy(), lca_ref)                      # bit-exact indices
    assert np.array_equal(cla.cpu().numpy(), cla_ref)                      # integer edge counts
    np.testing.assert_allclose(pat.cpu().numpy(), pat_ref, rtol=1e-12, atol=1e-13)
    p = pat.cpu().numpy()
    assert np.array_equal(p, p.T) and np.all(np.diag(p) == 0.0)            # exactly symmetric, zero diagonal


def test_patristic_rows_cols_subset(cuda):
    from draupnir_asr_b200 import ops
    t = S.random_tree(300, seed=3)
    rows = torch.from_numpy(t.leaves[:50].astype(np.int32))
    cols = torch.from_numpy(t.internal[:77].astype(np.int32))
    pat, _, lca = ops.tree_distances(torch.from_numpy(t.parent).to(cuda), torch.from_numpy(t.branch).to(cuda), rows=rows,
                                     cols=cols, want_lca=True)
    lca_ref, pat_ref, _ = P.patristic_c(t.parent, t.branch, t.depth, rows.numpy(), cols.numpy())
    assert np.array_equal(lca.cpu().numpy(), lca_ref)
    np.testing.assert_allclose(pat.cpu().numpy(), pat_ref, rtol=1e-12, atol=1e-13)


def _sampled_lca_check(cuda, n, seed, nrows, ncols, shape="agglomerate"):
    from draupnir_asr_b200 import ops
    t = S.random_tree(n, seed=seed, shape=shape)
    N = t.n_nodes
    rng = np.random.default_rng(seed)
    rows = np.sort(rng.choice(N, nrows, replace=False)).astype(np.int32)
    cols = np.sort(rng.choice(N, ncols, replace=False)).astype(np.int32)
    parent, branch = torch.from_numpy(t.parent).to(cuda), torch.from_numpy(t.branch).to(cuda)
    pat, cla, lca = ops.tree_distances(parent, branch, rows=torch.from_numpy(rows), cols=torch.from_numpy(cols), want_lca=True,
                                       want_cladistic=True)
    lca_ref, pat_ref, cla_ref = P.patristic_c(t.parent, t.branch, t.depth, rows, cols, want_cladistic=True)
    assert np.array_equal(lca.cpu().numpy(), lca_ref)
    assert np.array_equal(cla.cpu().numpy(), cla_ref)
    np.testing.assert_allclose(pat.cpu().numpy(), pat_ref, rtol=1e-12, atol=1e-13)
    return t, parent, branch


def test_patristic_lca_bit_exact_c4_and_c5_sizes(cuda):
    """BASELINE configs 4 and 5: ALL pairs of the 3,999-node tree of C4, and at the top of the C5 sweep (8,000 leaves,
    15,999 nodes) 96 complete rows plus a 512 x 512 random sample, bit-exact against the scalar C oracle."""
    from draupnir_asr_b200 import ops
    t = S.random_tree(2000, seed=2000)
    parent, branch = torch.from_numpy(t.parent).to(cuda), torch.from_numpy(t.branch).to(cuda)
    pat, _, lca = ops.tree_distances(parent, branch, want_lca=True)
    lca_ref, pat_ref, _ = P.patristic_c(t.parent, t.branch, t.depth)
    assert lca.shape == (3999, 3999) and np.array_equal(lca.cpu().numpy(), lca_ref)
    np.testing.assert_allclose(pat.cpu().numpy(), pat_ref, rtol=1e-12, atol=1e-13)
    del pat, lca
    t, parent, branch = _sampled_lca_check(cuda, 8000, 8000, 512, 512)
    rows = np.sort(np.random.default_rng(1).choice(t.n_nodes, 96, replace=False)).astype(np.int32)
    pat, _, lca = ops.tree_distances(parent, branch, rows=torch.from_numpy(rows), want_lca=True)
    lca_ref, pat_ref, _ = P.patristic_c(t.parent, t.branch, t.depth, rows, np.arange(t.n_nodes, dtype=np.int32))
    assert lca.shape == (96, 15999) and np.array_equal(lca.cpu().numpy(), lca_ref)
    np.testing.assert_allclose(pat.cpu().numpy(), pat_ref, rtol=1e-12, atol=1e-13)
    # the all-pairs matrix at that size is symmetric with a zero diagonal (2 GB; checked on the device)
    full = ops.tree_distances(parent, branch)[0]
    assert full.shape == (15999, 15999) and bool((full.diagonal() == 0).all())
    assert torch.equal(full[:4000, 12000:], full[12000:, :4000].T)
    # the all-pairs call (no row / column lists) against the 96-row call above: same rows, bit for bit, and the LCA ids of the
    # all-pairs call against the C oracle
    rows_t = torch.from_numpy(rows.astype(np.int64)).to(cuda)
    assert torch.equal(full[rows_t], pat)
    del full, pat
    lca_full = ops.tree_distances(parent, branch, want_lca=True, want_patristic=False)[2]
    assert np.array_equal(lca_full[rows_t].cpu().numpy(), lca_ref)


@pytest.mark.parametrize("n,shape", [(35000, "agglomerate"), (45000, "agglomerate"), (1400, "caterpillar")])
def test_patristic_beyond_the_table_limits(cuda, n, shape):
    """Nothing on this path is limited to 65,535 nodes: the preorder lookup table stores CHAIN indices (uint16, bounded by
    the depth, which is checked), not node ids.  69,999 nodes still use the table (one row per CTA), 89,999 nodes exceed
    its shared memory and take the binary-search kernel, a 1,399-deep caterpillar exceeds the shared-memory chain and
    spills it to the workspace - all bit-exact on a random sample of pairs."""
    _sampled_lca_check(cuda, n, n, 192, 256, shape=shape)


# ---------------------------------------------------------------- K2 ---------------------------------------------
@pytest.mark.parametrize("n,z", [(1, 1), (19, 30), (257, 7), (600, 30)])
def test_ou_cov_forward_backward(cuda, n, z):
    from draupnir_asr_b200 import ops
    D, sf, sn, lam, _ = _spd_problem(max(n, 2), z)
    D = D[:n, :n].contiguous()
    ps = [v.clone().requires_grad_(True) for v in (sf, sn, lam)]
    K_ref = O.ou_covariance(D, *ps)
    W = torch.randn(z, n, n, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    (K_ref * W).sum().backward()
    pg = [v.clone().to(cuda).requires_grad_(True) for v in (sf, sn, lam)]
    # strided view with a header row/column, as the model passes it
    Dh = torch.zeros(n + 1, n + 1, dtype=torch.float64)
    Dh[1:, 1:] = D
    K = ops.ou_cov(Dh.to(cuda)[1:, 1:], *pg)
    assert rel(K.detach().cpu(), K_ref.detach()) < 1e-14
    (K * W.to(cuda)).sum().backward()
    for a, b in zip(pg, ps):
        assert rel(a.grad.cpu(), b.grad) < 1e-10


# ---------------------------------------------------------------- K3 ---------------------------------------------
@pytest.mark.parametrize("n,z", [(1, 2), (5, 3), (64, 4), (65, 3), (200, 30), (513, 5)])
def test_potrf_matches_torch(cuda, n, z):
    from draupnir_asr_b200 import ops
    D, sf, sn, lam, _ = _spd_problem(max(n, 2), z, seed=n)
    K = O.ou_covariance(D[:n, :n], sf, sn, lam)
    L_ref = torch.linalg.cholesky(K)
    L = ops.potrf(K.to(cuda)).cpu()
    assert rel(L, L_ref) < 1e-11
    assert torch.equal(L, torch.tril(L))


def test_potrf_raises_linalg_error(cuda):
    from draupnir_asr_b200 import ops
    K = torch.eye(40, dtype=torch.float64).repeat(3, 1, 1)
    K[1, 17, 17] = -1.0
    with pytest.raises(torch.linalg.LinAlgError, match="Batch element 1"):
        ops.potrf(K.to(cuda))


def test_potri(cuda):
    from draupnir_asr_b200 import ops
    D, sf, sn, lam, _ = _spd_problem(150, 6)
    K = O.ou_covariance(D, sf, sn, lam)
    Kinv = ops.potri(K.to(cuda)).cpu()
    assert rel(Kinv, torch.linalg.inv(K)) < 1e-9


# ------------------------------------------------------------ K3+K4+K5 -------------------------------------------
@pytest.mark.parametrize("n,z", [(2, 1), (19, 30), (64, 5), (130, 30), (300, 8)])
def test_ou_prior_logp_and_gradients(cuda, n, z):
    from draupnir_asr_b200 import ops
    D, sf, sn, lam, x = _spd_problem(n, z, seed=n + 1)
    ps = [v.clone().requires_grad_(True) for v in (sf, sn, lam, x)]
    K = O.ou_covariance(D, *ps[:3])
    lp_ref = torch.distributions.MultivariateNormal(torch.zeros(1, n, dtype=torch.float64), K).log_prob(ps[3])
    w = torch.linspace(0.5, 1.5, z, dtype=torch.float64)
    (lp_ref * w).sum().backward()
    pg = [v.clone().to(cuda).requires_grad_(True) for v in (sf, sn, lam, x)]
    lp = ops.ou_prior_logp(D.to(cuda), *pg)
    assert rel(lp.detach().cpu(), lp_ref.detach()) < 1e-11
    (lp * w.to(cuda)).sum().backward()
    for a, b, name in zip(pg, ps, ("sigma_f", "sigma_n", "lambd", "latent_z")):
        assert rel(a.grad.cpu(), b.grad) < 1e-7, name
    with torch.no_grad():                                                   # forward-only path (no gradient block)
        lp2 = ops.ou_prior_logp(D.to(cuda), *[p.detach() for p in pg])
    assert rel(lp2.cpu(), lp_ref.detach()) < 1e-11


@pytest.mark.parametrize("n,z", [(5, 3), (64, 2), (130, 4), (450, 3)])
def test_ou_prior_two_phase(cuda, n, z):
    """Factor first (hyper-parameters only, possibly on another stream), complete with the latent vectors later: same
    log-densities and gradients as the oracle, and a factor is handed out only for the tensors it was made from."""
    from draupnir_asr_b200 import ops
    D, sf, sn, lam, x = _spd_problem(n, z, seed=n + 2)
    ps = [v.clone().requires_grad_(True) for v in (sf, sn, lam, x)]
    K = O.ou_covariance(D, *ps[:3])
    lp_ref = torch.distributions.MultivariateNormal(torch.zeros(1, n, dtype=torch.float64), K).log_prob(ps[3])
    w = torch.linspace(0.5, 1.5, z, dtype=torch.float64)
    (lp_ref * w).sum().backward()
    Dg = D.to(cuda)
    pg = [v.clone().to(cuda).requires_grad_(True) for v in (sf, sn, lam, x)]
    side = torch.cuda.Stream()
    f = ops.ou_prior_factor(Dg, *pg[:3], key=tuple(pg[:3]), stream=side)
    assert f.matches(Dg, tuple(pg[:3])) and not f.matches(Dg, (pg[0], pg[1], pg[2].clone()))
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        lp = ops.ou_prior_logp(Dg, *pg, factor=f)
    cur.wait_stream(side)
    assert f.used and not f.matches(Dg, tuple(pg[:3]))
    assert rel(lp.detach().cpu(), lp_ref.detach()) < 1e-10
    (lp * w.to(cuda)).sum().backward()
    torch.cuda.synchronize()
    for a, b, name in zip(pg, ps, ("sigma_f", "sigma_n", "lambd", "latent_z")):
        assert rel(a.grad.cpu(), b.grad) < 1e-7, name


def test_ou_prior_through_the_tensor_path_at_1100_leaves(cuda):
    """value AND gradients (sigma_f, sigma_n, lambd, latent_z) of the fused prior at a size whose trailing
    updates run on tcgen05 (R = 2n+1 = 2201 rows, three-level blocking, col_limit, ragged last panel), one-call and
    two-phase form, against autograd through torch's fp64 Cholesky on the CPU.  Gates: 1e-4 (ELBO) / 1e-3 (gradients)."""
    from draupnir_asr_b200 import ops
    n, z = 1100, 2
    D, sf, sn, lam, x = _spd_problem(n, z, seed=11)
    ps = [v.clone().requires_grad_(True) for v in (sf, sn, lam, x)]
    lp_ref = torch.distributions.MultivariateNormal(torch.zeros(1, n, dtype=torch.float64),
                                                    O.ou_covariance(D, *ps[:3])).log_prob(ps[3])
    w = torch.tensor([0.7, 1.3], dtype=torch.float64)
    (lp_ref * w).sum().backward()
    Dg = D.to(cuda)
    for two_phase in (False, True):
        pg = [v.clone().to(cuda).requires_grad_(True) for v in (sf, sn, lam, x)]
        f = ops.ou_prior_factor(Dg, *pg[:3], key=tuple(pg[:3])) if two_phase else None
        lp = ops.ou_prior_logp(Dg, *pg, factor=f)
        (lp * w.to(cuda)).sum().backward()
        assert rel(lp.detach().cpu(), lp_ref.detach()) < 1e-9, two_phase
        for a, b, name in zip(pg, ps, ("sigma_f", "sigma_n", "lambd", "latent_z")):
            assert rel(a.grad.cpu(), b.grad) < 1e-6, (two_phase, name, rel(a.grad.cpu(), b.grad))


def test_mvn_logprob_explicit_cov(cuda):
    from draupnir_asr_b200 import ops
    n, z = 90, 4
    D, sf, sn, lam, x = _spd_problem(n, z)
    K = O.ou_covariance(D, sf, sn, lam).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    lp_ref = torch.distributions.MultivariateNormal(torch.zeros(n, dtype=torch.float64), K).log_prob(xr)
    lp_ref.sum().backward()
    Kg, xg = K.detach().to(cuda).requires_grad_(True), x.to(cuda).requires_grad_(True)
    lp = ops.mvn_logprob(Kg, xg)
    lp.sum().backward()
    assert rel(lp.detach().cpu(), lp_ref.detach()) < 1e-11
    assert rel(xg.grad.cpu(), xr.grad) < 1e-8
    gref = 0.5 * (K.grad + K.grad.transpose(-1, -2))
    assert rel(Kg.grad.cpu(), gref) < 1e-7


# ---------------------------------------------------------------- K6 ---------------------------------------------
@pytest.mark.parametrize("n", [3, 19, 120])
def test_conditional_matches_precision_partition(cuda, n):
    from draupnir_asr_b200 import ops
    pr = S.make_problem(n, 8, seed=n)
    par = S.make_parameters(n, z_dim=6, seed=n)
    mean_ref, cov_ref = O.conditional_moments(par, pr.patristic_matrix_full, pr.internal_nodes, jitter=1e-6)
    pm = pr.patristic_matrix_full.to(cuda)
    leaf = torch.from_numpy(pr.tree.leaves)
    inte = torch.from_numpy(pr.tree.internal)
    mean, cov = ops.ou_conditional(pm[1:, 1:], leaf, inte, par["sigma_f"].to(cuda), par["sigma_n"].to(cuda),
                                   par["lambd"].to(cuda), par["latent_z"].to(cuda), jitter=1e-6, want_cov=True)
    assert rel(mean.cpu(), mean_ref) < 1e-9
    assert rel(cov.cpu(), cov_ref) < 1e-9
    mean2, none = ops.ou_conditional(pm[1:, 1:], leaf, inte, par["sigma_f"].to(cuda), par["sigma_n"].to(cuda),
                                     par["lambd"].to(cuda), par["latent_z"].to(cuda), want_cov=False)
    assert none is None and rel(mean2.cpu(), mean_ref) < 1e-9


@pytest.mark.parametrize("n,z", [(1100, 2), (800, 3)])
def test_conditional_full_size_through_the_tensor_path(cuda, n, z):
    """Row a10 at the sizes where the elimination of the leaf block runs on tcgen05 with `col_limit` (N = 2,199 and the
    1,599 nodes of C3): conditional mean and covariance against the reference's precision-partition arithmetic
    (three dense fp64 inverses on the CPU)."""
    from draupnir_asr_b200 import ops
    pr = S.make_problem(n, 4, seed=n)
    par = S.make_parameters(n, z_dim=z, seed=n)
    mean_ref, cov_ref = O.conditional_moments(par, pr.patristic_matrix_full, pr.internal_nodes, jitter=1e-6)
    pm = pr.patristic_matrix_full.to(cuda)
    leaf, inte = torch.from_numpy(pr.tree.leaves), torch.from_numpy(pr.tree.internal)
    hp = [par[k].to(cuda) for k in ("sigma_f", "sigma_n", "lambd", "latent_z")]
    mean, cov = ops.ou_conditional(pm[1:, 1:], leaf, inte, *hp, jitter=1e-6, want_cov=True)
    assert rel(mean.cpu(), mean_ref) < 1e-8, rel(mean.cpu(), mean_ref)
    assert rel(cov.cpu(), cov_ref) < 1e-8, rel(cov.cpu(), cov_ref)
    mean2, _ = ops.ou_conditional(pm[1:, 1:], leaf, inte, *hp, want_cov=False)          # the MAP shape: col_limit = n + 1
    assert rel(mean2.cpu(), mean_ref) < 1e-8


# ---------------------------------------------------------------- K10 --------------------------------------------
@pytest.mark.parametrize("rows,A", [(1, 21), (255, 21), (256, 21), (257, 21), (4275, 21), (1000, 24)])
def test_cat_loglik(cuda, rows, A):
    from draupnir_asr_b200 import ops
    g = torch.Generator().manual_seed(rows)
    s = (3 * torch.randn(rows, A, generator=g, dtype=torch.float64)).requires_grad_(True)
    obs = torch.randint(0, A, (rows,), generator=g).to(torch.float64)
    ref = torch.distributions.Categorical(logits=torch.log_softmax(s, -1)).log_prob(obs).sum()
    (2.5 * ref).backward()
    sg = s.detach().to(cuda).requires_grad_(True)
    out = ops.cat_loglik(sg, obs.to(cuda))
    (2.5 * out).backward()
    assert abs(float(out) - float(ref)) / abs(float(ref)) < 1e-13
    assert rel(sg.grad.cpu(), s.grad) < 1e-12
    out_i = ops.cat_loglik(sg.detach(), obs.long().to(cuda))                # int64 observations
    assert abs(float(out_i) - float(ref)) / abs(float(ref)) < 1e-13


def test_ops_refuse_cpu_tensors():
    from draupnir_asr_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.ou_cov(torch.zeros(3, 3, dtype=torch.float64), torch.ones(2, dtype=torch.float64),
                   torch.ones(2, dtype=torch.float64), torch.ones(2, dtype=torch.float64))
