"""Host logic of the Pyro-compatible runtime, on CPU with toy models (the DRAUPNIR classes themselves need the GPU)."""
import pytest
import torch

import draupnir_asr_b200.pyro_compat as pyro
from draupnir_asr_b200.pyro_compat import distributions as dist, poutine
from draupnir_asr_b200.pyro_compat.infer import SVI, Trace_ELBO
from draupnir_asr_b200.pyro_compat.infer.autoguide import AutoDelta, init_to_value
from draupnir_asr_b200.pyro_compat.nn import PyroModule, PyroParam
from draupnir_asr_b200.pyro_compat.optim import ClippedAdam
from oracle import draupnir_oracle as O

torch.set_default_dtype(torch.float64)


def _model(data):
    s = pyro.sample("s", dist.HalfNormal(torch.tensor(1.0)).expand_by([2]).to_event(1)) + 1e-6
    mu = pyro.sample("mu", dist.Normal(torch.zeros(()), s[0]))
    with pyro.plate("d", len(data)):
        pyro.sample("x", dist.Normal(mu, s[1]), obs=data)


def test_trace_replay_and_elbo_sum():
    pyro.clear_param_store()
    data = torch.tensor([0.3, 1.2, -0.4])
    init = {"s": torch.tensor([0.8, 1.1]), "mu": torch.tensor(0.25)}
    guide = AutoDelta(_model, init_loc_fn=init_to_value(values=init))
    loss = Trace_ELBO().differentiable_loss(_model, guide, data)
    want = (torch.distributions.HalfNormal(torch.tensor(1.0)).log_prob(init["s"]).sum()
            + torch.distributions.Normal(0.0, init["s"][0] + 1e-6).log_prob(init["mu"])
            + torch.distributions.Normal(init["mu"], init["s"][1] + 1e-6).log_prob(data).sum())
    assert torch.allclose(loss, -want, rtol=1e-14)
    # positive sites live in log space, real sites are stored as they are
    assert torch.allclose(guide.s_unconstrained.detach(), init["s"].log())
    assert torch.allclose(guide.mu_unconstrained.detach(), init["mu"])


def test_svi_step_equals_oracle_clipped_adam():
    pyro.clear_param_store()
    data = torch.tensor([0.3, 1.2, -0.4, 2.0])
    init = {"s": torch.tensor([0.8, 1.1]), "mu": torch.tensor(0.25)}
    guide = AutoDelta(_model, init_loc_fn=init_to_value(values=init))
    svi = SVI(_model, guide, ClippedAdam({"lr": 0.05, "clip_norm": 0.7, "lrd": 0.99}), Trace_ELBO())
    us, um = init["s"].log().clone().requires_grad_(True), init["mu"].clone().requires_grad_(True)
    ref = O.ClippedAdam([us, um], lr=0.05, clip_norm=0.7, lrd=0.99)
    for _ in range(4):
        got = svi.step(data)
        s = us.exp() + 1e-6
        want = -(torch.distributions.HalfNormal(torch.tensor(1.0)).log_prob(us.exp()).sum()
                 + torch.distributions.Normal(0.0, s[0]).log_prob(um)
                 + torch.distributions.Normal(um, s[1]).log_prob(data).sum())
        want.backward()
        ref.step()
        assert abs(got - float(want)) < 1e-12
    assert torch.allclose(guide.s_unconstrained.detach(), us.detach(), rtol=1e-12)
    assert torch.allclose(guide.mu_unconstrained.detach(), um.detach(), rtol=1e-12)


def test_sizeless_plate_is_a_noop_and_subsampling_is_refused():
    with pyro.plate("plate_batch", dim=-1):
        assert True
    with pytest.raises(NotImplementedError):
        pyro.plate("p", 10, subsample_size=3)


def test_delta_sites_score_zero_and_keep_shape():
    v = torch.tensor([[0.5], [1.5], [2.5]])
    d = dist.Delta(v).to_event(1)
    assert d.batch_shape == (3,) and d.event_shape == (1,)
    assert torch.equal(d.rsample(), v) and float(d.log_prob(v).sum()) == 0.0


def test_pyro_module_registers_parameters_and_svi_updates_them():
    pyro.clear_param_store()
    lin = torch.nn.Linear(2, 1)
    x, y = torch.randn(16, 2), torch.randn(16)

    def model(x, y):
        pyro.module("lin", lin)
        pyro.sample("y", dist.Normal(lin(x).squeeze(-1), torch.tensor(1.0)).to_event(1), obs=y)

    def guide(x, y):
        return {}

    before = lin.weight.detach().clone()
    SVI(model, guide, ClippedAdam({"lr": 0.1}), Trace_ELBO()).step(x, y)
    assert "lin$$$weight" in pyro.get_param_store() and not torch.equal(before, lin.weight.detach())
    assert lin.weight.grad is None                                    # gradients are zeroed after the step


def test_pyro_param_attribute_is_constrained_and_traced():
    class G(PyroModule):
        def __init__(self):
            super().__init__()
            self.alpha = PyroParam(torch.tensor([[0.5], [2.0]]), constraint=torch.distributions.constraints.positive)

    g = G()
    assert torch.allclose(g.alpha, torch.tensor([[0.5], [2.0]])) and "alpha_unconstrained" in dict(g.named_parameters())
    with poutine.trace() as tr:
        _ = g.alpha
    (key, site), = tr.trace.param_sites()
    assert site["unconstrained"] is g.alpha_unconstrained and key == "alpha@param"


def test_install_as_pyro_makes_reference_imports_work():
    assert pyro.install_as_pyro() in (True, False)
    import pyro as p2
    import pyro.distributions as d2
    from pyro.infer import SVI as S2, Trace_ELBO as T2  # noqa: F401
    from pyro.contrib.easyguide import EasyGuide  # noqa: F401
    assert hasattr(p2, "sample") and hasattr(d2, "HalfNormal")


def test_deferred_dict_behaves_like_a_dict_of_values():
    """utils.DeferredDict (the variational guide's return value, S/guides.py:123-133): deferred entries are evaluated once,
    on first access through any of the reading paths, and the dictionary pickles as a plain dict."""
    import pickle
    from draupnir_asr_b200.utils import Deferred, DeferredDict
    calls = []

    def make():
        return DeferredDict({"a": 1, "b": Deferred(lambda: (calls.append(1), 7)[1]), "c": 3})

    d = make()
    assert list(d) == ["a", "b", "c"] and len(d) == 3 and "b" in d and calls == []
    handed_on = DeferredDict({k: d.raw(k) for k in d})
    assert calls == [] and handed_on["b"] == 7 and d["b"] == 7 and calls == [1]          # shared, evaluated once
    assert d.get("b") == 7 and d.get("x", 9) == 9
    for read in (lambda m: dict(m), lambda m: {**m}, lambda m: dict(m.items()), lambda m: dict(zip(m, m.values())),
                 lambda m: pickle.loads(pickle.dumps(m)), lambda m: m.copy().materialize(), lambda m: {"b": m.pop("b"), **m}):
        got = read(make())
        assert got == {"a": 1, "b": 7, "c": 3}, got
        assert not any(isinstance(v, Deferred) for v in dict.values(got))
    assert type(pickle.loads(pickle.dumps(make()))) is dict


def test_guide_return_value_is_recorded_and_epoch_helpers():
    """The guide trace keeps the guide's return value (Pyro's `_RETURN` node): `epochs.run_epochs` takes `map_estimates`
    from the step it just ran instead of a second guide forward.  Plus the two statistics helpers on plain tensors."""
    from draupnir_asr_b200 import epochs
    from draupnir_asr_b200.backend import dist, pyro
    from draupnir_asr_b200.pyro_compat.infer import SVI, Trace_ELBO
    from draupnir_asr_b200.pyro_compat.optim import ClippedAdam
    pyro.clear_param_store()

    def model(x):
        z = pyro.sample("z", dist.Normal(torch.zeros(3, dtype=torch.float64), 1.0).to_event(1))
        pyro.sample("x", dist.Normal(z, 1.0).to_event(1), obs=x)

    def guide(x):
        loc = pyro.param("loc", torch.zeros(3, dtype=torch.float64))
        return {"latent": pyro.sample("z", dist.Normal(loc, 0.5).to_event(1))}

    svi = SVI(model, guide, ClippedAdam({"lr": 1e-2}), Trace_ELBO())
    x = torch.ones(3, dtype=torch.float64)
    svi.step(x)
    guide_trace = svi.loss.last_traces[1]
    assert guide_trace.return_value["latent"] is guide_trace.nodes["z"]["value"]
    assert epochs._guide_return(svi, guide, (x,)) is guide_trace.return_value
    n, L = 4, 6
    data = torch.zeros(n, L + 2, 30, dtype=torch.float64)
    data[:, 2:, 0] = torch.randint(0, 21, (n, L)).double()
    data[:, 0, 1] = torch.arange(n)
    pred = data[:, 2:, 0].clone()
    pred[:2, 0] += 1                                        # two sequences with one wrong site
    pid, std = epochs.calculate_percent_id(data, pred, L)
    per_seq = torch.tensor([100.0 * (L - 1) / L] * 2 + [100.0] * 2, dtype=torch.float64)
    assert pid == pytest.approx(float(per_seq.mean())) and std == pytest.approx(float(per_seq.std()))
    logits = torch.randn(n, L, 21, dtype=torch.float64)
    p = torch.exp(logits) / (1 + torch.exp(logits))
    ent = epochs.compute_sites_entropies(logits, data[:, 0, 1])
    assert ent.shape == (n, L + 1) and torch.equal(ent[:, 0], data[:, 0, 1])
    assert torch.allclose(ent[:, 1:], -(p * p.log()).sum(2), rtol=1e-12)


@pytest.mark.parametrize("kind", ["normal", "diagonal_normal"])
def test_mean_field_autoguides_elbo_is_the_explicit_formula(kind):
    """AutoNormal / AutoDiagonalNormal (S/main.py:517-518): u = loc + softplus(rho) eps in unconstrained space, value =
    bijection(u), log q(value) = log N(u; loc, scale) - log|d value / d u|; checked against the sum written out by hand."""
    from draupnir_asr_b200.pyro_compat.infer.autoguide import AutoDiagonalNormal, AutoNormal
    pyro.clear_param_store()
    data = torch.tensor([0.3, 1.2, -0.4])
    init = {"s": torch.tensor([0.8, 1.1]), "mu": torch.tensor(0.25)}
    cls = AutoNormal if kind == "normal" else AutoDiagonalNormal
    guide = cls(_model, init_loc_fn=init_to_value(values=init))
    guide(data)                                              # the prototype run draws from the prior: keep it off the seed
    torch.manual_seed(11)
    loss = Trace_ELBO().differentiable_loss(_model, guide, data)
    torch.manual_seed(11)
    eps_s, eps_mu = torch.randn(2), torch.randn(())          # the sites are drawn in model order, one normal per element
    if kind == "diagonal_normal":
        torch.manual_seed(11)
        e = torch.randn(3)
        eps_s, eps_mu = e[:2], e[2]
    scale = torch.tensor(0.1)
    u_s, u_mu = init["s"].log() + scale * eps_s, init["mu"] + scale * eps_mu
    s, mu = u_s.exp(), u_mu
    log_q = (torch.distributions.Normal(init["s"].log(), scale).log_prob(u_s).sum() - u_s.sum()      # log|ds/du| = u
             + torch.distributions.Normal(init["mu"], scale).log_prob(u_mu))
    log_p = (torch.distributions.HalfNormal(torch.tensor(1.0)).log_prob(s).sum()
             + torch.distributions.Normal(0.0, s[0] + 1e-6).log_prob(mu)
             + torch.distributions.Normal(mu, s[1] + 1e-6).log_prob(data).sum())
    assert torch.allclose(loss, -(log_p - log_q), rtol=1e-12)
    # parameters: locs start at the unconstrained initial values, scales at softplus^-1(0.1); both receive gradients
    loss.backward()
    params = dict(guide.named_parameters())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params.values())
    if kind == "normal":
        assert torch.allclose(params["locs.s"].detach(), init["s"].log())
        assert torch.allclose(torch.nn.functional.softplus(params["scales_unconstrained.mu"].detach()), scale)
    else:
        assert torch.allclose(params["loc"].detach(), torch.cat([init["s"].log(), init["mu"].reshape(1)]))
    med = guide.median(data)
    assert torch.allclose(med["s"], init["s"]) and torch.allclose(med["mu"], init["mu"])
    # a few SVI steps run and improve the bound on average
    svi = SVI(_model, guide, ClippedAdam({"lr": 0.05}), Trace_ELBO())
    torch.manual_seed(0)
    first = sum(svi.step(data) for _ in range(20)) / 20
    last = sum(svi.step(data) for _ in range(20)) / 20
    assert last < first


def test_site_log_prob_sums_equal_torch_and_delta_shortcut():
    """Round 2: the fused `log_prob_sum` entry points of HalfNormal / Normal fall back to torch on CPU tensors, the Delta
    shortcut (a guide site scored at its own tensor) equals the generic path, and the reference's [k,1] x [k] outer broadcast
    (guide value `[3,1]`, model scale `[3]`: S/guides.py:33-36 vs S/models.py:89) is what both compute."""
    v = torch.tensor([[0.4], [1.3], [0.2]])                                  # the variational guide's alpha: [3, 1]
    hn = dist.HalfNormal(torch.ones(())).expand_by([3]).to_event(1)
    want = torch.distributions.HalfNormal(torch.ones(3)).log_prob(v).sum()   # broadcasts to [3, 3]
    assert want.shape == () and torch.allclose(hn.log_prob_sum(v), want, rtol=1e-15)
    assert torch.allclose(hn.log_prob_sum(v), 3 * torch.distributions.HalfNormal(torch.ones(())).log_prob(v).sum(), rtol=1e-15)
    loc, scale = torch.randn(4, 5), torch.rand(4, 5) + 0.5
    x = torch.randn(4, 5)
    assert torch.allclose(dist.Normal(loc.T, scale.T).log_prob_sum(x.T), torch.distributions.Normal(loc, scale).log_prob(x).sum(),
                          rtol=1e-14)
    d = dist.Delta(v, event_dim=2)
    assert d.log_prob_sum(d.rsample()) == 0.0 and float(d.log_prob(v).sum()) == 0.0
    assert float(d.log_prob_sum(v + 1.0)) == float("-inf")                  # another value: the generic path
    dj = dist.Delta(torch.tensor([0.5, 0.7]), log_density=torch.tensor([-1.0, -2.0]), event_dim=0)
    assert float(dj.log_prob_sum(dj.v)) == -3.0 == float(dj.log_prob(dj.v).sum())
    assert float(dist.Delta(torch.zeros(4, 2), log_density=0.25, event_dim=1).log_prob_sum(torch.zeros(4, 2))) == 1.0  # other storage
    dd = dist.Delta(torch.zeros(4, 2), log_density=0.25, event_dim=1)
    assert dd.log_prob_sum(dd.v) == 1.0 == float(dd.log_prob(dd.v).sum())


def test_svi_without_cuda_arguments_stays_eager_and_checks_before_the_update():
    """CPU tensors never enter the CUDA-graph path; the step still evaluates, reduces and updates in the order
    evaluate -> (flags) -> optimiser, and leaves `.grad` cleared."""
    pyro.clear_param_store()
    data = torch.tensor([0.3, 1.2, -0.4])
    init = {"s": torch.tensor([0.8, 1.1]), "mu": torch.tensor(0.25)}
    guide = AutoDelta(_model, init_loc_fn=init_to_value(values=init))
    svi = SVI(_model, guide, ClippedAdam({"lr": 0.01}), Trace_ELBO(), cuda_graph=True)
    l0, l1, l2, l3 = (svi.step(data) for _ in range(4))
    assert not svi.last_step_was_replay and svi._graph_key((data,)) == (None, None, None) and not svi._graphs
    assert l3 < l0 and all(p.grad is None for p in guide.parameters())
    assert guide.prototype_trace is not None
    assert all(not (isinstance(s.get("value"), torch.Tensor) and s["value"].requires_grad)
               for _, s in guide.prototype_trace.sample_sites())           # taken under no_grad
