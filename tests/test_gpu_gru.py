"""GPU parity of the bidirectional GRU recurrence kernels (csrc/gru.cu, through the C ABI) against torch's fp64 nn.GRU
on the host: hidden states and every gradient (input, h0, W_ih, W_hh, b_ih, b_hh).  Tolerance: fp64 re-association only
(1e-11 relative); the decoder's structured input path must equal the materialised-input path."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_and_ours(n, L, I, seed, cuda):
    from draupnir_asr_b200 import ops
    torch.manual_seed(seed)
    rnn = torch.nn.GRU(input_size=I, hidden_size=60, batch_first=True, bidirectional=True, num_layers=1).double()
    x = torch.randn(n, L, I, dtype=torch.float64, requires_grad=True)
    h0 = torch.randn(2, n, 60, dtype=torch.float64, requires_grad=True)
    w = torch.randn(n, L, 120, dtype=torch.float64)
    out_ref, hn_ref = rnn(x, h0)
    ((out_ref * w).sum() + hn_ref.sum()).backward()
    ref = {"out": out_ref.detach(), "hn": hn_ref.detach(), "x": x.grad.clone(), "h0": h0.grad.clone()}
    ref.update({k: p.grad.clone() for k, p in rnn.named_parameters()})
    rnn_g = torch.nn.GRU(input_size=I, hidden_size=60, batch_first=True, bidirectional=True, num_layers=1).double().to(cuda)
    rnn_g.load_state_dict(rnn.state_dict())
    xg = x.detach().to(cuda).requires_grad_(True)
    hg = h0.detach().to(cuda).requires_grad_(True)
    out, hn = ops.gru_module_forward(rnn_g, xg, hg)
    ((out * w.to(cuda)).sum() + hn.sum()).backward()
    got = {"out": out.detach().cpu(), "hn": hn.detach().cpu(), "x": xg.grad.cpu(), "h0": hg.grad.cpu()}
    got.update({k: p.grad.cpu() for k, p in rnn_g.named_parameters()})
    return ref, got


def _rows_per_cta(monkeypatch, rows):
    """DRP_GRU_ROWS overrides the one-wave heuristic (8 / 16 / 32 sequences per CTA, classic kernels); "pair" leaves the
    choice to the library, which runs this few rows on CTA pairs (gru_fwdc_kernel / gru_bwdc_kernel)."""
    if rows == "pair":
        monkeypatch.delenv("DRP_GRU_ROWS", raising=False)
    else:
        monkeypatch.setenv("DRP_GRU_ROWS", str(rows))


@pytest.mark.parametrize("rows", [8, 16, 32, "pair"])
@pytest.mark.parametrize("n,L,I", [(1, 1, 21), (5, 7, 51), (37, 23, 51), (64, 40, 21), (130, 33, 51)])
def test_gru_matches_torch(cuda, monkeypatch, n, L, I, rows):
    _rows_per_cta(monkeypatch, rows)
    ref, got = _ref_and_ours(n, L, I, seed=n + L, cuda=cuda)
    for k in ref:
        scale = ref[k].abs().max().clamp_min(1e-30)
        err = float((ref[k] - got[k]).abs().max() / scale)
        assert err < 1e-11, (k, err)


@pytest.mark.parametrize("n,L,I", [(8, 2, 21), (9, 3, 51), (16, 4, 21), (250, 61, 51), (296, 5, 21)])
def test_gru_on_cta_pairs_edge_shapes(cuda, monkeypatch, n, L, I):
    """The CTA-pair kernels at the edges of their pipelines: one, two and three steps past the prologue (two TMA stages,
    two operand buffers), full and ragged last row groups, the largest size the pairs are chosen for (296 rows)."""
    _rows_per_cta(monkeypatch, "pair")
    ref, got = _ref_and_ours(n, L, I, seed=n + L, cuda=cuda)
    for k in ref:
        scale = ref[k].abs().max().clamp_min(1e-30)
        err = float((ref[k] - got[k]).abs().max() / scale)
        assert err < 1e-11, (k, err)


@pytest.mark.parametrize("rows", [8, 16, 32, "pair"])
def test_no_grad_path_and_structured_decoder(cuda, monkeypatch, rows):
    _rows_per_cta(monkeypatch, rows)
    from draupnir_asr_b200.models_utils import RNNDecoder_Tiling
    torch.manual_seed(3)
    n, L, z, A = 45, 31, 30, 21
    dec = RNNDecoder_Tiling(L, A, 60, z, z + A, 5, 1, None).double().to(cuda)
    latent = torch.randn(n, z, dtype=torch.float64, device=cuda, requires_grad=True)
    blosum = torch.rand(L, A, dtype=torch.float64, device=cuda, requires_grad=True)
    hidden = torch.randn(60, dtype=torch.float64, device=cuda).expand(2, n, 60).contiguous()
    x = torch.cat((latent[:, None, :].expand(n, L, z), blosum[None].expand(n, L, A)), dim=2)
    a = dec.scores(x, hidden)
    b = dec.scores_structured(latent, blosum, hidden)
    assert float((a - b).abs().max()) < 1e-12
    ga = torch.autograd.grad(a.square().sum(), (latent, blosum) + tuple(dec.parameters()))
    gb = torch.autograd.grad(b.square().sum(), (latent, blosum) + tuple(dec.parameters()))
    for u, v in zip(ga, gb):
        assert float((u - v).abs().max() / u.abs().max().clamp_min(1e-30)) < 1e-11
    with torch.no_grad():
        c = dec.forward_structured(latent, blosum, hidden)
    assert float((c - torch.log_softmax(a, -1)).abs().max()) < 1e-12
    # library nn.GRU (cuDNN) on the same device agrees as well
    ref, _ = dec.rnn(x, hidden)
    assert float((dec.linear_probs(dec.fc1(ref)) - a).abs().max()) < 1e-10


@pytest.mark.parametrize("n,L,I", [(3, 1, 21), (37, 23, 21), (150, 40, 21)])
def test_encoder_last_position_path(cuda, n, L, I):
    """`gru_encoder_last`: same hidden states as nn.GRU, and for a loss that only sees the last position the gradients of
    the input, h0 and every GRU parameter equal torch's (forward direction: full BPTT; reverse direction: one step)."""
    from draupnir_asr_b200 import ops
    torch.manual_seed(n + L)
    rnn = torch.nn.GRU(input_size=I, hidden_size=60, batch_first=True, bidirectional=True, num_layers=1).double()
    x = torch.randn(n, L, I, dtype=torch.float64, requires_grad=True)
    h0 = torch.randn(2, n, 60, dtype=torch.float64, requires_grad=True)
    w = torch.randn(n, 120, dtype=torch.float64)
    out_ref, hn_ref = rnn(x, h0)
    (out_ref[:, -1] * w).sum().backward()
    rnn_g = torch.nn.GRU(input_size=I, hidden_size=60, batch_first=True, bidirectional=True, num_layers=1).double().to(cuda)
    rnn_g.load_state_dict(rnn.state_dict())
    xg = x.detach().to(cuda).requires_grad_(True)
    hg = h0.detach().to(cuda).requires_grad_(True)
    last, full = ops.gru_encoder_last(rnn_g, xg, hg)
    (last * w.to(cuda)).sum().backward()
    grads = {k: p.grad.clone() for k, p in rnn_g.named_parameters()}
    # the deferred complete pass uses the weights of the call, even if they are updated before somebody looks
    with torch.no_grad():
        for p in rnn_g.parameters():
            p.add_(0.25)
    out, hn = full.value()
    assert not out.requires_grad and full.value()[0] is out
    pairs = [("last", out_ref[:, -1].detach(), last.detach().cpu()), ("out", out_ref.detach(), out.cpu()),
             ("hn", hn_ref.detach(), hn.cpu()), ("x", x.grad, xg.grad.cpu()), ("h0", h0.grad, hg.grad.cpu())]
    pairs += [(k, p.grad, grads[k].cpu()) for k, p in rnn.named_parameters()]
    for k, a, b in pairs:
        err = float((a - b).abs().max() / a.abs().max().clamp_min(1e-30))
        assert err < 1e-11, (k, err)


def test_encoder_dictionary_entries_are_deferred_and_exact(cuda):
    """RNNEncoder.forward (S/models_utils.py:52-65): z_loc / z_scale / rnn_final_hidden_state are computed; the entries
    that need the complete bidirectional pass are evaluated on first access and equal nn.GRU's."""
    import pickle
    from draupnir_asr_b200.models_utils import RNNEncoder
    from draupnir_asr_b200.utils import Deferred, DeferredDict
    torch.manual_seed(4)
    n, L = 21, 17
    enc = RNNEncoder(L, 21, n, 60, 30, 21, 0, 1, None).double().to(cuda)
    x = torch.randn(n, L, 21, dtype=torch.float64, device=cuda)
    h0 = torch.randn(2, n, 60, dtype=torch.float64, device=cuda)
    out = enc(x, h0)
    assert isinstance(out, DeferredDict) and list(out) == ["z_loc", "z_scale", "rnn_final_bidirectional", "rnn_hidden_states",
                                                          "rnn_final_hidden_state"]
    assert isinstance(out.raw("rnn_hidden_states"), Deferred) and isinstance(out.raw("rnn_final_bidirectional"), Deferred)
    states_ref, hn_ref = enc.rnn(x, h0)
    last_ref = enc.fc1(states_ref[:, -1, :60] + states_ref[:, -1, 60:])
    assert float((out["rnn_final_hidden_state"] - last_ref).abs().max()) < 1e-11
    assert float((out["z_loc"] - enc.linear_means(last_ref)).abs().max()) < 1e-11
    assert float((out["z_scale"] - enc.softplus(enc.linear_std(last_ref))).abs().max()) < 1e-11
    assert float((out["rnn_hidden_states"] - (states_ref[:, :, :60] + states_ref[:, :, 60:])).abs().max()) < 1e-11
    assert not isinstance(out.raw("rnn_hidden_states"), Deferred)
    back = pickle.loads(pickle.dumps(out))
    assert type(back) is dict and float((back["rnn_final_bidirectional"] - hn_ref).abs().max()) < 1e-11
    assert all(torch.is_tensor(v) for v in out.values())


@pytest.mark.parametrize("K,M,N", [(1, 1, 1), (31, 21, 120), (32, 21, 120), (1000, 21, 120), (4099, 180, 21), (70000, 21, 120),
                                   (513, 64, 64)])
def test_gram_tn(cuda, K, M, N):
    from draupnir_asr_b200 import ops
    g = torch.Generator().manual_seed(K + M)
    A = torch.randn(K, M, generator=g, dtype=torch.float64)
    B = torch.randn(K, N, generator=g, dtype=torch.float64)
    C, cs = ops.gram_tn(A.to(cuda), B.to(cuda))
    C_ref, cs_ref = A.t() @ B, A.sum(0)
    assert float((C.cpu() - C_ref).abs().max() / C_ref.abs().max().clamp_min(1e-30)) < 1e-12
    assert float((cs.cpu() - cs_ref).abs().max() / cs_ref.abs().max().clamp_min(1e-30)) < 1e-12


@pytest.mark.parametrize("n,L,C", [(1, 1, 360), (37, 5, 360), (300, 17, 360), (2500, 3, 24), (250, 500, 360), (9, 131, 360)])
def test_input_sums_one_pass(cuda, n, L, C):
    """drp_gru_input_sums_f64: dU = sum of the position-chunk partials = dgi.sum(1), dV = sum of the row-block partials =
    dgi.sum(0); (250, 500) is a leaf-sharded rank's shape (two rows per CTA, positions split over the grid)."""
    from draupnir_asr_b200.ops import _p, _rc, _stream
    from draupnir_asr_b200._lib import lib
    g = torch.Generator().manual_seed(n + L)
    dgi = torch.randn(n, L, C, generator=g, dtype=torch.float64)
    d = dgi.to(cuda)
    nblk = lib.drp_gru_input_sums_blocks(n)
    assert 1 <= nblk <= n
    nch = lib.drp_gru_input_sums_chunks(n, L)
    assert 1 <= nch <= 16 and (n > 100 or L < 64 or nch > 1)
    dU = torch.full((nch, n, C), float("nan"), dtype=torch.float64, device=cuda)
    dVp = torch.full((nblk, L, C), float("nan"), dtype=torch.float64, device=cuda)
    _rc(lib.drp_gru_input_sums_f64(_p(d), n, L, C, _p(dU), _p(dVp), _stream()), "drp_gru_input_sums_f64")
    assert torch.allclose(dU.sum(0).cpu(), dgi.sum(1), rtol=1e-13, atol=1e-12)
    assert torch.allclose(dVp.sum(0).cpu(), dgi.sum(0), rtol=1e-13, atol=1e-12)


@pytest.mark.parametrize("M,K,N", [(1, 21, 180), (63, 21, 180), (64, 21, 21), (1000, 120, 21), (4099, 180, 21), (130, 21, 120),
                                   (70001, 21, 180), (20000, 120, 21), (257, 192, 184), (300, 4, 8)])
def test_tall_gemm(cuda, M, K, N):
    """drp_tall_gemm_f64 (the dense layers on n*L rows) against torch, with and without bias; and through `affine_tall`
    the gradients of input, weight and bias."""
    from draupnir_asr_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g, dtype=torch.float64)
    W = torch.randn(N, K, generator=g, dtype=torch.float64)
    b = torch.randn(N, generator=g, dtype=torch.float64)
    xg, Wg, bg = x.to(cuda), W.to(cuda), b.to(cuda)
    ref = x @ W.t() + b
    scale = float(ref.abs().max())
    assert float((ops.tall_gemm(xg, Wg.t().contiguous(), bg).cpu() - ref).abs().max()) < 1e-12 * scale
    assert float((ops.tall_gemm(xg, Wg.t().contiguous()).cpu() - x @ W.t()).abs().max()) < 1e-12 * scale
    xr, Wr, br = (t.clone().requires_grad_(True) for t in (x, W, b))
    w = torch.randn(M, N, generator=g, dtype=torch.float64)
    ((xr @ Wr.t() + br) * w).sum().backward()
    xq, Wq, bq = (t.clone().requires_grad_(True) for t in (xg, Wg, bg))
    (ops.affine_tall(xq, Wq, bq) * w.to(cuda)).sum().backward()
    for a, c in ((xr.grad, xq.grad), (Wr.grad, Wq.grad), (br.grad, bq.grad)):
        assert float((a - c.cpu()).abs().max() / a.abs().max().clamp_min(1e-30)) < 1e-12
