"""TEST INFRASTRUCTURE (see oracle/__init__.py) — fp64 CPU torch restatement of the reference hot path.

S/ = /root/reference/draupnir/src/draupnir/.  Every function cites the lines it restates.  No Pyro: the ELBO is
the explicit sum of `log_prob`s that `Trace_ELBO(num_particles=1)` forms for these fully reparameterised,
un-subsampled sites.  Distribution objects are built with the *same shapes* the reference builds
them with, so broadcasting quirks (e.g. the `[3,1]`-shaped `alpha` PyroParam of the variational guide being scored
against a `[3]`-shaped HalfNormal, S/guides.py:33 vs S/models.py:89) arise from torch itself, not from hand-coding.
"""
from __future__ import annotations

import math
from collections import namedtuple
from typing import Dict, Optional

import torch
import torch.nn as nn
from torch.distributions import Categorical, HalfNormal, Independent, MultivariateNormal, Normal

SamplingOutput = namedtuple("SamplingOutput", ["aa_sequences", "latent_space", "logits", "phis", "psis", "mean_phi",
                                               "mean_psi", "kappa_phi", "kappa_psi"])  # S/models.py:17


# ----------------------------------------------------------------------------------------------------------------
# numerical blocks (S/models_utils.py)
# ----------------------------------------------------------------------------------------------------------------
def squeeze_tensor(required_ndims: int, tensor: torch.Tensor) -> torch.Tensor:
    """S/utils.py:1874-1902 — drop size-1 dims until `required_ndims` remain."""
    while tensor.dim() > required_ndims:
        if tensor.shape[0] == 1:
            tensor = tensor.squeeze(0)
        elif 1 in tensor.shape:
            tensor = tensor.squeeze(list(tensor.shape).index(1))
        else:
            break
    return tensor


def ou_covariance(t: torch.Tensor, sigma_f: torch.Tensor, sigma_n: torch.Tensor, lamb: torch.Tensor) -> torch.Tensor:
    """`OUKernel_Fast.forward` S/models_utils.py:303-311: K_z = sf_z^2 exp(-t/lam_z) + sn_z^2 I, `[z,n,n]`."""
    first_term = (sigma_f ** 2).unsqueeze(-1).unsqueeze(-1)
    lamb_ = lamb.unsqueeze(-1).unsqueeze(-1)
    second_term = torch.exp(-t / lamb_)
    noise = torch.eye(t.shape[0], dtype=t.dtype)
    return first_term * second_term + sigma_n.unsqueeze(-1).unsqueeze(-1) ** 2 * noise


class RNNDecoder_Tiling(nn.Module):
    """S/models_utils.py:67-99 (same parameter names => state_dicts interchange with the product module)."""

    def __init__(self, aa_prob: int, gru_hidden_dim: int, rnn_input_size: int):
        super().__init__()
        self.fc1 = nn.Linear(2 * gru_hidden_dim, gru_hidden_dim)
        self.linear_probs = nn.Linear(gru_hidden_dim, aa_prob)
        self.rnn = nn.GRU(input_size=rnn_input_size, hidden_size=gru_hidden_dim, batch_first=True,
                          bidirectional=True, num_layers=1, dropout=0.0)
        self.logsoftmax = nn.LogSoftmax(dim=-1)
        self.num_layers = 1

    def forward(self, input, hidden):
        rnn_output, _ = self.rnn(input, hidden)
        return self.logsoftmax(self.linear_probs(self.fc1(rnn_output)))


class EmbedComplex(nn.Module):
    """S/models_utils.py:226-237 (also `EmbedComplexEncoder` :238-249, identical arithmetic)."""

    def __init__(self, aa_probs: int, embedding_dim: int):
        super().__init__()
        self.fc1 = nn.Linear(aa_probs, embedding_dim)
        self.fc2 = nn.Linear(embedding_dim, aa_probs)
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, input):
        return self.softmax(self.fc2(self.fc1(input)))


class RNNEncoder(nn.Module):
    """S/models_utils.py:22-65."""

    def __init__(self, aa_prob: int, gru_hidden_dim: int, z_dim: int):
        super().__init__()
        self.gru_hidden_dim = gru_hidden_dim
        nndim = int(gru_hidden_dim / 2)
        self.fc1 = nn.Linear(gru_hidden_dim, nndim)
        self.linear_means = nn.Linear(nndim, z_dim)
        self.linear_std = nn.Linear(nndim, z_dim)
        self.softplus = nn.Softplus()
        self.rnn = nn.GRU(input_size=aa_prob, hidden_size=gru_hidden_dim, batch_first=True, bidirectional=True,
                          num_layers=1, dropout=0.0)
        self.num_layers = 1

    def forward(self, input, hidden):
        states, final = self.rnn(input, hidden)
        H = self.gru_hidden_dim
        states = states[:, :, :H] + states[:, :, H:]
        last = self.fc1(states[:, -1])
        return {"z_loc": self.linear_means(last), "z_scale": self.softplus(self.linear_std(last)),
                "rnn_final_bidirectional": final, "rnn_hidden_states": states, "rnn_final_hidden_state": last}


class Networks(nn.Module):
    """All trainable networks + the two constant initial hidden states, default-initialised from one seed."""

    def __init__(self, z_dim=30, gru_hidden_dim=60, aa_probs=21, embedding_dim=50, seed=0, variational=True):
        super().__init__()
        gen_state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        dt = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        try:
            self.decoder = RNNDecoder_Tiling(aa_probs, gru_hidden_dim, z_dim + aa_probs)   # S/models.py:296
            self.embed = EmbedComplex(aa_probs, embedding_dim)                             # S/models.py:298
            self.h_0_MODEL = torch.randn(gru_hidden_dim)                                   # S/models.py:66 (constant)
            if variational:
                self.encoder = RNNEncoder(aa_probs, gru_hidden_dim, z_dim)                 # S/guides.py:31
                self.embeddingencoder = EmbedComplex(aa_probs, embedding_dim)              # S/guides.py:32
                self.h_0_GUIDE = torch.randn(gru_hidden_dim)                               # S/guides.py:30 (constant)
        finally:
            torch.set_default_dtype(dt)
            torch.random.set_rng_state(gen_state)


# ----------------------------------------------------------------------------------------------------------------
# model / guide log-densities
# ----------------------------------------------------------------------------------------------------------------
def gp_prior(values: Dict[str, torch.Tensor], patristic_matrix_sorted: torch.Tensor, z_dim: int):
    """S/models.py:85-121.  Returns (latent_space [n,z], dict of the five site log-probs)."""
    one = torch.tensor(1.0, dtype=torch.float64)
    lp = {}
    lp["alpha"] = Independent(HalfNormal(one).expand([3]), 1).log_prob(values["alpha"]).sum()
    alpha = values["alpha"] + 1e-6
    for k, name in enumerate(("sigma_f", "sigma_n", "lambd")):
        base = HalfNormal(alpha[k])
        d = Independent(base.expand(torch.Size([z_dim]) + base.batch_shape), 1)
        lp[name] = d.log_prob(values[name]).sum()
    sigma_f = squeeze_tensor(1, values["sigma_f"] + 1e-6)
    sigma_n = squeeze_tensor(1, values["sigma_n"] + 1e-6)
    lambd = squeeze_tensor(1, values["lambd"] + 1e-6)
    patristic_matrix = patristic_matrix_sorted[1:, 1:]
    cov = ou_covariance(patristic_matrix, sigma_f, sigma_n, lambd)
    mean = torch.zeros((patristic_matrix.shape[0],), dtype=torch.float64).unsqueeze(0)
    lp["latent_z"] = Independent(MultivariateNormal(mean, cov), 1).log_prob(values["latent_z"]).sum()
    return values["latent_z"].T, lp


def decoder_logits(nets: Networks, latent_space: torch.Tensor, blosum_weighted: torch.Tensor) -> torch.Tensor:
    """Tile / embed / cat / biGRU / 2 Linear / LogSoftmax — S/models.py:339-349, S/models_utils.py:93-99."""
    n, z = latent_space.shape
    L, A = blosum_weighted.shape
    ls = latent_space.repeat(1, L).reshape(n, L, z)
    blosum = blosum_weighted.repeat(n, 1).reshape(n, L, A)
    blosum = nets.embed(blosum)
    x = torch.cat((ls, blosum), dim=2)
    h0 = nets.h_0_MODEL.expand(2, n, nets.h_0_MODEL.shape[0]).contiguous()
    return nets.decoder(x, h0)


def model_log_joint(nets: Networks, values, family_data, patristic_matrix_sorted, blosum_weighted, z_dim: int):
    """S/models.py:300-352 (`model_variational` and `model_delta_map` score the same sites; the plates are
    size-less or un-subsampled => scale 1).  Returns (sum, per-site dict, logits)."""
    aminoacid_sequences = family_data[:, 2:, 0]
    latent_space, lp = gp_prior(values, patristic_matrix_sorted, z_dim)
    logits = decoder_logits(nets, latent_space, blosum_weighted)
    lp["aa_sequences"] = Categorical(logits=logits).log_prob(aminoacid_sequences).sum()
    return sum(lp.values()), lp, logits


def guide_batch(nets: Networks, guide_params: Dict[str, torch.Tensor], data_blosum: torch.Tensor, eps: torch.Tensor):
    """S/guides.py:95-133.  `guide_params` holds the constrained `[3,1]/[z,1]` PyroParams; `eps [z,n]` is the
    standard-normal draw of `Normal.rsample`.  Returns (dict like the reference's, log q)."""
    x = nets.embeddingencoder(data_blosum)
    h0 = nets.h_0_GUIDE.expand(2, x.shape[0], nets.h_0_GUIDE.shape[0]).contiguous()
    out = nets.encoder(x, h0)
    z_loc, z_scale = out["z_loc"], out["z_scale"]
    q = Normal(z_loc.T, z_scale.T)
    latent_z = q.loc + eps * q.scale           # torch Normal.rsample
    log_q = q.log_prob(latent_z).sum()         # Delta sites contribute 0
    res = {"alpha": guide_params["alpha"], "sigma_n": guide_params["sigma_n"], "sigma_f": guide_params["sigma_f"],
           "lambd": guide_params["lambd"], "z_loc": z_loc, "z_scale": z_scale, "latent_z": latent_z,
           "rnn_final_bidirectional": out["rnn_final_bidirectional"],
           "rnn_final_hidden_state": out["rnn_final_hidden_state"], "rnn_hidden_states": out["rnn_hidden_states"]}
    return res, log_q


# ----------------------------------------------------------------------------------------------------------------
# one SVI step (Trace_ELBO + ClippedAdam), S/train.py:58-79, S/main.py:1021-1059
# ----------------------------------------------------------------------------------------------------------------
class ClippedAdam:
    """pyro.optim.ClippedAdam as remembered from pyro-ppl 1.8/1.9 (not vendored): lr decays by
    `lrd` at the top of every step, gradients are clamped element-wise to +-clip_norm, Adam with eps added outside
    the bias correction."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_norm=10.0, lrd=1.0):
        self.params = list(params)
        self.lr, self.betas, self.eps, self.wd, self.clip, self.lrd = lr, betas, eps, weight_decay, clip_norm, lrd
        self.state = [dict(step=0, m=torch.zeros_like(p), v=torch.zeros_like(p)) for p in self.params]

    @torch.no_grad()
    def step(self):
        self.lr *= self.lrd
        b1, b2 = self.betas
        for p, st in zip(self.params, self.state):
            if p.grad is None:
                continue
            g = p.grad.clamp(-self.clip, self.clip)
            st["step"] += 1
            if self.wd != 0:
                g = g.add(p, alpha=self.wd)
            st["m"].mul_(b1).add_(g, alpha=1 - b1)
            st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = st["v"].sqrt().add_(self.eps)
            step_size = self.lr * math.sqrt(1 - b2 ** st["step"]) / (1 - b1 ** st["step"])
            p.addcdiv_(st["m"], denom, value=-step_size)
            p.grad = None


class OracleSVI:
    """The optimisable state of one run and `svi.step` for both guides.

    delta_map:   AutoDelta (S/main.py:516) keeps one parameter per latent site, positive sites in log space.
    variational: DRAUPNIRGUIDES PyroParams `[3,1]/[z,1]` in log space (S/guides.py:33-36) + encoder networks.
    Parameters are injected (`init`), never drawn, so both sides of a parity test start identically.
    """

    POSITIVE = ("alpha", "sigma_f", "sigma_n", "lambd")

    def __init__(self, nets: Networks, init: Dict[str, torch.Tensor], guide: str, z_dim=30, **adam):
        assert guide in ("delta_map", "variational")
        self.nets, self.guide, self.z_dim = nets, guide, z_dim
        self.u = {}
        for k in self.POSITIVE:
            v = init[k].detach().clone().to(torch.float64)
            if guide == "variational":
                v = v.reshape(-1, 1)
            self.u[k] = v.log().requires_grad_(True)
        if guide == "delta_map":
            self.u["latent_z"] = init["latent_z"].detach().clone().requires_grad_(True)
        self.optim = ClippedAdam(list(self.u.values()) + [p for p in nets.parameters()], **adam)

    def constrained(self):
        vals = {k: self.u[k].exp() for k in self.POSITIVE}
        if self.guide == "delta_map":
            vals["latent_z"] = self.u["latent_z"]
        return vals

    def loss(self, family_data, patristic_matrix_sorted, data_blosum, blosum_weighted, eps=None):
        vals = self.constrained()
        log_q = torch.zeros((), dtype=torch.float64)
        if self.guide == "variational":
            res, log_q = guide_batch(self.nets, vals, data_blosum, eps)
            vals = dict(vals, latent_z=res["latent_z"])
        log_p, terms, logits = model_log_joint(self.nets, vals, family_data, patristic_matrix_sorted, blosum_weighted,
                                               self.z_dim)
        terms = dict(terms, log_q=log_q)
        return -(log_p - log_q), terms, logits

    def step(self, family_data, patristic_matrix_sorted, data_blosum, blosum_weighted, eps=None) -> float:
        loss, _, _ = self.loss(family_data, patristic_matrix_sorted, data_blosum, blosum_weighted, eps)
        loss.backward()
        self.optim.step()
        return float(loss.detach())


# ----------------------------------------------------------------------------------------------------------------
# ancestral reconstruction
# ----------------------------------------------------------------------------------------------------------------
def conditional_moments(map_estimates, patristic_matrix, internal_nodes, jitter: float = 0.0):
    """Mean `[z,ni]` and covariance `[z,ni,ni]` of p(z_internal | z_leaves) through the PRECISION-matrix partition,
    exactly as S/models.py:153-193 (jitter=1e-6, added to every entry :193) and :200-239 (jitter=0)."""
    sigma_f = squeeze_tensor(1, map_estimates["sigma_f"])
    sigma_n = squeeze_tensor(1, map_estimates["sigma_n"])
    lambd = squeeze_tensor(1, map_estimates["lambd"])
    internal_indexes = (patristic_matrix[1:, 0][..., None] == internal_nodes).any(-1)
    cov_full = ou_covariance(patristic_matrix[1:, 1:], sigma_f, sigma_n, lambd)
    inverse_full = torch.linalg.inv(cov_full)
    inverse_internal = inverse_full[:, internal_indexes, :][:, :, internal_indexes]
    inverse_internal_leaves = inverse_full[:, internal_indexes][:, :, ~internal_indexes]
    xb = map_estimates["latent_z"]
    part1 = torch.matmul(torch.linalg.inv(inverse_internal), inverse_internal_leaves)
    mean = (0.0 - torch.matmul(part1, xb[:, :, None])).squeeze(-1)
    return mean, torch.linalg.inv(inverse_internal) + jitter


def conditional_sampling(map_estimates, patristic_matrix, internal_nodes, eps: torch.Tensor) -> torch.Tensor:
    """S/models.py:149-196: one draw `[ni,z]`; `eps [z,ni]` is the standard normal torch's MVN.sample would draw."""
    mean, cov = conditional_moments(map_estimates, patristic_matrix, internal_nodes, jitter=1e-6)
    L = torch.linalg.cholesky(cov)
    return (mean + torch.matmul(L, eps[:, :, None]).squeeze(-1)).T


def conditional_samplingMAP(map_estimates, patristic_matrix, internal_nodes) -> torch.Tensor:
    """S/models.py:197-242: returns the conditional MEAN `[ni,z]` (the draw at :239 is discarded)."""
    mean, _ = conditional_moments(map_estimates, patristic_matrix, internal_nodes, jitter=0.0)
    return mean.T


def sample(nets: Networks, latent_space: torch.Tensor, blosum_weighted: torch.Tensor, use_argmax=True,
           n_samples=1, generator: Optional[torch.Generator] = None) -> SamplingOutput:
    """S/models.py:361-401 after the latent has been chosen (leaves / conditional draw / conditional mean)."""
    logits = decoder_logits(nets, latent_space, blosum_weighted)
    if use_argmax:
        aa = torch.argmax(logits, dim=2).unsqueeze(0)
    else:
        probs = logits.exp().reshape(-1, logits.shape[-1])
        aa = torch.multinomial(probs, n_samples, replacement=True, generator=generator).T.reshape(n_samples, *logits.shape[:2])
    return SamplingOutput(aa.detach(), latent_space.detach(), logits.detach(), None, None, None, None, None, None)
