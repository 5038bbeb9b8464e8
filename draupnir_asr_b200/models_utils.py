"""Numerical building blocks with the reference's names and constructor signatures (S/models_utils.py, S/ =
/root/reference/draupnir/src/draupnir/).  Parameter names equal the reference's so `Model_state_dict.p` /
`Guide_state_dict.p` checkpoints interchange."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
import torch.nn as nn

from . import ops
from .utils import Deferred, DeferredDict


class RNNEncoder(nn.Module):
    """S/models_utils.py:22-65 — biGRU encoder of the variational guide."""

    def __init__(self, align_seq_len, aa_prob, n_leaves, gru_hidden_dim, z_dim, rnn_input_size, kappa_addition, num_layers,
                 pretrained_params):
        super().__init__()
        self.gru_hidden_dim, self.z_dim, self.n_leaves = gru_hidden_dim, z_dim, n_leaves
        self.rnn_input_size, self.align_seq_len, self.aa_prob = rnn_input_size, align_seq_len, aa_prob
        self.num_layers, self.kappa_addition = num_layers, kappa_addition
        self.nndim = int(self.gru_hidden_dim / 2)
        self.fc1 = nn.Linear(self.gru_hidden_dim, self.nndim)
        self.linear_means = nn.Linear(self.nndim, self.z_dim)
        self.linear_std = nn.Linear(self.nndim, self.z_dim)
        self.softplus = nn.Softplus()
        self.rnn = nn.GRU(input_size=self.rnn_input_size, hidden_size=self.gru_hidden_dim, batch_first=True,
                          bidirectional=True, num_layers=self.num_layers, dropout=0.0)

    def forward(self, input, hidden):
        # only the hidden state at the LAST position reaches the loss (S/models_utils.py:57): that is all that is computed
        # here; the full sequence of hidden states and h_n (needing the whole reverse pass) are dictionary entries that
        # are evaluated when somebody reads them (utils.DeferredDict) — svi.step never does
        last, full = ops.gru_encoder_last(self.rnn, input, hidden)
        H = self.gru_hidden_dim
        rnn_final_hidden_state = self.fc1(last[:, :H] + last[:, H:])
        z_loc = self.linear_means(rnn_final_hidden_state)
        z_scale = self.softplus(self.linear_std(rnn_final_hidden_state))
        return DeferredDict({"z_loc": z_loc, "z_scale": z_scale,
                             "rnn_final_bidirectional": Deferred(lambda: full.value()[1]),
                             "rnn_hidden_states": Deferred(lambda: full.value()[0][:, :, :H] + full.value()[0][:, :, H:]),
                             "rnn_final_hidden_state": rnn_final_hidden_state})


class RNNDecoder_Tiling(nn.Module):
    """S/models_utils.py:67-99 — biGRU -> fc1 -> linear_probs -> LogSoftmax."""

    def __init__(self, align_seq_len, aa_prob, gru_hidden_dim, z_dim, rnn_input_size, kappa_addition, num_layers,
                 pretrained_params):
        super().__init__()
        self.gru_hidden_dim, self.z_dim, self.rnn_input_size = gru_hidden_dim, z_dim, rnn_input_size
        self.align_seq_len, self.aa_prob, self.num_layers, self.kappa_addition = align_seq_len, aa_prob, num_layers, kappa_addition
        self.logsoftmax = nn.LogSoftmax(dim=-1)
        self.fc1 = nn.Linear(2 * self.gru_hidden_dim, self.gru_hidden_dim)
        self.linear_probs = nn.Linear(self.gru_hidden_dim, self.aa_prob)
        self.rnn = nn.GRU(input_size=self.rnn_input_size, hidden_size=self.gru_hidden_dim, batch_first=True,
                          bidirectional=True, num_layers=self.num_layers, dropout=0.0)

    def scores(self, input, hidden):
        """Un-normalised per-site scores `[n,L,A]`; the training path feeds these straight to the fused
        categorical log-likelihood kernel (log_softmax is idempotent, so the ELBO is unchanged)."""
        rnn_output, _ = ops.gru_module_forward(self.rnn, input, hidden)
        return self._project(rnn_output)

    def _project(self, rnn_output):
        """`linear_probs(fc1(rnn_output))` (S/models_utils.py:98).  There is no non-linearity between the two layers, so
        they are applied as ONE `[2H -> aa_prob]` affine map (W2 W1, W2 b1 + b2): a third of the flops and no
        `[n,L,H]` intermediate; autograd still delivers the gradients of fc1 / linear_probs themselves."""
        W = self.linear_probs.weight @ self.fc1.weight                                   # [A, 2H]
        b = torch.addmv(self.linear_probs.bias, self.linear_probs.weight, self.fc1.bias)  # [A]
        n, L, F = rnn_output.shape
        return ops.affine_tall(rnn_output.reshape(n * L, F), W, b).reshape(n, L, -1)

    def scores_structured(self, latent_space, blosum, hidden):
        """`scores(cat(latent_space tiled over L, blosum tiled over n), hidden)` without building the `[n,L,z+A]` input
        (S/models.py:339-342): since x[i,l] = [latent_i ; blosum_l], W_ih x + b_ih = W_z latent_i + (W_b blosum_l + b_ih)
        is a broadcast sum of an `[n,3H]` and an `[L,3H]` term per direction."""
        H, z = self.gru_hidden_dim, latent_space.shape[1]
        ops.require_gru_hidden(H)
        wih, whh, bih, bhh = ops.gru_stacked_params(self.rnn)
        U = latent_space @ wih[:, :, :z].reshape(6 * H, z).t()                               # [n, 2*3H]
        V = torch.addmm(bih.reshape(-1), blosum, wih[:, :, z:].reshape(6 * H, -1).t())      # [L, 2*3H]
        rnn_output = ops.gru_bidirectional_uv(U, V, whh, bhh, hidden)
        return self._project(rnn_output)

    def forward(self, input, hidden):
        return self.logsoftmax(self.scores(input, hidden))

    def forward_structured(self, latent_space, blosum, hidden):
        return self.logsoftmax(self.scores_structured(latent_space, blosum, hidden))


class EmbedComplex(nn.Module):
    """S/models_utils.py:226-237."""

    def __init__(self, aa_probs, embedding_dim, pretrained_params):
        super().__init__()
        self.aa_probs, self.embedding_dim = aa_probs, embedding_dim
        self.softmax = nn.Softmax(dim=-1)
        self.fc1 = nn.Linear(self.aa_probs, self.embedding_dim)
        self.fc2 = nn.Linear(self.embedding_dim, self.aa_probs)

    def forward(self, input):
        # fc2(fc1(x)) has no non-linearity in between (S/models_utils.py:235-236): applied as ONE [A -> A] affine map
        # (W2 W1, W2 b1 + b2) — the encoder's copy runs on n*L rows, where this saves the [n*L, embedding_dim]
        # intermediate and two thirds of the flops; autograd still delivers the gradients of fc1 / fc2 themselves
        if not input.is_cuda:
            raise RuntimeError("EmbedComplex: expected a CUDA tensor (draupnir_asr_b200 has no CPU path)")
        W = self.fc2.weight @ self.fc1.weight
        b = torch.addmv(self.fc2.bias, self.fc2.weight, self.fc1.bias)
        A = input.shape[-1]
        return self.softmax(ops.affine_tall(input.reshape(-1, A), W, b).reshape(input.shape[:-1] + (W.shape[0],)))


class EmbedComplexEncoder(EmbedComplex):
    """S/models_utils.py:238-249 (same arithmetic as EmbedComplex)."""


class GPKernel(ABC):
    """S/models_utils.py:251-257."""

    @abstractmethod
    def preforward(self, t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    @abstractmethod
    def forward(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class OUKernel_Fast(GPKernel):
    """S/models_utils.py:286-311 — K_z = sigma_f_z^2 exp(-t / lamb_z) + sigma_n_z^2 I as ONE fused kernel
    (`drp_ou_cov_fwd_f64`) with an analytic backward, instead of six element-wise passes."""

    def __init__(self, sigma_f, sigma_n, lamb):
        self.sigma_f, self.sigma_n, self.lamb = sigma_f, sigma_n, lamb

    def preforward(self, t1: torch.Tensor, t2: torch.Tensor) -> torch.Tensor:
        return torch.zeros((1,), device=t1.device)

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        return ops.ou_cov(t, self.sigma_f, self.sigma_n, self.lamb)


def compute_sites_entropies(logits, node_names):
    """S/models_utils.py:409-427 (consumer of the decoder's logits; kept for API completeness)."""
    prob = torch.exp(logits) / (1 + torch.exp(logits))
    seq_entropies = -torch.sum(prob * torch.log(prob), dim=2)
    return torch.cat((node_names[:, None], seq_entropies), dim=1)
