"""Variational guide with the reference's class name, constructor and return dictionary (S/guides.py, S/ =
/root/reference/draupnir/src/draupnir/)."""
from __future__ import annotations

import torch
import torch.distributions.constraints as constraints
import torch.nn as nn

from .backend import pyro, dist, USING_REAL_PYRO
from .models_utils import EmbedComplexEncoder, RNNEncoder
from .utils import DeferredDict

if USING_REAL_PYRO:  # pragma: no cover
    from pyro.contrib.easyguide import EasyGuide
    from pyro.nn import PyroParam
else:
    from .pyro_compat.contrib.easyguide import EasyGuide
    from .pyro_compat.nn import PyroParam


class DRAUPNIRGUIDES(EasyGuide):
    """S/guides.py:18-171."""

    def __init__(self, draupnir_model, ModelLoad, Draupnir):
        super(DRAUPNIRGUIDES, self).__init__(draupnir_model)
        self.guide_type = ModelLoad.args.select_guide
        self.__dict__["draupnir"] = Draupnir          # plain attribute: the model is not a sub-module of the guide
        self.encoder_rnn_input_size = Draupnir.aa_probs
        self.dataset_train_blosum = Draupnir.dataset_train_blosum
        self.batch_size = Draupnir.batch_size
        self.batch_by_clade = Draupnir.batch_by_clade
        dev = Draupnir.device
        if Draupnir.pretrained_params is not None:
            h0 = Draupnir.pretrained_params["h_0_GUIDE"]
        else:
            h0 = torch.randn(Draupnir.gru_hidden_dim, dtype=torch.float64)
        # S/guides.py:28-30 — never registered with pyro.module, i.e. a constant of the guide
        self.h_0_GUIDE = nn.Parameter(h0.to(dev), requires_grad=False)
        self.encoder = RNNEncoder(Draupnir.align_seq_len, Draupnir.aa_probs, Draupnir.n_leaves, Draupnir.gru_hidden_dim,
                                  Draupnir.z_dim, self.encoder_rnn_input_size, Draupnir.kappa_addition, Draupnir.num_layers,
                                  Draupnir.pretrained_params)
        self.embeddingencoder = EmbedComplexEncoder(Draupnir.aa_probs, Draupnir.embedding_dim, Draupnir.pretrained_params)
        one = torch.tensor([1.0], dtype=torch.float64)
        hn = torch.distributions.HalfNormal(one)
        self.alpha = PyroParam(hn.sample([3]), constraint=constraints.positive, event_dim=0)
        self.sigma_n = PyroParam(hn.sample([Draupnir.z_dim]), constraint=constraints.positive, event_dim=0)
        self.sigma_f = PyroParam(hn.sample([Draupnir.z_dim]), constraint=constraints.positive, event_dim=0)
        self.lambd = PyroParam(hn.sample([Draupnir.z_dim]), constraint=constraints.positive, event_dim=0)
        self.to(device=dev, dtype=torch.float64)

    def guide(self, family_data, patristic_matrix, cladistic_matrix, data_blosum, batch_blosum=None, map_estimates=None):
        """Dispatcher of S/guides.py:41-56."""
        if self.batch_size is None or self.batch_size > 1:
            if self.batch_by_clade:
                raise NotImplementedError("clade batching splits the OU process into blocks (README-discouraged) and is "
                                          "outside the accelerated whole-tree path")
            return self.guide_batch(family_data, patristic_matrix, cladistic_matrix, data_blosum, batch_blosum=None)
        return self.guide_noplating(family_data, patristic_matrix, cladistic_matrix, data_blosum, batch_blosum=None)

    def _encode(self, data_blosum):
        hook = self.__dict__.get("_before_encode")      # a host-fed step waits for the BLOSUM-encoded data set only here
        if hook is not None:
            hook()
        shard = getattr(self.draupnir, "shard", None)
        n = data_blosum.shape[0]
        if shard is not None:        # encoder is data-parallel over contiguous blocks of leaves
            data_blosum = data_blosum[shard.leaf_lo:shard.leaf_hi]
        aminoacid_sequences = self.embeddingencoder(data_blosum)
        encoder_h_0 = self.h_0_GUIDE.expand(self.encoder.num_layers * 2, aminoacid_sequences.shape[0],
                                            self.draupnir.gru_hidden_dim).contiguous()
        out = self.encoder(aminoacid_sequences, encoder_h_0)
        if shard is not None:        # one all-gather of [z_loc | z_scale] rebuilds the [n, z] posterior parameters
            z = out["z_loc"].shape[1]
            both = shard.gather_leaf_rows(torch.cat((out["z_loc"], out["z_scale"]), dim=1))
            out = DeferredDict({k: out.raw(k) for k in out})
            out["z_loc"], out["z_scale"] = both[:, :z], both[:, z:]
        return out, n

    def _latent_site(self, out):
        shard = getattr(self.draupnir, "shard", None)
        q = dist.Normal(out["z_loc"].T, out["z_scale"].T)
        if shard is None:
            return pyro.sample("latent_z", q)
        with pyro.poutine.scale(scale=shard.replicated_scale):   # q(z) is replicated: scored on rank 0 only
            return pyro.sample("latent_z", q)

    @staticmethod
    def _result(alpha, sigma_n, sigma_f, lambd, latent_z, out):
        # same ten keys in the same order as S/guides.py:123-133; the encoder's full hidden-state entries stay deferred
        return DeferredDict({"alpha": alpha, "sigma_n": sigma_n, "sigma_f": sigma_f, "lambd": lambd, "z_loc": out["z_loc"],
                             "z_scale": out["z_scale"], "latent_z": latent_z,
                             "rnn_final_bidirectional": out.raw("rnn_final_bidirectional"),
                             "rnn_final_hidden_state": out["rnn_final_hidden_state"],
                             "rnn_hidden_states": out.raw("rnn_hidden_states")})

    def guide_noplating(self, family_data, patristic_matrix_sorted, cladistic_matrix, data_blosum, batch_blosum=None,
                        map_estimates=None):
        """S/guides.py:58-93 (the hyper-parameters are not sample sites here, exactly as in the reference)."""
        alpha, sigma_n, sigma_f, lambd = self.alpha, self.sigma_n, self.sigma_f, self.lambd
        pyro.module("encoder", self.encoder)
        pyro.module("embeddingsencoder", self.embeddingencoder)
        with pyro.plate("plate_batch", dim=-1, device=self.draupnir.device):
            out, n = self._encode(self.dataset_train_blosum)
            latent_z = self._latent_site(out)
            assert latent_z.shape == (self.draupnir.z_dim, n)
        return self._result(alpha, sigma_n, sigma_f, lambd, latent_z, out)

    def guide_batch(self, family_data, patristic_matrix_sorted, cladistic_matrix, data_blosum, batch_blosum=None,
                    map_estimates=None):
        """S/guides.py:95-133 — the guide the whole-tree run uses (batch_size == number of sequences)."""
        pyro.module("encoder", self.encoder)
        pyro.module("embeddingsencoder", self.embeddingencoder)
        with pyro.plate("plate_batch", dim=-1, device=self.draupnir.device):
            alpha = pyro.sample("alpha", dist.Delta(self.alpha).to_event(1))
            sigma_n = pyro.sample("sigma_n", dist.Delta(self.sigma_n).to_event(1))
            sigma_f = pyro.sample("sigma_f", dist.Delta(self.sigma_f).to_event(1))
            lambd = pyro.sample("lambd", dist.Delta(self.lambd).to_event(1))
            # the model's OU covariances depend on these three values only: their factorisation starts now
            self.draupnir.prefetch_prior(patristic_matrix_sorted, sigma_f, sigma_n, lambd)
            out, n = self._encode(data_blosum)
            latent_z = self._latent_site(out)
            assert latent_z.shape == (self.draupnir.z_dim, n)
        return self._result(alpha, sigma_n, sigma_f, lambd, latent_z, out)
