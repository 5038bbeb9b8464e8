"""DRAUPNIR model classes with the reference's names, constructor, sample sites and method signatures
(S/models.py, S/ = /root/reference/draupnir/src/draupnir/), computing on the sm_100a kernels.

What changes under the API:
* `gp_prior`: the `latent_z` site is an `OUProcessPrior` — covariance assembly, batched Cholesky, triangular solve,
  log-determinant and every gradient are one fused pipeline; the `[z,n,n]` covariance is never materialised.
* conditional sampling uses the Schur complement on the leaf block instead of three dense inversions.
* the `aa_sequences` site scores un-normalised decoder scores with one fused log-sum-exp/gather/reduce kernel.
* the BLOSUM embedding is evaluated once per alignment column instead of once per (node, column).
"""
from __future__ import annotations

from abc import abstractmethod
from collections import namedtuple

import torch
import torch.nn as nn

from . import ops
from .backend import pyro, dist
from .models_utils import *  # noqa: F401,F403  (the reference does the same, S/models.py:13)
from .models_utils import EmbedComplex, OUKernel_Fast, RNNDecoder_Tiling
from .utils import require_cuda, squeeze_tensor

SamplingOutput = namedtuple("SamplingOutput", ["aa_sequences", "latent_space", "logits", "phis", "psis", "mean_phi",
                                               "mean_psi", "kappa_phi", "kappa_psi"])  # S/models.py:17


class DRAUPNIRModelClass(nn.Module):
    """S/models.py:19-288."""

    def __init__(self, ModelLoad):
        super(DRAUPNIRModelClass, self).__init__()
        self.args = ModelLoad.args
        self.num_epochs = ModelLoad.args.num_epochs
        self.gru_hidden_dim = ModelLoad.gru_hidden_dim
        self.embedding_dim = ModelLoad.args.embedding_dim
        self.position_embedding_dim = ModelLoad.args.position_embedding_dim
        self.pretrained_params = ModelLoad.pretrained_params
        self.z_dim = ModelLoad.z_dim
        self.leaves_nodes = ModelLoad.leaves_nodes
        self.n_tree_levels = ModelLoad.n_tree_levels
        self.align_seq_len = ModelLoad.align_seq_len
        self.aa_probs = ModelLoad.build_config.aa_probs
        self.edge_info = ModelLoad.graph_coo
        self.nodes_representations_array = ModelLoad.nodes_representations_array
        self.dgl_graph = ModelLoad.dgl_graph
        self.children_dict = ModelLoad.children_dict
        self.closest_leaves_dict = ModelLoad.closest_leaves_dict
        self.descendants_dict = ModelLoad.descendants_dict
        self.clades_dict_all = ModelLoad.clades_dict_all
        self.max_indel_size = ModelLoad.args.max_indel_size
        self.rnn_input_size = self.z_dim
        self.use_attention = False
        self.batch_first = True
        self.leaves_testing = ModelLoad.leaves_testing
        self.batch_by_clade = ModelLoad.args.batch_by_clade
        self.device = require_cuda(ModelLoad.device)
        self.kappa_addition = ModelLoad.args.kappa_addition
        self.aa_frequencies = ModelLoad.aa_frequencies
        self.blosum = ModelLoad.blosum
        self.blosum_max = ModelLoad.blosum_max
        self.blosum_weighted = ModelLoad.blosum_weighted
        self.dataset_train_blosum = ModelLoad.dataset_train_blosum
        self.variable_score = ModelLoad.variable_score
        self.internal_nodes = ModelLoad.internal_nodes
        self.batch_size = ModelLoad.build_config.batch_size
        self.plating = ModelLoad.args.plating
        self.plate_size = ModelLoad.build_config.plate_subsample_size
        self.plate_unordered = ModelLoad.plate_unordered
        self.one_hot_encoding = ModelLoad.one_hot_encoding
        self.n_leaves_batch = self.batch_size
        self.n_internal_batch = self.batch_size
        self.n_leaves = len(self.leaves_nodes)
        self.n_internal = len(self.internal_nodes)
        self.n_all = self.n_leaves + self.n_internal
        self.num_layers = 1
        # S/models.py:66 — never registered with pyro.module, i.e. a constant of the model
        self.h_0_MODEL = nn.Parameter(torch.randn(self.gru_hidden_dim, dtype=torch.float64), requires_grad=False)
        self._one = None
        self._pos_cache = {}
        self._cond_cache = None
        self._prefetched = None
        self._hyper_sites = ()
        self._before_prior = None  # hook: called on the stream that is about to read the patristic matrix (host-fed steps)
        self.shard = None          # zshard.ShardContext when the run is spread over several GPUs
        self.map_consumes_rng = True   # conditional_samplingMAP advances the generator like the reference (S/models.py:239)
        for name in ("blosum_weighted", "dataset_train_blosum", "internal_nodes", "leaves_nodes", "aa_frequencies",
                     "blosum", "blosum_max"):
            t = getattr(self, name)
            if isinstance(t, torch.Tensor):
                setattr(self, name, t.to(self.device))
        self.to(self.device)

    @abstractmethod
    def guide(self, family_data, patristic_matrix, cladistic_matrix, data_blosum, batch_blosum, guide_map_estimates):
        raise NotImplementedError

    @abstractmethod
    def model(self, family_data, patristic_matrix, cladistic_matrix, data_blosum, batch_blosum, guide_map_estimates):
        raise NotImplementedError

    @abstractmethod
    def sample(self, map_estimates, n_samples, family_data_test, patristic_matrix, cladistic_matrix, use_test=True):
        raise NotImplementedError

    def get_class(self):
        full_name = self.__class__
        return str(full_name).split(".")[-1].replace("'>", "")

    # -------------------------------------------------------------------------------------------------------------
    def _hyper_priors(self, eps: float):
        """The four HalfNormal sites of S/models.py:89-92 (sampled/replayed through pyro, +eps as the reference)."""
        if self._one is None or self._one.device != self.h_0_MODEL.device:
            self._one = torch.ones((), dtype=torch.float64, device=self.h_0_MODEL.device)
        if self.shard is not None:   # replicated sites are scored on rank 0 only (zshard.py)
            with pyro.poutine.scale(scale=self.shard.replicated_scale):
                return self._hyper_prior_sites(eps)
        return self._hyper_prior_sites(eps)

    def _hyper_prior_sites(self, eps: float):
        alpha = pyro.sample("alpha", dist.HalfNormal(self._one).expand_by([3, ]).to_event(1)) + eps
        sf_site = pyro.sample("sigma_f", dist.HalfNormal(alpha[0]).expand_by([self.z_dim, ]).to_event(1))
        sn_site = pyro.sample("sigma_n", dist.HalfNormal(alpha[1]).expand_by([self.z_dim, ]).to_event(1))
        lam_site = pyro.sample("lambd", dist.HalfNormal(alpha[2]).expand_by([self.z_dim, ]).to_event(1))
        self._hyper_sites = (sf_site, sn_site, lam_site)         # what a prefetched factorisation is matched against
        return (squeeze_tensor(1, alpha), squeeze_tensor(1, sf_site + eps), squeeze_tensor(1, sn_site + eps),
                squeeze_tensor(1, lam_site + eps))

    def prefetch_prior(self, patristic_matrix_sorted, sigma_f, sigma_n, lambd):
        """Called by the guide as soon as ITS values of the three kernel hyper-parameters exist (under `Trace_ELBO` the
        model's sites are replayed from them): starts assembling and factoring the OU covariances on the side stream, so
        that the part of the prior that costs time runs while the encoder, and in a host-fed step the copy of the
        BLOSUM-encoded data set, are still under way.  `gp_prior` picks the factorisation up if, and only if, it is handed
        the very same tensors; otherwise it is dropped unused."""
        self._prefetched = None
        n = patristic_matrix_sorted.shape[0] - 1
        if not (ops.PRIOR_STREAM_ENABLED and ops.in_elbo_scope() and patristic_matrix_sorted.is_cuda and n >= 400
                and torch.is_grad_enabled()):
            return
        sf, sn, lam = (squeeze_tensor(1, t + 1e-6) for t in (sigma_f, sigma_n, lambd))
        if self.shard is not None:
            lo, hi = self.shard.z_lo, self.shard.z_hi
            if hi <= lo:
                return
            sf, sn, lam = sf[lo:hi], sn[lo:hi], lam[lo:hi]
        self._prefetched = ops.ou_prior_factor(patristic_matrix_sorted[1:, 1:], sf, sn, lam, key=(sigma_f, sigma_n, lambd),
                                               stream=ops.side_stream(patristic_matrix_sorted.device),
                                               before_launch=self._before_prior)

    def gp_prior(self, patristic_matrix_sorted):
        """Ornstein-Uhlenbeck process prior over the latent space (S/models.py:85-121).  Returns `[n, z]`."""
        alpha, sigma_f, sigma_n, lambd = self._hyper_priors(1e-6)
        if self._before_prior is not None:
            self._before_prior()
        patristic_matrix = patristic_matrix_sorted[1:, 1:]
        n_expected = self.n_all if self.leaves_testing else self.n_leaves
        assert patristic_matrix.shape == (n_expected, n_expected), \
            f"Expected shape {(self.z_dim, n_expected, n_expected)}, got {(self.z_dim,) + tuple(patristic_matrix.shape)}"
        z_slice = None if self.shard is None else (self.shard.z_lo, self.shard.z_hi)
        factor, self._prefetched = self._prefetched, None
        if factor is not None and not factor.matches(patristic_matrix, self._hyper_sites):
            factor = None
        self._hyper_sites = ()             # no reference to this step's autograd graph survives the step
        latent_space = pyro.sample("latent_z", dist.OUProcessPrior(patristic_matrix, sigma_f, sigma_n, lambd,
                                                                   z_slice=z_slice, factor=factor).to_event(1))  # [z, n]
        return latent_space.T

    def gp_prior_batched(self, patristic_matrix_sorted):
        """S/models.py:122-142 — same prior on a batch slice of the tree, without the +1e-6."""
        alpha, sigma_f, sigma_n, lambd = self._hyper_priors(0.0)
        patristic_matrix = patristic_matrix_sorted[1:, 1:]
        assert patristic_matrix.shape == (self.n_leaves_batch, self.n_leaves_batch)
        latent_space = pyro.sample("latent_z",
                                   dist.OUProcessPrior(patristic_matrix, sigma_f, sigma_n, lambd).to_event(1))
        return latent_space.T

    def map_sampling(self, map_estimates, patristic_matrix_full):
        """S/models.py:143-148."""
        test_indexes = (patristic_matrix_full[1:, 0][..., None] == self.internal_nodes).any(-1)
        latent_space_internal = map_estimates["latent_z"][:, test_indexes].T
        assert latent_space_internal.shape == (self.n_internal, self.z_dim)
        return latent_space_internal

    # -------------------------------------------------------------------------------------------------------------
    def _positions(self, patristic_matrix, internal_nodes):
        """Row positions of the internal / leaf nodes inside `patristic_matrix[1:,1:]` (S/models.py:157); cached
        because the mask is a constant of the data set and `nonzero` synchronises."""
        key = (patristic_matrix.data_ptr(), tuple(patristic_matrix.shape), internal_nodes.data_ptr())
        hit = self._pos_cache.get(key)
        if hit is None:
            ids = patristic_matrix[1:, 0]
            internal_indexes = (ids[..., None] == internal_nodes.to(ids.device)).any(-1)
            hit = (torch.nonzero(~internal_indexes).reshape(-1), torch.nonzero(internal_indexes).reshape(-1))
            self._pos_cache = {key: hit}
        return hit

    def _conditional(self, map_estimates, patristic_matrix, internal_nodes, n_internal, eps, jitter, want_cov):
        sigma_f = squeeze_tensor(1, map_estimates["sigma_f"]) + eps
        sigma_n = squeeze_tensor(1, map_estimates["sigma_n"]) + eps
        lambd = squeeze_tensor(1, map_estimates["lambd"]) + eps
        leaf_pos, int_pos = self._positions(patristic_matrix, internal_nodes)
        assert int_pos.numel() == n_internal
        xb = map_estimates["latent_z"]
        if self.leaves_testing:
            leaves_indexes = (patristic_matrix[1:, 0][..., None] == self.leaves_nodes.to(xb.device)).any(-1)
            xb = xb[:, leaves_indexes]
        sigma_f, sigma_n, lambd, xb = sigma_f.detach(), sigma_n.detach(), lambd.detach(), xb.detach()
        if self._z_sharded():
            # the z conditional posteriors are as independent as the z priors.  This rank eliminates the leaf
            # block of ITS latent dimensions only; one all-gather rebuilds the means [z, ni].  The covariance (needed for
            # draws) stays sharded: `cov` holds the rows z_lo:z_hi and `conditional_sampling_many` gathers the draws.
            lo, hi = self.shard.z_lo, self.shard.z_hi
            if hi > lo:
                mean, cov = ops.ou_conditional(patristic_matrix[1:, 1:], leaf_pos, int_pos, sigma_f[lo:hi], sigma_n[lo:hi],
                                               lambd[lo:hi], xb[lo:hi].contiguous(), jitter=jitter, want_cov=want_cov)
            else:
                mean = xb.new_zeros((0, n_internal))
                cov = xb.new_zeros((0, n_internal, n_internal)) if want_cov else None
            return self.shard.gather_z_rows(mean), cov
        return ops.ou_conditional(patristic_matrix[1:, 1:], leaf_pos, int_pos, sigma_f, sigma_n, lambd, xb, jitter=jitter,
                                  want_cov=want_cov)

    def _z_sharded(self) -> bool:
        return self.shard is not None and self.shard.world > 1

    def _conditional_factor(self, map_estimates, patristic_matrix):
        """Conditional mean `[z,ni]` and Cholesky factor `[z,ni,ni]` of the conditional covariance (+1e-6 on every
        entry as S/models.py:193), kept for as long as the SAME estimate tensors come back: the reference's marginal
        sampling loop (S/main.py:1337-1350) calls `sample(..., use_test=True)` once per draw with unchanged
        `map_estimates` and re-derives both every time.  The key holds the tensors themselves (identity + version
        counter), so a recycled allocation can never alias a stale entry."""
        keys = tuple(map_estimates[k] for k in ("sigma_f", "sigma_n", "lambd", "latent_z")) + (patristic_matrix,)
        stamp = tuple(int(t._version) for t in keys)
        hit = self._cond_cache
        if hit is not None and len(hit[0]) == len(keys) and all(a is b for a, b in zip(hit[0], keys)) and hit[1] == stamp:
            return hit[2], hit[3]
        mean, cov = self._conditional(map_estimates, patristic_matrix, self.internal_nodes, self.n_internal, 0.0, 1e-6, True)
        # the Schur complement of an OU covariance inherits lambda_min >= sigma_n^2 and a norm below the joint trace
        bound = ops.ou_cond_bound(squeeze_tensor(1, map_estimates["sigma_f"]), squeeze_tensor(1, map_estimates["sigma_n"]),
                                  self.n_all)
        if self._z_sharded():                       # cov holds this rank's latent dimensions only
            bound = bound[self.shard.z_lo:self.shard.z_hi]
        tril = ops.potrf(cov, cond_bound=bound) if cov.shape[0] else cov
        del cov
        self._cond_cache = (keys, stamp, mean, tril)
        return mean, tril

    def conditional_sampling(self, map_estimates, patristic_matrix):
        """One draw of the internal nodes given the leaves (S/models.py:149-196; Bishop B.49-50).  The reference adds
        1e-6 to every entry of the conditional covariance (:193); so does this."""
        return self.conditional_sampling_many(map_estimates, patristic_matrix, 1)[0]

    def conditional_sampling_many(self, map_estimates, patristic_matrix, n_samples: int, eps=None):
        """`n_samples` independent draws `[S, ni, z]` of the internal nodes given the leaves from ONE factorisation:
        `mean + L eps` as a single batched product (replaces S x `conditional_sampling`).  `eps`
        (`[z, ni, S]` standard-normal numbers) can be supplied to reproduce a draw."""
        assert patristic_matrix[1:, 1:].shape == (self.n_all, self.n_all), \
            "Remember to use the entire/full patristic matrix for conditional sampling!"
        mean, tril = self._conditional_factor(map_estimates, patristic_matrix)
        if eps is None:
            eps = torch.randn((self.z_dim, self.n_internal, int(n_samples)), dtype=mean.dtype, device=mean.device)
            if self._z_sharded():
                eps = self.shard.broadcast_from_rank0(eps)                      # all ranks must draw the same numbers
        assert eps.shape == (self.z_dim, self.n_internal, int(n_samples))
        if self._z_sharded():                       # this rank's latent dimensions, then one all-gather of the draws
            lo, hi = self.shard.z_lo, self.shard.z_hi
            draws = self.shard.gather_z_rows(mean[lo:hi].unsqueeze(-1) + torch.bmm(tril, eps[lo:hi]))
        else:
            draws = mean.unsqueeze(-1) + torch.bmm(tril, eps)                   # [z, ni, S]
        latent_space = draws.permute(2, 1, 0).contiguous()
        assert latent_space.shape == (int(n_samples), self.n_internal, self.z_dim)
        return latent_space

    def conditional_samplingMAP(self, map_estimates, patristic_matrix):
        """Conditional MEAN of the internal nodes (S/models.py:197-242).  The reference also draws one sample from the
        conditional distribution and discards it (:239); all that is left of that draw is that the device's random
        generator advances by `[z, n_internal]` standard-normal numbers, which `map_consumes_rng` (default True = the
        reference's behaviour) reproduces, so that whatever is sampled AFTER a MAP call sees the same generator state."""
        assert patristic_matrix[1:, 1:].shape == (self.n_all, self.n_all), \
            "Remember to use the entire/full patristic matrix for conditional sampling!"
        mean, _ = self._conditional(map_estimates, patristic_matrix, self.internal_nodes, self.n_internal, 0.0, 0.0, False)
        assert mean.shape == (self.z_dim, self.n_internal)
        if self.map_consumes_rng:        # MultivariateNormal.sample() = loc + L @ _standard_normal(loc.shape)
            torch.distributions.utils._standard_normal(mean.shape, dtype=mean.dtype, device=mean.device)
        return mean.T

    def conditional_sampling_batch(self, map_estimates, patristic_matrix):
        """S/models.py:243-287 (batched variant: +1e-6 on the hyper-parameters, no jitter on the covariance)."""
        internal = self.internal_nodes_batch
        assert not self._z_sharded(), "the batched variant is outside the sharded path"
        mean, cov = self._conditional(map_estimates, patristic_matrix, internal, len(internal), 1e-6, 0.0, True)
        return dist.MultivariateNormal(mean, cov).to_event(1).sample().T


class DRAUPNIRModel_classic(DRAUPNIRModelClass):
    """S/models.py:289-401 — whole-tree model, GRU mapping function, BLOSUM embeddings."""

    def __init__(self, ModelLoad):
        DRAUPNIRModelClass.__init__(self, ModelLoad)
        self.rnn_input_size = self.z_dim + self.aa_probs
        self.num_layers = 1
        self.decoder = RNNDecoder_Tiling(self.align_seq_len, self.aa_probs, self.gru_hidden_dim, self.z_dim,
                                         self.rnn_input_size, self.kappa_addition, self.num_layers, self.pretrained_params)
        self.embed = EmbedComplex(self.aa_probs, self.embedding_dim, self.pretrained_params)
        self.to(device=self.device, dtype=torch.float64)

    def _decoder_input(self, latent_space):
        """`[n, L, z + aa_probs]` GRU input of S/models.py:339-342; the embedding runs once on `[L, aa_probs]`."""
        n = latent_space.shape[0]
        blosum = self.embed(self.blosum_weighted)                                     # [L, A]
        return torch.cat((latent_space[:, None, :].expand(n, self.align_seq_len, self.z_dim),
                          blosum[None].expand(n, self.align_seq_len, self.aa_probs)), dim=2)

    def _hidden(self, n):
        return self.h_0_MODEL.expand(self.decoder.num_layers * 2, n, self.gru_hidden_dim).contiguous()

    def _likelihood(self, latent_space, aminoacid_sequences):
        if self.shard is not None:   # decoder is data-parallel over contiguous blocks of leaves
            lo, hi = self.shard.leaf_lo, self.shard.leaf_hi
            latent_space, aminoacid_sequences = latent_space[lo:hi], aminoacid_sequences[lo:hi]
        scores = self.decoder.scores_structured(latent_space, self.embed(self.blosum_weighted),
                                                hidden=self._hidden(latent_space.shape[0]))
        pyro.sample("aa_sequences", dist.Categorical(logits=scores), obs=aminoacid_sequences)

    def model_variational(self, family_data, patristic_matrix_sorted, cladistic_matrix, data_blosum, batch_blosum=None,
                          guide_map_estimates=None):
        """S/models.py:300-325."""
        aminoacid_sequences = family_data[:, 2:, 0]
        pyro.module("embeddings", self.embed)
        pyro.module("decoder", self.decoder)
        with pyro.plate("plate_batch", dim=-1, device=self.device):
            latent_space = self.gp_prior(patristic_matrix_sorted)
            with pyro.plate("plate_len", dim=-2):
                self._likelihood(latent_space, aminoacid_sequences)

    def model_delta_map(self, family_data, patristic_matrix_sorted, cladistic_matrix, data_blosum, batch_blosum=None,
                        guide_map_estimates=None):
        """S/models.py:328-352."""
        aminoacid_sequences = family_data[:, 2:, 0]
        pyro.module("embeddings", self.embed)
        pyro.module("decoder", self.decoder)
        latent_space = self.gp_prior(patristic_matrix_sorted)
        with pyro.plate("plate_len", aminoacid_sequences.shape[1], dim=-1), \
                pyro.plate("plate_seq", aminoacid_sequences.shape[0], dim=-2):
            self._likelihood(latent_space, aminoacid_sequences)

    def model(self, family_data, patristic_matrix_sorted, cladistic_matrix, data_blosum, batch_blosum, guide_map_estimates):
        """S/models.py:355-359."""
        if self.args.select_guide == "delta_map":
            self.model_delta_map(family_data, patristic_matrix_sorted, cladistic_matrix, data_blosum, batch_blosum,
                                 guide_map_estimates)
        else:
            self.model_variational(family_data, patristic_matrix_sorted, cladistic_matrix, data_blosum, batch_blosum,
                                   guide_map_estimates)

    @torch.no_grad()
    def sample(self, map_estimates, n_samples, family_data_test, patristic_matrix, cladistic_matrix, use_argmax=False,
               use_test=True, use_test2=False):
        """S/models.py:361-401."""
        if use_test2:
            assert patristic_matrix[1:, 1:].shape == (self.n_all, self.n_all)
            latent_space = self.conditional_samplingMAP(map_estimates, patristic_matrix)
            n_nodes = self.n_internal
        elif use_test:
            assert patristic_matrix[1:, 1:].shape == (self.n_all, self.n_all)
            latent_space = self.conditional_sampling(map_estimates, patristic_matrix)
            n_nodes = self.n_internal
        else:
            latent_space = map_estimates["latent_z"].T
            assert latent_space.shape == (self.n_leaves, self.z_dim)
            n_nodes = self.n_leaves
        logits = self.decoder.forward_structured(latent_space, self.embed(self.blosum_weighted), hidden=self._hidden(n_nodes))
        if use_argmax:
            aa_sequences = torch.argmax(logits, dim=2).unsqueeze(0)
        else:
            aa_sequences = dist.Categorical(logits=logits).sample([n_samples])
        return SamplingOutput(aa_sequences=aa_sequences.detach(), latent_space=latent_space.detach(),
                              logits=logits.detach(), phis=None, psis=None, mean_phi=None, mean_psi=None, kappa_phi=None,
                              kappa_psi=None)

    @torch.no_grad()
    def sample_marginal(self, map_estimates, n_samples, patristic_matrix, nodes_per_call: int = 8192):
        """The marginal ancestral samples of S/main.py:1334-1350 in one call: `n_samples` draws of the internal nodes'
        latent vectors from one conditional factorisation, each decoded and followed by one categorical draw — what
        the reference obtains from `n_samples` x `sample(map_estimates, 1, ..., use_test=True)`.  Returns a
        `SamplingOutput` with `aa_sequences [S, ni, L]`, `latent_space [S, ni, z]`, `logits [S, ni, L, A]`; the decoder
        runs over all draws as one batch of sequences (`nodes_per_call` bounds the rows per launch)."""
        S = int(n_samples)
        latent = self.conditional_sampling_many(map_estimates, patristic_matrix, S)          # [S, ni, z]
        flat = latent.reshape(S * self.n_internal, self.z_dim)
        blosum = self.embed(self.blosum_weighted)
        logits = torch.empty((flat.shape[0], self.align_seq_len, self.aa_probs), dtype=flat.dtype, device=flat.device)
        for lo in range(0, flat.shape[0], nodes_per_call):
            hi = min(lo + nodes_per_call, flat.shape[0])
            logits[lo:hi] = self.decoder.forward_structured(flat[lo:hi], blosum, hidden=self._hidden(hi - lo))
        aa = dist.Categorical(logits=logits).sample()
        return SamplingOutput(aa_sequences=aa.reshape(S, self.n_internal, self.align_seq_len),
                              latent_space=latent, logits=logits.reshape(S, self.n_internal, self.align_seq_len, self.aa_probs),
                              phis=None, psis=None, mean_phi=None, mean_psi=None, kappa_phi=None, kappa_psi=None)
