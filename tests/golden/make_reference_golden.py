"""Generates tests/golden/ref_*.npz by running the REFERENCE'S OWN code (read in place from /root/reference, nothing
copied) on the CPU in fp64 — see tests/reference_harness.py for exactly what is real (models.py, models_utils.py,
guides.py, train.py, utils.py and torch under them) and what is substituted (`pyro` -> draupnir_asr_b200.pyro_compat,
since pyro-ppl is not installed; unused plotting / tree libraries -> empty stand-ins).

Run in the build container (the only place /root/reference exists), from the repo root:
    python tests/golden/make_reference_golden.py

Every case is seeded; inputs are regenerated from the seeds by the tests (draupnir_asr_b200.synthetic, oracle.Networks),
only the standard-normal draws and the reference's OUTPUTS are stored:
  loss0, term_<site>            -ELBO and per-site log-prob sums of the first evaluation (Trace_ELBO over the
                                reference's model / guide)
  grad_<site> / gradnorm_<net>  gradients w.r.t. the constrained-space-independent (unconstrained) guide parameters,
                                and gradient norms per network
  losses                        three consecutive `train(svi, ...)` calls (S/train.py:58-79)
  cond_mean                     `Draupnir.conditional_samplingMAP` (S/models.py:197-242)         [ni, z]
  cond_draw (+ cond_eps)        `Draupnir.conditional_sampling` (S/models.py:149-196) for the stored draw
  argmax_internal / _leaves     `Draupnir.sample(..., use_argmax=True, use_test2=True / use_test=False)` (S/models.py:361-401)
  logits_internal_lse0          a checksum of the internal-node logits
  guide_keys                    keys of the dictionary the guide returns (S/guides.py:123-133 / AutoDelta)
  blosum_weighted, blosum_max, variable_score, aa_frequencies, dataset_train_blosum_sum
                                `process_blosum`, `calculate_aa_frequencies_torch`, `blosum_embedding_encoder`
                                (S/utils.py:1752-1760,1820-1851) on the synthetic alignment
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reference_harness as H  # noqa: E402
from draupnir_asr_b200 import synthetic as S  # noqa: E402
from oracle import draupnir_oracle as O  # noqa: E402

CASES = {  # name: (n_leaves, sites, guide, problem seed)
    "ref_c1_delta_map": (19, 225, "delta_map", 0),
    "ref_c1_variational": (19, 225, "variational", 0),
    "ref_n48_delta_map": (48, 64, "delta_map", 3),
    "ref_n48_variational": (48, 64, "variational", 3),
}


def _clone_state(st):
    return {k: ({kk: vv.detach().clone() for kk, vv in v.items()} if isinstance(v, dict) else v.detach().clone())
            for k, v in st.items()}


def run_case(n, L, guide, seed):
    torch.set_default_dtype(torch.float64)
    pr = S.make_problem(n, L, seed=seed)
    init = S.make_parameters(n, seed=seed)
    nets = O.Networks(variational=(guide == "variational"), seed=seed)       # only the source of the initial weights
    run = H.build_reference_run(pr, guide, init, _clone_state(H.nets_state_of(nets)))
    ref, D = run.ref, run.Draupnir
    ni = pr.n_internal
    eps = torch.randn(30, n, generator=torch.Generator().manual_seed(5))
    cond_eps = torch.randn(30, ni, generator=torch.Generator().manual_seed(6))
    out = {"eps": eps.numpy(), "cond_eps": cond_eps.numpy()}

    # ---- first ELBO evaluation: loss, per-site terms, gradients -------------------------------------------------
    with H.FixedNoise(eps):
        loss = run.svi.loss.differentiable_loss(run.svi.model, run.svi.guide, *run.args)
    model_trace, guide_trace = run.svi.loss.last_traces
    out["loss0"] = float(loss.detach())
    for name, site in model_trace.sample_sites():
        out[f"term_{name}"] = float(site["log_prob_sum"].detach())
    out["term_log_q"] = float(sum(float(torch.as_tensor(site["log_prob_sum"]).detach())
                                  for _, site in guide_trace.sample_sites()))
    loss.backward()
    for k in ("alpha", "sigma_f", "sigma_n", "lambd") + (("latent_z",) if guide == "delta_map" else ()):
        u = getattr(run.guide, f"{k}_unconstrained")
        out[f"grad_{k}"] = u.grad.detach().reshape(-1).numpy().copy() if k != "latent_z" else u.grad.detach().numpy().copy()
        u.grad = None
    groups = {"decoder": D.decoder, "embed": D.embed}
    if guide == "variational":
        groups.update(encoder=run.guide.encoder, embeddingencoder=run.guide.embeddingencoder)
    for k, mod in groups.items():
        out[f"gradnorm_{k}"] = float(torch.cat([p.grad.reshape(-1) for p in mod.parameters()]).norm())
        for p in mod.parameters():
            p.grad = None

    # ---- three training epochs through the reference's own loop --------------------------------------------------
    losses = []
    for _ in range(3):
        with H.FixedNoise(eps):
            losses.append(ref.train.train(run.svi, run.train_input))
    out["losses"] = np.array(losses)

    # ---- the guide's dictionary and the ancestral reconstruction (S/main.py:1097,1113-1170) -----------------------
    with torch.no_grad(), H.FixedNoise(eps):
        map_estimates = run.guide(*run.args)
    out["guide_keys"] = np.array(sorted(map_estimates.keys()))
    map_estimates = {k: v.detach() for k, v in map_estimates.items()}
    out["map_latent_z"] = map_estimates["latent_z"].numpy().copy()
    with torch.no_grad():
        out["cond_mean"] = D.conditional_samplingMAP(map_estimates, pr.patristic_matrix_full).numpy().copy()
        with H.FixedNoise(cond_eps):
            out["cond_draw"] = D.conditional_sampling(map_estimates, pr.patristic_matrix_full).numpy().copy()
        so = D.sample(map_estimates, 1, pr.dataset_test, pr.patristic_matrix_full, None, use_argmax=True, use_test=False,
                      use_test2=True)
        out["argmax_internal"] = so.aa_sequences.numpy().copy()
        out["logits_internal_lse0"] = float(torch.logsumexp(so.logits[..., 0].reshape(-1), 0))
        so = D.sample(map_estimates, 1, pr.dataset_train, pr.patristic_matrix_full, None, use_argmax=True, use_test=False,
                      use_test2=False)
        out["argmax_leaves"] = so.aa_sequences.numpy().copy()

    # ---- input pipeline helpers the synthetic generator restates ---------------------------------------------------
    U = ref.utils
    residues = pr.dataset_train[:, 2:, 0]
    freqs = U.calculate_aa_frequencies_torch(residues, pr.aa_probs)
    bmax, bw, vs = U.process_blosum(pr.blosum, freqs, pr.align_seq_len, pr.aa_probs)
    enc = U.blosum_embedding_encoder(pr.blosum, freqs, pr.align_seq_len, pr.aa_probs, pr.dataset_train, False)
    out.update(aa_frequencies=freqs.numpy(), blosum_max=bmax.numpy(), blosum_weighted=bw.numpy(),
               variable_score=vs.numpy(), dataset_train_blosum_sum=enc.sum(-1).numpy())
    return out


if __name__ == "__main__":
    torch.set_num_threads(1)          # bitwise-stable reductions for the committed vectors
    for name, (n, L, guide, seed) in CASES.items():
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **run_case(n, L, guide, seed))
        print("wrote", name)
