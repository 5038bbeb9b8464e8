"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot be imported in this image — no pyro /
ignite / Bio / ete3 / R — so the oracle, pinned by the closed-form and cross-derivation checks in
tests/test_oracle.py, is what produces the vectors).  Run from the repo root:  python tests/golden/make_golden.py

Fixtures (all seeded, fp64):
  c1_delta_map.npz / c1_variational.npz : Randall-CFP shape (19 leaves, 225 sites, z=30): per-site log-probs, -ELBO,
        gradient norms per parameter group, three SVI losses, conditional mean, argmax ancestral sequences
  tree_64.npz : 64-leaf tree: parent / branch / LCA matrix / patristic / cladistic matrices
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from draupnir_asr_b200 import synthetic as S  # noqa: E402
from oracle import draupnir_oracle as O  # noqa: E402
from oracle import patristic_oracle as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def c1(guide):
    pr = S.make_problem(19, 225, seed=0)
    init = S.make_parameters(19, seed=0)
    nets = O.Networks(variational=(guide == "variational"), seed=0)
    svi = O.OracleSVI(nets, init, guide)
    eps = torch.randn(30, 19, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    args = (pr.dataset_train, pr.patristic_matrix_train, pr.dataset_train_blosum, pr.blosum_weighted, eps)
    loss, terms, logits = svi.loss(*args)
    loss.backward()
    out = {"loss0": float(loss.detach()), "eps": eps.numpy(), "logits_sum": float(logits.sum())}
    out.update({f"term_{k}": float(v) for k, v in terms.items()})
    out.update({f"gradnorm_{k}": float(u.grad.norm()) for k, u in svi.u.items()})
    out["gradnorm_decoder"] = float(torch.cat([p.grad.reshape(-1) for p in nets.decoder.parameters()]).norm())
    out["gradnorm_embed"] = float(torch.cat([p.grad.reshape(-1) for p in nets.embed.parameters()]).norm())
    for u in svi.u.values():
        u.grad = None
    for p in nets.parameters():
        p.grad = None
    out["losses"] = np.array([svi.step(*args) for _ in range(3)])
    vals = {k: v.detach() for k, v in svi.constrained().items()}
    if guide == "variational":
        vals["latent_z"] = init["latent_z"]
        vals = {k: O.squeeze_tensor(1, v) if k != "latent_z" else v for k, v in vals.items()}
    mean = O.conditional_samplingMAP(vals, pr.patristic_matrix_full, pr.internal_nodes)
    out["cond_mean"] = mean.numpy()
    out["argmax_internal"] = O.sample(nets, mean, pr.blosum_weighted).aa_sequences.numpy()
    np.savez_compressed(os.path.join(HERE, f"c1_{guide}.npz"), **out)


def tree():
    t = S.random_tree(64, seed=11)
    lca, pat, cla = P.lca_parent_walk(t.parent, t.branch, t.depth, range(t.n_nodes), range(t.n_nodes))
    np.savez_compressed(os.path.join(HERE, "tree_64.npz"), parent=t.parent, branch=t.branch, depth=t.depth, lca=lca,
                        patristic=pat, cladistic=cla)


if __name__ == "__main__":
    torch.set_num_threads(1)          # bitwise-stable reductions for the committed vectors
    c1("delta_map")
    c1("variational")
    tree()
    print("golden fixtures written to", HERE)
