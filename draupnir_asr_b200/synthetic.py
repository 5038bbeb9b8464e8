"""Seeded synthetic inputs in the reference's tensor layouts.

Nothing here computes anything on the hot path: it only manufactures the tree,
the alignment and the host-side tensors that `S/main.py:load_data` +
`S/load_utils.py:pretreatment` would hand to the model (S/ =
/root/reference/draupnir/src/draupnir/).  Layouts reproduced:

* node ids are BFS level-order positions, root = 0   (S/utils.py:864,879)
* `dataset_train [n, L+2, 30]`: `[:,0,0]` sequence length, `[:,0,1]` node id,
  `[:,0,2]` branch length, `[:,2:,0]` residues 0..20  (S/utils.py:866-884)
* patristic matrices `(N+1)x(N+1)` with a node-id header row/column and -inf
  in the corner, sorted ascending by node id (S/utils.py:971-979,
  S/load_utils.py:452-466); the train matrix is trimmed to the leaves
* `blosum [22,22]` with the same header convention (S/utils.py:129-131),
  residue order `-RHKDESTNQCGPAVILMFYW` (S/utils.py:97)
* `aa_frequencies [L,21]` (S/utils.py:1744-1751), `blosum_max`,
  `blosum_weighted [L,21]`, `variable_score [L]` (S/utils.py:1820-1840),
  `dataset_train_blosum [n,L,21]` (S/utils.py:1841-1851)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

AA_ORDER = "-RHKDESTNQCGPAVILMFYW"  # S/utils.py:97 (21 symbols, gap = 0)

# BLOSUM62 in the usual NCBI order; '*' row/col is what the reference uses for gaps
# (S/utils.py:110-112).
_B62_ORDER = "ARNDCQEGHILKMFPSTWYV*"
_B62_ROWS = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -4
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -4
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3 -4
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3 -4
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -4
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2 -4
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2 -4
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -4
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3 -4
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -4
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2 -4
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -4
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -4
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -4
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2 -4
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -4
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -4
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -4
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""


def blosum62_array(aa_probs: int = 21) -> np.ndarray:
    """`[aa_probs+1, aa_probs+1]` BLOSUM62 with the id header (S/utils.py:105-131)."""
    assert aa_probs == 21, "only the 21-symbol alphabet of the reference's default path"
    raw = np.array([[int(v) for v in row.split()] for row in _B62_ROWS.strip().splitlines()], dtype=np.float64)
    pos = {c: i for i, c in enumerate(_B62_ORDER)}
    order = [pos["*" if c == "-" else c] for c in AA_ORDER]
    sub = raw[np.ix_(order, order)]
    out = np.zeros((aa_probs + 1, aa_probs + 1), dtype=np.float64)
    out[0, 0] = -np.inf
    out[0, 1:] = np.arange(aa_probs)
    out[1:, 0] = np.arange(aa_probs)
    out[1:, 1:] = sub
    return out


@dataclass
class SyntheticTree:
    """Rooted binary tree, nodes numbered in BFS level order (root = 0)."""
    parent: np.ndarray      # int32 [N], parent[0] = -1, parent[i] < i
    branch: np.ndarray      # float64 [N], length of the edge to the parent (root: 0)
    depth: np.ndarray       # int32 [N], edge count from the root
    root_dist: np.ndarray   # float64 [N], sum of branch lengths from the root
    is_leaf: np.ndarray     # bool [N]

    @property
    def n_nodes(self) -> int:
        return int(self.parent.shape[0])

    @property
    def leaves(self) -> np.ndarray:
        return np.nonzero(self.is_leaf)[0].astype(np.int64)

    @property
    def internal(self) -> np.ndarray:
        return np.nonzero(~self.is_leaf)[0].astype(np.int64)


def random_tree(n_leaves: int, seed: int = 0, shape: str = "agglomerate",
                branch_lo: float = 0.01, branch_hi: float = 0.5) -> SyntheticTree:
    """Random rooted binary tree with `n_leaves` leaves and `2n-1` nodes.

    shape = "agglomerate" (join two uniformly chosen roots),
            "caterpillar" (maximally unbalanced; depth n-1, an edge case for the LCA walk) or
            "balanced".
    """
    assert n_leaves >= 2
    rng = np.random.default_rng(seed)
    n_nodes = 2 * n_leaves - 1
    # build with temporary labels: leaves 0..n-1, internal n..2n-2, last created = root
    children = {}
    roots = list(range(n_leaves))
    nxt = n_leaves
    while len(roots) > 1:
        if shape == "agglomerate":
            i, j = rng.choice(len(roots), size=2, replace=False)
        elif shape == "caterpillar":
            i, j = len(roots) - 1, 0
        elif shape == "balanced":
            i, j = 0, 1
        else:
            raise ValueError(shape)
        a, b = roots[i], roots[j]
        for k in sorted((int(i), int(j)), reverse=True):
            roots.pop(k)
        children[nxt] = (a, b)
        roots.append(nxt)
        nxt += 1
    root = roots[0]
    # BFS relabel: position in level order == node id (S/utils.py:864)
    order = [root]
    head = 0
    tmp_parent = {root: -1}
    while head < len(order):
        v = order[head]
        head += 1
        for c in children.get(v, ()):
            tmp_parent[c] = v
            order.append(c)
    new_id = {v: k for k, v in enumerate(order)}
    parent = np.full(n_nodes, -1, dtype=np.int32)
    is_leaf = np.zeros(n_nodes, dtype=bool)
    for v in order:
        k = new_id[v]
        parent[k] = -1 if tmp_parent[v] < 0 else new_id[tmp_parent[v]]
        is_leaf[k] = v not in children
    branch = rng.uniform(branch_lo, branch_hi, size=n_nodes)
    branch[0] = 0.0
    depth = np.zeros(n_nodes, dtype=np.int32)
    root_dist = np.zeros(n_nodes, dtype=np.float64)
    for k in range(1, n_nodes):
        depth[k] = depth[parent[k]] + 1
        root_dist[k] = root_dist[parent[k]] + branch[k]
    return SyntheticTree(parent=parent, branch=branch, depth=depth, root_dist=root_dist, is_leaf=is_leaf)


def evolve_alignment(tree: SyntheticTree, seq_len: int, seed: int = 0, gap_p: float = 0.05) -> np.ndarray:
    """Residues 0..20 for every node `[N, L]` (gaps only at leaves), evolved down the tree."""
    rng = np.random.default_rng(seed + 1)
    N = tree.n_nodes
    x = np.zeros((N, seq_len), dtype=np.int64)
    x[0] = rng.integers(1, 21, size=seq_len)
    for k in range(1, N):
        p_sub = 1.0 - np.exp(-tree.branch[k])
        mut = rng.random(seq_len) < p_sub
        x[k] = np.where(mut, rng.integers(1, 21, size=seq_len), x[tree.parent[k]])
    gaps = rng.random((N, seq_len)) < gap_p
    gaps[~tree.is_leaf] = False
    x[gaps] = 0
    return x


def patristic_host(tree: SyntheticTree) -> np.ndarray:
    """All-pairs path length `[N,N]` by propagating down the tree (host helper for building inputs;
    the product builder is the CUDA kernel in `patristic.py`, the checker is `oracle/`)."""
    N = tree.n_nodes
    d = np.zeros((N, N), dtype=np.float64)
    for k in range(1, N):
        p = tree.parent[k]
        d[k, :k] = d[p, :k] + tree.branch[k]
        d[:k, k] = d[k, :k]
    return d


def with_id_header(matrix: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Prepend the node-id header row/column, -inf in the corner (S/utils.py:971-979)."""
    m = matrix.shape[0]
    out = np.empty((m + 1, m + 1), dtype=np.float64)
    out[0, 0] = -np.inf
    out[0, 1:] = ids
    out[1:, 0] = ids
    out[1:, 1:] = matrix
    return out


def process_blosum(blosum: torch.Tensor, aa_freqs: torch.Tensor, align_seq_len: int, aa_probs: int):
    """`blosum_max`, `blosum_weighted`, `variable_score` with the semantics of S/utils.py:1820-1840."""
    sub = blosum[1:, 1:]                                             # [A, A]
    top = torch.argmax(aa_freqs, dim=1)                              # [L]
    blosum_max = sub[top]                                            # row of the most frequent residue
    blosum_weighted = (aa_freqs[:, :, None] * sub[None]).mean(dim=1)  # [L, A]
    variable_score = torch.count_nonzero(aa_freqs, dim=1) / aa_probs
    return blosum_max, blosum_weighted, variable_score


def blosum_embedding_encoder(blosum: torch.Tensor, residues: torch.Tensor) -> torch.Tensor:
    """Each residue replaced by its BLOSUM column `[n, L, A]` (S/utils.py:1841-1851)."""
    sub = blosum[1:, 1:]
    return sub.T[residues.to(torch.int64)].contiguous()


@dataclass
class SyntheticProblem:
    tree: SyntheticTree
    n_leaves: int
    n_internal: int
    align_seq_len: int
    aa_probs: int
    leaves_nodes: torch.Tensor              # fp64 [n]   = dataset_train[:,0,1]
    internal_nodes: torch.Tensor            # fp64 [ni]  = patristic_matrix_test[1:,0]
    dataset_train: torch.Tensor             # fp64 [n, L+2, 30]
    dataset_test: torch.Tensor              # fp64 [ni, L+2, 30] (true ancestral residues)
    patristic_matrix_train: torch.Tensor    # fp64 [n+1, n+1]
    patristic_matrix_full: torch.Tensor     # fp64 [N+1, N+1]
    blosum: torch.Tensor                    # fp64 [22, 22]
    aa_frequencies: torch.Tensor            # fp64 [L, 21]
    blosum_max: torch.Tensor
    blosum_weighted: torch.Tensor           # fp64 [L, 21]
    variable_score: torch.Tensor
    dataset_train_blosum: torch.Tensor      # fp64 [n, L, 21]


def _dataset(ids: np.ndarray, residues: np.ndarray, tree: SyntheticTree) -> torch.Tensor:
    n, L = residues.shape
    d = np.zeros((n, L + 2, 30), dtype=np.float64)
    d[:, 0, 0] = (residues != 0).sum(1)
    d[:, 0, 1] = ids
    d[:, 0, 2] = tree.branch[ids]
    d[:, 2:, 0] = residues
    return torch.from_numpy(d)


def patristic_device(tree: SyntheticTree, device="cuda") -> np.ndarray:
    """The same `[N,N]` matrix from the product's builder (K1: `ops.tree_distances`, the LCA kernel that replaces the
    reference's ete3 / R path, S/utils.py:496-579), brought back to the host in the layout `make_problem` stores."""
    from . import ops
    parent = torch.from_numpy(tree.parent).to(device)
    branch = torch.from_numpy(tree.branch).to(device)
    return ops.tree_distances(parent, branch)[0].cpu().numpy()


def make_problem(n_leaves: int, seq_len: int, seed: int = 0, shape: str = "agglomerate",
                 aa_probs: int = 21, patristic_full: Optional[np.ndarray] = None, builder: str = "host",
                 device="cuda") -> SyntheticProblem:
    """Tree + alignment + every host tensor the model/guide constructors and `svi.step` consume.  `builder` = "host"
    (numpy propagation; CPU tests, the oracle's and the reference arm's inputs) or "device" (the CUDA patristic builder:
    what bench.py's own arm and smoke() feed the model with)."""
    tree = random_tree(n_leaves, seed=seed, shape=shape)
    x = evolve_alignment(tree, seq_len, seed=seed)
    leaves, internal = tree.leaves, tree.internal
    if patristic_full is not None:
        d_full = patristic_full
    elif builder == "device":
        d_full = patristic_device(tree, device)
    else:
        assert builder == "host", builder
        d_full = patristic_host(tree)
    pm_full = torch.from_numpy(with_id_header(d_full, np.arange(tree.n_nodes)))
    pm_train = torch.from_numpy(with_id_header(d_full[np.ix_(leaves, leaves)], leaves))
    ds_train = _dataset(leaves, x[leaves], tree)
    ds_test = _dataset(internal, x[internal], tree)
    blosum = torch.from_numpy(blosum62_array(aa_probs))
    res = torch.from_numpy(x[leaves])
    freqs = torch.stack([torch.bincount(res[:, l], minlength=aa_probs) for l in range(seq_len)]).to(torch.float64) / n_leaves
    b_max, b_w, v = process_blosum(blosum, freqs, seq_len, aa_probs)
    return SyntheticProblem(
        tree=tree, n_leaves=n_leaves, n_internal=len(internal), align_seq_len=seq_len, aa_probs=aa_probs,
        leaves_nodes=ds_train[:, 0, 1].clone(), internal_nodes=pm_full[1:, 0][torch.from_numpy(internal)].clone(),
        dataset_train=ds_train, dataset_test=ds_test, patristic_matrix_train=pm_train, patristic_matrix_full=pm_full,
        blosum=blosum, aa_frequencies=freqs, blosum_max=b_max, blosum_weighted=b_w, variable_score=v,
        dataset_train_blosum=blosum_embedding_encoder(blosum, res))


def make_parameters(n_leaves: int, z_dim: int = 30, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Latent-site values injected identically into oracle and CUDA path."""
    g = torch.Generator().manual_seed(seed + 7)
    u = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64) + 0.5
    return {
        "alpha": torch.ones(3, dtype=torch.float64),
        "sigma_f": u(z_dim), "sigma_n": u(z_dim), "lambd": u(z_dim),
        "latent_z": torch.randn(z_dim, n_leaves, generator=g, dtype=torch.float64),
    }
