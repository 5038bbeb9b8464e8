"""Tree and distance-matrix FILE FORMATS either side of the patristic builder.

The reference turns a Newick tree into `<name>_patristic_distance_matrix.csv` (+ `_cladistic_`) and
`<name>_tree_levelorder_info.csv` with ete3 + an O(N^2 depth) pandas loop or an R/ape subprocess
(S/utils.py:496-579, 857-897; S/ = /root/reference/draupnir/src/draupnir/), and reads them back in `load_data`
(S/main.py:196-207) into the `(N+1) x (N+1)` matrix with a node-id header that the model consumes
(S/utils.py:945-979, S/load_utils.py:452-466).  This module reads the Newick text directly into level-order parent /
branch arrays, builds the matrices with the CUDA LCA kernel (`ops.tree_distances`) and writes / reads the same files, so
the GPU builder slots into the reference's data preparation without ete3, R or a CSV round trip in between.

Conventions taken from the reference (ete3 itself is not available in this image, so they are restated, not imported):
  * node ids are positions in ete3's default `tree.traverse()` order = breadth-first level order, root 0, children in
    file order (S/utils.py:864);
  * unnamed internal nodes are called `A<k>`, k counting up from the number of leaves in level order (S/utils.py:638-651);
  * the matrix files list the internal nodes (level order) first, then the leaves in ete3's `get_leaf_names()` order =
    depth-first pre-order (S/utils.py:855-856, 895);
  * a branch without a length is 1.0 and the root's own branch is 0.0 (ete3 defaults).
The small-tree branch of the reference fills only the upper triangle and symmetrises on load (S/utils.py:572-575, 939);
the files written here hold the full symmetric matrix, which that loader maps to itself.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

DEFAULT_DIST = 1.0


@dataclass
class LevelOrderTree:
    """Rooted tree (any arity) with nodes in breadth-first level order (root = 0, `parent[i] < i`)."""
    names: List[str]
    parent: np.ndarray          # int32 [N], root: -1
    branch: np.ndarray          # float64 [N], length of the edge to the parent
    children: List[List[int]]

    @property
    def n_nodes(self) -> int:
        return len(self.names)

    @property
    def is_leaf(self) -> np.ndarray:
        return np.array([len(c) == 0 for c in self.children], dtype=bool)

    @property
    def internal_ids(self) -> np.ndarray:
        return np.nonzero(~self.is_leaf)[0].astype(np.int32)

    @property
    def leaf_ids_preorder(self) -> np.ndarray:
        """Leaves in depth-first pre-order (ete3 `get_leaf_names()` order)."""
        out, stack = [], [0]
        while stack:
            v = stack.pop()
            if not self.children[v]:
                out.append(v)
            else:
                stack.extend(reversed(self.children[v]))
        return np.asarray(out, dtype=np.int32)

    def ancestors(self, i: int) -> List[int]:
        """Parent, grand-parent, ..., root (ete3 `node.get_ancestors()`)."""
        out = []
        p = int(self.parent[i])
        while p >= 0:
            out.append(p)
            p = int(self.parent[p])
        return out

    def file_order(self) -> np.ndarray:
        """Row/column order of the matrix files: internal nodes in level order, then leaves in pre-order."""
        return np.concatenate((self.internal_ids, self.leaf_ids_preorder)).astype(np.int32)


# ----------------------------------------------------------------------------------------------------------------------
# Newick
# ----------------------------------------------------------------------------------------------------------------------
class NewickError(ValueError):
    pass


def _tokenise(text: str):
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c.isspace():
            i += 1
        elif c in "(),:;":
            yield c, None
            i += 1
        elif c == "[":                                   # comment / NHX block
            j = text.find("]", i)
            if j < 0:
                raise NewickError("unterminated [comment]")
            i = j + 1
        elif c == "'":                                   # quoted label, '' escapes a quote
            j, buf = i + 1, []
            while True:
                if j >= n:
                    raise NewickError("unterminated quoted label")
                if text[j] == "'":
                    if j + 1 < n and text[j + 1] == "'":
                        buf.append("'")
                        j += 2
                        continue
                    break
                buf.append(text[j])
                j += 1
            yield "label", "".join(buf)
            i = j + 1
        else:
            j = i
            while j < n and text[j] not in "(),:;[" and not text[j].isspace():
                j += 1
            yield "label", text[i:j]
            i = j


def read_newick(source: str) -> LevelOrderTree:
    """Parse Newick text (or the file at `source`) into level-order arrays.  Labels of internal nodes are kept as names."""
    text = source
    if "(" not in source and os.path.exists(source):
        with open(source) as f:
            text = f.read()
    # build a pointer tree in file order
    kids: List[List[int]] = [[]]
    name: List[str] = [""]
    dist: List[Optional[float]] = [None]
    up: List[int] = [-1]
    cur = 0                 # node whose label / length is being read
    stack: List[int] = []   # open internal nodes
    expect_len = False
    seen_any, finished = False, False

    def new_child(of: int) -> int:
        kids.append([])
        name.append("")
        dist.append(None)
        up.append(of)
        kids[of].append(len(kids) - 1)
        return len(kids) - 1

    for tok, val in _tokenise(text):
        if finished:
            raise NewickError("text after the closing ';'")
        seen_any = True
        if tok == "(":
            if expect_len:
                raise NewickError("'(' where a branch length was expected")
            if name[cur] or kids[cur]:
                raise NewickError("'(' after a label")
            stack.append(cur)
            cur = new_child(cur)
        elif tok == ",":
            if not stack:
                raise NewickError("',' outside parentheses")
            cur = new_child(stack[-1])
            expect_len = False
        elif tok == ")":
            if not stack:
                raise NewickError("unbalanced ')'")
            cur = stack.pop()
            expect_len = False
        elif tok == ":":
            expect_len = True
        elif tok == "label":
            if expect_len:
                try:
                    dist[cur] = float(val)
                except ValueError:
                    raise NewickError(f"bad branch length {val!r}") from None
                expect_len = False
            else:
                name[cur] = val
        elif tok == ";":
            if stack:
                raise NewickError("unbalanced '('")
            finished = True
    if not seen_any or not finished:
        raise NewickError("a Newick tree ends with ';'")
    # breadth-first relabelling
    order, head = [0], 0
    while head < len(order):
        order.extend(kids[order[head]])
        head += 1
    new_id = {v: k for k, v in enumerate(order)}
    N = len(order)
    parent = np.full(N, -1, dtype=np.int32)
    branch = np.zeros(N, dtype=np.float64)
    for v in order:
        k = new_id[v]
        parent[k] = -1 if up[v] < 0 else new_id[up[v]]
        branch[k] = (0.0 if up[v] < 0 else DEFAULT_DIST) if dist[v] is None else dist[v]
    return LevelOrderTree(names=[name[v] for v in order], parent=parent, branch=branch,
                          children=[[new_id[c] for c in kids[v]] for v in order])


def rename_tree_internal_nodes(tree: LevelOrderTree) -> LevelOrderTree:
    """Internal nodes become `A<k>`, k = n_leaves, n_leaves + 1, ... in level order (S/utils.py:638-651)."""
    edge = int(tree.is_leaf.sum())
    names = list(tree.names)
    for i in range(tree.n_nodes):
        if tree.children[i]:
            names[i] = "A%d" % edge
            edge += 1
    return LevelOrderTree(names=names, parent=tree.parent, branch=tree.branch, children=tree.children)


def _quote(label: str) -> str:
    return "'" + label.replace("'", "''") + "'" if any(c in label for c in "(),:;[]' \t\n") else label


def to_newick(tree: LevelOrderTree, internal_names: bool = True, lengths: bool = True, root_name: bool = False) -> str:
    """Newick text; (`True`, `True`) ~ ete3 format 1, (`True`, `False`) ~ format 8 (S/utils.py:527-530)."""
    def fmt(x: float) -> str:
        return repr(float(x))

    out: List[str] = []
    stack: List[Tuple[int, int]] = [(0, 0)]        # (node, next child index)
    while stack:
        v, k = stack.pop()
        ch = tree.children[v]
        if ch and k == 0:
            out.append("(")
        if k < len(ch):
            if k:
                out.append(",")
            stack.append((v, k + 1))
            stack.append((ch[k], 0))
            continue
        if ch:
            out.append(")")
        if not ch or (internal_names and (v != 0 or root_name)):
            out.append(_quote(tree.names[v]))
        if lengths and v != 0:
            out.append(":" + fmt(tree.branch[v]))
    return "".join(out) + ";"


# ----------------------------------------------------------------------------------------------------------------------
# matrices
# ----------------------------------------------------------------------------------------------------------------------
def distance_matrices(tree: LevelOrderTree, device="cuda", want_cladistic: bool = False):
    """Patristic (and cladistic = edge-count) matrices in FILE order as device tensors, built by the LCA kernel
    (replaces the ete3 double loop / `Calculate_Patristic.R`, S/utils.py:509-575).  Returns `(order, patristic,
    cladistic_or_None)`; `order[k]` is the level-order id of row/column k."""
    import torch
    from . import ops
    from .utils import require_cuda
    dev = require_cuda(device)
    order = tree.file_order()
    parent = torch.from_numpy(tree.parent).to(dev)
    branch = torch.from_numpy(tree.branch).to(dev)
    idx = torch.from_numpy(order).to(dev)
    pat, cla, _ = ops.tree_distances(parent, branch, rows=idx, cols=idx, want_cladistic=want_cladistic)
    return order, pat, cla


def _write_matrix_csv(path: str, labels: Sequence[str], matrix: np.ndarray):
    import pandas as pd
    I = pd.Index(list(labels), name="rows")
    C = pd.Index(list(labels), name="columns")
    pd.DataFrame(data=matrix, index=I, columns=C).to_csv(path)         # same call as S/utils.py:574-575


def write_tree_levelorder_info(path: str, tree: LevelOrderTree):
    """`<name>_tree_levelorder_info.csv`: one row per node in level order — name, branch length, names of its ancestors
    (parent first), padded with empty cells; tab separated with pandas' default index/header (S/utils.py:857-863, 893-895)."""
    import pandas as pd
    rows = []
    for i in range(tree.n_nodes):
        rows.append([tree.names[i].replace("'", ""), float(tree.branch[i])] + [tree.names[a].replace("'", "") for a in tree.ancestors(i)])
    width = max(map(len, rows))
    info = np.array([r + [None] * (width - len(r)) for r in rows], dtype=object)
    pd.DataFrame(info).to_csv(path, sep="\t")


def write_dataset_tree_files(newick: str, storage_folder: str, name_file: str, device="cuda", rename_internal_nodes: bool = True,
                             cladistic: Optional[bool] = None) -> LevelOrderTree:
    """What `create_dataset` leaves on disk for the tree (S/utils.py:849-897): `<name>.newick`, `<name>.format8newick`,
    `<name>_tree_levelorder_info.csv`, `<name>_patristic_distance_matrix.csv` and — up to 200 leaves, like the
    reference — `<name>_cladistic_distance_matrix.csv`.  The matrices come from the GPU builder."""
    tree = read_newick(newick)
    if rename_internal_nodes:
        tree = rename_tree_internal_nodes(tree)
    os.makedirs(storage_folder, exist_ok=True)
    base = os.path.join(storage_folder, name_file)
    n_leaves = int(tree.is_leaf.sum())
    if cladistic is None:
        cladistic = n_leaves <= 200
    order, pat, cla = distance_matrices(tree, device, want_cladistic=cladistic)
    labels = [tree.names[i] for i in order]
    _write_matrix_csv(base + "_patristic_distance_matrix.csv", labels, pat.cpu().numpy())
    if cla is not None:
        _write_matrix_csv(base + "_cladistic_distance_matrix.csv", labels, cla.cpu().numpy())
    write_tree_levelorder_info(base + "_tree_levelorder_info.csv", tree)
    with open(base + ".newick", "w") as f:
        f.write(to_newick(tree, internal_names=True, lengths=True) + "\n")
    with open(base + ".format8newick", "w") as f:
        f.write(to_newick(tree, internal_names=True, lengths=False, root_name=True) + "\n")
    return tree


def read_tree_levelorder_names(path: str) -> np.ndarray:
    """Node names in level order from `<name>_tree_levelorder_info.csv` (S/main.py:203-207)."""
    import pandas as pd
    info = pd.read_csv(path, sep="\t", index_col=False, low_memory=False)
    info["0"] = info["0"].astype(str)
    return np.asarray(info["0"].tolist())


def load_distance_matrix(csv_path: str, tree_levelorder_names: np.ndarray) -> np.ndarray:
    """The `(N+1) x (N+1)` float64 matrix the model consumes: read the CSV (S/main.py:196-198), symmetrise
    (S/utils.py:939), rename rows/columns to level-order ids (S/utils.py:945-953) and prepend the id header with -inf in
    the corner (S/utils.py:971-979).  Row order is the file's; `sort_by_node_id` gives S/load_utils.py:452-458."""
    import pandas as pd
    m = pd.read_csv(csv_path, low_memory=False, float_precision="round_trip")   # pandas' default parser may be 1 ulp off
    m = m.rename(columns={"Unnamed: 0": "rows"})
    m.set_index("rows", inplace=True)
    m.index = m.index.astype(str)
    vals = m.to_numpy(dtype=np.float64)
    vals = np.maximum(vals, vals.T)
    pos = {str(nm): k for k, nm in enumerate(tree_levelorder_names)}
    try:
        rows = np.array([pos[str(r)] for r in m.index], dtype=np.float64)
        cols = np.array([pos[str(c)] for c in m.columns], dtype=np.float64)
    except KeyError as e:
        raise KeyError(f"node {e.args[0]!r} of {csv_path} is not in the tree level-order file") from None
    out = np.empty((len(rows) + 1, len(cols) + 1), dtype=np.float64)
    out[0, 0] = -np.inf
    out[0, 1:] = cols
    out[1:, 0] = rows
    out[1:, 1:] = vals
    return out


def sort_by_node_id(matrix: np.ndarray) -> np.ndarray:
    """`matrix_sort` of S/load_utils.py:452-458: rows and columns in ascending node id (the -inf corner stays first)."""
    idx = np.argsort(matrix[:, 0], kind="stable")
    return matrix[idx][:, idx]


def header_matrix_on_device(tree: LevelOrderTree, device="cuda"):
    """The sorted `(N+1) x (N+1)` patristic tensor straight from the tree, no files in between: what
    `load_distance_matrix` + `sort_by_node_id` return for the files `write_dataset_tree_files` writes."""
    import torch
    from . import ops
    from .utils import require_cuda
    dev = require_cuda(device)
    N = tree.n_nodes
    pat, _, _ = ops.tree_distances(torch.from_numpy(tree.parent).to(dev), torch.from_numpy(tree.branch).to(dev))
    out = torch.empty((N + 1, N + 1), dtype=torch.float64, device=dev)
    ids = torch.arange(N, dtype=torch.float64, device=dev)
    out[0, 0] = float("-inf")
    out[0, 1:] = ids
    out[1:, 0] = ids
    out[1:, 1:] = pat
    return out
