"""Tree / matrix file formats either side of the patristic builder (`draupnir_asr_b200/tree_io.py`).

CPU part: Newick parsing, ete3's orders (level order ids, pre-order leaves, `A<k>` internal names), the level-order info
file and the loader (S/main.py:196-207, S/utils.py:939-979) with matrices from the oracle.  GPU part: the files written by
the CUDA builder load back to the oracle's matrix."""
import os

import numpy as np
import pytest

from draupnir_asr_b200 import synthetic as S
from draupnir_asr_b200 import tree_io as T
from oracle import patristic_oracle as P

NEWICK = "((a:0.1,'b c':0.2)X:0.3,(d:0.4,(e:0.5,f:0.6):0.7):0.8,g:0.9);"


def _depth(tree):
    d = np.zeros(tree.n_nodes, dtype=np.int32)
    for i in range(1, tree.n_nodes):
        d[i] = d[tree.parent[i]] + 1
    return d


def test_newick_level_order_and_names():
    t = T.read_newick(NEWICK)
    # breadth first, children in file order: root | X, (d,(e,f)), g | a, 'b c', d, (e,f) | e, f
    assert t.names == ["", "X", "", "g", "a", "b c", "d", "", "e", "f"]
    assert t.parent.tolist() == [-1, 0, 0, 0, 1, 1, 2, 2, 7, 7]
    assert np.allclose(t.branch, [0.0, 0.3, 0.8, 0.9, 0.1, 0.2, 0.4, 0.7, 0.5, 0.6])
    assert t.leaf_ids_preorder.tolist() == [4, 5, 6, 8, 9, 3]                 # a, b c, d, e, f, g
    assert t.internal_ids.tolist() == [0, 1, 2, 7]
    assert t.ancestors(8) == [7, 2, 0] and t.ancestors(0) == []
    r = T.rename_tree_internal_nodes(t)
    assert [r.names[i] for i in r.internal_ids] == ["A6", "A7", "A8", "A9"]   # 6 leaves -> A6.. in level order
    assert [r.names[i] for i in r.file_order()] == ["A6", "A7", "A8", "A9", "a", "b c", "d", "e", "f", "g"]
    # round trip through text (quoted label, lengths as shortest round-trip repr)
    again = T.read_newick(T.to_newick(r, root_name=True))
    assert again.names == r.names and again.parent.tolist() == r.parent.tolist() and np.array_equal(again.branch, r.branch)
    assert T.to_newick(t, internal_names=False, lengths=False) == "((a,'b c'),(d,(e,f)),g);"


def test_newick_defaults_comments_and_errors():
    t = T.read_newick("(A,(B,C)[&&NHX:S=x]:2)root;")
    assert t.names == ["root", "A", "", "B", "C"] and t.branch.tolist() == [0.0, 1.0, 2.0, 1.0, 1.0]
    assert T.read_newick("A;").names == ["A"]
    for bad in ("", "(A,B)", "(A,B));", "((A,B);", "(A:x,B);", "(A,B);(C,D);", "(A,B)'unterminated;"):
        with pytest.raises(T.NewickError):
            T.read_newick(bad)


def _random_newick(n, seed):
    """A random binary tree as Newick text plus the SyntheticTree it came from (ids are level order in both)."""
    st = S.random_tree(n, seed=seed)
    kids = [[] for _ in range(st.n_nodes)]
    for i in range(1, st.n_nodes):
        kids[st.parent[i]].append(i)

    def rec(v):
        own = "" if kids[v] else f"L{v}"
        body = "(" + ",".join(rec(c) for c in kids[v]) + ")" if kids[v] else ""
        return body + own + (f":{float(st.branch[v])!r}" if v else "")

    import sys
    sys.setrecursionlimit(max(10000, 4 * n))
    return rec(0) + ";", st


@pytest.mark.parametrize("n", [2, 7, 60])
def test_files_round_trip_with_oracle_matrices(tmp_path, n):
    """Writer formats + loader: the header matrix loaded from the files equals the oracle's all-pairs matrix in id order."""
    text, st = _random_newick(n, seed=n)
    t = T.rename_tree_internal_nodes(T.read_newick(text))
    assert t.parent.tolist() == st.parent.tolist() and np.array_equal(t.branch, st.branch)     # ids ARE level order
    order = t.file_order()
    _, pat, cla = P.patristic_c(t.parent, t.branch, _depth(t), order, order, want_cladistic=True)
    base = os.path.join(tmp_path, "fam")
    labels = [t.names[i] for i in order]
    upper = np.triu(pat)                                       # the reference's small-tree branch writes i < j only
    T._write_matrix_csv(base + "_patristic_distance_matrix.csv", labels, upper)
    T._write_matrix_csv(base + "_cladistic_distance_matrix.csv", labels, cla)
    T.write_tree_levelorder_info(base + "_tree_levelorder_info.csv", t)
    head = open(base + "_patristic_distance_matrix.csv").readline().rstrip("\n").split(",")
    assert head == ["rows"] + labels
    names = T.read_tree_levelorder_names(base + "_tree_levelorder_info.csv")
    assert names.tolist() == t.names
    line1 = open(base + "_tree_levelorder_info.csv").read().splitlines()[1].split("\t")
    assert line1[:3] == ["0", t.names[0], "0.0"]
    m = T.sort_by_node_id(T.load_distance_matrix(base + "_patristic_distance_matrix.csv", names))
    N = t.n_nodes
    assert m.shape == (N + 1, N + 1) and np.isneginf(m[0, 0])
    assert m[0, 1:].tolist() == list(range(N)) and m[1:, 0].tolist() == list(range(N))
    _, full, _ = P.patristic_c(t.parent, t.branch, _depth(t))
    assert np.array_equal(m[1:, 1:], full)                     # CSV text round-trips float64 exactly
    c = T.sort_by_node_id(T.load_distance_matrix(base + "_cladistic_distance_matrix.csv", names))
    d = _depth(t)
    assert c[1 + 0, 1:].tolist() == d.tolist()                 # edge count to the root = depth


@pytest.mark.gpu
@pytest.mark.parametrize("n", [5, 150, 260])
def test_gpu_builder_writes_the_reference_files(cuda, tmp_path, n):
    text, st = _random_newick(n, seed=100 + n)
    t = T.write_dataset_tree_files(text, str(tmp_path), "fam", device=cuda)
    base = os.path.join(tmp_path, "fam")
    assert os.path.exists(base + "_cladistic_distance_matrix.csv") == (n <= 200)      # S/utils.py:509-511
    names = T.read_tree_levelorder_names(base + "_tree_levelorder_info.csv")
    m = T.sort_by_node_id(T.load_distance_matrix(base + "_patristic_distance_matrix.csv", names))
    _, full, _ = P.patristic_c(t.parent, t.branch, _depth(t))
    assert np.allclose(m[1:, 1:], full, rtol=1e-12, atol=0.0)
    direct = T.header_matrix_on_device(t, cuda).cpu().numpy()
    assert np.array_equal(direct, m)
    again = T.read_newick(base + ".newick")
    assert again.names[1:] == t.names[1:] and np.array_equal(again.branch, t.branch)
    assert T.read_newick(base + ".format8newick").names == t.names
