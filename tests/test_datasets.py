"""cal_b200.datasets: the vectorised SPMotif generator, the bias split (utils.py:121-159) and the TU text reader
(tu_dataset.py:73-88 + feature_expansion.py) -- CPU tests."""
import os

import numpy as np
import pytest
import torch

from cal_b200.data import MOTIFS, Batch
from cal_b200.datasets import FlatGraphs, dataset_bias_split, expand_features, generate_spmotif, read_tu_dataset


def _graph(fg, g):
    n0, n1, e0, e1 = int(fg.node_ptr[g]), int(fg.node_ptr[g + 1]), int(fg.edge_ptr[g]), int(fg.edge_ptr[g + 1])
    return n1 - n0, fg.edge_index[:, e0:e1], fg.feat[n0:n1]


def _components(n, ei):
    parent = list(range(n))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for a, b in ei.t().tolist():
        parent[find(a)] = find(b)
    return len({find(a) for a in range(n)})


def test_generator_structure_small_graphs():
    fg = generate_spmotif(40, base_nodes=(12, 28), ba_m=1, noise=0.1, seed=3)
    assert len(fg) == 4 * 2 * 40 and fg.num_features == 10
    nodes = (fg.node_ptr[1:] - fg.node_ptr[:-1])
    motif_nodes = torch.tensor([5, 6, 6, 6])[fg.y]
    assert int((nodes - motif_nodes).min()) >= 12 and int((nodes - motif_nodes).max()) < 28
    for g in range(0, len(fg), 7):
        n, ei, feat = _graph(fg, g)
        assert int(ei.min()) >= 0 and int(ei.max()) < n and bool((ei[0] != ei[1]).all())
        pairs = set(map(tuple, ei.t().tolist()))
        assert len(pairs) == ei.size(1) and all((b, a) in pairs for a, b in pairs)      # simple, symmetric
        key = ei[0] * 4096 + ei[1]
        assert bool((key[1:] > key[:-1]).all())                                          # from_networkx order
        assert _components(n, ei) == 1
        deg = torch.bincount(ei[0], minlength=n)
        assert torch.equal(feat.argmax(1), deg.clamp(max=9)) and bool((feat.sum(1) == 1).all())
        und = ei.size(1) // 2
        base = n - int(motif_nodes[g])
        motif_e = [6, 6, 7, 8][int(fg.y[g])]
        base_e = base - 1                                                               # tree, and BA with m = 1
        assert und == int((base_e + motif_e + 1) * 1.1)                                  # + attach, + 10 % noise


def test_generator_reference_sizes_and_determinism():
    a = generate_spmotif(6, node_num=15, seed=9)                 # utils.py:62-63: tree 15-ary height 2, BA 225 nodes m = 2
    b = generate_spmotif(6, node_num=15, seed=9)
    assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.feat, b.feat)
    nodes = a.node_ptr[1:] - a.node_ptr[:-1]
    tree = a.context == 0
    assert set((nodes[tree] - torch.tensor([5, 6, 6, 6])[a.y[tree]]).tolist()) == {1 + 15 + 225}
    assert set((nodes[~tree] - torch.tensor([5, 6, 6, 6])[a.y[~tree]]).tolist()) == {225}
    n, ei, _ = _graph(a, int(torch.nonzero(~tree)[0]))
    assert _components(n, ei) == 1
    # BA with m = 2: 2 * (n - 3) + 2 base edges (star on 3 nodes, then 2 per node)
    g = int(torch.nonzero((~tree) & (a.y == 0))[0])
    n, ei, _ = _graph(a, g)
    assert ei.size(1) // 2 == int((2 * (225 - 3) + 2 + 6 + 1) * 1.1)


def test_bias_split_matches_reference_counts():
    fg = generate_spmotif(120, base_nodes=(8, 12), ba_m=1, noise=0.0, seed=1)
    tr, va, te, the = dataset_bias_split(fg, bias=0.9, split=(7, 1, 2), total=400)
    assert (len(tr), len(va), len(te)) == (4 * (63 + 7), 4 * (9 + 1), 4 * 20) or len(tr) + len(va) + len(te) <= 400
    assert len(set(tr.tolist()) & set(te.tolist())) == 0 and len(set(tr.tolist()) & set(va.tolist())) == 0
    y, ctx = fg.y[tr], fg.context[tr]
    for k in range(4):
        frac_tree = float(((y == k) & (ctx == 0)).sum()) / float((y == k).sum())
        assert abs(frac_tree - (0.9 if k == 0 else 0.1)) < 0.02                          # utils.py:126
    yt, ct = fg.y[te], fg.context[te]
    for k in range(4):
        assert int(((yt == k) & (ct == 0)).sum()) == int(((yt == k) & (ct == 1)).sum())  # test split is unbiased
    assert the > 0
    sub = fg.select(tr[:9])
    ref = Batch.from_data_list([fg.to_data_list()[i] for i in tr[:9].tolist()])
    off = torch.repeat_interleave(sub.node_ptr[:-1], sub.edge_ptr[1:] - sub.edge_ptr[:-1])
    assert torch.equal(sub.edge_index + off, ref.edge_index) and torch.equal(sub.feat, ref.feat)


def test_tu_reader_and_feature_expansion(tmp_path):
    # two graphs: a triangle (3 nodes) and a path of 4 nodes; node labels in {0, 1, 2}; duplicate + self-loop lines
    name = "TOY"
    edges = [(1, 2), (2, 1), (2, 3), (3, 2), (1, 3), (3, 1), (1, 1), (1, 2),
             (4, 5), (5, 4), (5, 6), (6, 5), (6, 7), (7, 6)]
    (tmp_path / (name + "_A.txt")).write_text("\n".join("%d, %d" % e for e in edges) + "\n")
    (tmp_path / (name + "_graph_indicator.txt")).write_text("\n".join(map(str, [1, 1, 1, 2, 2, 2, 2])) + "\n")
    (tmp_path / (name + "_graph_labels.txt")).write_text("1\n-1\n")
    (tmp_path / (name + "_node_labels.txt")).write_text("\n".join(map(str, [0, 1, 2, 0, 0, 1, 2])) + "\n")
    fg = read_tu_dataset(str(tmp_path), name, degree=True, onehot_maxdeg=100)
    assert len(fg) == 2 and fg.node_ptr.tolist() == [0, 3, 7] and fg.edge_ptr.tolist() == [0, 6, 12]
    assert fg.y.tolist() == [1, 0]                                                       # labels -> 0 .. C-1
    assert fg.edge_index[:, :6].t().tolist() == [[0, 1], [0, 2], [1, 0], [1, 2], [2, 0], [2, 1]]
    assert fg.edge_index[:, 6:].t().tolist() == [[0, 1], [1, 0], [1, 2], [2, 1], [2, 3], [3, 2]]
    assert fg.num_features == 3 + 1 + 101                                                # MUTAG recipe: 7 + 1 + 101 = 109
    deg = torch.tensor([2, 2, 2, 1, 2, 2, 1])
    assert torch.equal(fg.feat[:, 3], deg.float())
    assert torch.equal(fg.feat[:, 4:].argmax(1), deg) and torch.equal(fg.feat[:, :3].argmax(1), torch.tensor([0, 1, 2, 0, 0, 1, 2]))
    dl = fg.to_data_list()
    assert dl[1].feat.shape == (4, 105) and dl[1].edge_index.shape == (2, 6)
