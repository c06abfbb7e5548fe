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


# ---- against the reference's own generator code (synthetic_structsim.py needs only networkx + numpy) ----
REF_SS = "/root/reference/synthetic_structsim.py"


def _ref_structsim():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_synthetic_structsim", REF_SS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _nx_graph(n, ei, lo=0, hi=None):
    import networkx as nx
    hi = n if hi is None else hi
    G = nx.Graph()
    G.add_nodes_from(range(lo, hi))
    G.add_edges_from((a, b) for a, b in ei.t().tolist() if lo <= a < hi and lo <= b < hi)
    return G


@pytest.mark.skipif(not os.path.isfile(REF_SS), reason="needs /root/reference (build container only)")
def test_generated_graphs_match_the_reference_build_graph():
    """Every graph against ``synthetic_structsim.build_graph`` run with the arguments of gengraph.py:60-75 /
    utils.py:62-63: same node and edge counts; the motif (its first node marked: that is where the attach edge
    ends) is isomorphic to the reference's shape as a ROOTED graph; the tree base is the reference's balanced tree;
    exactly one edge joins base and motif; the plugin node is spread over the base like np.random.choice."""
    import networkx as nx
    ss = _ref_structsim()
    np.random.seed(0)
    node_num = 4
    fg = generate_spmotif(30, node_num=node_num, noise=0.0, seed=5)
    shapes = {"house": [["house"]], "cycle": [["cycle", 6]], "grid": [["grid"]], "diamond": [["diamond"]]}
    plug = {0: [], 1: []}
    for g in range(len(fg)):
        n, ei, _ = _graph(fg, g)
        ctx, motif = int(fg.context[g]), MOTIFS[int(fg.y[g])]
        width, m = (2, node_num) if ctx == 0 else (node_num ** 2, 2)               # settings_dict, utils.py:62-63
        R, _, ref_plugins = ss.build_graph(width, "tree" if ctx == 0 else "ba", shapes[motif], rdm_basis_plugins=True,
                                           start=0, m=m)
        assert n == R.number_of_nodes() and ei.size(1) == 2 * R.number_of_edges(), (g, motif, ctx)
        nb = n - {"house": 5, "cycle": 6, "grid": 6, "diamond": 6}[motif]
        ours_m, ref_m = _nx_graph(n, ei, nb, n), R.subgraph(range(nb, n)).copy()
        for G_, in ((ours_m,), (ref_m,)):
            nx.set_node_attributes(G_, {v: int(v == nb) for v in G_.nodes}, "root")
        assert nx.is_isomorphic(ours_m, ref_m, node_match=lambda a, b: a["root"] == b["root"]), (g, motif)
        ours_b, ref_b = _nx_graph(n, ei, 0, nb), R.subgraph(range(nb)).copy()
        if ctx == 0:
            assert sorted(ours_b.edges) == sorted(tuple(sorted(e)) for e in ref_b.edges)   # the same labelled tree
        else:
            assert ours_b.number_of_edges() == ref_b.number_of_edges() and nx.is_connected(ours_b)
        cross = [(a, b) for a, b in ei.t().tolist() if a < nb <= b]
        assert len(cross) == 1 and cross[0][1] == nb                                       # base node -- motif node 0
        plug[ctx].append(cross[0][0] / (nb - 1))
        assert 0 <= int(ref_plugins[0]) < nb
    for ctx in (0, 1):                                                                     # uniform over the base nodes
        p = np.array(plug[ctx])
        assert abs(p.mean() - 0.5) < 0.1 and p.min() < 0.15 and p.max() > 0.85 and len(set(p.tolist())) > 10


@pytest.mark.skipif(not os.path.isfile(REF_SS), reason="needs /root/reference (build container only)")
def test_batched_barabasi_albert_has_the_degree_statistics_of_the_reference_base():
    """The BA base grown for all graphs at once (one multinomial draw without replacement per new node) against
    ``synthetic_structsim.ba`` = ``nx.barabasi_albert_graph``: same edge count, and the same degree statistics
    (second moment, hub degree, leaf fraction) within sampling error over 200 graphs of 100 nodes."""
    from cal_b200.datasets import _ba_edges_batched
    ss = _ref_structsim()
    n, m, G = 100, 2, 200
    s, d, gid = _ba_edges_batched(torch.full((G,), n), m, torch.Generator().manual_seed(3), torch.device("cpu"))
    deg = torch.zeros(G, n)
    deg.index_put_((gid, s), torch.ones(s.numel()), accumulate=True)
    deg.index_put_((gid, d), torch.ones(s.numel()), accumulate=True)
    assert s.numel() == G * m * (n - m) and bool((s != d).all())
    key = (gid * n + torch.minimum(s, d)) * n + torch.maximum(s, d)
    assert torch.unique(key).numel() == key.numel()                                        # simple graphs
    import random
    random.seed(1)
    np.random.seed(1)
    ref = np.array([[dg for _, dg in ss.ba(0, n, m=m)[0].degree()] for _ in range(G)], dtype=np.float64)
    assert ref.sum() == float(deg.sum())
    ours = deg.numpy().astype(np.float64)
    for stat in (lambda a: (a ** 2).mean(1), lambda a: a.max(1), lambda a: (a == m).mean(1)):
        a, b = stat(ours), stat(ref)
        se = np.sqrt(a.var() / G + b.var() / G)
        assert abs(a.mean() - b.mean()) < 4 * se + 1e-9, (a.mean(), b.mean(), se)


def test_constant_and_gaussian_feature_modes():
    """utils.py:46-47: with feature_dim > 0 every node of a graph carries the same U(0, 1) vector."""
    fg = generate_spmotif(3, base_nodes=(6, 9), ba_m=1, noise=0.0, feature_dim=7, seed=2)
    for g in range(len(fg)):
        _, _, feat = _graph(fg, g)
        assert feat.shape[1] == 7 and bool((feat == feat[0]).all()) and 0.0 <= float(feat.min()) and float(feat.max()) < 1.0
    assert not torch.equal(_graph(fg, 0)[2][0], _graph(fg, 1)[2][0])
    fgn = generate_spmotif(3, base_nodes=(6, 9), ba_m=1, noise=0.0, feature_dim=7, seed=2, gaussian_features=True)
    assert float(fgn.feat.std()) > 0.8 and not bool((_graph(fgn, 0)[2] == _graph(fgn, 0)[2][0]).all())


@pytest.mark.skipif(not os.path.isfile("/root/reference/utils.py"), reason="needs /root/reference (build container only)")
def test_bias_split_selects_the_graphs_the_reference_function_selects(capsys):
    """The UNMODIFIED body of ``dataset_bias_split`` (utils.py:121-159; extracted with ast -- the module itself
    imports matplotlib / PyG) on the same graphs in the reference's dict-of-lists layout: identical train / val /
    test membership and identical ``the``."""
    import argparse
    import ast
    import random
    src = open("/root/reference/utils.py").read()
    lines = src.splitlines()
    code = "\n\n".join("\n".join(lines[n.lineno - 1:n.end_lineno]) for n in ast.parse(src).body
                       if isinstance(n, ast.FunctionDef) and n.name in ("dataset_bias_split", "print_graph_info"))
    ns = {"random": random}
    exec(compile(code, "reference_utils_excerpt", "exec"), ns)

    class G:
        def __init__(self, idx, nodes, edges):
            self.idx, self.num_nodes, self.num_edges = idx, nodes, edges

    fg = generate_spmotif(150, base_nodes=(8, 12), ba_m=1, noise=0.1, seed=4)
    nodes, edges = fg.node_ptr[1:] - fg.node_ptr[:-1], fg.edge_ptr[1:] - fg.edge_ptr[:-1]
    ds = {"tree": {}, "ba": {}}
    for k, shape in enumerate(MOTIFS):
        for c, name in enumerate(("tree", "ba")):
            idx = torch.nonzero((fg.y == k) & (fg.context == c)).view(-1).tolist()
            ds[name][shape] = [G(i, int(nodes[i]), int(edges[i])) for i in idx]
    for bias, total in ((0.9, 400), (0.5, 240), (0.7, 520)):
        random.seed(0)
        r_tr, r_va, r_te, r_the = ns["dataset_bias_split"](ds, argparse.Namespace(num_classes=4), bias=bias,
                                                           split=[7, 1, 2], total=total)
        tr, va, te, the = dataset_bias_split(fg, bias=bias, split=(7, 1, 2), total=total)
        assert sorted(g.idx for g in r_tr) == sorted(tr.tolist())
        assert sorted(g.idx for g in r_va) == sorted(va.tolist())
        assert sorted(g.idx for g in r_te) == sorted(te.tolist())
        assert the == pytest.approx(r_the)
    capsys.readouterr()                                          # (the reference prints a table per class)


_FE_CHECK = r"""
import os, sys
root, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [os.path.join(root, "oracle", "pyg_shim"), ref, root]
import torch
import feature_expansion                                          # the reference's file, unmodified
from cal_b200.datasets import FlatGraphs, expand_features

class D:                                                          # what FeatureExpander.transform touches of a PyG Data
    def __init__(self, x, edge_index):
        self.x, self.edge_index = x, edge_index
    num_nodes = property(lambda s: s.x.size(0))
    num_edges = property(lambda s: s.edge_index.size(1))

gen = torch.Generator().manual_seed(7)
G, L = 12, 7
n = torch.randint(2, 30, (G,), generator=gen)
node_ptr = torch.zeros(G + 1, dtype=torch.long); node_ptr[1:] = torch.cumsum(n, 0)
eis, labels = [], torch.randint(0, L, (int(n.sum()),), generator=gen)
for g in range(G):
    e = int(torch.randint(1, 120, (1,), generator=gen)) if g != 3 else 300      # graph 3: degrees beyond the cap
    a = torch.randint(0, int(n[g]), (2, e), generator=gen)
    a = a[:, a[0] != a[1]]
    a = torch.unique(torch.cat([a, a.flip(0)], 1), dim=1)                       # symmetric, coalesced (TU / PyG)
    eis.append(a)
edge_ptr = torch.zeros(G + 1, dtype=torch.long); edge_ptr[1:] = torch.cumsum(torch.tensor([a.size(1) for a in eis]), 0)
for maxdeg in (100, 10, 3):                                                      # opts.py:122,133: odeg100 / odeg10
    fg = FlatGraphs(node_ptr, edge_ptr, torch.cat(eis, 1), torch.zeros(int(n.sum()), 0), torch.zeros(G, dtype=torch.long))
    fg = expand_features(fg, num_node_labels=L, node_labels=labels, degree=True, onehot_maxdeg=maxdeg)
    fe = feature_expansion.FeatureExpander(degree=True, onehot_maxdeg=maxdeg, AK=0)
    for g in range(G):
        x0 = torch.nn.functional.one_hot(labels[node_ptr[g]:node_ptr[g + 1]], L).float()
        want = fe.transform(D(x0, eis[g])).x
        got = fg.feat[node_ptr[g]:node_ptr[g + 1]]
        assert got.shape == want.shape == (int(n[g]), L + 1 + maxdeg + 1), (got.shape, want.shape)
        assert torch.equal(got, want), (g, maxdeg)
print("feature expansion identical")
"""


@pytest.mark.skipif(not os.path.isfile("/root/reference/feature_expansion.py"), reason="needs /root/reference (build container only)")
def test_feature_expansion_equals_the_reference_feature_expander():
    """The reference's unmodified ``FeatureExpander(degree=True, onehot_maxdeg=100 | 10, AK=0).transform``
    (feature_expansion.py:42-59,100-113; the 'deg+odeg100' recipe of opts.py:122 that gives MUTAG its 109 columns)
    over the stand-in PyG, graph by graph, against ``expand_features`` on the flat arrays: bit-identical."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _FE_CHECK, root, "/root/reference"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "feature expansion identical" in r.stdout
