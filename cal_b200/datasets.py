"""Dataset construction for the CAL hot path, vectorised.

The reference builds its synthetic SPMotif data one networkx graph at a time
(utils.py:38-89 ``creat_one_pyg_graph`` / ``graph_dataset_generate``,
gengraph.py:13-33 ``perturb``, synthetic_structsim.py:207-288 ``build_graph``) --
minutes of Python per dataset -- and reads TU datasets through PyG
(tu_dataset.py:73-88 ``read_tu_data``, feature_expansion.py:56-106).  Here

* :func:`generate_spmotif` builds ALL graphs of a dataset at once with torch tensor
  ops (on the GPU when ``device`` says so): the base graphs (balanced ``r``-ary tree of
  height ``h`` / Barabasi-Albert with ``m`` attachments, grown one node per step for
  every graph simultaneously), the motif (house / cycle / grid / diamond) attached by
  one edge, the fraction ``noise`` of random extra edges (rejection-sampled for the
  whole batch of graphs), degree one-hot features -- straight into the flat arrays of
  :class:`FlatGraphs`, the layout ``GraphStore`` uploads (no per-graph Python objects);
* :func:`dataset_bias_split` is utils.py:121-159 on those arrays (train / val biased
  towards tree+house and BA+others, test unbiased, class balanced);
* :func:`read_tu_dataset` parses the TU text format (``<name>_A.txt``,
  ``_graph_indicator.txt``, ``_graph_labels.txt``, ``_node_labels.txt``,
  ``_node_attributes.txt``) with numpy and applies the reference's feature expansion
  (node-label one-hot | degree | one-hot degree, feature_expansion.py:56-59,101-106).

Everything here is offline data preparation: no kernel of the hot path depends on it.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .data import Data, MOTIFS

__all__ = ["FlatGraphs", "generate_spmotif", "dataset_bias_split", "read_tu_dataset", "expand_features"]


class FlatGraphs:
    """A set of graphs as flat arrays (CPU or GPU tensors):
    ``node_ptr`` i64[G+1], ``edge_ptr`` i64[G+1], ``edge_index`` i64[2, E] (graph-local node ids, both
    directions, sorted by (source, target) inside a graph; ``from_networkx`` emits the same columns grouped by
    source in node order, the targets of a source in edge-insertion order -- only the summation order differs),
    ``feat`` f32[N, F], ``y`` i64[G], ``context`` i64[G] (0 = tree base, 1 = BA base; -1 unknown)."""

    def __init__(self, node_ptr, edge_ptr, edge_index, feat, y, context=None):
        self.node_ptr, self.edge_ptr, self.edge_index, self.feat, self.y = node_ptr, edge_ptr, edge_index, feat, y
        self.context = context if context is not None else torch.full_like(y, -1)

    def __len__(self):
        return int(self.y.numel())

    @property
    def num_features(self):
        return int(self.feat.size(1))

    def to(self, device):
        return FlatGraphs(*[t.to(device) for t in (self.node_ptr, self.edge_ptr, self.edge_index, self.feat, self.y, self.context)])

    def select(self, idx):
        """The graphs ``idx`` (1-d index tensor / array), re-packed."""
        idx = torch.as_tensor(idx, dtype=torch.long, device=self.y.device)
        n0, n1 = self.node_ptr[idx], self.node_ptr[idx + 1]
        e0, e1 = self.edge_ptr[idx], self.edge_ptr[idx + 1]
        nn, ne = n1 - n0, e1 - e0
        node_ptr = torch.zeros(idx.numel() + 1, dtype=torch.long, device=idx.device)
        edge_ptr = torch.zeros_like(node_ptr)
        node_ptr[1:] = torch.cumsum(nn, 0)
        edge_ptr[1:] = torch.cumsum(ne, 0)
        nsel = torch.repeat_interleave(n0 - node_ptr[:-1], nn) + torch.arange(int(node_ptr[-1]), device=idx.device)
        esel = torch.repeat_interleave(e0 - edge_ptr[:-1], ne) + torch.arange(int(edge_ptr[-1]), device=idx.device)
        return FlatGraphs(node_ptr, edge_ptr, self.edge_index[:, esel], self.feat[nsel], self.y[idx], self.context[idx])

    def to_data_list(self):
        """``list[Data]`` (CPU) for loaders that want per-graph objects (``cal_b200.data.DataLoader`` / ``GraphStore``)."""
        f = self.to("cpu")
        out = []
        for g in range(len(f)):
            n0, n1, e0, e1 = int(f.node_ptr[g]), int(f.node_ptr[g + 1]), int(f.edge_ptr[g]), int(f.edge_ptr[g + 1])
            out.append(Data(feat=f.feat[n0:n1].clone(), edge_index=f.edge_index[:, e0:e1].clone(), y=f.y[g:g + 1].clone()))
        return out


# ------------------------------------------------------------------------------------------------
# vectorised SPMotif generator
# ------------------------------------------------------------------------------------------------

def _motif_table(device):
    """Padded motif edge table: nodes[k], edges[k, 8, 2] (-1 padded), nedges[k] (synthetic_structsim.py:49-69,114-125,169-204)."""
    spec = {
        "house": (5, [(0, 1), (1, 2), (2, 3), (3, 0), (4, 0), (4, 1)]),
        "cycle": (6, [(i, (i + 1) % 6) for i in range(6)]),
        "grid": (6, [(0, 1), (1, 2), (3, 4), (4, 5), (0, 3), (1, 4), (2, 5)]),
        "diamond": (6, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0), (5, 1), (4, 2)]),
    }
    nodes = torch.tensor([spec[m][0] for m in MOTIFS], device=device)
    ned = torch.tensor([len(spec[m][1]) for m in MOTIFS], device=device)
    edges = torch.full((len(MOTIFS), 8, 2), -1, dtype=torch.long, device=device)
    for k, m in enumerate(MOTIFS):
        edges[k, :len(spec[m][1])] = torch.tensor(spec[m][1], device=device)
    return nodes, edges, ned


def _ba_edges_batched(n_nodes, m, gen, device):
    """Barabasi-Albert graphs for a whole batch at once (networkx convention: star on m + 1 nodes, then every new
    node attaches to m distinct targets drawn proportionally to degree).  ``n_nodes`` i64[G].  Returns
    (src, dst, graph id) of the undirected edges.  One step per node index, every graph in parallel: the
    degree-proportional draw is a multinomial without replacement over the current degree table."""
    G = int(n_nodes.numel())
    nmax = int(n_nodes.max())
    m = max(1, min(m, nmax - 1))
    deg = torch.zeros(G, nmax, device=device)
    deg[:, 0] = m
    deg[:, 1:m + 1] = 1
    gid0 = torch.arange(G, device=device)
    src = [torch.zeros(G * m, dtype=torch.long, device=device)]
    dst = [torch.arange(1, m + 1, device=device).repeat(G)]
    gid = [gid0.repeat_interleave(m)]
    for v in range(m + 1, nmax):
        live = n_nodes > v
        if not bool(live.any()):
            break
        w = deg[live, :v]
        tg = torch.multinomial(w, m, replacement=False, generator=gen)          # [live, m] distinct targets
        g_live = gid0[live]
        src.append(tg.reshape(-1))
        dst.append(torch.full((tg.numel(),), v, dtype=torch.long, device=device))
        gid.append(g_live.repeat_interleave(m))
        deg[g_live.repeat_interleave(m), tg.reshape(-1)] += 1
        deg[g_live, v] = m
    return torch.cat(src), torch.cat(dst), torch.cat(gid)


def generate_spmotif(num_per_class, node_num=15, tree_height=2, ba_m=2, noise=0.1, max_degree=10, feature_dim=-1,
                     contexts=("tree", "ba"), base_nodes=None, seed=666, device="cpu", random_plugin=True,
                     gaussian_features=False):
    """All graphs of ``graph_dataset_generate`` (utils.py:59-89) at once: for every motif class and every context
    ``num_per_class`` graphs.  Reference sizes: tree = balanced ``node_num``-ary tree of height ``tree_height``
    (settings_dict, utils.py:62-63), BA = ``node_num ** 2`` nodes with ``m = 2``.  ``base_nodes=(lo, hi)`` instead
    draws every base size uniformly from [lo, hi) (a tree then is the first n nodes of the balanced binary tree):
    the small-graph workloads of BASELINE.json.  The motif's first node is joined to ONE base node drawn uniformly
    (``generate_graph`` passes ``rdm_basis_plugins=True``, gengraph.py:70-75, synthetic_structsim.py:247-249;
    ``random_plugin=False``: base node 0).  ``feature_dim == -1``: one-hot(min(degree, max_degree - 1))
    (featgen.py:19-28), else one U(0, 1) vector per graph repeated on every node (utils.py:46-47,
    ``ConstFeatureGen``; ``gaussian_features=True``: N(0, 1) per node, the cfg 5 recipe of SURVEY.md 8d).  Returns :class:`FlatGraphs` on ``device``, ordered
    (class-major, then context, then index) like the reference's dict of lists."""
    device = torch.device(device)
    gen = torch.Generator(device=device).manual_seed(int(seed))
    K, nctx = len(MOTIFS), len(contexts)
    G = K * nctx * num_per_class
    label = torch.arange(K, device=device).repeat_interleave(nctx * num_per_class)
    ctx = torch.tensor([0 if c == "tree" else 1 for c in contexts], device=device).repeat_interleave(num_per_class).repeat(K)
    m_nodes, m_edges, m_ned = _motif_table(device)
    # ---- base sizes ----
    if base_nodes is not None:
        nb = torch.randint(int(base_nodes[0]), int(base_nodes[1]), (G,), generator=gen, device=device)
        tree_r = 2
    else:
        tree_r = int(node_num)
        n_tree = sum(tree_r ** i for i in range(tree_height + 1))
        nb = torch.where(ctx == 0, torch.full((G,), n_tree, device=device), torch.full((G,), node_num ** 2, device=device))
    nm = m_nodes[label]
    n = nb + nm
    node_ptr = torch.zeros(G + 1, dtype=torch.long, device=device)
    node_ptr[1:] = torch.cumsum(n, 0)
    gid_all = torch.arange(G, device=device)
    # ---- undirected edges as (graph, a, b) triples ----
    parts = []
    is_tree = ctx == 0
    if bool(is_tree.any()):                                  # balanced r-ary tree: node i > 0 hangs under (i - 1) // r
        gt = gid_all[is_tree]
        cnt = nb[is_tree] - 1
        g_rep = gt.repeat_interleave(cnt)
        start = torch.cumsum(cnt, 0) - cnt
        child = torch.arange(int(cnt.sum()), device=device) - start.repeat_interleave(cnt) + 1
        parts.append((g_rep, (child - 1) // tree_r, child))
    if bool((~is_tree).any()):
        gb = gid_all[~is_tree]
        s, d, g_loc = _ba_edges_batched(nb[~is_tree], ba_m, gen, device)
        parts.append((gb[g_loc], s, d))
    # motif edges, shifted behind the base; the attach edge (motif node 0 -- the plugin, a base node drawn uniformly:
    # np.random.choice(n_basis, 1), synthetic_structsim.py:247-264 with rdm_basis_plugins=True)
    me = m_edges[label]                                      # [G, 8, 2]
    ok = me[:, :, 0] >= 0
    g_rep = gid_all.unsqueeze(1).expand(-1, 8)[ok]
    parts.append((g_rep, me[:, :, 0][ok] + nb[g_rep], me[:, :, 1][ok] + nb[g_rep]))
    if random_plugin:
        plugin = (torch.rand(G, generator=gen, device=device, dtype=torch.float64) * nb).long().clamp(max=nb - 1)
    else:
        plugin = torch.zeros(G, dtype=torch.long, device=device)
    parts.append((gid_all, plugin, nb.clone()))
    g_e = torch.cat([p[0] for p in parts])
    a_e = torch.cat([p[1] for p in parts])
    b_e = torch.cat([p[2] for p in parts])
    lo, hi = torch.minimum(a_e, b_e), torch.maximum(a_e, b_e)
    nmax = int(n.max())
    key = (g_e * nmax + lo) * nmax + hi
    key = torch.unique(key)
    # ---- noise: int(noise * #edges) random extra edges per graph between unconnected distinct nodes (gengraph.py:13-33)
    if noise > 0:
        e_cnt = torch.bincount(torch.div(key, nmax * nmax, rounding_mode="floor"), minlength=G)
        want = (e_cnt.to(torch.float64) * noise).floor().long()
        have = torch.zeros_like(want)
        for _ in range(64):                                  # rejection rounds for the whole batch
            need = want - have
            if not bool((need > 0).any()):
                break
            g_rep = gid_all.repeat_interleave(need.clamp(min=0))
            u = (torch.rand(g_rep.numel(), generator=gen, device=device) * n[g_rep]).long()
            v = (torch.rand(g_rep.numel(), generator=gen, device=device) * n[g_rep]).long()
            k_new = (g_rep * nmax + torch.minimum(u, v)) * nmax + torch.maximum(u, v)
            k_new = torch.unique(k_new[u != v])
            k_new = k_new[~torch.isin(k_new, key)]
            # a round may propose more than a graph still needs after de-duplication: keep the first `need` per graph
            g_new = torch.div(k_new, nmax * nmax, rounding_mode="floor")
            first = torch.searchsorted(g_new, gid_all)
            rank = torch.arange(k_new.numel(), device=device) - first[g_new]
            k_new = k_new[rank < need[g_new]]
            have = have + torch.bincount(torch.div(k_new, nmax * nmax, rounding_mode="floor"), minlength=G)
            key = torch.cat([key, k_new])
        key = torch.sort(key).values
    # ---- both directions, sorted by (graph, source, target) ----
    g_u = torch.div(key, nmax * nmax, rounding_mode="floor")
    lo = torch.div(key, nmax, rounding_mode="floor") % nmax
    hi = key % nmax
    g2 = torch.cat([g_u, g_u])
    s2 = torch.cat([lo, hi])
    d2 = torch.cat([hi, lo])
    order = torch.argsort((g2 * nmax + s2) * nmax + d2)
    g2, s2, d2 = g2[order], s2[order], d2[order]
    edge_ptr = torch.zeros(G + 1, dtype=torch.long, device=device)
    edge_ptr[1:] = torch.cumsum(torch.bincount(g2, minlength=G), 0)
    edge_index = torch.stack([s2, d2])
    # ---- features ----
    N = int(node_ptr[-1])
    if feature_dim == -1:
        deg = torch.bincount(node_ptr[g2] + s2, minlength=N)
        feat = torch.nn.functional.one_hot(deg.clamp(max=max_degree - 1), max_degree).to(torch.float32)
    elif gaussian_features:
        feat = torch.randn(N, int(feature_dim), generator=gen, device=device)
    else:
        feat = torch.rand(G, int(feature_dim), generator=gen, device=device).repeat_interleave(n, dim=0)
    return FlatGraphs(node_ptr, edge_ptr, edge_index, feat, label.clone(), ctx.clone())


def dataset_bias_split(graphs, bias=0.9, split=(7, 1, 2), total=20000, seed=666):
    """utils.py:121-159: class-balanced train / val / test index sets.  Train and val take a fraction ``bias`` of
    the house class from the tree context and ``1 - bias`` of every other class (the rest from BA); test is 50 / 50.
    ``graphs`` must come from :func:`generate_spmotif` (class-major, context, index order).  Returns the three
    shuffled index tensors and ``the`` (mean edge count of the first tree / BA graph of every class, used by the
    reference to tell the contexts apart, utils.py:156-158)."""
    K = len(MOTIFS)
    tr, va, te = [float(s) / 10 for s in split]
    assert abs(tr + va + te - 1.0) < 1e-9
    per = [total * f / K for f in (tr, va, te)]
    y, ctx = graphs.y.cpu(), graphs.context.cpu()
    out = [[], [], []]
    edges_num = 0
    for k, shape in enumerate(MOTIFS):
        b = bias if shape == "house" else 1.0 - bias
        t_idx = torch.nonzero((y == k) & (ctx == 0)).view(-1)
        b_idx = torch.nonzero((y == k) & (ctx == 1)).view(-1)
        ntr_t, ntr_b = int(per[0] * b), int(per[0] * (1 - b))
        nva_t, nva_b = int(per[1] * b), int(per[1] * (1 - b))
        nte_t, nte_b = int(per[2] * 0.5), int(per[2] * 0.5)
        if ntr_t + nva_t + nte_t > t_idx.numel() or ntr_b + nva_b + nte_b > b_idx.numel():
            raise ValueError("dataset_bias_split: not enough graphs of class %s for total=%d" % (shape, total))
        out[0] += [t_idx[:ntr_t], b_idx[:ntr_b]]
        out[1] += [t_idx[ntr_t:ntr_t + nva_t], b_idx[ntr_b:ntr_b + nva_b]]
        out[2] += [t_idx[ntr_t + nva_t:ntr_t + nva_t + nte_t], b_idx[ntr_b + nva_b:ntr_b + nva_b + nte_b]]
        ep = graphs.edge_ptr.cpu()
        edges_num += int(ep[t_idx[0] + 1] - ep[t_idx[0]]) + int(ep[b_idx[0] + 1] - ep[b_idx[0]])
    g = torch.Generator().manual_seed(int(seed))
    res = []
    for parts in out:
        idx = torch.cat(parts)
        res.append(idx[torch.randperm(idx.numel(), generator=g)])
    return res[0], res[1], res[2], float(edges_num) / (K * 2)


# ------------------------------------------------------------------------------------------------
# TU text format
# ------------------------------------------------------------------------------------------------

def _read_txt(path, dtype):
    with open(path) as f:
        rows = [ln.replace(",", " ").split() for ln in f if ln.strip()]
    return np.asarray(rows, dtype=dtype)


def expand_features(graphs, num_node_labels=0, node_labels=None, degree=True, onehot_maxdeg=100):
    """feature_expansion.py:56-59,101-106 as used by main_real.py: [node-label one-hot | degree | one-hot(min(degree,
    onehot_maxdeg)) (onehot_maxdeg + 1 columns)].  MUTAG: 7 + 1 + 101 = 109 features."""
    dev = graphs.y.device
    N = int(graphs.node_ptr[-1])
    gid = torch.repeat_interleave(torch.arange(len(graphs), device=dev), graphs.edge_ptr[1:] - graphs.edge_ptr[:-1])
    deg = torch.bincount(graphs.node_ptr[gid] + graphs.edge_index[0], minlength=N)
    cols = []
    if node_labels is not None and num_node_labels > 0:
        cols.append(torch.nn.functional.one_hot(torch.as_tensor(node_labels, device=dev).long(), num_node_labels).float())
    elif graphs.feat is not None and graphs.feat.numel() > 0:
        cols.append(graphs.feat)
    if degree:
        cols.append(deg.float().unsqueeze(1))
    if onehot_maxdeg is not None and onehot_maxdeg > 0:
        cols.append(torch.nn.functional.one_hot(deg.clamp(max=onehot_maxdeg), onehot_maxdeg + 1).float())
    graphs.feat = torch.cat(cols, dim=1)
    return graphs


def read_tu_dataset(folder, name, degree=True, onehot_maxdeg=100, device="cpu"):
    """``read_tu_data`` (tu_dataset.py:73-88 -> torch_geometric.io) + the reference's feature expansion, vectorised.
    Node ids in ``<name>_A.txt`` are 1-based and global; graphs are contiguous in ``_graph_indicator.txt``."""
    p = lambda s: os.path.join(folder, "%s_%s.txt" % (name, s))
    A = _read_txt(p("A"), np.int64) - 1                                   # [E, 2] (row, col)
    ind = _read_txt(p("graph_indicator"), np.int64).reshape(-1) - 1       # [N]
    ylab = _read_txt(p("graph_labels"), np.int64).reshape(-1)
    G, N = int(ind.max()) + 1, ind.shape[0]
    if np.any(np.diff(ind) < 0):
        raise ValueError("read_tu_dataset: graph_indicator is not sorted")
    _, y = np.unique(ylab, return_inverse=True)                            # labels -> 0 .. C-1 (PyG does the same)
    node_ptr = np.zeros(G + 1, dtype=np.int64)
    node_ptr[1:] = np.cumsum(np.bincount(ind, minlength=G))
    # PyG: remove self loops, coalesce (sort by (row, col), drop duplicates)
    A = A[A[:, 0] != A[:, 1]]
    A = np.unique(A, axis=0)
    ge = ind[A[:, 0]]
    order = np.lexsort((A[:, 1], A[:, 0], ge))
    A, ge = A[order], ge[order]
    edge_ptr = np.zeros(G + 1, dtype=np.int64)
    edge_ptr[1:] = np.cumsum(np.bincount(ge, minlength=G))
    local = A - node_ptr[ge][:, None]
    feats = []
    nl = None
    n_labels = 0
    if os.path.exists(p("node_attributes")):
        feats.append(_read_txt(p("node_attributes"), np.float32).reshape(N, -1))
    if os.path.exists(p("node_labels")):
        nl = _read_txt(p("node_labels"), np.int64).reshape(N, -1)[:, 0]
        nl = nl - nl.min()
        n_labels = int(nl.max()) + 1
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)
    fg = FlatGraphs(t(node_ptr, torch.long), t(edge_ptr, torch.long), t(local.T, torch.long),
                    t(np.concatenate(feats, 1) if feats else np.zeros((N, 0), np.float32), torch.float32), t(y, torch.long))
    if nl is not None:
        lab = torch.nn.functional.one_hot(t(nl, torch.long), n_labels).float()
        fg.feat = torch.cat([fg.feat, lab], dim=1) if fg.feat.numel() else lab
    return expand_features(fg, degree=degree, onehot_maxdeg=onehot_maxdeg)
