"""Mini-graph containers and synthetic SPMotif-style batches.

The reference consumes PyG ``Data`` / ``Batch`` objects produced by
``torch_geometric.data.DataLoader`` (train_causal.py:13-15,171-176).  PyG is not
installable here, so this module provides duck-typed equivalents with the same
attribute names and the same collate rule (nodes concatenated, ``edge_index``
offset by the running node count, ``batch`` = graph id per node, ``y``
concatenated), which is all the hot path reads (model.py:87-89).

The generator follows the *shape statistics* of the reference's synthetic data
(utils.py:38-159, synthetic_structsim.py:49-204, gengraph.py:13-33,
featgen.py:19-28): a tree or Barabasi-Albert base, one motif (house / cycle-6 /
3x2 grid / diamond) attached by one edge, a fraction of random extra edges,
one-hot(min(degree, max_degree-1)) node features, label = motif id, base type
correlated with the label through ``bias``.  Edges are stored in both
directions, sorted by source like ``from_networkx`` emits them (utils.py:55).
It is plain numpy -- offline CPU data preparation, out of the hot path.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["Data", "Batch", "DataLoader", "spmotif_graph", "make_dataset",
           "make_batches", "CONFIGS"]


class Data:
    """One graph: ``x``/``feat`` f32[n,F], ``edge_index`` i64[2,e], ``y`` i64[1]."""

    def __init__(self, x=None, edge_index=None, y=None, feat=None, **kw):
        self.x, self.feat, self.edge_index, self.y = x, feat, edge_index, y
        self.batch = None
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        t = self.x if self.x is not None else self.feat
        return int(t.size(0))

    @property
    def num_edges(self):
        return int(self.edge_index.size(1))

    def _tensors(self):
        return [k for k, v in self.__dict__.items() if torch.is_tensor(v)]

    def to(self, device, non_blocking=False):
        out = self.__class__.__new__(self.__class__)
        out.__dict__.update(self.__dict__)
        for k in self._tensors():
            setattr(out, k, getattr(self, k).to(device, non_blocking=non_blocking))
        return out

    def pin_memory(self):
        for k in self._tensors():
            setattr(self, k, getattr(self, k).pin_memory())
        return self


class Batch(Data):
    """Disjoint union of graphs (PyG ``Batch.from_data_list`` collate rule)."""

    num_graphs: int = 0

    @staticmethod
    def from_data_list(graphs):
        xs, feats, eis, ys, bs = [], [], [], [], []
        off = 0
        for g, d in enumerate(graphs):
            n = d.num_nodes
            if d.x is not None:
                xs.append(d.x)
            if d.feat is not None:
                feats.append(d.feat)
            eis.append(d.edge_index + off)
            ys.append(d.y.view(-1))
            bs.append(torch.full((n,), g, dtype=torch.long))
            off += n
        b = Batch(x=torch.cat(xs) if xs else None,
                  feat=torch.cat(feats) if feats else None,
                  edge_index=torch.cat(eis, dim=1), y=torch.cat(ys))
        b.batch = torch.cat(bs)
        b.num_graphs = len(graphs)
        return b


class DataLoader:
    """``torch_geometric.data.DataLoader(dataset, batch_size, shuffle)`` stand-in.

    ``rank`` / ``world_size`` select graphs ``rank::world_size`` of the epoch
    permutation (DistributedSampler-style)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, seed=0, rank=0, world_size=1,
                 drop_last=False):
        self.dataset, self.batch_size, self.shuffle = list(dataset), batch_size, shuffle
        self.rank, self.world_size, self.drop_last = rank, world_size, drop_last
        self._rng = np.random.RandomState(seed)

    def __len__(self):
        n = len(range(self.rank, len(self.dataset), self.world_size))
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def __iter__(self):
        idx = np.arange(len(self.dataset))
        if self.shuffle:
            self._rng.shuffle(idx)
        idx = idx[self.rank::self.world_size]
        for s in range(0, len(idx), self.batch_size):
            chunk = idx[s:s + self.batch_size]
            if self.drop_last and len(chunk) < self.batch_size:
                return
            yield Batch.from_data_list([self.dataset[i] for i in chunk])


# --------------------------------------------------------------------------
# SPMotif-style generator
# --------------------------------------------------------------------------

MOTIFS = ("house", "cycle", "grid", "diamond")          # utils.py:61 class_list


def _motif_edges(kind):
    """(num_nodes, undirected edge list); synthetic_structsim.py:49-69,114-125,169-204."""
    if kind == "house":
        return 5, [(0, 1), (1, 2), (2, 3), (3, 0), (4, 0), (4, 1)]
    if kind == "cycle":
        return 6, [(i, (i + 1) % 6) for i in range(6)]
    if kind == "grid":                                   # nx.grid_graph([3, 2])
        return 6, [(0, 1), (1, 2), (3, 4), (4, 5), (0, 3), (1, 4), (2, 5)]
    if kind == "diamond":
        return 6, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0), (5, 1), (4, 2)]
    raise ValueError(kind)


def _tree_edges(n):
    """Balanced binary tree truncated to n nodes (nx.balanced_tree(2, h) prefix)."""
    return [((i - 1) // 2, i) for i in range(1, n)]


def _ba_edges(n, m, rng):
    """Barabasi-Albert preferential attachment (networkx 3 convention: star on
    m+1 nodes, then every new node attaches to m distinct degree-weighted targets)."""
    m = max(1, min(m, n - 1))
    edges = [(0, i) for i in range(1, m + 1)]
    rep = [0] * m + list(range(1, m + 1))
    for v in range(m + 1, n):
        tg = set()
        while len(tg) < m:
            tg.add(rep[rng.randint(len(rep))])
        for t in tg:
            edges.append((t, v))
        rep.extend(tg)
        rep.extend([v] * m)
    return edges


def spmotif_graph(rng, base, motif, n_base, noise=0.1, ba_m=2, max_degree=10,
                  feature_dim=None):
    """One SPMotif-style graph as a ``Data`` (label = motif id)."""
    und = _tree_edges(n_base) if base == "tree" else _ba_edges(n_base, ba_m, rng)
    nm, me = _motif_edges(motif)
    und += [(a + n_base, b + n_base) for a, b in me]
    und.append((n_base, int(rng.randint(n_base))))       # attach; synthetic_structsim.py:264
    n = n_base + nm
    es = {(min(a, b), max(a, b)) for a, b in und}
    for _ in range(int(len(es) * noise)):                # gengraph.py:13-33
        while True:
            u, v = int(rng.randint(n)), int(rng.randint(n))
            if u != v and (min(u, v), max(u, v)) not in es:
                break
        es.add((min(u, v), max(u, v)))
    e = np.array(sorted(es), dtype=np.int64)
    ei = np.concatenate([e, e[:, ::-1]], axis=0)
    ei = ei[np.lexsort((ei[:, 1], ei[:, 0]))].T.copy()   # row-sorted, both directions
    if feature_dim is None:                              # one-hot degree; featgen.py:19-28
        deg = np.bincount(ei[0], minlength=n)
        feat = np.eye(max_degree, dtype=np.float32)[np.minimum(deg, max_degree - 1)]
    elif feature_dim == "mutag":
        # main_real.py's MUTAG features: 7-way node-label one-hot | degree | one-hot(min(degree,
        # 100)) = 109 columns (feature_expansion.py:56-59,101-106); the node labels are synthetic (uniform)
        deg = np.bincount(ei[0], minlength=n)
        lab = np.eye(7, dtype=np.float32)[rng.randint(7, size=n)]
        feat = np.concatenate([lab, deg[:, None].astype(np.float32),
                               np.eye(101, dtype=np.float32)[np.minimum(deg, 100)]], axis=1)
    else:
        feat = rng.standard_normal((n, feature_dim)).astype(np.float32)
    return Data(feat=torch.from_numpy(feat), edge_index=torch.from_numpy(ei),
                y=torch.tensor([MOTIFS.index(motif)], dtype=torch.long))


# Named workloads (BASELINE.json `configs`).
CONFIGS = {
    # cfg 1/2/4: SPMotif bias=0.9, ~25 nodes / ~50-55 edge_index columns per graph, F=10.
    "spmotif": dict(avg_nodes=25, ba_m=1, noise=0.1, feature_dim=None, max_degree=10,
                    num_classes=4, batch_size=128),
    # reference-default generator size (opts.py:18: node_num=15 -> ~240 nodes), context only.
    "spmotif_refsize": dict(avg_nodes=240, ba_m=2, noise=0.1, feature_dim=None, max_degree=10,
                            num_classes=4, batch_size=128),
    # cfg 3: MUTAG-shaped (17.93 nodes, 39.6 directed edges), F=109, 2 classes.
    "mutag": dict(avg_nodes=18, ba_m=1, noise=0.1, feature_dim="mutag", max_degree=10,
                  num_classes=2, batch_size=128),
    # cfg 5: large synthetic, ~200 nodes / ~800 edge columns, 64-d features, batch 512.
    # (BA(m=2)-based graphs only -> 4 edge_index columns per node)
    "large": dict(avg_nodes=200, ba_m=2, noise=0.0, feature_dim=64, max_degree=10,
                  num_classes=4, batch_size=512, base="ba"),
}


def make_dataset(num_graphs, seed=666, bias=0.9, avg_nodes=25, ba_m=1, noise=0.1,
                 feature_dim=None, max_degree=10, num_classes=4, base=None, **_):
    """``num_graphs`` graphs with balanced labels; P(tree base | house) = bias,
    P(tree base | other motif) = 1 - bias (utils.py:126,142-150)."""
    rng = np.random.RandomState(seed)
    out = []
    for i in range(num_graphs):
        label = i % num_classes
        motif = MOTIFS[label % len(MOTIFS)]
        p_tree = bias if motif == "house" else 1.0 - bias
        base_kind = "tree" if rng.rand() < p_tree else "ba"
        if base is not None:                                 # a workload made of one base type only
            base_kind = base
        nm = _motif_edges(motif)[0]
        lo = max(3, int(round(0.6 * (avg_nodes - nm))))
        hi = max(lo + 1, int(round(1.4 * (avg_nodes - nm))) + 1)
        g = spmotif_graph(rng, base_kind, motif, int(rng.randint(lo, hi)), noise, ba_m,
                          max_degree, feature_dim)
        g.y = torch.tensor([label], dtype=torch.long)
        out.append(g)
    order = rng.permutation(num_graphs)
    return [out[i] for i in order]


def make_batches(workload="spmotif", num_batches=16, seed=666, batch_size=None, **over):
    """Pre-collated batches of a named workload (list of ``Batch``)."""
    cfg = dict(CONFIGS[workload])
    cfg.update(over)
    bs = batch_size or cfg["batch_size"]
    ds = make_dataset(num_batches * bs, seed=seed, **{k: v for k, v in cfg.items() if k != "batch_size"})
    return [Batch.from_data_list(ds[i * bs:(i + 1) * bs]) for i in range(num_batches)]
