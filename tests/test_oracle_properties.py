"""The oracle's restatements of the THIRD-PARTY primitives (PyG 1.x propagate / GCN normalisation call
sites gcn_conv.py:44-104, GATConv model.py:340, GINConv model.py:187-193, global_add_pool model.py:115-116,
torch_scatter.scatter_add gcn_conv.py:66) against independent closed forms.

The golden vectors (tests/golden) pin the in-repo arithmetic to the reference's own model.py, but both the oracle
and the stand-in PyG used to freeze them restate the third-party pieces ("parity unpinned" at that boundary,
DESIGN.md section 2).  SURVEY.md section 8(c) names what pins those instead: closed-form small-graph cases, fp64
gradient checks and property tests.  These are they: dense-matrix / per-node-loop formulas written from the
published definitions (Kipf & Welling's D^-1/2 (A + I) D^-1/2, Velickovic's attention, Xu's GIN sum), never
sharing code with oracle/cal_oracle.py, compared in fp64.  CPU only.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F
pytest.importorskip("hypothesis")                     # (a missing package must skip this file, not abort a collection)
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import cal_oracle as O
from tests.util import random_case

TOL64 = 1e-11
SET = dict(max_examples=30, deadline=None, derandomize=True)


# ---------------------------------------------------------------------------------------------------
# graph strategy: directed multigraphs with self loops, duplicate columns, one-way edges, isolated nodes
# ---------------------------------------------------------------------------------------------------
@st.composite
def graphs(draw, max_nodes=9, max_edges=24):
    n = draw(st.integers(1, max_nodes))
    e = draw(st.integers(0, max_edges))
    src = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    dst = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    seed = draw(st.integers(0, 2 ** 16))
    ei = torch.tensor([src, dst], dtype=torch.long).view(2, e)
    return n, ei, seed


def dense_gcn_matrix(n, ei, w):
    """M[r, c] of 'out[c] = sum_r M[r, c] h[r]': self loops dropped, parallel edges add, unit loops added,
    degree = ROW sum (gcn_conv.py:66 sums by edge_index[0]), inf -> 0.  Plain Python loops."""
    A = [[0.0] * n for _ in range(n)]
    for k in range(ei.size(1)):
        r, c = int(ei[0, k]), int(ei[1, k])
        if r != c:
            A[r][c] += float(w[k])
    for i in range(n):
        A[i][i] += 1.0
    deg = [sum(A[r]) for r in range(n)]
    dis = [(1.0 / math.sqrt(d)) if d > 0 else 0.0 for d in deg]
    return torch.tensor([[dis[r] * A[r][c] * dis[c] for c in range(n)] for r in range(n)], dtype=torch.float64)


# ---------------------------------------------------------------------------------------------------
# GCN normalisation and message passing
# ---------------------------------------------------------------------------------------------------
def test_gcn_norm_closed_form_on_a_path():
    """Path 0 - 1 - 2 stored both ways: degrees with the self loop 2, 3, 2."""
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    ei2, norm = O.gcn_norm(ei, 3, None, False, torch.float64)
    assert ei2.tolist() == [[0, 1, 1, 2, 0, 1, 2], [1, 0, 2, 1, 0, 1, 2]]            # loops appended LAST
    s6 = 1.0 / math.sqrt(6.0)
    want = torch.tensor([s6, s6, s6, s6, 0.5, 1.0 / 3.0, 0.5], dtype=torch.float64)
    assert torch.allclose(norm, want, atol=1e-15)
    # the aggregate of constant features: row c of out = sum of column c of M
    conv = O.GCNConv(1, 1).double()
    with torch.no_grad():
        conv.weight.fill_(1.0)
        conv.bias.fill_(0.25)
    out = conv(torch.ones(3, 1, dtype=torch.float64), ei)
    assert torch.allclose(out.view(-1), torch.tensor([0.5 + s6, 1.0 / 3.0 + 2 * s6, 0.5 + s6], dtype=torch.float64) + 0.25)


def test_gcn_norm_self_loops_are_replaced_not_kept():
    """gcn_conv.py:56-57: an input self loop (and its weight) is dropped; the appended loop has weight 1."""
    ei = torch.tensor([[0, 0, 1], [0, 1, 0]])
    w = torch.tensor([7.0, 2.0, 3.0], dtype=torch.float64)
    ei2, norm = O.gcn_norm(ei, 2, w, False, torch.float64)
    assert ei2.tolist() == [[0, 1, 0, 1], [1, 0, 0, 1]]
    d0, d1 = 2.0 + 1.0, 3.0 + 1.0
    want = torch.tensor([2.0 / math.sqrt(d0 * d1), 3.0 / math.sqrt(d0 * d1), 1.0 / d0, 1.0 / d1], dtype=torch.float64)
    assert torch.allclose(norm, want, atol=1e-15)


def test_gcn_norm_isolated_node_parallel_edges_and_one_way_edges_by_hand():
    """Node 2 is isolated (degree 1 from its appended loop: it keeps exactly its own transformed row); the column
    0 -> 1 appears twice (both count in node 0's ROW degree); 1 -> 0 is absent (a one-way edge: node 1's degree is
    its loop only, yet it RECEIVES from node 0 -- degree by source, aggregation at target, gcn_conv.py:66,92-97)."""
    ei = torch.tensor([[0, 0], [1, 1]])
    _, norm = O.gcn_norm(ei, 3, None, False, torch.float64)
    d0, d1, d2 = 3.0, 1.0, 1.0
    want = torch.tensor([1 / math.sqrt(d0 * d1)] * 2 + [1 / d0, 1 / d1, 1 / d2], dtype=torch.float64)
    assert torch.allclose(norm, want, atol=1e-15)
    conv = O.GCNConv(2, 2).double()
    with torch.no_grad():
        conv.weight.copy_(torch.eye(2))
    x = torch.tensor([[1.0, 2.0], [10.0, 20.0], [100.0, 200.0]], dtype=torch.float64)
    out = conv(x, ei)
    assert torch.allclose(out[2], x[2])                                            # isolated: itself / 1
    assert torch.allclose(out[0], x[0] / 3.0)                                      # nothing arrives at node 0
    assert torch.allclose(out[1], x[1] + 2.0 * x[0] / math.sqrt(3.0))              # two parallel messages


def test_oracle_ops_pass_torch_gradcheck():
    """torch.autograd.gradcheck (fp64) of the oracle's GCNConv w.r.t. features, EDGE WEIGHTS and parameters, of the
    per-target softmax and of GATConv w.r.t. features."""
    gen = torch.Generator().manual_seed(11)
    n, e = 5, 9
    ei = torch.randint(0, n, (2, e), generator=gen)
    x = torch.randn(n, 3, generator=gen, dtype=torch.float64, requires_grad=True)
    w = (0.3 + torch.rand(e, generator=gen, dtype=torch.float64)).requires_grad_(True)
    conv = O.GCNConv(3, 2).double()
    # (W and b are the module's own tensors: gradcheck perturbs them in place)
    assert torch.autograd.gradcheck(lambda xx, ww, W, b: conv(xx, ei, ww), (x, w, conv.weight, conv.bias))
    sc = torch.randn(e, 2, generator=gen, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda s_: O.segment_softmax(s_, ei[1], n), (sc,))
    gat = O.GATConv(3, 2, heads=2, dropout=0.0).double()
    assert torch.autograd.gradcheck(lambda xx: gat(xx, ei), (x,))


@settings(**SET)
@given(graphs(), st.booleans())
def test_gcnconv_equals_dense_normalised_adjacency(g, weighted):
    n, ei, seed = g
    gen = torch.Generator().manual_seed(seed)
    w = (0.05 + torch.rand(ei.size(1), generator=gen, dtype=torch.float64)) if weighted else None
    x = torch.randn(n, 5, generator=gen, dtype=torch.float64)
    conv = O.GCNConv(5, 4).double()
    with torch.no_grad():
        conv.bias.copy_(torch.randn(4, generator=gen, dtype=torch.float64))
    M = dense_gcn_matrix(n, ei, w if weighted else torch.ones(ei.size(1)))
    want = M.t() @ (x @ conv.weight.detach()) + conv.bias.detach()
    got = conv(x, ei, w)
    assert (got - want).abs().max() < TOL64 * max(1.0, float(want.abs().max()))


@settings(**SET)
@given(graphs())
def test_gcnconv_gfn_is_the_bare_transform(g):
    """gcn_conv.py:76-77: conv_feat (gfn=True) returns x @ W -- no aggregation and NO bias."""
    n, ei, seed = g
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, generator=gen, dtype=torch.float64)
    conv = O.GCNConv(3, 4, gfn=True).double()
    with torch.no_grad():
        conv.bias.fill_(5.0)
    assert torch.equal(conv(x, ei), x @ conv.weight)


@settings(max_examples=12, deadline=None, derandomize=True)
@given(graphs(max_nodes=6, max_edges=10))
def test_gradient_flows_through_the_weighted_normalisation(g):
    """model.py:112-113: edge_att enters the masked convs through `norm`, i.e. through BOTH endpoints'
    degrees.  Autograd of the oracle against (a) autograd of the dense closed form and (b) central differences."""
    n, ei, seed = g
    if ei.size(1) == 0:
        return
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, generator=gen, dtype=torch.float64)
    conv = O.GCNConv(3, 2).double()
    coef = torch.randn(n, 2, generator=gen, dtype=torch.float64)
    w = (0.2 + torch.rand(ei.size(1), generator=gen, dtype=torch.float64)).requires_grad_(True)

    def dense(wv):
        A = torch.zeros(n, n, dtype=torch.float64)
        keep = ei[0] != ei[1]
        A = A.index_put((ei[0][keep], ei[1][keep]), wv[keep], accumulate=True) + torch.eye(n, dtype=torch.float64)
        dis = A.sum(1).pow(-0.5)
        M = dis.view(-1, 1) * A * dis.view(1, -1)
        return ((M.t() @ (x @ conv.weight) + conv.bias) * coef).sum()

    g_or, = torch.autograd.grad((conv(x, ei, w) * coef).sum(), w)
    g_dn, = torch.autograd.grad(dense(w), w)
    assert (g_or - g_dn).abs().max() < 1e-10 * max(1.0, float(g_dn.abs().max()))
    eps = 1e-6
    for k in range(min(ei.size(1), 4)):
        d = torch.zeros_like(w)
        d[k] = eps
        with torch.no_grad():
            fd = (float((conv(x, ei, w + d) * coef).sum()) - float((conv(x, ei, w - d) * coef).sum())) / (2 * eps)
        assert abs(fd - float(g_or[k])) < 1e-6 * max(1.0, abs(fd))


# ---------------------------------------------------------------------------------------------------
# GATConv (1.x) and GINConv
# ---------------------------------------------------------------------------------------------------
def loop_gat(n, ei, xw, att, bias, heads, ch, slope=0.2):
    """Velickovic et al. with PyG 1.x conventions, one target node at a time: incoming edges = the non-loop
    columns with edge_index[1] == i (parallel edges stay separate terms) + ONE self loop; e = leaky_relu(a_l . x'_i
    + a_r . x'_j), att = [a_l || a_r] with the TARGET first; softmax over the incoming edges."""
    out = torch.zeros(n, heads * ch, dtype=torch.float64)
    xh = xw.view(n, heads, ch)
    for i in range(n):
        srcs = [int(ei[0, k]) for k in range(ei.size(1)) if int(ei[1, k]) == i and int(ei[0, k]) != i] + [i]
        for h in range(heads):
            a_l, a_r = att[0, h, :ch], att[0, h, ch:]
            e = torch.stack([F.leaky_relu((xh[i, h] * a_l).sum() + (xh[j, h] * a_r).sum(), slope) for j in srcs])
            p = torch.exp(e - e.max())
            p = p / (p.sum() + 1e-16)
            for pj, j in zip(p, srcs):
                out[i, h * ch:(h + 1) * ch] += pj * xh[j, h]
    return out + bias


@settings(max_examples=20, deadline=None, derandomize=True)
@given(graphs(max_nodes=7, max_edges=16), st.sampled_from([(1, 4), (2, 3), (4, 2)]))
def test_gatconv_equals_per_node_attention(g, hc):
    n, ei, seed = g
    heads, ch = hc
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 5, generator=gen, dtype=torch.float64)
    conv = O.GATConv(5, ch, heads=heads, dropout=0.0).double()
    with torch.no_grad():
        conv.bias.copy_(torch.randn(heads * ch, generator=gen, dtype=torch.float64))
    want = loop_gat(n, ei, x @ conv.weight.detach(), conv.att.detach(), conv.bias.detach(), heads, ch)
    got = conv(x, ei)
    assert (got - want).abs().max() < TOL64 * max(1.0, float(want.abs().max()))


def test_gat_attention_sums_to_one_and_dropout_mask_rescales():
    """softmax over the edges of a target sums to 1 (up to the 1e-16 of PyG 1.x); an injected keep mask
    multiplies alpha by mask / (1 - p) -- what F.dropout does with the same mask (model.py:340 dropout=0.2)."""
    gen = torch.Generator().manual_seed(5)
    n, e = 6, 14
    ei = torch.randint(0, n, (2, e), generator=gen)
    score = torch.randn(e, 3, generator=gen, dtype=torch.float64) * 4
    alpha = O.segment_softmax(score, ei[1], n)
    sums = torch.zeros(n, 3, dtype=torch.float64).index_add_(0, ei[1], alpha)
    present = torch.zeros(n, dtype=torch.bool)
    present[ei[1]] = True
    assert torch.allclose(sums[present], torch.ones_like(sums[present]), atol=1e-14)
    assert torch.allclose(O.segment_softmax(score + 100.0, ei[1], n), alpha, atol=1e-13)     # shift invariance

    x = torch.randn(n, 4, generator=gen, dtype=torch.float64)
    conv = O.GATConv(4, 2, heads=2, dropout=0.5).double().train()
    ei_nl, _ = O.remove_self_loops(ei)
    e1 = ei_nl.size(1) + n
    conv.dropout_mask = torch.ones(e1, 2)
    full = conv(x, ei)
    conv.dropout = 0.0
    conv.dropout_mask = None
    assert torch.allclose(full, 2.0 * conv(x, ei) - conv.bias, atol=1e-13)                 # all kept: alpha / (1 - 0.5)


@settings(**SET)
@given(graphs())
def test_ginconv_equals_dense_sum(g):
    """model.py:187-193: nn((1 + 0) x_i + sum_j x_j) over the non-loop columns (parallel edges count twice)."""
    n, ei, seed = g
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 4, generator=gen, dtype=torch.float64)
    conv = O.GINConv(torch.nn.Identity())
    C = torch.zeros(n, n, dtype=torch.float64)
    for k in range(ei.size(1)):
        r, c = int(ei[0, k]), int(ei[1, k])
        if r != c:
            C[c, r] += 1.0
    want = x + C @ x
    assert (conv(x, ei) - want).abs().max() < TOL64 * max(1.0, float(want.abs().max()))


# ---------------------------------------------------------------------------------------------------
# pooling, scatter, loss
# ---------------------------------------------------------------------------------------------------
def test_global_add_pool_and_scatter_add_are_segment_sums_with_empty_segments():
    x = torch.arange(12, dtype=torch.float64).view(6, 2)
    batch = torch.tensor([0, 0, 2, 2, 2, 3])
    got = O.global_add_pool(x, batch, 5)                       # graph 1 and graph 4 have no nodes
    want = torch.tensor([[2.0, 4.0], [0.0, 0.0], [18.0, 21.0], [10.0, 11.0], [0.0, 0.0]], dtype=torch.float64)
    assert torch.equal(got, want)
    assert O.global_add_pool(x, batch).shape == (4, 2)        # size = batch.max() + 1 (model.py:115)
    assert torch.equal(O.scatter_add(torch.ones(4), torch.tensor([1, 1, 1, 0]), 0, 3), torch.tensor([1.0, 3.0, 0.0]))


def test_causal_loss_closed_form():
    """train_causal.py:178-183: KL(uniform || softmax) batchmean = mean_b sum_k (1/C) (log(1/C) - logp_bk); 0 when
    the causal head is uniform; NLL = -mean logp[y]; weights 0.5 / 1 / 0.5 (opts.py:43-45)."""
    gen = torch.Generator().manual_seed(3)
    B, C = 7, 4
    lp = [F.log_softmax(torch.randn(B, C, generator=gen, dtype=torch.float64), -1) for _ in range(3)]
    y = torch.randint(0, C, (B,), generator=gen)
    loss, c_loss, o_loss, co_loss = O.causal_loss(*lp, y, C)
    kl = sum((1.0 / C) * (math.log(1.0 / C) - float(lp[0][b, k])) for b in range(B) for k in range(C)) / B
    assert abs(float(c_loss) - kl) < 1e-12
    assert abs(float(o_loss) + float(sum(lp[1][b, y[b]] for b in range(B))) / B) < 1e-12
    assert abs(float(co_loss) + float(sum(lp[2][b, y[b]] for b in range(B))) / B) < 1e-12
    assert abs(float(loss) - (0.5 * float(c_loss) + float(o_loss) + 0.5 * float(co_loss))) < 1e-12
    uni = torch.full((B, C), math.log(1.0 / C), dtype=torch.float64)
    assert abs(float(O.causal_loss(uni, lp[1], lp[2], y, C)[1])) < 1e-15


# ---------------------------------------------------------------------------------------------------
# whole-model symmetries (all three backbones, fp64)
# ---------------------------------------------------------------------------------------------------
def _model_case(kind, seed, train):
    net, b, perm = random_case(seed=seed, kind=kind, hidden=16, layers=2, batch_size=6, avg_nodes=9)
    net = net.double().train(train)
    b.feat = b.feat.double()
    return net, b, perm


def _relabel_within_graphs(b, gen):
    """A node permutation that keeps `batch` sorted: nodes are shuffled inside every graph."""
    N = b.batch.numel()
    new_of_old = torch.empty(N, dtype=torch.long)
    for g in range(int(b.batch.max()) + 1):
        idx = (b.batch == g).nonzero().view(-1)
        new_of_old[idx] = idx[torch.randperm(idx.numel(), generator=gen)]
    return new_of_old


class _B:
    def __init__(self, feat, edge_index, batch, y, num_graphs):
        self.x, self.feat, self.edge_index, self.batch, self.y, self.num_graphs = None, feat, edge_index, batch, y, num_graphs


@pytest.mark.parametrize("train", [True, False], ids=["train", "eval"])
@pytest.mark.parametrize("kind", ["CausalGCN", "CausalGAT", "CausalGIN"])
def test_model_is_invariant_to_node_relabelling_and_edge_column_order(kind, train):
    """Graph-level outputs do not depend on the node numbering inside a graph nor on the order of the edge_index
    columns (sum aggregation, per-target softmax, BatchNorm over all rows, sum pooling)."""
    net, b, perm = _model_case(kind, 31, train)
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        ref = net(b, eval_random=True, perm=perm)
        new_of_old = _relabel_within_graphs(b, gen)
        old_of_new = torch.empty_like(new_of_old)
        old_of_new[new_of_old] = torch.arange(new_of_old.numel())
        cols = torch.randperm(b.edge_index.size(1), generator=gen)
        b2 = _B(b.feat[old_of_new], new_of_old[b.edge_index][:, cols], b.batch.clone(), b.y, b.num_graphs)
        got = net(b2, eval_random=True, perm=perm)
    for a, w in zip(got, ref):
        assert (a - w).abs().max() < 1e-9


@pytest.mark.parametrize("kind", ["CausalGCN", "CausalGAT", "CausalGIN"])
def test_model_is_equivariant_to_the_order_of_the_graphs(kind):
    """Reordering the graphs of a batch reorders the rows of the three outputs; the random-intervention
    permutation (model.py:147-160: x = xc[perm] + xo) is conjugated with the reordering."""
    net, b, perm = _model_case(kind, 32, True)
    B = b.num_graphs
    gen = torch.Generator().manual_seed(9)
    order = torch.randperm(B, generator=gen)                        # new graph k = old graph order[k]
    counts = torch.bincount(b.batch, minlength=B)
    starts = torch.cumsum(counts, 0) - counts
    old_nodes = torch.cat([torch.arange(int(starts[g]), int(starts[g] + counts[g])) for g in order.tolist()])
    new_of_old = torch.empty_like(old_nodes)
    new_of_old[old_nodes] = torch.arange(old_nodes.numel())
    new_batch = torch.repeat_interleave(torch.arange(B), counts[order])
    inv = torch.empty_like(order)
    inv[order] = torch.arange(B)
    perm2 = inv[perm[order]]                                         # row k mixes xc of old graph perm[order[k]]
    with torch.no_grad():
        ref = net(b, eval_random=True, perm=perm)
        b2 = _B(b.feat[old_nodes], new_of_old[b.edge_index], new_batch, b.y[order], B)
        got = net(b2, eval_random=True, perm=perm2)
    for a, w in zip(got, ref):
        assert (a - w[order]).abs().max() < 1e-9


def test_random_readout_identity_permutation_and_ablations():
    """model.py:145-164 / 97-108: with the identity permutation the `co` head sees xc + xo; with both attention
    ablations the two branches carry 0.5 x each and share the edge weights 0.5."""
    net, b, _ = _model_case("CausalGCN", 33, False)
    net.without_node_attention = net.without_edge_attention = True
    with torch.no_grad():
        c, o, co = net(b, eval_random=False)
        x = net.backbone(b.feat, b.edge_index)
        half = torch.full((b.edge_index.size(1),), 0.5, dtype=torch.float64)
        xc = F.relu(net.context_convs(net.bnc(0.5 * x), b.edge_index, half))
        xo = F.relu(net.objects_convs(net.bno(0.5 * x), b.edge_index, half))
        pc, po = O.global_add_pool(xc, b.batch, b.num_graphs), O.global_add_pool(xo, b.batch, b.num_graphs)
        assert torch.allclose(c, net._readout(pc, "c"), atol=1e-12)
        assert torch.allclose(o, net._readout(po, "o"), atol=1e-12)
        assert torch.allclose(co, net._readout(pc + po, "co"), atol=1e-12)


def test_full_model_gradients_match_central_differences():
    """fp64 central differences of the training loss for entries of the parameters on the longest gradient paths
    (edge attention -> normalisation of both masked convs; the input transform; a backbone layer)."""
    net, b, perm = _model_case("CausalGCN", 34, True)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.momentum = 0.0                                          # (the probes must not move the running statistics)

    def loss():
        return O.causal_loss(*net(b, eval_random=True, perm=perm), b.y, net.num_classes)[0]

    net.zero_grad()
    loss().backward()
    eps = 1e-6
    rng = np.random.RandomState(0)
    for name in ["edge_att_mlp.weight", "node_att_mlp.weight", "conv_feat.weight", "convs.0.weight", "bnc.weight",
                 "context_convs.weight", "fc1_co.weight"]:
        p = dict(net.named_parameters())[name]
        flat, gflat = p.data.view(-1), p.grad.view(-1)
        for k in rng.choice(flat.numel(), size=3, replace=False):
            old = float(flat[k])
            with torch.no_grad():
                flat[k] = old + eps
                up = float(loss())
                flat[k] = old - eps
                dn = float(loss())
                flat[k] = old
            fd = (up - dn) / (2 * eps)
            assert abs(fd - float(gflat[k])) < 2e-6 * max(1e-2, abs(fd)), (name, int(k), fd, float(gflat[k]))


# ---------------------------------------------------------------------------------------------------
# the same closed forms against the code the golden vectors were frozen with: the reference's OWN gcn_conv.py
# (only where /root/reference exists) running over the stand-in PyG of oracle/pyg_shim, and the stand-in's
# GATConv / GINConv.  Run in a subprocess: the stand-in installs itself as `torch_geometric` / `torch_scatter`.
# ---------------------------------------------------------------------------------------------------
_SHIM_CHECK = r"""
import os, sys
root, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [os.path.join(root, "oracle", "pyg_shim"), root]
import torch
from torch_geometric.nn import GATConv, GINConv, global_add_pool
from tests.test_oracle_properties import dense_gcn_matrix, loop_gat
have_ref = os.path.isfile(os.path.join(ref, "gcn_conv.py"))
if have_ref:
    sys.path.insert(0, ref)
    import gcn_conv                                            # the reference's file, unmodified
torch.set_default_dtype(torch.float64)
gen = torch.Generator().manual_seed(1234)
worst = 0.0
for case in range(40):
    n = int(torch.randint(1, 9, (1,), generator=gen))
    e = int(torch.randint(0, 22, (1,), generator=gen))
    ei = torch.randint(0, n, (2, e), generator=gen)
    x = torch.randn(n, 5, generator=gen)
    w = 0.05 + torch.rand(e, generator=gen)
    if have_ref:
        for weighted in (False, True):
            conv = gcn_conv.GCNConv(5, 4)
            with torch.no_grad():
                conv.bias.copy_(torch.randn(4, generator=gen))
            M = dense_gcn_matrix(n, ei, w if weighted else torch.ones(e))
            want = M.t() @ (x @ conv.weight.detach()) + conv.bias.detach()
            got = conv(x, ei, w if weighted else None)
            worst = max(worst, float((got - want).abs().max()) / max(1.0, float(want.abs().max())))
    heads, ch = [(1, 4), (2, 3), (4, 2)][case % 3]
    gat = GATConv(5, ch, heads=heads, dropout=0.0)
    with torch.no_grad():
        gat.bias.copy_(torch.randn(heads * ch, generator=gen))
    want = loop_gat(n, ei, x @ gat.weight.detach(), gat.att.detach(), gat.bias.detach(), heads, ch)
    worst = max(worst, float((gat(x, ei) - want).abs().max()) / max(1.0, float(want.abs().max())))
    gin = GINConv(torch.nn.Identity())
    C = torch.zeros(n, n)
    for k in range(e):
        r, c = int(ei[0, k]), int(ei[1, k])
        if r != c:
            C[c, r] += 1.0
    worst = max(worst, float((gin(x, ei) - (x + C @ x)).abs().max()))
    batch = torch.sort(torch.randint(0, 3, (n,), generator=gen)).values
    pooled = global_add_pool(x, batch, 3)
    for g in range(3):
        worst = max(worst, float((pooled[g] - x[batch == g].sum(0)).abs().max()))
print("have_ref=%d worst=%.3e" % (have_ref, worst))
assert worst < 1e-11
"""


def test_reference_gcnconv_and_the_standin_pyg_match_the_closed_forms():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _SHIM_CHECK, root, "/root/reference"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "worst=" in r.stdout
