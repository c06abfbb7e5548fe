"""CPU restatement of the CAL hot path (CausalGCN / CausalGAT / CausalGIN fwd + loss + bwd).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Pure PyTorch on the
CPU, no dependency on torch_geometric / torch_scatter (neither is installable
here).  Every function cites the reference file:line it follows (paths relative
to the upstream repo, yongduosui/CAL).

Parity status
-------------
* The in-repo arithmetic (``gcn_conv.py``, ``model.py``, ``train_causal.py``)
  is PINNED: ``tests/golden/make_golden.py`` imports the *unmodified* reference
  ``model.py`` / ``gcn_conv.py`` in the build container on top of the minimal
  third-party shim in ``oracle/pyg_shim`` and freezes inputs / outputs / grads
  as fixtures under ``tests/golden``; ``tests/test_oracle_golden.py`` checks
  this file against them.
* The third-party pieces (PyG ``MessagePassing.propagate``, ``GATConv``,
  ``global_add_pool``, ``remove_self_loops`` / ``add_self_loops``, ``glorot``;
  torch_scatter ``scatter_add``) are restated from their published 1.x
  behaviour -- the reference ships no tests or golden vectors of its own and
  the real packages are absent, so at THAT boundary parity is unpinned
  ("parity unpinned": torch-geometric 1.x / torch-scatter 1.x, README.md:22-26).

Two evaluation orders are provided for the sparse ops:
``materialize=True`` replays the op sequence the reference executes on the CPU
(index_select -> mul -> index_add_, norm recomputed in every conv); this is the
one timed as the CPU baseline.  The numerical result is identical either way.
"""
from __future__ import annotations

import math
import random
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import BatchNorm1d, Linear, Parameter

# --------------------------------------------------------------------------
# third-party primitives, restated (PyG 1.x / torch_scatter 1.x semantics)
# --------------------------------------------------------------------------


def scatter_add(src, index, dim=0, dim_size=None):
    """torch_scatter.scatter_add (call sites gcn_conv.py:66, PyG propagate).

    zeros(dim_size) then a sequential ``index_add_`` in element order."""
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def remove_self_loops(edge_index, edge_attr=None):
    """torch_geometric.utils.remove_self_loops (gcn_conv.py:56): keep row != col."""
    row, col = edge_index
    mask = row != col
    edge_attr = edge_attr if edge_attr is None else edge_attr[mask]
    return edge_index[:, mask], edge_attr


def add_self_loops(edge_index, num_nodes):
    """torch_geometric.utils.add_self_loops (gcn_conv.py:57): append [i, i]."""
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1), None


def global_add_pool(x, batch, size=None):
    """torch_geometric.nn.global_add_pool (model.py:27,115-116)."""
    size = int(batch.max()) + 1 if size is None else size
    return scatter_add(x, batch, dim=0, dim_size=size)


def glorot(t):
    """torch_geometric.nn.inits.glorot (gcn_conv.py:40): U(+-sqrt(6/(fan_in+fan_out)))."""
    if t is not None:
        stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        t.data.uniform_(-stdv, stdv)


def zeros(t):
    """torch_geometric.nn.inits.zeros (gcn_conv.py:41)."""
    if t is not None:
        t.data.fill_(0)


def segment_softmax(src, index, num_nodes):
    """torch_geometric.utils.softmax (1.x): exp(x - max_i) / (sum_i + 1e-16)."""
    mx = src.new_full((num_nodes,) + tuple(src.shape[1:]), float("-inf"))
    mx = mx.scatter_reduce(0, index.view(-1, *([1] * (src.dim() - 1))).expand_as(src), src,
                           reduce="amax", include_self=True)
    out = (src - mx[index]).exp()
    return out / (scatter_add(out, index, 0, num_nodes)[index] + 1e-16)


# --------------------------------------------------------------------------
# GCNConv  (gcn_conv.py:10-108)
# --------------------------------------------------------------------------


def gcn_norm(edge_index, num_nodes, edge_weight=None, improved=False, dtype=None):
    """GCNConv.norm, gcn_conv.py:44-70.

    Degree is summed by ROW (source) while messages are aggregated at COL
    (target); self loops are removed (edges and weights), then N unit-weight
    loops are appended LAST; inf -> 0."""
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    edge_weight = edge_weight.view(-1)
    assert edge_weight.size(0) == edge_index.size(1)
    edge_index, edge_weight = remove_self_loops(edge_index, edge_weight)
    edge_index, _ = add_self_loops(edge_index, num_nodes)
    loop_weight = torch.full((num_nodes,), 1 if not improved else 2,
                             dtype=edge_weight.dtype, device=edge_weight.device)
    edge_weight = torch.cat([edge_weight, loop_weight], dim=0)
    row, col = edge_index
    deg = scatter_add(edge_weight, row, dim=0, dim_size=num_nodes)
    dis = deg.pow(-0.5)
    dis = torch.where(dis == float("inf"), torch.zeros_like(dis), dis)
    return edge_index, dis[row] * edge_weight * dis[col]


def propagate_add(edge_index, x, norm, num_nodes):
    """PyG MessagePassing('add').propagate, flow source_to_target
    (gcn_conv.py:92-97): x_j = x[edge_index[0]]; out[edge_index[1]] += norm*x_j."""
    x_j = x.index_select(0, edge_index[0])
    msg = norm.view(-1, 1) * x_j if norm is not None else x_j
    return scatter_add(msg, edge_index[1], dim=0, dim_size=num_nodes)


class GCNConv(nn.Module):
    """gcn_conv.py:10-108 (constructor order and init preserved)."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False,
                 bias=True, edge_norm=True, gfn=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.edge_norm, self.gfn = improved, edge_norm, gfn
        self.weight = Parameter(torch.empty(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        glorot(self.weight)
        zeros(self.bias)

    def forward(self, x, edge_index, edge_weight=None):
        x = torch.matmul(x, self.weight)          # gcn_conv.py:75
        if self.gfn:                              # gcn_conv.py:76-77 (no bias)
            return x
        if self.edge_norm:
            edge_index, norm = gcn_norm(edge_index, x.size(0), edge_weight, self.improved, x.dtype)
        else:
            norm = None
        out = propagate_add(edge_index, x, norm, x.size(0))
        if self.bias is not None:                 # gcn_conv.py:101-104
            out = out + self.bias
        return out


# --------------------------------------------------------------------------
# GATConv  (PyG 1.x; call site model.py:340)
# --------------------------------------------------------------------------


class GATConv(nn.Module):
    """torch_geometric.nn.GATConv, 1.x semantics (concat=True).

    x' = xW; e_ij = leaky_relu_0.2(<x'_i, a_i> + <x'_j, a_j>) with
    att = [a_i || a_j], i = target = edge_index[1], j = source = edge_index[0];
    alpha = softmax over edges sharing the target; dropout(alpha) in training;
    out_i = concat_h sum_j alpha_ij x'_j + bias.  Self loops removed then re-added.

    ``dropout_mask`` (optional, [E', heads] of 0/1) replaces the RNG so that a
    GPU run and the oracle can share the identical mask."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True,
                 negative_slope=0.2, dropout=0.0, bias=True):
        super().__init__()
        assert concat
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope, self.dropout = negative_slope, dropout
        self.weight = Parameter(torch.empty(in_channels, heads * out_channels))
        self.att = Parameter(torch.empty(1, heads, 2 * out_channels))
        self.bias = Parameter(torch.empty(heads * out_channels)) if bias else None
        glorot(self.weight)
        glorot(self.att)
        zeros(self.bias)
        self.dropout_mask = None

    def forward(self, x, edge_index):
        n = x.size(0)
        edge_index, _ = remove_self_loops(edge_index)
        edge_index, _ = add_self_loops(edge_index, n)
        x = torch.matmul(x, self.weight)
        src, dst = edge_index
        xh = x.view(-1, self.heads, self.out_channels)
        x_i, x_j = xh.index_select(0, dst), xh.index_select(0, src)
        alpha = (torch.cat([x_i, x_j], dim=-1) * self.att).sum(dim=-1)
        alpha = F.leaky_relu(alpha, self.negative_slope)
        alpha = segment_softmax(alpha, dst, n)
        if self.dropout_mask is not None and self.training:
            alpha = alpha * self.dropout_mask.to(alpha.dtype) / (1.0 - self.dropout)
        else:
            alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        out = scatter_add(x_j * alpha.view(-1, self.heads, 1), dst, 0, n)
        out = out.view(-1, self.heads * self.out_channels)
        if self.bias is not None:
            out = out + self.bias
        return out


# --------------------------------------------------------------------------
# CausalGCN / CausalGAT  (model.py:12-164, 315-450)
# --------------------------------------------------------------------------


# Test hook: ReLU is not differentiable at 0, so two correct fp32 evaluations whose pre-activations
# differ by one ulp around 0 back-propagate through different activation patterns.  The parity
# tests therefore compare gradients AT THE SAME PATTERN: with RELU_OVERRIDE = {tag: bool mask} the
# oracle uses the given pattern (x * mask) instead of relu(x) for that activation.  Tags: "x1" ..
# "x{L+1}" (backbone), "zc" / "zo" (masked convs), "h1_c" / "h1_o" / "h1_co" (readout fc1).
RELU_OVERRIDE = None


def _relu(x, tag):
    if RELU_OVERRIDE is not None and tag in RELU_OVERRIDE:
        return x * RELU_OVERRIDE[tag].to(x.dtype)
    return F.relu(x)


class _CausalBase(nn.Module):
    """Everything CausalGCN and CausalGAT share (model.py:47-83 / 342-378 for
    construction order; model.py:97-164 / 392-450 for the forward tail)."""

    def _build_tail(self, hidden, num_classes, GConv):
        self.edge_att_mlp = nn.Linear(hidden * 2, 2)
        self.node_att_mlp = nn.Linear(hidden, 2)
        self.bnc = BatchNorm1d(hidden)
        self.bno = BatchNorm1d(hidden)
        self.context_convs = GConv(hidden, hidden)
        self.objects_convs = GConv(hidden, hidden)
        self.fc1_bn_c = BatchNorm1d(hidden)
        self.fc1_c = Linear(hidden, hidden)
        self.fc2_bn_c = BatchNorm1d(hidden)
        self.fc2_c = Linear(hidden, num_classes)
        self.fc1_bn_o = BatchNorm1d(hidden)
        self.fc1_o = Linear(hidden, hidden)
        self.fc2_bn_o = BatchNorm1d(hidden)
        self.fc2_o = Linear(hidden, num_classes)
        if self.args.cat_or_add == "cat":
            self.fc1_bn_co = BatchNorm1d(hidden * 2)
            self.fc1_co = Linear(hidden * 2, hidden)
        elif self.args.cat_or_add == "add":
            self.fc1_bn_co = BatchNorm1d(hidden)
            self.fc1_co = Linear(hidden, hidden)
        else:
            assert False
        self.fc2_bn_co = BatchNorm1d(hidden)
        self.fc2_co = Linear(hidden, num_classes)
        for m in self.modules():                  # model.py:80-83
            if isinstance(m, BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0.0001)

    # model.py:85-122 / 380-409
    def forward(self, data, eval_random=True, perm=None, train_type="base"):
        x = data.x if getattr(data, "x", None) is not None else data.feat
        edge_index, batch = data.edge_index, data.batch
        row, col = edge_index
        x = self.backbone(x, edge_index)
        edge_rep = torch.cat([x[row], x[col]], dim=-1)
        if self.without_edge_attention:
            edge_att = 0.5 * torch.ones(edge_rep.shape[0], 2, dtype=x.dtype, device=x.device)
        else:
            edge_att = F.softmax(self.edge_att_mlp(edge_rep), dim=-1)
        edge_weight_c, edge_weight_o = edge_att[:, 0], edge_att[:, 1]
        if self.without_node_attention:
            node_att = 0.5 * torch.ones(x.shape[0], 2, dtype=x.dtype, device=x.device)
        else:
            node_att = F.softmax(self.node_att_mlp(x), dim=-1)
        xc = node_att[:, 0].view(-1, 1) * x
        xo = node_att[:, 1].view(-1, 1) * x
        xc = _relu(self.context_convs(self.bnc(xc), edge_index, edge_weight_c), "zc")
        xo = _relu(self.objects_convs(self.bno(xo), edge_index, edge_weight_o), "zo")
        num_graphs = getattr(data, "num_graphs", None)
        xc = global_add_pool(xc, batch, num_graphs)
        xo = global_add_pool(xo, batch, num_graphs)
        xc_logis = self._readout(xc, "c")
        xo_logis = self._readout(xo, "o", raw=train_type == "irm")     # CausalGIN, model.py:288-290: (x, x_logis)
        xco_logis = self.random_readout_layer(xc, xo, eval_random, perm)
        return xc_logis, xo_logis, xco_logis

    def backbone(self, x, edge_index):            # model.py:90-95 / 385-390
        x = self.bn_feat(x)
        x = _relu(self.conv_feat(x, edge_index), "x1")
        for i, conv in enumerate(self.convs):
            x = self.bns_conv[i](x)
            x = _relu(conv(x, edge_index), "x%d" % (i + 2))
        return x

    def _readout(self, x, tag, raw=False):        # model.py:125-143; raw: CausalGIN's train_type "irm", model.py:281-292
        x = getattr(self, "fc1_bn_" + tag)(x)
        x = _relu(getattr(self, "fc1_" + tag)(x), "h1_" + tag)
        x = getattr(self, "fc2_bn_" + tag)(x)
        x = getattr(self, "fc2_" + tag)(x)
        return (x, F.log_softmax(x, dim=-1)) if raw else F.log_softmax(x, dim=-1)

    def _shuffles(self, eval_random):
        raise NotImplementedError

    def random_readout_layer(self, xc, xo, eval_random, perm=None):   # model.py:145-164
        num = xc.shape[0]
        if perm is None:
            l = [i for i in range(num)]
            if self._shuffles(eval_random):
                random.shuffle(l)
            perm = torch.tensor(l)
        perm = torch.as_tensor(perm, dtype=torch.long)
        if self.args.cat_or_add == "cat":
            x = torch.cat((xc[perm], xo), dim=1)
        else:
            x = xc[perm] + xo
        return self._readout(x, "co")


class CausalGCN(_CausalBase):
    """model.py:12-164."""

    def __init__(self, num_features, num_classes, args, gfn=False, collapse=False,
                 residual=False, res_branch="BNConvReLU", global_pool="sum",
                 dropout=0, edge_norm=True):
        super().__init__()
        hidden = args.hidden
        self.args = args
        self.dropout = dropout
        self.with_random = args.with_random
        self.without_node_attention = args.without_node_attention
        self.without_edge_attention = args.without_edge_attention
        GConv = partial(GCNConv, edge_norm=edge_norm, gfn=gfn)
        self.num_classes = num_classes
        self.fc_num = args.fc_num
        self.bn_feat = BatchNorm1d(num_features)
        self.conv_feat = GCNConv(num_features, hidden, gfn=True)
        self.bns_conv = nn.ModuleList()
        self.convs = nn.ModuleList()
        for _ in range(args.layers):
            self.bns_conv.append(BatchNorm1d(hidden))
            self.convs.append(GConv(hidden, hidden))
        self._build_tail(hidden, num_classes, GConv)

    def _shuffles(self, eval_random):             # model.py:149-151
        return bool(self.with_random and eval_random)


class CausalGAT(_CausalBase):
    """model.py:315-450."""

    def __init__(self, num_features, num_classes, args, head=4, dropout=0.2):
        super().__init__()
        hidden = args.hidden
        self.args = args
        self.dropout = dropout
        self.without_node_attention = False       # CausalGAT has no ablation switches
        self.without_edge_attention = False
        GConv = partial(GCNConv, edge_norm=True, gfn=False)
        self.num_classes = num_classes
        self.fc_num = args.fc_num
        self.bn_feat = BatchNorm1d(num_features)
        self.conv_feat = GCNConv(num_features, hidden, gfn=True)
        self.bns_conv = nn.ModuleList()
        self.convs = nn.ModuleList()
        for _ in range(args.layers):
            self.bns_conv.append(BatchNorm1d(hidden))
            self.convs.append(GATConv(hidden, int(hidden / head), heads=head, dropout=dropout))
        self._build_tail(hidden, num_classes, GConv)

    def _shuffles(self, eval_random):             # model.py:435
        return bool(eval_random)


class GINConv(nn.Module):
    """torch_geometric.nn.GINConv, 1.x semantics (call site model.py:187-193):
    out = nn((1 + eps) * x + sum_{j in N(i)} x_j), eps = 0, self loops removed first."""

    def __init__(self, nn_module, eps=0.0):
        super().__init__()
        self.nn = nn_module
        self.register_buffer("eps", torch.tensor([float(eps)]))     # PyG 1.x keeps eps as a buffer (train_eps=False)

    def aggregate(self, x, edge_index):
        edge_index, _ = remove_self_loops(edge_index)
        return (1 + self.eps) * x + propagate_add(edge_index, x, None, x.size(0))

    def forward(self, x, edge_index):
        return self.nn(self.aggregate(x, edge_index))


class CausalGIN(_CausalBase):
    """model.py:166-313.  Backbone layer = GINConv(Linear -> BatchNorm1d -> ReLU -> Linear -> ReLU)
    (model.py:187-193); no BatchNorm between the layers (bns_conv stays empty, model.py:185,237-238);
    the permutation is drawn whenever eval_random is set (model.py:296-297); no ablation switches."""

    def __init__(self, num_features, num_classes, args, gfn=False, edge_norm=True):
        super().__init__()
        hidden = args.hidden
        self.args = args
        self.dropout = 0.0
        self.without_node_attention = False
        self.without_edge_attention = False
        GConv = partial(GCNConv, edge_norm=edge_norm, gfn=gfn)
        self.num_classes = num_classes
        self.fc_num = args.fc_num
        self.bn_feat = BatchNorm1d(num_features)
        self.conv_feat = GCNConv(num_features, hidden, gfn=True)
        self.bns_conv = nn.ModuleList()
        self.convs = nn.ModuleList()
        for _ in range(args.layers):
            self.convs.append(GINConv(nn.Sequential(Linear(hidden, hidden), BatchNorm1d(hidden), nn.ReLU(),
                                                    Linear(hidden, hidden), nn.ReLU())))
        self._build_tail(hidden, num_classes, GConv)

    def _shuffles(self, eval_random):             # model.py:296-297
        return bool(eval_random)

    def backbone(self, x, edge_index):
        """model.py:233-238 with the two ReLUs of every layer routed through the test hook
        (tags "r{i+1}" = after the inner BatchNorm, "x{i+2}" = layer output)."""
        x = self.bn_feat(x)
        x = _relu(self.conv_feat(x, edge_index), "x1")
        for i, conv in enumerate(self.convs):
            lin1, bn, _, lin2, _ = conv.nn
            h = lin1(conv.aggregate(x, edge_index))
            r = _relu(bn(h), "r%d" % (i + 1))
            x = _relu(lin2(r), "x%d" % (i + 2))
        return x


# --------------------------------------------------------------------------
# loss  (train_causal.py:176-183)
# --------------------------------------------------------------------------


def causal_loss(c_logs, o_logs, co_logs, y, num_classes, c=0.5, o=1.0, co=0.5):
    """train_causal.py:178-183: KL(uniform || exp(c_logs)) batchmean + 2 NLL."""
    y = y.view(-1)
    uniform_target = torch.ones_like(c_logs) / num_classes
    c_loss = F.kl_div(c_logs, uniform_target, reduction="batchmean")
    o_loss = F.nll_loss(o_logs, y)
    co_loss = F.nll_loss(co_logs, y)
    loss = c * c_loss + o * o_loss + co * co_loss
    return loss, c_loss, o_loss, co_loss


def train_step(model, data, perm=None, c=0.5, o=1.0, co=0.5, eval_random=True):
    """One iteration of train_causal_epoch minus the optimizer
    (train_causal.py:173-187).  Returns the loss parts; grads are left on the
    parameters."""
    model.zero_grad()
    outs = model(data, eval_random=eval_random, perm=perm)
    loss, c_loss, o_loss, co_loss = causal_loss(*outs, data.y, model.num_classes, c, o, co)
    correct_o = int(outs[1].max(1)[1].eq(data.y.view(-1)).sum())
    loss.backward()
    return outs, (loss, c_loss, o_loss, co_loss), correct_o
