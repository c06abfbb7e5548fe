"""Shared helpers for the tests: golden-fixture loading and oracle construction."""
import argparse
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def make_args(**kw):
    d = dict(layers=3, hidden=32, with_random=True, without_node_attention=False,
             without_edge_attention=False, fc_num="222", cat_or_add="add",
             c=0.5, o=1.0, co=0.5, eval_random=False)
    d.update(kw)
    return argparse.Namespace(**d)


class GoldenCase:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        kind, workload, train, cat, layers, hidden, C, dropout = [str(s) for s in z["meta"]]
        self.kind, self.workload, self.train = kind, workload, train == "1"
        self.args = make_args(cat_or_add=cat, layers=int(layers), hidden=int(hidden))
        self.num_classes, self.dropout = int(C), float(dropout)
        self.params = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
        self.grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
        self.after = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("after/")}
        self.perm = torch.from_numpy(z["perm"])
        self.outs = [torch.from_numpy(z[k]) for k in ("c_logs", "o_logs", "co_logs")]
        self.loss = z["loss"] if "loss" in z.files else None

    def batch(self):
        from cal_b200.data import Batch
        z = self.z
        b = Batch(feat=torch.from_numpy(z["feat"]), edge_index=torch.from_numpy(z["edge_index"]),
                  y=torch.from_numpy(z["y"]))
        b.batch = torch.from_numpy(z["batch"])
        b.num_graphs = int(z["y"].shape[0])
        return b

    def build(self, module, dtype=torch.float32):
        """Instantiate module.CausalGCN / CausalGAT and load the golden parameters."""
        F_in = self.z["feat"].shape[1]
        if self.kind == "CausalGCN":
            net = module.CausalGCN(F_in, self.num_classes, self.args)
        elif self.kind == "CausalGIN":
            net = module.CausalGIN(F_in, self.num_classes, self.args)
        else:
            net = module.CausalGAT(F_in, self.num_classes, self.args, dropout=self.dropout)
        missing = net.load_state_dict(self.params, strict=True)
        net = net.to(dtype)
        net.train(self.train)
        return net


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def oracle_trace(net, data, perm, loss_weights=(0.5, 1.0, 0.5)):
    """Run the oracle's forward (+ loss + backward when net.training) step by step, keeping every
    intermediate the CUDA workspace also holds.  Returns (dict of tensors, dict of their grads)."""
    import torch.nn.functional as F
    from oracle import cal_oracle as O
    t, keep = {}, {}

    def rec(name, v):
        t[name] = v
        if v.requires_grad:
            v.retain_grad()
            keep[name] = v
        return v

    x = data.x if getattr(data, "x", None) is not None else data.feat
    ei, batch = data.edge_index, data.batch
    row, col = ei
    x = net.bn_feat(x)
    x = rec("x1", F.relu(net.conv_feat(x, ei)))
    for i, conv in enumerate(net.convs):
        y = rec("y%d" % (i + 1), net.bns_conv[i](x))
        x = rec("x%d" % (i + 2), F.relu(conv(y, ei)))
    edge_rep = torch.cat([x[row], x[col]], dim=-1)
    if net.without_edge_attention:
        edge_att = 0.5 * torch.ones(edge_rep.shape[0], 2)
    else:
        edge_att = F.softmax(net.edge_att_mlp(edge_rep), dim=-1)
    rec("edge_att", edge_att)
    if net.without_node_attention:
        node_att = 0.5 * torch.ones(x.shape[0], 2)
    else:
        node_att = F.softmax(net.node_att_mlp(x), dim=-1)
    rec("node_att", node_att)
    xc = node_att[:, 0].view(-1, 1) * x
    xo = node_att[:, 1].view(-1, 1) * x
    yc = rec("yc", net.bnc(xc))
    yo = rec("yo", net.bno(xo))
    zc = rec("zc", F.relu(net.context_convs(yc, ei, edge_att[:, 0])))
    zo = rec("zo", F.relu(net.objects_convs(yo, ei, edge_att[:, 1])))
    B = int(data.y.numel())
    gc = rec("gc", O.global_add_pool(zc, batch, B))
    go = rec("go", O.global_add_pool(zo, batch, B))
    outs = [net._readout(gc, "c"), net._readout(go, "o"), net.random_readout_layer(gc, go, True, perm)]
    for n, o in zip(("c", "o", "co"), outs):
        rec("logp_" + n, o)
    if net.training:
        loss, *_ = O.causal_loss(*outs, data.y, net.num_classes, *loss_weights)
        t["loss"] = loss.detach()
        net.zero_grad()
        loss.backward()
    grads = {k: v.grad for k, v in keep.items() if v.grad is not None}
    return {k: v.detach() for k, v in t.items()}, grads


def ref_prep(edge_index, batch, num_graphs):
    """numpy restatement of the structure the reference builds inside every GCNConv.norm call
    (gcn_conv.py:56-57: remove_self_loops, then add_self_loops appends [i, i] LAST) arranged as the
    two CSR orderings cal_prep emits; within a row entries follow edge_index column order, which is
    the order the CPU scatter_add accumulates in."""
    ei = np.asarray(edge_index)
    N = int(np.asarray(batch).shape[0])
    E = ei.shape[1]
    keep = ei[0] != ei[1]
    ids = np.nonzero(keep)[0]
    rows = np.concatenate([ei[0][ids], np.arange(N)])
    cols = np.concatenate([ei[1][ids], np.arange(N)])
    keys = np.concatenate([ids, E + np.arange(N)])
    o_in = np.lexsort((keys, cols))
    o_out = np.lexsort((keys, rows))
    in_ptr = np.zeros(N + 1, np.int32)
    np.add.at(in_ptr, cols + 1, 1)
    in_ptr = np.cumsum(in_ptr).astype(np.int32)
    out_ptr = np.zeros(N + 1, np.int32)
    np.add.at(out_ptr, rows + 1, 1)
    out_ptr = np.cumsum(out_ptr).astype(np.int32)
    pos_of_key = {int(k): p for p, k in enumerate(keys[o_in])}
    out_pos = np.array([pos_of_key[int(k)] for k in keys[o_out]], np.int32)
    b = np.asarray(batch)
    graph_ptr = np.searchsorted(b, np.arange(num_graphs + 1), side="left").astype(np.int32)
    deg = np.bincount(rows, minlength=N).astype(np.float32)
    # gcn_conv.py:67 deg.pow(-0.5): torch evaluates the exponent -0.5 as 1 / sqrt(x), correctly rounded twice (checked
    # for every degree below 2000); numpy's float32 power differs from it in the last bit from degree 17 on
    dis = (np.float32(1.0) / np.sqrt(deg)).astype(np.float32)
    return dict(in_ptr=in_ptr, in_src=rows[o_in].astype(np.int32), in_key=keys[o_in].astype(np.int32),
                out_ptr=out_ptr, out_dst=cols[o_out].astype(np.int32), out_key=keys[o_out].astype(np.int32),
                out_pos=out_pos, graph_ptr=graph_ptr, dis=dis,
                in_norm=(dis[rows[o_in]] * dis[cols[o_in]]).astype(np.float32))


# ----------------------------------------------------------------------------------------------
# seeded random cases (oracle vs CUDA on identical inputs)
# ----------------------------------------------------------------------------------------------

def random_case(seed=0, kind="CausalGCN", hidden=32, features=10, classes=4, layers=3, cat="add",
                batch_size=12, avg_nodes=25, ba_m=1, noise=0.1, dropout=0.0, gaussian_feat=None, **arg_over):
    """(oracle model, CPU batch, perm): a seeded model with every 1-d parameter moved off its init
    (so all gradient paths are exercised) and one SPMotif-style batch."""
    from cal_b200.data import make_batches
    from oracle import cal_oracle
    args = make_args(cat_or_add=cat, layers=layers, hidden=hidden, **arg_over)
    fd = features if (gaussian_feat if gaussian_feat is not None else features != 10) else None
    b = make_batches("spmotif", num_batches=1, seed=seed, batch_size=batch_size, avg_nodes=avg_nodes,
                     ba_m=ba_m, noise=noise, feature_dim=fd, num_classes=classes)[0]
    g = torch.Generator().manual_seed(seed + 1)
    if fd is None:
        b.feat = b.feat + 0.25 * torch.randn(b.feat.shape, generator=g)
    torch.manual_seed(seed + 2)
    if kind == "CausalGCN":
        net = cal_oracle.CausalGCN(b.feat.size(1), classes, args)
    elif kind == "CausalGIN":
        net = cal_oracle.CausalGIN(b.feat.size(1), classes, args)
    else:
        net = cal_oracle.CausalGAT(b.feat.size(1), classes, args, dropout=dropout)
    with torch.no_grad():
        for _, p in net.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    perm = torch.randperm(batch_size, generator=g)
    return net, b, perm


def clone_to_cuda(oracle_net, module, device="cuda:0"):
    """A cal_b200 module with the oracle's constructor arguments and state_dict."""
    from oracle import cal_oracle
    F_in = oracle_net.bn_feat.num_features
    if isinstance(oracle_net, cal_oracle.CausalGCN):
        net = module.CausalGCN(F_in, oracle_net.num_classes, oracle_net.args)
    elif isinstance(oracle_net, cal_oracle.CausalGIN):
        net = module.CausalGIN(F_in, oracle_net.num_classes, oracle_net.args)
    else:
        net = module.CausalGAT(F_in, oracle_net.num_classes, oracle_net.args, dropout=oracle_net.dropout)
    net.load_state_dict(oracle_net.state_dict(), strict=True)
    net.train(oracle_net.training)
    return net.to(device)


def grad_or_zero(p):
    return p.grad if p.grad is not None else torch.zeros_like(p)
