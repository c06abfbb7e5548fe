"""Freeze golden vectors from the UNMODIFIED reference model code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own ``model.py`` / ``gcn_conv.py`` on top of the
third-party stand-ins in ``oracle/pyg_shim`` (torch_geometric / torch_scatter
are not installable here), runs forward + loss (train_causal.py:176-183) +
backward on small seeded batches and stores inputs, parameters, outputs, loss
parts, every parameter gradient and the BatchNorm running statistics after the
step as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` then checks
``oracle/cal_oracle.py`` against these files; the GPU tests check the CUDA path
against them too.
"""
import argparse
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CAL_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import model as ref_model  # noqa: E402  (the reference's model.py)

from cal_b200.data import make_batches  # noqa: E402


def ns(**kw):
    d = dict(layers=3, hidden=32, with_random=True, without_node_attention=False,
             without_edge_attention=False, fc_num="222", cat_or_add="add",
             c=0.5, o=1.0, co=0.5)
    d.update(kw)
    return argparse.Namespace(**d)


CASES = {
    # name: (model, workload, batch_size, args overrides, train?, ctor kwargs)
    "gcn_add_h32": ("CausalGCN", "spmotif", 12, dict(), True, {}),
    "gcn_cat_h32": ("CausalGCN", "spmotif", 9, dict(cat_or_add="cat", layers=2), True, {}),
    "gcn_add_h128": ("CausalGCN", "spmotif", 16, dict(hidden=128), True, {}),
    "gcn_eval_h32": ("CausalGCN", "spmotif", 10, dict(), False, {}),
    "gat_add_h32": ("CausalGAT", "mutag", 10, dict(layers=2), True, dict(dropout=0.0)),
    "gat_eval_h32": ("CausalGAT", "mutag", 7, dict(layers=2), False, dict(dropout=0.2)),
    "gin_add_h32": ("CausalGIN", "spmotif", 11, dict(layers=2), True, {}),
    "gin_cat_h64": ("CausalGIN", "spmotif", 8, dict(layers=3, hidden=64, cat_or_add="cat"), True, {}),
    "gin_eval_h32": ("CausalGIN", "spmotif", 9, dict(layers=2), False, {}),
}


def add_oddities(batch, seed):
    """Inject what the reference tolerates: a self loop, a duplicate edge, an
    isolated node (only its appended self loop), an asymmetric (one-way) edge."""
    g = torch.Generator().manual_seed(seed)
    ei = batch.edge_index
    n = batch.batch.numel()
    first = int((batch.batch == 0).sum())
    extra = torch.tensor([[1, 0, 2, 0], [1, 2, 2, 3]])       # loops + dup-ish + one-way inside graph 0
    extra = extra[:, (extra < first).all(0)]
    k = int(torch.randint(0, ei.size(1), (1,), generator=g))
    ei = torch.cat([ei[:, :k], extra, ei[:, k:], ei[:, :2]], dim=1)   # also duplicates two edges
    # isolate the last node of the last graph
    keep = (ei[0] != n - 1) & (ei[1] != n - 1)
    batch.edge_index = ei[:, keep].contiguous()
    return batch


def run_case(name):
    kind, workload, bs, over, train, ctor = CASES[name]
    args = ns(**over)
    batch = make_batches(workload, num_batches=1, seed=1234 + len(name), batch_size=bs)[0]
    batch = add_oddities(batch, 7)
    feat = batch.feat
    # non-degenerate features for BN (one-hot columns can be all-zero in a tiny batch)
    g = torch.Generator().manual_seed(99)
    feat = feat + 0.25 * torch.randn(feat.shape, generator=g)
    batch.feat = feat
    F_in = feat.size(1)
    C = 4 if workload != "mutag" else 2
    torch.manual_seed(666)
    net = getattr(ref_model, kind)(F_in, C, args, **ctor)
    # move BN / bias params off their init so every gradient path is exercised
    with torch.no_grad():
        g = torch.Generator().manual_seed(5)
        for n_, p in net.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    net.train(train)
    state0 = {k: v.detach().clone() for k, v in net.state_dict().items()}

    class D:  # the duck-typed batch the reference forward reads (model.py:87-89)
        x = None
    D.feat, D.edge_index, D.batch, D.y = batch.feat, batch.edge_index, batch.batch, batch.y

    random.seed(42)
    perm = list(range(bs))
    if train:
        st = random.getstate()
        random.shuffle(perm)             # the draw model.py:151 / 436 will make
        random.setstate(st)
    out = {}
    if train:
        c_logs, o_logs, co_logs = net(D, eval_random=True)
        y = batch.y.view(-1)
        uniform = torch.ones_like(c_logs, dtype=torch.float) / net.num_classes
        c_loss = F.kl_div(c_logs, uniform, reduction="batchmean")
        o_loss = F.nll_loss(o_logs, y)
        co_loss = F.nll_loss(co_logs, y)
        loss = args.c * c_loss + args.o * o_loss + args.co * co_loss
        loss.backward()
        out["loss"] = np.array([loss.item(), c_loss.item(), o_loss.item(), co_loss.item()], np.float64)
        for n_, p in net.named_parameters():
            out["grad/" + n_] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
            out["hasgrad/" + n_] = np.array(p.grad is not None)
    else:
        with torch.no_grad():
            c_logs, o_logs, co_logs = net(D, eval_random=False)
    out["c_logs"], out["o_logs"], out["co_logs"] = (t.detach().numpy() for t in (c_logs, o_logs, co_logs))
    for k, v in state0.items():
        out["param/" + k] = v.numpy()
    for k, v in net.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["after/" + k] = v.numpy()
    out["feat"], out["edge_index"] = batch.feat.numpy(), batch.edge_index.numpy()
    out["batch"], out["y"] = batch.batch.numpy(), batch.y.numpy()
    out["perm"] = np.array(perm, np.int64)
    out["meta"] = np.array([kind, workload, str(int(train)), args.cat_or_add, str(args.layers),
                            str(args.hidden), str(C), str(ctor.get("dropout", 0.2))])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "N=%d E=%d B=%d" % (batch.batch.numel(), batch.edge_index.size(1), bs),
          "loss=%s" % (out.get("loss"),))


if __name__ == "__main__":
    for n in (sys.argv[1:] or CASES):
        run_case(n)
