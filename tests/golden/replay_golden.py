"""Replay the committed golden fixtures through the UNMODIFIED reference model code.

Run in the build container only (needs /root/reference):

    python tests/golden/replay_golden.py [case ...]

For every ``tests/golden/<case>.npz`` the reference's own ``model.py`` class (imported over the stand-in PyG of
``oracle/pyg_shim``, exactly as ``make_golden.py`` does) is constructed, loaded with the STORED parameters and
buffers, fed the STORED batch and permutation, and its outputs, loss parts, parameter gradients and BatchNorm
running statistics are compared with the stored ones.  This keeps the fixtures honest independently of
``make_golden.py``'s input generators (which may change): whatever inputs a fixture holds, its outputs are what
the reference computes for them.  ``tests/test_oracle_golden.py`` runs it where the reference exists.
"""
import argparse
import glob
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

import model as ref_model  # noqa: E402  (the reference's model.py)

TOL = 2e-6


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) if a.size else 0.0


def replay(path):
    z = np.load(path)
    kind, workload, train, cat, layers, hidden, C, dropout = [str(s) for s in z["meta"]]
    train = train == "1"
    args = argparse.Namespace(layers=int(layers), hidden=int(hidden), with_random=True, without_node_attention=False,
                              without_edge_attention=False, fc_num="222", cat_or_add=cat, c=0.5, o=1.0, co=0.5)
    ctor = dict(dropout=float(dropout)) if kind == "CausalGAT" else {}
    net = getattr(ref_model, kind)(int(z["feat"].shape[1]), int(C), args, **ctor)
    state = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    net.load_state_dict(state, strict=True)
    net.train(train)

    class D:                                   # the duck-typed batch the reference forward reads (model.py:87-89)
        x = None
    D.feat, D.edge_index = torch.from_numpy(z["feat"]), torch.from_numpy(z["edge_index"])
    D.batch, D.y = torch.from_numpy(z["batch"]), torch.from_numpy(z["y"])
    perm = z["perm"].tolist()
    worst = 0.0
    real_shuffle = random.shuffle

    def stored_shuffle(l):                     # the draw of model.py:151 / 297 / 436, replaced by the stored one
        assert len(l) == len(perm)
        l[:] = perm
    random.shuffle = stored_shuffle
    try:
        if train:
            outs = net(D, eval_random=True)
            y = D.y.view(-1)
            uniform = torch.ones_like(outs[0], dtype=torch.float) / net.num_classes
            c_loss = F.kl_div(outs[0], uniform, reduction="batchmean")
            o_loss, co_loss = F.nll_loss(outs[1], y), F.nll_loss(outs[2], y)
            loss = args.c * c_loss + args.o * o_loss + args.co * co_loss
            loss.backward()
            worst = max(worst, rel([loss.item(), c_loss.item(), o_loss.item(), co_loss.item()], z["loss"]))
            for n_, p in net.named_parameters():
                assert bool(z["hasgrad/" + n_]) == (p.grad is not None), n_
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                worst = max(worst, float(np.abs(g.numpy() - z["grad/" + n_]).max()) /
                            max(max(float(np.abs(z[k]).max()) for k in z.files if k.startswith("grad/")), 1e-30))
        else:
            with torch.no_grad():
                outs = net(D, eval_random=False)
    finally:
        random.shuffle = real_shuffle
    for t, k in zip(outs, ("c_logs", "o_logs", "co_logs")):
        worst = max(worst, rel(t.detach().numpy(), z[k]))
    for k, v in net.state_dict().items():
        if "running" in k:
            worst = max(worst, rel(v.numpy(), z["after/" + k]))
        if "num_batches" in k:
            assert int(v) == int(z["after/" + k]), k
    return worst


if __name__ == "__main__":
    names = sys.argv[1:] or sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "*.npz")))
    bad = 0
    for n in names:
        w = replay(os.path.join(HERE, n + ".npz"))
        print("%-14s worst relative deviation from the stored vectors %.2e %s" % (n, w, "ok" if w < TOL else "MISMATCH"))
        bad += w >= TOL
    sys.exit(1 if bad else 0)
