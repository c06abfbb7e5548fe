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
        else:
            net = module.CausalGAT(F_in, self.num_classes, self.args, dropout=self.dropout)
        missing = net.load_state_dict(self.params, strict=True)
        net = net.to(dtype)
        net.train(self.train)
        return net


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
