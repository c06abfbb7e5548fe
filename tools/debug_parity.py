"""GPU debugging aid: run one golden case through the CUDA path and print, stage by stage, the
error of every workspace region against the oracle's intermediates (forward and backward)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cal_b200  # noqa: E402
from oracle import cal_oracle  # noqa: E402
from tests.util import GoldenCase, oracle_trace, ref_prep, rel_err  # noqa: E402


def show(name, got, want):
    got, want = got.detach().cpu().double(), want.detach().cpu().double()
    if got.shape != want.shape:
        print("%-28s SHAPE %s vs %s" % (name, tuple(got.shape), tuple(want.shape)))
        return
    bad = "" if torch.isfinite(got).all() else "  NONFINITE"
    print("%-28s rel %.3e  (max|ref| %.3e)%s" % (name, rel_err(got, want), float(want.abs().max()), bad))


def main(name):
    gc = GoldenCase(name)
    ora = gc.build(cal_oracle)
    b = gc.batch()
    tr, gr = oracle_trace(ora, b, gc.perm)
    dev = torch.device("cuda:0")
    net = gc.build(cal_b200).to(dev)
    eng = net.engine
    bd = b.to(dev)
    outs = net(bd, eval_random=True, perm=gc.perm.tolist())
    torch.cuda.synchronize()
    print("status", eng.status())
    N, E, B = b.batch.numel(), b.edge_index.size(1), b.y.numel()
    H, L, Nm, Bm = eng.H, eng.L, eng.caps.max_nodes, eng.caps.max_graphs
    rp = ref_prep(b.edge_index.numpy(), b.batch.numpy(), B)
    EPn = int(rp["in_ptr"][-1])
    for k, n in (("IN_PTR", N + 1), ("IN_SRC", EPn), ("IN_KEY", EPn), ("OUT_PTR", N + 1), ("OUT_DST", EPn),
                 ("OUT_KEY", EPn), ("OUT_POS", EPn), ("GRAPH_PTR", B + 1)):
        got = eng.region(k, torch.int32)[:n].cpu().numpy()
        print("%-28s %s" % (k, "exact" if np.array_equal(got, rp[k.lower()]) else "MISMATCH"))
    show("DIS", eng.region("DIS")[:N], torch.from_numpy(rp["dis"]))
    show("IN_NORM", eng.region("IN_NORM")[:EPn], torch.from_numpy(rp["in_norm"]))
    X = eng.region("X").view(L + 1, Nm, H)
    for l in range(L + 1):
        show("X[%d] (x%d)" % (l, l + 1), X[l, :N], tr["x%d" % (l + 1)])
    show("NODE_ATT", eng.region("NODE_ATT").view(Nm, 2)[:N], tr["node_att"])
    key = torch.from_numpy(rp["in_key"]).long()
    watt = eng.region("EDGE_ATT").view(-1, 2)[:EPn].cpu()
    m = key < E
    show("EDGE_ATT", watt[m], tr["edge_att"][key[m]])
    Z = eng.region("Z").view(2, Nm, H)
    show("Z[0] (zc)", Z[0, :N], tr["zc"])
    show("Z[1] (zo)", Z[1, :N], tr["zo"])
    P = eng.region("POOLED").view(2, Bm, H)
    show("POOLED[0]", P[0, :B], tr["gc"])
    show("POOLED[1]", P[1, :B], tr["go"])
    for i, n in enumerate(("c", "o", "co")):
        show("logp_" + n, outs[i], tr["logp_" + n])
    if not gc.train:
        return
    loss, *_ = cal_oracle.causal_loss(*outs, bd.y, net.num_classes)
    print("loss gpu %.7f  oracle %.7f" % (float(loss), float(tr["loss"])))
    loss.backward()
    torch.cuda.synchronize()
    DU = eng.region("DU").view(3, Bm, 2 * H)           # d(readout inputs): head 0 = d gc (c head only), head 1 = d go
    show("DU[0] (d gc, c head)", DU[0, :B, :H], gr["gc"])
    DY = eng.region("DYM").view(2, Nm, H)
    show("DYM[0] (d yc)", DY[0, :N], gr["yc"])
    show("DYM[1] (d yo)", DY[1, :N], gr["yo"])
    dnrm = eng.region("DNRM").view(-1, 2)[:EPn].cpu()
    dt = eng.region("DT").view(-1, 2)[:EPn].cpu()
    print("dnrm finite", bool(torch.isfinite(dnrm).all()), " dt finite", bool(torch.isfinite(dt).all()))
    D = eng.region("D").view(2, Nm, H)
    show("D[0] (d y1)", D[0, :N], gr["y1"])
    if L > 1:
        show("D[1] (d y2)", D[1, :N], gr["y2"])
    for n, p in net.named_parameters():
        want = gc.grads[n]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        show("grad " + n, got, want)
    sd = net.state_dict()
    for k, want in gc.after.items():
        show("after " + k, sd[k].float(), want.float())


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["gcn_add_h32"]):
        print("=====", nm)
        main(nm)
