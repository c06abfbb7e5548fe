"""GPU debugging aid: for seeded random cases compare the backward intermediates held in the
workspace (DU, DAGG via dW, DYM, D) with the oracle's autograd values."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cal_b200
from oracle import cal_oracle as O
from tests.util import random_case, clone_to_cuda, rel_err, grad_or_zero, oracle_trace

def run(case):
    ora, b, perm = random_case(**case)
    tr, gr = oracle_trace(ora, b, perm)
    net = clone_to_cuda(ora, cal_b200)
    eng = net.engine
    bd = b.to("cuda:0")
    outs = net(bd, eval_random=True, perm=perm.tolist())
    O.causal_loss(*outs, bd.y, net.num_classes)[0].backward()
    torch.cuda.synchronize()
    N, B = b.batch.numel(), b.y.numel()
    H, L, Nm, Bm = eng.H, eng.L, eng.caps.max_nodes, eng.caps.max_graphs
    print("case", case, "N", N, "Nm", Nm, "Bm", Bm, "g_tile?", min((Nm + 31) // 32, 148))
    def show(name, got, want):
        got, want = got.detach().cpu().double(), want.detach().cpu().double()
        e = (got - want).abs()
        rows = (e.view(e.shape[0], -1).max(1)[0] > 1e-4 * want.abs().max()).nonzero().view(-1)
        print("  %-18s rel %.3e  bad rows %d %s" % (name, rel_err(got, want), rows.numel(), rows[:8].tolist()))
    DU = eng.region("DU").view(3, Bm, 2 * H)
    show("DU[0] (c head d gc)", DU[0, :B, :H], gr["gc"])
    DY = eng.region("DYM").view(2, Nm, H)
    show("DYM[0]", DY[0, :N], gr["yc"]); show("DYM[1]", DY[1, :N], gr["yo"])
    Z = eng.region("Z").view(2, Nm, H)
    show("Z[0]", Z[0, :N], tr["zc"]); show("Z[1]", Z[1, :N], tr["zo"])
    # expected dz / dagg of the two masked convs from the oracle's pooled gradients
    for k, (zn, gn, wn) in enumerate((("zc", "gc", "context_convs.weight"), ("zo", "go", "objects_convs.weight"))):
        dz = (tr[zn] > 0).float() * gr[gn][b.batch]
        W = dict(ora.named_parameters())[wn].detach()
        show("DAGG[%d]" % k, eng.region("DAGG").view(2, Nm, H)[k, :N], dz @ W.t())
    D = eng.region("D").view(2, Nm, H)
    show("D[1] (dx4/dy2)", D[1, :N], gr["y2"])
    show("D[0] (dy1)", D[0, :N], gr["y1"])
    for n in ("context_convs.weight", "context_convs.bias", "objects_convs.weight", "objects_convs.bias", "bnc.weight", "bno.weight"):
        p = dict(net.named_parameters())[n]; q = dict(ora.named_parameters())[n]
        print("  grad %-22s rel %.3e" % (n, rel_err(p.grad.cpu(), q.grad)))

cases = {
  "a": dict(seed=12, hidden=128, layers=3, batch_size=256, avg_nodes=30),
  "b": dict(seed=13, hidden=128, layers=3, batch_size=256, avg_nodes=14),
  "c": dict(seed=14, hidden=128, layers=3, batch_size=128, avg_nodes=60),
  "d": dict(seed=15, hidden=32, layers=3, batch_size=128, avg_nodes=60),
}
for k in sys.argv[1:]:
    run(cases[k])
