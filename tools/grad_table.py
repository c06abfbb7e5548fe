"""GPU debugging aid: per-parameter gradient error table (vs fp32 oracle, vs fp64 oracle, and the
fp32 oracle's own error vs fp64) for seeded random cases."""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cal_b200
from oracle import cal_oracle as O
from tests.util import random_case, clone_to_cuda, rel_err, grad_or_zero
from tests.test_gpu_parity import _oracle_step, CASES

def run(case):
    ora, b, perm = random_case(**case)
    o32, l32, g32, _, _ = _oracle_step(ora, b, perm)
    o64, l64, g64, _, _ = _oracle_step(ora, b, perm, torch.float64)
    net = clone_to_cuda(ora, cal_b200)
    bd = b.to("cuda:0")
    outs = net(bd, eval_random=True, perm=perm.tolist())
    O.causal_loss(*outs, bd.y, net.num_classes)[0].backward()
    torch.cuda.synchronize()
    print("case", case, "N", b.batch.numel())
    for i in range(3):
        print("  out%d gpu-vs-32 %.2e gpu-vs-64 %.2e ref32-vs-64 %.2e" % (i, rel_err(outs[i].detach().cpu(), o32[i]), rel_err(outs[i].detach().cpu(), o64[i]), rel_err(o32[i], o64[i])))
    rows = []
    for n, p in net.named_parameters():
        g = grad_or_zero(p).cpu()
        rows.append((rel_err(g, g64[n]), rel_err(g, g32[n]), rel_err(g32[n], g64[n]), n, float(g64[n].abs().max())))
    rows.sort(reverse=True)
    for r in rows[:12]:
        print("  %-24s gpu-vs-64 %.2e gpu-vs-32 %.2e ref32-vs-64 %.2e  max|g| %.2e" % (r[3], r[0], r[1], r[2], r[4]))

for seed in [int(x) for x in sys.argv[1:]]:
    for c in CASES:
        if c["seed"] == seed:
            run(c)
    if seed == 91:
        run(dict(seed=91, hidden=128, layers=3, batch_size=512, features=64, avg_nodes=200, ba_m=2, noise=0.0))
