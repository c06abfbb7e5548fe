"""GPU experiment: 3-step Adam trajectories (Trainer, eager) of the fused small-graph path vs the tiled kernels vs the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cal_b200 as M
from oracle import cal_oracle as O
from tests.util import clone_to_cuda, random_case, rel_err

EPS = float(os.environ.get("ADAM_EPS", "1e-8"))
ora, b0, _ = random_case(seed=299, hidden=128, batch_size=24)
SEED0 = int(os.environ.get("SEED0", "300"))
batches = [random_case(seed=SEED0 + i, hidden=128, batch_size=24)[1] for i in range(2)]
res = {}
for mode in ("off", "auto"):
    net = clone_to_cuda(ora, M)
    net.engine.fsg_mode = mode
    tr = M.Trainer(net, M.batch_caps(batches), lr=1e-3, eps=EPS, use_graph=False)
    print(mode, "fused:", tr.fused_small_graphs)
    grads = []
    for s in range(3):
        b = batches[s % 2]
        tr.step_host(tr.pack(b, perm=list(range(b.num_graphs))))
        torch.cuda.synchronize()
        grads.append(net.engine.flat_grad.clone().cpu())
    res[mode] = ({n: p.detach().cpu().clone() for n, p in net.named_parameters()}, grads, net.engine)
import copy
o = copy.deepcopy(ora)
opt = torch.optim.Adam(o.parameters(), lr=1e-3, eps=EPS)
init = {n: p.detach().clone() for n, p in ora.named_parameters()}
og = []
for s in range(3):
    b = batches[s % 2]
    O.train_step(o, b, perm=torch.arange(b.num_graphs))
    og.append({n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in o.named_parameters()})
    opt.step()
eng = res["auto"][2]
for n, p in o.named_parameters():
    a, f = res["off"][0][n], res["auto"][0][n]
    line = "%-28s param err tiled %.2e fused %.2e | update err tiled %.2e fused %.2e |" % (
        n, rel_err(a, p.detach()), rel_err(f, p.detach()), rel_err(a - init[n], p.detach() - init[n]), rel_err(f - init[n], p.detach() - init[n]))
    off = eng.param_offs[n]
    for s in range(3):
        gt = res["off"][1][s][off:off + p.numel()].view(p.shape)
        gf = res["auto"][1][s][off:off + p.numel()].view(p.shape)
        line += " g%d tiled %.1e fused %.1e (|g|max %.1e)" % (s, rel_err(gt, og[s][n]), rel_err(gf, og[s][n]), float(og[s][n].abs().max()))
    print(line)
