"""Per-parameter gradient error table of the CUDA path (run on the GPU box):

    python tools/grad_table.py > profiles/grad_errors_rNN.txt

For every golden case (free-running reference gradients frozen from the unmodified reference model code)
and every seeded case of tests/test_gpu_parity.py it prints, per parameter tensor, the tensor's largest
gradient entry relative to the model's largest (|g|/gmax) and the max-norm errors -- divided by
the tensor's scale (tests/test_gpu_parity.py::grad_errors: its own largest entry, with the named floors), exactly
the rule the tests assert --
of the GPU gradient against the fp32 reference (e32), against the fp64 oracle (e64), and of the fp32
reference against the fp64 oracle (ref).  A line is marked PASS when e32 < 1e-5 or e64 < 1e-5."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cal_b200  # noqa: E402
from oracle import cal_oracle as O  # noqa: E402
from tests import test_gpu_parity as T  # noqa: E402
from tests.util import GoldenCase, clone_to_cuda, golden_names, grad_or_zero, random_case, rel_err  # noqa: E402

DEV = "cuda:0"


def table(title, gpu, g32, g64, outs_err):
    rows = T.grad_errors(gpu, g32, g64)
    bad = [r for r in rows if not (r[2] < T.TOL or r[3] < T.TOL)]
    print("== %s: outputs rel err %s; %d tensors, worst min(e32, e64) %.2e, %d above 1e-5"
          % (title, " ".join("%.1e" % e for e in outs_err), len(rows), max(min(r[2], r[3]) for r in rows), len(bad)))
    for n, rel, e32, e64, ref in sorted(rows, key=lambda r: -min(r[2], r[3])):
        print("   %-26s |g|/gmax %.1e  e32 %.2e  e64 %.2e  ref %.2e  %s"
              % (n, rel, e32, e64, ref, "PASS" if (e32 < T.TOL or e64 < T.TOL) else
                 ("illcond" if e64 <= 3 * ref else "FAIL")))


def golden(name):
    gc = GoldenCase(name)
    if not gc.train:
        return
    net = gc.build(cal_b200).to(DEV)
    b = gc.batch().to(DEV)
    outs = net(b, eval_random=True, perm=gc.perm.tolist())
    O.causal_loss(*outs, b.y, net.num_classes)[0].backward()
    torch.cuda.synchronize()
    _, _, g64, _, _ = T._oracle_step(gc.build(O), gc.batch(), gc.perm, torch.float64)
    gpu = {n: grad_or_zero(p) for n, p in net.named_parameters()}
    table("golden %s (B=%d, free-running)" % (name, b.y.numel()), gpu, gc.grads, g64,
          [rel_err(o.detach().cpu(), w) for o, w in zip(outs, gc.outs)])


def seeded(case, free=False):
    case = dict(case)
    ora, b, perm = random_case(**case)
    net = clone_to_cuda(ora, cal_b200)
    if case.get("kind") == "CausalGAT" and case.get("dropout", 0.0) > 0:
        net.dropout_mask = T._gat_masks(ora, b, case["seed"], case["dropout"])
    bd = b.to(DEV)
    outs = net(bd, eval_random=True, perm=perm.tolist())
    O.causal_loss(*outs, bd.y, net.num_classes)[0].backward()
    torch.cuda.synchronize()
    eng = net.engine
    masks = None if free else T._gpu_relu_masks(eng, b.batch.numel(), b.y.numel())
    o32, _, g32, _, _ = T._oracle_step(ora, b, perm, torch.float32, masks)
    _, _, g64, _, _ = T._oracle_step(ora, b, perm, torch.float64, masks)
    gpu = {n: grad_or_zero(p) for n, p in net.named_parameters()}
    table("seed %d %s (N=%d, B=%d, %s)" % (case["seed"], {k: v for k, v in case.items() if k != "seed"},
                                           b.batch.numel(), b.y.numel(), "free-running" if free else "GPU ReLU pattern"),
          gpu, g32, g64, [rel_err(o.detach().cpu(), w) for o, w in zip(outs, o32)])


def main():
    for name in golden_names():
        golden(name)
    for case in T.CASES + T.GIN_CASES + T.GAT_CASES:
        seeded(case)
    seeded(dict(seed=81, hidden=128, layers=3, batch_size=128))
    seeded(dict(seed=81, hidden=128, layers=3, batch_size=128), free=True)
    seeded(dict(seed=3, hidden=128, layers=3, batch_size=128), free=True)
    if "--large" in sys.argv:
        seeded(dict(seed=91, hidden=128, layers=3, batch_size=512, features=64, avg_nodes=200, ba_m=2, noise=0.0))


if __name__ == "__main__":
    main()
