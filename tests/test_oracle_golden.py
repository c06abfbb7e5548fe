"""The oracle (oracle/cal_oracle.py) against the golden vectors frozen from the
unmodified reference model.py / gcn_conv.py (tests/golden/make_golden.py)."""
import os

import pytest
import torch

from oracle import cal_oracle
from tests.util import GoldenCase, golden_names, rel_err

TOL = 2e-6   # same fp32 op sequence on the same CPU -> essentially bit-identical


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    gc = GoldenCase(name)
    net = gc.build(cal_oracle)
    b = gc.batch()
    if gc.train:
        outs, losses, _ = cal_oracle.train_step(net, b, perm=gc.perm)
        for got, want in zip(outs, gc.outs):
            assert rel_err(got.detach(), want) < TOL
        for got, want in zip(losses, gc.loss):
            assert abs(float(got.detach()) - want) < TOL * max(1.0, abs(want))     # (threaded CPU reductions: order varies with load)
        for n, p in net.named_parameters():
            want = gc.grads[n]
            got = p.grad if p.grad is not None else torch.zeros_like(p)
            assert bool(gc.z["hasgrad/" + n]) == (p.grad is not None), n
            assert rel_err(got, want) < 2e-5, n
        sd = net.state_dict()
        for k, want in gc.after.items():
            assert rel_err(sd[k].double(), want.double()) < TOL, k
    else:
        with torch.no_grad():
            outs = net(b, eval_random=False)
        for got, want in zip(outs, gc.outs):
            assert rel_err(got, want) < TOL


def test_state_dict_keys_match_reference():
    gc = GoldenCase("gcn_add_h32")
    net = gc.build(cal_oracle)
    assert list(net.state_dict().keys()) == list(gc.params.keys())
    gc = GoldenCase("gat_add_h32")
    net = gc.build(cal_oracle)
    assert list(net.state_dict().keys()) == list(gc.params.keys())


def test_init_matches_reference_rng_order():
    """Constructing the oracle under the same torch seed reproduces the reference init
    (make_golden.py seeds 666, builds, then perturbs 1-d params; 2-d weights are untouched)."""
    for name in ("gcn_add_h32", "gat_add_h32"):
        gc = GoldenCase(name)
        torch.manual_seed(666)
        F_in = gc.z["feat"].shape[1]
        if gc.kind == "CausalGCN":
            net = cal_oracle.CausalGCN(F_in, gc.num_classes, gc.args)
        else:
            net = cal_oracle.CausalGAT(F_in, gc.num_classes, gc.args, dropout=gc.dropout)
        for n, p in net.named_parameters():
            if p.dim() >= 2:
                assert torch.equal(p.detach(), gc.params[n]), n


@pytest.mark.skipif(not os.path.isfile("/root/reference/model.py"), reason="needs /root/reference (build container only)")
def test_committed_fixtures_replay_through_the_unmodified_reference():
    """Every committed fixture, fed back through the reference's own model.py (stored parameters, stored batch,
    stored permutation; tests/golden/replay_golden.py): outputs, loss parts, every parameter gradient and the
    running statistics are what the file holds.  The fixtures stay pinned to the reference even when the input
    generators of make_golden.py move on."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "golden", "replay_golden.py")], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(" ok") == len(golden_names()) and "MISMATCH" not in r.stdout
