#!/usr/bin/env python
"""Probe of cal_selftest_umma on a B200: every (kind, M, N, K, variant) in its own process, so that a
configuration the hardware rejects (illegal instruction / descriptor) cannot poison the others.
Prints one line per configuration: max-norm relative error vs a float64 product."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(kind, M, N, K, variant):
    import ctypes as C
    import torch
    from cal_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(1000 * kind + M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.cal_selftest_umma(kind, M, N, K, Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), variant,
                               torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = A.double() @ B.double().t()
    got = D.cpu().double()
    err = float((got - want).abs().max() / want.abs().max())
    f32 = float(((A @ B.t()).double() - want).abs().max() / want.abs().max())
    print("kind %d M %3d N %3d K %4d variant %d rc %d  rel err %.3e   (torch fp32 matmul on CPU: %.3e)"
          % (kind, M, N, K, variant, rc, err, f32), flush=True)


def main():
    if len(sys.argv) > 1:
        one(*[int(v) for v in sys.argv[1:6]])
        return
    cfgs = []
    for variant in (0, 1):
        for kind in (0, 1, 2):
            cfgs += [(kind, 128, 128, 128, variant), (kind, 128, 32, 32, variant), (kind, 100, 48, 40, variant)]
    cfgs += [(0, 128, 16, 128, 0), (0, 128, 24, 128, 0), (0, 128, 8, 128, 0), (0, 128, 200, 64, 0), (0, 128, 256, 256, 0),
             (0, 128, 128, 1024, 0), (2, 128, 128, 1024, 0), (0, 7, 5, 3, 0),
             (0, 128, 24, 128, 2), (0, 128, 8, 64, 2), (0, 128, 40, 64, 2), (2, 128, 24, 64, 2),
             (0, 128, 128, 128, 4), (0, 128, 32, 96, 4), (2, 128, 64, 128, 4), (0, 128, 24, 128, 6)]
    for c in cfgs:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + [str(v) for v in c], capture_output=True, text=True,
                           timeout=120)
        out = (r.stdout.strip().splitlines() or ["(no output)"])[-1]
        if r.returncode != 0:
            out = "cfg %s FAILED rc %d: %s" % (c, r.returncode, (r.stderr.strip().splitlines() or ["?"])[-1][:200])
        print(out, flush=True)


if __name__ == "__main__":
    main()
