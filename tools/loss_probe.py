#!/usr/bin/env python
"""Times the fused loss kernels at one shape for several data regimes and kernel variants (GPU box).

    python tools/loss_probe.py [--B 6144] [--M 0] [--d 10] [--p 2] [--iters 30] [--out gpurun_out/loss_probe.json]
Regimes: init (tiny spread around an offset, what a freshly initialised encoder emits), sphere (unit sphere, what a
trained one emits), gauss (std 0.7 Gaussian: the norm bound of the p = 2 dot form fails for part of the pairs).
Variants: CLICA_LPNCE_DOT in {0, 1} x CLICA_LPNCE_R4 in {0, 1} (p = 2 only)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=6144)
    ap.add_argument("--M", type=int, default=0)
    ap.add_argument("--d", type=int, default=10)
    ap.add_argument("--p", type=float, default=2.0)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--once", action="store_true", help="one fwd+bwd per variant (for ncu)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from clica_b200 import functional as F
    dev = torch.device("cuda:0")
    B, d, p = args.B, args.d, args.p
    M = args.M or B
    rng = np.random.RandomState(0)
    res = []
    variants = [("1", "1"), ("1", "0"), ("0", "1"), ("0", "0")] if p == 2.0 else [("0", "0")]
    for regime in ("init", "sphere", "gauss"):
        z = rng.randn(max(B, M), d).astype(np.float32)
        if regime == "sphere":
            z /= np.linalg.norm(z, axis=1, keepdims=True)
        elif regime == "init":
            z = (0.3 + 0.02 * z).astype(np.float32)
        else:
            z *= 0.7
        z1 = torch.tensor(z[:B], device=dev, requires_grad=True)
        z2 = (z1.detach() + 0.05 * torch.randn(B, d, device=dev)).requires_grad_(True)
        for dot, r4 in variants:
            os.environ["CLICA_LPNCE_DOT"], os.environ["CLICA_LPNCE_R4"] = dot, r4
            for sym in (True, False):
                def fwd():
                    z3 = torch.roll(z1, 1, 0) if sym else z3_leaf
                    return F.lp_infonce(z1, z2, z3, p, 1.0, 0.5, True)
                z3_leaf = torch.tensor(z[:M], device=dev, requires_grad=True)
                n_it = 1 if args.once else args.iters
                for _ in range(0 if args.once else 3):
                    fwd()[0].backward()
                torch.cuda.synchronize()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                tf = tb = 0.0
                for _ in range(n_it):
                    z1.grad = z2.grad = None
                    e[0].record()
                    out = fwd()
                    e[1].record()
                    out[0].backward()
                    e[2].record()
                    torch.cuda.synchronize()
                    tf += e[0].elapsed_time(e[1])
                    tb += e[1].elapsed_time(e[2])
                row = dict(regime=regime, dot=dot, r4=r4, symmetric=sym, B=B, M=M, d=d, p=p,
                           fwd_us=1e3 * tf / n_it, bwd_us=1e3 * tb / n_it, loss=float(out[0]))
                res.append(row)
                print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
