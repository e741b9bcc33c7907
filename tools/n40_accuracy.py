#!/usr/bin/env python
"""Accuracy of the 3xTF32 encoder stack at n = 40 (K up to 2000) against the numpy fp64 oracle, with and without K-chunked
accumulation (CLICA_TC_KCHUNK): relative errors (to the tensor max) of the output, the input gradient and every dW / db.

    python tools/n40_accuracy.py [--n 40] [--M 1536]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--M", type=int, default=1536)
    args = ap.parse_args()
    import torch
    from clica_b200 import functional as F
    from oracle import mlp_oracle
    from test_gpu_mlp import _safe_rows, _rel
    n, M = args.n, args.M
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(n + M)
    widths = [n, 10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n, n]
    Wn = [(rng.uniform(-1, 1, size=(widths[i + 1], widths[i])) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
    bn = [(rng.uniform(-1, 1, size=(widths[i + 1],)) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
    xn = rng.randn(M, n).astype(np.float32)
    xn = xn[_safe_rows(xn, Wn, bn)]
    M = len(xn)
    gyn = rng.randn(M, n).astype(np.float32)
    y_ref, acts, pre = mlp_oracle.mlp_forward(xn, Wn, bn, slope=0.01)
    dWs, dbs, dx = mlp_oracle.mlp_backward(gyn, Wn, acts, pre, slope=0.01, need_dx=True)
    for kc in ("0", "16", "8", "4"):
        os.environ["CLICA_TC_KCHUNK"] = kc
        Ws = [torch.tensor(w, device=dev, requires_grad=True) for w in Wn]
        bs = [torch.tensor(b, device=dev, requires_grad=True) for b in bn]
        x = torch.tensor(xn, device=dev, requires_grad=True)
        y = F.mlp_forward(x, Ws, bs, slope=0.01, mode=0)
        y.backward(torch.tensor(gyn, device=dev))
        torch.cuda.synchronize()
        errs = {"y": _rel(y.detach().cpu().numpy(), y_ref), "dx": _rel(x.grad.cpu().numpy(), dx)}
        errs["dW_max"] = max(_rel(Ws[l].grad.cpu().numpy(), dWs[l]) for l in range(7))
        errs["db_max"] = max(_rel(bs[l].grad.cpu().numpy(), dbs[l]) for l in range(7))
        print(f"n={n} M={M} CLICA_TC_KCHUNK={kc}: " + "  ".join(f"{k} {v:.2e}" for k, v in errs.items()), flush=True)


if __name__ == "__main__":
    main()
