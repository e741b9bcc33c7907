"""Debug probe: per-tensor relative error of the whole-stack encoder path against the numpy fp64 oracle."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clica_b200 import functional as F
from oracle import mlp_oracle
dev = torch.device("cuda:0")
def rel(a, b): return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / (np.abs(b).max() + 1e-30))
for (n, M) in [(5, 200), (10, 1000), (16, 333), (10, 6144)]:
    for mode_name, mode in (("fp32", 3), ("3xtf32", 0), ("tf32", 1)):
        rng = np.random.RandomState(n + M)
        widths = [n, 10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n, n]
        Wn = [(rng.uniform(-1, 1, size=(widths[i + 1], widths[i])) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
        bn = [(rng.uniform(-1, 1, size=(widths[i + 1],)) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
        xn = rng.randn(M, n).astype(np.float32); gyn = rng.randn(M, n).astype(np.float32)
        Ws = [torch.tensor(w, device=dev, requires_grad=True) for w in Wn]
        bs = [torch.tensor(b, device=dev, requires_grad=True) for b in bn]
        x = torch.tensor(xn, device=dev, requires_grad=True)
        y = F.mlp_forward(x, Ws, bs, slope=0.01, mode=mode)
        y.backward(torch.tensor(gyn, device=dev))
        y_ref, acts, pre = mlp_oracle.mlp_forward(xn, Wn, bn, slope=0.01)
        dWs, dbs, dx = mlp_oracle.mlp_backward(gyn, Wn, acts, pre, slope=0.01, need_dx=True)
        errs = ["y %.1e" % rel(y.detach().cpu().numpy(), y_ref), "dx %.1e" % rel(x.grad.cpu().numpy(), dx)]
        errs += ["dW%d %.1e" % (l, rel(Ws[l].grad.cpu().numpy(), dWs[l])) for l in range(7)]
        errs += ["db%d %.1e" % (l, rel(bs[l].grad.cpu().numpy(), dbs[l])) for l in range(7)]
        print(f"n={n} M={M} {mode_name}: " + " ".join(errs), flush=True)
