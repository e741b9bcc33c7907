import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clica_b200 import functional as F
from oracle import mlp_oracle
dev = torch.device("cuda:0")
def rel(a, b): return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / (np.abs(b).max() + 1e-30))
n = 10
for env in ({}, {"CLICA_TC_BN": "256"}, {"CLICA_TC_BN": "128"}, {"CLICA_TC_STAGES": "2"}):
    for k in ("CLICA_TC_BN", "CLICA_TC_STAGES"): os.environ.pop(k, None)
    os.environ.update(env)
    for M in (1000, 4096, 6144):
        for mode_name, mode in (("3xtf32", 0), ("tf32", 1)):
            rng = np.random.RandomState(n + M)
            widths = [n, 10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n, n]
            Wn = [(rng.uniform(-1, 1, size=(widths[i + 1], widths[i])) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
            bn = [(rng.uniform(-1, 1, size=(widths[i + 1],)) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
            xn = rng.randn(M, n).astype(np.float32)
            y_ref, acts, pre = mlp_oracle.mlp_forward(xn, Wn, bn, slope=0.01)
            Ws = [torch.tensor(w, device=dev) for w in Wn]; bs = [torch.tensor(b, device=dev) for b in bn]
            x = torch.tensor(xn, device=dev)
            errs = []
            for rep in range(3):
                y = F.mlp_forward(x, Ws, bs, slope=0.01, mode=mode)
                d = np.abs(y.cpu().numpy().astype(np.float64) - y_ref)
                bad_rows = np.where(d.max(axis=1) > 2e-3 * np.abs(y_ref).max())[0]
                errs.append("%.1e(bad rows %d, first %s)" % (d.max() / np.abs(y_ref).max(), len(bad_rows), bad_rows[:6].tolist()))
            print(env, f"M={M} {mode_name}:", " ".join(errs), flush=True)
