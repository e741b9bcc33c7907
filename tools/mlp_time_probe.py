"""Whole-stack encoder timing (clica_mlp_fwd + clica_mlp_bwd) per kernel family, under tile-width overrides."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clica_b200 import _lib, functional as F
lib = _lib.load()
dev = torch.device("cuda:0")
cases = [(10, 6144), (10, 12288), (40, 8192), (40, 1024)]
bns = [int(a) for a in sys.argv[1:]] or [0, 256, 192, 128]
for (n, M) in cases:
    widths = [n, 10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n, n]
    torch.manual_seed(0)
    Ws = [(torch.rand(widths[i + 1], widths[i], device=dev) * 2 - 1).div_(widths[i] ** 0.5).requires_grad_() for i in range(7)]
    bs = [(torch.rand(widths[i + 1], device=dev) * 2 - 1).div_(widths[i] ** 0.5).requires_grad_() for i in range(7)]
    x = torch.randn(M, n, device=dev)
    gy = torch.randn(M, n, device=dev)
    for bn in bns:
        os.environ["CLICA_TC_BN"] = str(bn)
        def step():
            for t in Ws + bs: t.grad = None
            y = F.mlp_forward(x, Ws, bs, slope=0.01, mode=0)
            y.backward(gy)
        for _ in range(3): step()
        torch.cuda.synchronize()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): step()
        e1.record(); torch.cuda.synchronize()
        wall = e0.elapsed_time(e1) / reps
        lib.clica_prof_enable(1)
        for _ in range(reps): step()
        ms = (ctypes.c_float * 7)(); cnt = (ctypes.c_int * 7)()
        lib.clica_prof_collect(ms, cnt); lib.clica_prof_enable(0)
        print(f"n={n} M={M} BN={bn}: fwd+bwd {wall * 1e3:8.1f} us | tc {ms[3] / reps * 1e3:8.1f} us ({cnt[3] // reps} launches) "
              f"simt {ms[4] / reps * 1e3:7.1f} us misc {ms[6] / reps * 1e3:6.1f} us", flush=True)
