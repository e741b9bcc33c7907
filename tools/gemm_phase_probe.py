"""Phase breakdown of the tcgen05 GEMM kernel (clock64 stamps written by the kernel itself when
CLICA_TC_TIMING_PTR is set) for the encoder's layer shapes, under tile/stage overrides.

stamps: 0 entry | 1 setup done | 2 first TMA issued | 3 first stage landed | 4 last MMA issued |
        5 first accumulator ready (epilogue) | 6 epilogue done | 7 exit
"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clica_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
st = lambda: torch.cuda.current_stream().cuda_stream
MHZ = 1965.0
tbuf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)

def run(fn, reps=10):
    os.environ.pop("CLICA_TC_TIMING_PTR", None)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    os.environ["CLICA_TC_TIMING_PTR"] = str(tbuf.data_ptr())
    tbuf.zero_(); fn(); torch.cuda.synchronize()
    os.environ.pop("CLICA_TC_TIMING_PTR", None)
    t = tbuf.view(148, 8).cpu().double()
    t = t[t[:, 7] > 0]
    d = lambda a, b: ((t[:, b] - t[:, a]).mean().item() / MHZ)
    return us, len(t), d(0, 1), d(1, 3), d(3, 4), d(4, 5), d(5, 6), d(0, 7)

shapes = [(6144, 500, 500), (12288, 500, 500), (6144, 100, 500), (8192, 2000, 2000)]
mode = 0
for (M, K, N) in shapes:
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev); dy = torch.randn(M, N, device=dev); dx = torch.empty(M, K, device=dev)
    dW = torch.empty(N, K, device=dev); db = torch.empty(N, device=dev)
    ws = torch.empty(lib.clica_linear_workspace_bytes(M, N, K, mode), dtype=torch.uint8, device=dev)
    f = lambda: lib.clica_linear_act_fwd(x.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), y.data_ptr(), N, M, K, N, 0.01, mode, ws.data_ptr(), ws.numel(), st())
    g = lambda: lib.clica_linear_act_bwd_data(dy.data_ptr(), N, W.data_ptr(), K, x.data_ptr(), K, 0.01, dx.data_ptr(), K, M, K, N, mode, ws.data_ptr(), ws.numel(), st())
    h = lambda: lib.clica_linear_bwd_weight(dy.data_ptr(), N, x.data_ptr(), K, dW.data_ptr(), K, db.data_ptr(), M, K, N, mode, ws.data_ptr(), ws.numel(), st())
    for bn, stages in ((256, 0), (256, 1), (128, 0), (128, 2)):
        os.environ["CLICA_TC_BN"] = str(bn)
        os.environ["CLICA_TC_STAGES"] = str(stages)
        for name, fn in (("fwd", f), ("dX", g), ("dW", h)):
            us, n, setup, first, main, drain, epi, tot = run(fn)
            print(f"M={M} K={K} N={N} BN={bn} st={stages} {name:3s}: call {us:7.1f} us (incl. plane split) | ctas {n:3d} setup {setup:5.2f} "
                  f"first-load {first:5.2f} mainloop {main:6.2f} drain {drain:5.2f} epilogue {epi:6.2f} total {tot:6.2f} us", flush=True)
