"""Per-shape timing of the tcgen05 GEMM kernel alone (library-side CUDA events around the kernel launch)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clica_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
st = lambda: torch.cuda.current_stream().cuda_stream
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    lib.clica_prof_enable(1)
    for _ in range(reps): fn()
    ms = (ctypes.c_float * 7)(); n = (ctypes.c_int * 7)()
    lib.clica_prof_collect(ms, n); lib.clica_prof_enable(0)
    return ms[3] / max(n[3], 1) * 1e3, ms[6] / reps * 1e3     # us per tc-gemm launch, misc us per call
shapes = [(12288, 100, 500), (12288, 500, 500), (12288, 500, 100), (16384, 2000, 2000), (16384, 400, 2000)]
modes = (("3xtf32", 0), ("tf32", 1)) if "--all-modes" in sys.argv else (("3xtf32", 0),)
print("CLICA_TC_PAIR =", os.environ.get("CLICA_TC_PAIR", "(default)"), flush=True)
for mode_name, mode in modes:
    for (M, K, N) in shapes:
        x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
        y = torch.empty(M, N, device=dev); dy = torch.randn(M, N, device=dev); dx = torch.empty(M, K, device=dev)
        dW = torch.empty(N, K, device=dev); db = torch.empty(N, device=dev)
        ws = torch.empty(lib.clica_linear_workspace_bytes(M, N, K, mode), dtype=torch.uint8, device=dev)
        f = lambda: lib.clica_linear_act_fwd(x.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), y.data_ptr(), N, M, K, N, 0.01, mode, ws.data_ptr(), ws.numel(), st())
        g = lambda: lib.clica_linear_act_bwd_data(dy.data_ptr(), N, W.data_ptr(), K, x.data_ptr(), K, 0.01, dx.data_ptr(), K, M, K, N, mode, ws.data_ptr(), ws.numel(), st())
        h = lambda: lib.clica_linear_bwd_weight(dy.data_ptr(), N, x.data_ptr(), K, dW.data_ptr(), K, db.data_ptr(), M, K, N, mode, ws.data_ptr(), ws.numel(), st())
        fl = 2.0 * M * K * N
        out = []
        for name, fn in (("fwd", f), ("dX", g), ("dW", h)):
            us, misc = timed(fn)
            out.append(f"{name} {us:7.1f} us {fl / us / 1e6:6.1f} TF/s")
        print(f"{mode_name} M={M} K={K} N={N}: " + " | ".join(out), flush=True)
