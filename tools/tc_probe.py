"""Debug probe: try MN-major descriptor / TMA-swizzle combinations for the tcgen05 TF32 GEMM (bwd_data path)."""
import itertools, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from clica_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
M, K, N, mode = 128, 128, 128, 1
rng = np.random.RandomState(0)
dy, W = rng.randn(M, N).astype(np.float32), rng.randn(N, K).astype(np.float32)
dyd, Wd = torch.tensor(dy, device=dev), torch.tensor(W, device=dev)
dx = torch.full((M, K), float("nan"), device=dev)
ws = torch.empty(lib.clica_linear_workspace_bytes(M, N, K, mode), dtype=torch.uint8, device=dev)
rc = lib.clica_linear_act_bwd_data(dyd.data_ptr(), N, Wd.data_ptr(), K, None, 0, 1.0, dx.data_ptr(), K, M, K, N, mode,
                                   ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
ref = dy.astype(np.float64) @ W.astype(np.float64)
got = dx.cpu().numpy()
print("rc", rc, "relerr %%.3e" %% (np.abs(got - ref).max() / np.abs(ref).max()), "nonzero", int((got != 0).sum()), "nan", int(np.isnan(got).sum()))
''' % ROOT
combos = [(1, 4, 512, 4096), (1, 4, 1024, 4096), (1, 4, 4096, 512), (1, 4, 4096, 1024), (2, 3, 1024, 4096), (2, 3, 4096, 1024),
          (1, 3, 512, 4096), (2, 4, 1024, 4096), (1, 4, 256, 4096), (1, 4, 512, 2048)]
for lt, swz, sbo, lbo in combos:
    env = dict(os.environ, CLICA_TC_MN_LT=str(lt), CLICA_TC_MN_SWZ=str(swz), CLICA_TC_MN_SBO=str(sbo), CLICA_TC_MN_LBO=str(lbo))
    try:
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=90)
        print(f"lt={lt} swz={swz} sbo={sbo} lbo={lbo}: {out.stdout.strip()} {out.stderr.strip()[-200:]}", flush=True)
    except subprocess.TimeoutExpired:
        print(f"lt={lt} swz={swz} sbo={sbo} lbo={lbo}: TIMEOUT", flush=True)
