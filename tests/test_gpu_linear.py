"""GPU (-m gpu): the per-layer C-ABI entry points (clica_linear_act_fwd / _bwd_data / _bwd_weight) in every GEMM
mode against float64 numpy.  Shapes cover tile-aligned, ragged M / N / K, split-K reductions and the
CUDA-core route for skinny layers.  Tolerances as in tests/test_gpu_mlp.py.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MODES = {"fp32": (3, 2e-5), "3xtf32": (0, 2e-5), "tf32": (1, 5e-3)}
SHAPES = [(128, 128, 128), (256, 64, 96), (200, 250, 50), (1000, 500, 500), (333, 100, 500), (6144, 500, 100),
          (77, 10, 100), (640, 100, 10)]


def _rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / (np.abs(b).max() + 1e-30))


def _ws(lib, M, N, K, mode, dev):
    n = lib.clica_linear_workspace_bytes(M, N, K, mode)
    return torch.empty(max(n, 16), dtype=torch.uint8, device=dev)


@pytest.mark.parametrize("mode_name", list(MODES))
@pytest.mark.parametrize("M,K,N", SHAPES)
def test_linear_forward(mode_name, M, K, N, cuda_device):
    from clica_b200 import _lib
    lib = _lib.load()
    mode, tol = MODES[mode_name]
    rng = np.random.RandomState(M + K + N)
    x, W, b = rng.randn(M, K).astype(np.float32), (rng.randn(N, K) / np.sqrt(K)).astype(np.float32), rng.randn(N).astype(np.float32)
    xd, Wd, bd = (torch.tensor(a, device=cuda_device) for a in (x, W, b))
    y = torch.full((M, N), float("nan"), device=cuda_device)
    ws = _ws(lib, M, N, K, mode, cuda_device)
    rc = lib.clica_linear_act_fwd(xd.data_ptr(), K, Wd.data_ptr(), K, bd.data_ptr(), y.data_ptr(), N, M, K, N, 0.01, mode,
                                  ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "clica_linear_act_fwd")
    z = x.astype(np.float64) @ W.astype(np.float64).T + b
    ref = np.where(z > 0, z, 0.01 * z)
    assert _rel(y.cpu().numpy(), ref) <= tol


@pytest.mark.parametrize("mode_name", list(MODES))
@pytest.mark.parametrize("M,K,N", SHAPES)
def test_linear_backward_data(mode_name, M, K, N, cuda_device):
    from clica_b200 import _lib
    lib = _lib.load()
    mode, tol = MODES[mode_name]
    rng = np.random.RandomState(M + 2 * K + N)
    dy, W = rng.randn(M, N).astype(np.float32), (rng.randn(N, K) / np.sqrt(K)).astype(np.float32)
    xa = rng.randn(M, K).astype(np.float32)
    dyd, Wd, xad = (torch.tensor(a, device=cuda_device) for a in (dy, W, xa))
    dx = torch.full((M, K), float("nan"), device=cuda_device)
    ws = _ws(lib, M, N, K, mode, cuda_device)
    rc = lib.clica_linear_act_bwd_data(dyd.data_ptr(), N, Wd.data_ptr(), K, xad.data_ptr(), K, 0.01, dx.data_ptr(), K,
                                       M, K, N, mode, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "clica_linear_act_bwd_data")
    ref = (dy.astype(np.float64) @ W.astype(np.float64)) * np.where(xa > 0, 1.0, 0.01)
    assert _rel(dx.cpu().numpy(), ref) <= tol


@pytest.mark.parametrize("mode_name", list(MODES))
@pytest.mark.parametrize("M,K,N", SHAPES)
def test_linear_backward_weight(mode_name, M, K, N, cuda_device):
    from clica_b200 import _lib
    lib = _lib.load()
    mode, tol = MODES[mode_name]
    rng = np.random.RandomState(M + K + 3 * N)
    dy, x = rng.randn(M, N).astype(np.float32), rng.randn(M, K).astype(np.float32)
    dyd, xd = torch.tensor(dy, device=cuda_device), torch.tensor(x, device=cuda_device)
    dW = torch.full((N, K), float("nan"), device=cuda_device)
    db = torch.full((N,), float("nan"), device=cuda_device)
    ws = _ws(lib, M, N, K, mode, cuda_device)
    rc = lib.clica_linear_bwd_weight(dyd.data_ptr(), N, xd.data_ptr(), K, dW.data_ptr(), K, db.data_ptr(), M, K, N, mode,
                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "clica_linear_bwd_weight")
    assert _rel(dW.cpu().numpy(), dy.astype(np.float64).T @ x.astype(np.float64)) <= tol
    assert _rel(db.cpu().numpy(), dy.astype(np.float64).sum(0)) <= 2e-5
