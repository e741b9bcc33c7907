"""GPU (-m gpu): encoder kernels through the C ABI vs the numpy fp64 oracle and the reference golden vectors.

Tolerances (relative to the tensor's max magnitude, through the whole 7-layer stack forward + backward):
CLICA_GEMM_FP32 (exact fp32 products, fp32 FFMA accumulation) 2e-5 -- the band the reference's own fp32 cuBLAS path
occupies (2e-6..1e-5); CLICA_GEMM_3XTF32 (hi/lo split, ~2^-21 per product, but the tensor
core's fp32 accumulator truncates when it aligns addends, a bias that grows with the number of accumulated k-steps:
measured 7e-6 at n = 10, 2.3e-5 at n = 16 with K = 800) 5e-5; CLICA_GEMM_TF32 is the labelled fast mode: 5e-3.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

MODES = {"fp32": (3, 2e-5), "3xtf32": (0, 5e-5), "tf32": (1, 5e-3)}


def _rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / (np.abs(b).max() + 1e-30))


def _cos(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))


def _safe_rows(x, Ws, bs, slope=0.01, thresh=2e-5):
    """Rows whose hidden pre-activations all stay away from 0.

    LeakyReLU'(z) jumps at z = 0: a pre-activation within rounding distance of 0 can take a different branch in
    fp32 than in the fp64 oracle, which changes that sample's whole gradient (this happens to the reference's
    own fp32 run as well).  Such samples say nothing about kernel accuracy, so they are dropped from the batch."""
    from oracle import mlp_oracle
    _, _, pre = mlp_oracle.mlp_forward(x, Ws, bs, slope=slope)
    ok = np.ones(len(x), dtype=bool)
    for z in pre[:-1]:
        ok &= np.abs(z).min(axis=1) > thresh * np.abs(z).max()
    return ok


@pytest.mark.parametrize("mode_name", list(MODES))
def test_golden_small_mlp(mode_name, cuda_device):
    from clica_b200 import functional as F
    mode, tol = MODES[mode_name]
    g = load_golden("mlp_small")
    keys = [str(k) for k in g["keys"]]
    wk = [k for k in keys if k.endswith("weight")]
    bk = [k for k in keys if k.endswith("bias")]
    Ws = [torch.tensor(g["param_" + k], device=cuda_device, requires_grad=True) for k in wk]
    bs = [torch.tensor(g["param_" + k], device=cuda_device, requires_grad=True) for k in bk]
    x = torch.tensor(g["x"], device=cuda_device, requires_grad=True)
    y = F.mlp_forward(x, Ws, bs, slope=0.01, mode=mode)
    y.backward(torch.tensor(g["gy"], device=cuda_device))
    assert _rel(y.detach().cpu().numpy(), g["y_64"]) <= tol
    assert _rel(x.grad.cpu().numpy(), g["dx_64"]) <= tol
    for k, W in zip(wk, Ws):
        assert _rel(W.grad.cpu().numpy(), g["grad64_" + k]) <= tol, k
    for k, b in zip(bk, bs):
        assert _rel(b.grad.cpu().numpy(), g["grad64_" + k]) <= tol, k


@pytest.mark.parametrize("mode_name", list(MODES))
@pytest.mark.parametrize("n,M", [(5, 200), (10, 1000), (16, 333), (10, 6144)])
def test_encoder_stack_against_numpy_oracle(mode_name, n, M, cuda_device):
    """The real encoder shape n -> 10n -> 50n x4 -> 10n -> n (main_mlp.py:297-309), ragged M."""
    from clica_b200 import functional as F
    from oracle import mlp_oracle
    mode, tol = MODES[mode_name]
    rng = np.random.RandomState(n + M)
    widths = [n, 10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n, n]
    Wn = [(rng.uniform(-1, 1, size=(widths[i + 1], widths[i])) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
    bn = [(rng.uniform(-1, 1, size=(widths[i + 1],)) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
    xn = rng.randn(M, n).astype(np.float32)
    if mode_name != "tf32":
        xn = xn[_safe_rows(xn, Wn, bn)]
        assert len(xn) > M // 3
        M = len(xn)
    gyn = rng.randn(M, n).astype(np.float32)
    Ws = [torch.tensor(w, device=cuda_device, requires_grad=True) for w in Wn]
    bs = [torch.tensor(b, device=cuda_device, requires_grad=True) for b in bn]
    x = torch.tensor(xn, device=cuda_device, requires_grad=True)
    y = F.mlp_forward(x, Ws, bs, slope=0.01, mode=mode)
    y.backward(torch.tensor(gyn, device=cuda_device))
    y_ref, acts, pre = mlp_oracle.mlp_forward(xn, Wn, bn, slope=0.01)
    dWs, dbs, dx = mlp_oracle.mlp_backward(gyn, Wn, acts, pre, slope=0.01, need_dx=True)
    assert _rel(y.detach().cpu().numpy(), y_ref) <= tol
    if mode_name == "tf32":
        # single-pass TF32 is the labelled fast mode: ~5e-4 per layer lets LeakyReLU masks flip for small
        # pre-activations, so gradients are compared by direction, not element-wise
        assert _cos(x.grad.cpu().numpy(), dx) >= 0.98
        for l in range(7):
            assert _cos(Ws[l].grad.cpu().numpy(), dWs[l]) >= 0.98, f"dW{l}"
        return
    assert _rel(x.grad.cpu().numpy(), dx) <= tol
    for l in range(7):
        assert _rel(Ws[l].grad.cpu().numpy(), dWs[l]) <= tol, f"dW{l}"
        assert _rel(bs[l].grad.cpu().numpy(), dbs[l]) <= tol, f"db{l}"


@pytest.mark.parametrize("mode_name,tol", [("fp32", 2e-5), ("3xtf32", 3e-5)])
def test_encoder_stack_n40(mode_name, tol, cuda_device):
    """BASELINE config 3's encoder (n = 40: 40 -> 400 -> 2000 x4 -> 400 -> 40).  The first / last layers take the
    KMAX = 48 skinny kernels; the 2000-wide tensor-core layers accumulate 63 k-blocks, where the tensor core's
    truncating fp32 accumulator would cost 3xTF32 ~5e-5: the kernel therefore accumulates at most 16 k-blocks in TMEM and
    adds the chunks in fp32 registers (gemm_tc.cu, CLICA_TC_KCHUNK): measured 1.4e-5 / 1.7e-5, stated 3e-5; exact-fp32
    mode 2e-5."""
    from clica_b200 import functional as F
    from oracle import mlp_oracle
    mode = MODES[mode_name][0]
    n, M = 40, 384
    rng = np.random.RandomState(40)
    widths = [n, 10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n, n]
    Wn = [(rng.uniform(-1, 1, size=(widths[i + 1], widths[i])) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
    bn = [(rng.uniform(-1, 1, size=(widths[i + 1],)) / np.sqrt(widths[i])).astype(np.float32) for i in range(7)]
    xn = rng.randn(M, n).astype(np.float32)
    xn = xn[_safe_rows(xn, Wn, bn)]
    M = len(xn)
    assert M >= 32
    gyn = rng.randn(M, n).astype(np.float32)
    Ws = [torch.tensor(w, device=cuda_device, requires_grad=True) for w in Wn]
    bs = [torch.tensor(b, device=cuda_device, requires_grad=True) for b in bn]
    x = torch.tensor(xn, device=cuda_device, requires_grad=True)
    y = F.mlp_forward(x, Ws, bs, slope=0.01, mode=mode)
    y.backward(torch.tensor(gyn, device=cuda_device))
    y_ref, acts, pre = mlp_oracle.mlp_forward(xn, Wn, bn, slope=0.01)
    dWs, dbs, dx = mlp_oracle.mlp_backward(gyn, Wn, acts, pre, slope=0.01, need_dx=True)
    errs = {"y": _rel(y.detach().cpu().numpy(), y_ref), "dx": _rel(x.grad.cpu().numpy(), dx)}
    for l in range(7):
        errs[f"dW{l}"] = _rel(Ws[l].grad.cpu().numpy(), dWs[l])
        errs[f"db{l}"] = _rel(bs[l].grad.cpu().numpy(), dbs[l])
    print(f"n=40 {mode_name}: max rel err {max(errs.values()):.2e} ({max(errs, key=errs.get)})")
    assert max(errs.values()) <= tol, errs


def test_dropin_get_mlp_on_cuda_matches_torch_modules(cuda_device):
    """The FusedMLP returned by the drop-in get_mlp must agree with the same nn modules run by torch."""
    import sys
    import clica_b200
    sys.path.insert(0, clica_b200.DROPIN_DIR)
    import encoders
    torch.manual_seed(5)
    n = 10
    f = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n])
    x = torch.randn(777, n)
    lin = [m for m in f if isinstance(m, torch.nn.Linear)]
    keep = _safe_rows(x.numpy(), [m.weight.detach().numpy() for m in lin], [m.bias.detach().numpy() for m in lin])
    f = f.to(cuda_device)
    x = x[torch.tensor(keep)].to(cuda_device)
    y = f(x)
    y_t = torch.nn.Sequential.forward(f, x)          # plain torch execution of the very same modules (fp32 cuBLAS)
    assert (y - y_t).abs().max().item() <= 2e-5 * y_t.abs().max().item()
    (y ** 2).mean().backward()
    g_fused = [p.grad.clone() for p in f.parameters()]
    f.zero_grad()
    (y_t ** 2).mean().backward()
    for gf, p in zip(g_fused, f.parameters()):
        assert (gf - p.grad).abs().max().item() <= 5e-5 * (p.grad.abs().max().item() + 1e-30)
    assert list(f.state_dict().keys())[:2] == ["0.weight", "0.bias"]


def test_fused_adam_matches_torch_adam(cuda_device):
    from clica_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(100, 10), (100,), (500, 100), (500,), (7,), (1, 1)]
    p_ref = [torch.randn(s, device=cuda_device, requires_grad=True) for s in shapes]
    p_fus = [p.detach().clone().requires_grad_(True) for p in p_ref]
    o_ref = torch.optim.Adam(p_ref, lr=1e-3)
    o_fus = FusedAdam(p_fus, lr=1e-3)
    for _ in range(5):
        for a, b in zip(p_ref, p_fus):
            g = torch.randn_like(a)
            a.grad, b.grad = g.clone(), g.clone()
        o_ref.step(), o_fus.step()
    for a, b in zip(p_ref, p_fus):
        assert (a - b).abs().max().item() <= 2e-6 * max(1.0, a.abs().max().item())


def test_packed_weight_cache_is_keyed_by_tensor_identity(cuda_device):
    """A NEW weight tensor that reuses a freed tensor's address (and version 0) must not hit the packed-weight
    cache of the old one (regression: stale planes made the 3rd same-shaped encoder of a process compute with
    the 1st one's weights)."""
    from clica_b200 import functional as F
    widths = [10, 100, 500, 500, 100, 10]
    x = torch.randn(640, widths[0], device=cuda_device)

    def run(seed):
        g = torch.Generator(device="cpu").manual_seed(seed)
        Ws = [(torch.rand(widths[i + 1], widths[i], generator=g) * 2 - 1).div_(widths[i] ** 0.5).to(cuda_device) for i in range(5)]
        bs = [torch.zeros(widths[i + 1], device=cuda_device) for i in range(5)]
        ptr0 = Ws[0].data_ptr()
        y = F.mlp_forward(x, Ws, bs, slope=0.01, mode=_lib_mode("3xtf32"))
        ref = x
        for l, (W, b) in enumerate(zip(Ws, bs)):
            ref = ref.double() @ W.double().t() + b.double()
            if l < 4:
                ref = torch.nn.functional.leaky_relu(ref, 0.01)
        return ptr0, (y.double() - ref).abs().max().item() / ref.abs().max().item()

    ptrs = set()
    for seed in range(4):
        ptr0, err = run(seed)
        ptrs.add(ptr0)
        assert err <= 5e-5, (seed, err)
    # (the caching allocator normally hands the same block back, which is exactly the hazardous case)


def _lib_mode(name):
    return MODES[name][0]
