"""GPU (-m gpu): the chained tcgen05 GEMM launch (one persistent grid per forward / backward of the hidden layers, tiles
ordered by per-row-block arrival counters; gemm_tc.cu ChainParams) and the Adam step that keeps the packed weight planes
current (clica_adam_step_capturable_packed).

Reference behaviour: nn.Linear + nn.LeakyReLU of encoders.py:38-48 and their autograd nodes (numpy fp64 oracle), and
torch.optim.Adam of main_mlp.py:312 followed by the next step's encoder forward.
"""
import copy
import ctypes
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _encoder(n, dev, seed=0):
    import clica_b200
    if clica_b200.DROPIN_DIR not in sys.path:
        sys.path.insert(0, clica_b200.DROPIN_DIR)
    import encoders
    torch.manual_seed(seed)
    return encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)


def _run_stack(f, x, gy):
    for p in f.parameters():
        p.grad = None
    xin = x.clone().requires_grad_(True)
    y = f(xin)
    y.backward(gy)
    return y.detach(), xin.grad.detach(), [p.grad.detach().clone() for p in f.parameters()]


@pytest.mark.parametrize("n,M", [(10, 12288), (10, 700), (16, 2049), (40, 1024), (5, 4096)])
def test_chained_launch_equals_per_gemm_launches(n, M, cuda_device, monkeypatch):
    """Same tiles, same k order, same epilogues: the chained launch must reproduce the per-GEMM launches bit for bit in
    everything a single tile produces (activations, input gradient) and to split-K summation order in dW / db."""
    f = _encoder(n, cuda_device)
    g = torch.Generator(device="cpu").manual_seed(n + M)
    x = torch.randn(M, n, generator=g).to(cuda_device)
    gy = (torch.randn(M, n, generator=g) / M).to(cuda_device)
    monkeypatch.setenv("CLICA_TC_CHAIN", "0")
    y0, gx0, gp0 = _run_stack(f, x, gy)
    monkeypatch.setenv("CLICA_TC_CHAIN", "1")
    y1, gx1, gp1 = _run_stack(f, x, gy)
    assert torch.equal(y0, y1)
    assert torch.equal(gx0, gx1)
    for a, b in zip(gp0, gp1):
        scale = a.abs().max().item() + 1e-30
        assert (a - b).abs().max().item() <= 2e-6 * scale      # TMA reduce-adds of the k-splits land in a different order


@pytest.mark.parametrize("n,M", [(10, 6144), (40, 1536)])
def test_chained_stack_against_the_fp64_oracle(n, M, cuda_device, monkeypatch):
    """Whole stack forward + backward through the chained launches against the numpy fp64 encoder oracle (rows whose
    pre-activations sit within rounding distance of 0 are dropped first: tests/test_gpu_mlp.py::_safe_rows)."""
    from oracle import mlp_oracle
    from test_gpu_mlp import _rel, _safe_rows
    monkeypatch.setenv("CLICA_TC_CHAIN", "1")
    f = _encoder(n, cuda_device, seed=3)
    lin = [m for m in f if isinstance(m, torch.nn.Linear)]
    Ws = [m.weight.detach().cpu().numpy() for m in lin]
    bs = [m.bias.detach().cpu().numpy() for m in lin]
    rng = np.random.RandomState(n)
    x = rng.randn(M, n).astype(np.float32)
    x = x[_safe_rows(x, Ws, bs)]
    assert len(x) > M // 3
    M = len(x)
    gy = (rng.randn(M, n) / M).astype(np.float32)
    y, gx, gp = _run_stack(f, torch.tensor(x, device=cuda_device), torch.tensor(gy, device=cuda_device))
    y_ref, acts, pre = mlp_oracle.mlp_forward(x, Ws, bs, slope=0.01)
    dWs, dbs, gx_ref = mlp_oracle.mlp_backward(gy, Ws, acts, pre, slope=0.01, need_dx=True)
    tol = 5e-5 if n <= 16 else 3e-5                           # 3xTF32 through 7 layers (tolerances)
    assert _rel(y.cpu().numpy(), y_ref) <= tol
    assert _rel(gx.cpu().numpy(), gx_ref) <= tol
    for l in range(len(Ws)):
        assert _rel(gp[2 * l].cpu().numpy(), dWs[l]) <= tol, f"dW{l}"
        assert _rel(gp[2 * l + 1].cpu().numpy(), dbs[l]) <= tol, f"db{l}"


def test_adam_keeps_the_packed_planes_current(cuda_device):
    """clica_adam_step_capturable_packed: after the update the (hi, lo) planes hold exactly what clica_mlp_pack_weights
    produces from the updated weights, and the parameters equal the plain capturable step's."""
    from clica_b200 import _lib, functional as F
    lib = _lib.load()
    mode = _lib.GEMM_MODES["3xtf32"]
    torch.manual_seed(0)
    widths = [12, 100, 500, 100, 12]
    Ws = [torch.randn(widths[i + 1], widths[i], device=cuda_device) * 0.1 for i in range(4)]
    Ws2 = [w.clone() for w in Ws]
    grads = [torch.randn_like(w) for w in Ws]
    m1, v1 = [torch.zeros_like(w) for w in Ws], [torch.zeros_like(w) for w in Ws]
    m2, v2 = [torch.zeros_like(w) for w in Ws], [torch.zeros_like(w) for w in Ws]
    L = len(Ws)
    cw = (ctypes.c_int * (L + 1))(*widths)
    nbytes = lib.clica_mlp_packed_weight_bytes(L, cw, mode)
    buf = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=cuda_device)
    ptr = (buf.data_ptr() + 1023) // 1024 * 1024
    F.repack_weights_into(ptr, Ws, mode)                       # initial planes (zero padding columns)
    targets = F.packed_weight_targets(Ws, mode, ptr)
    assert 2 <= len(targets) <= 4                               # the layers with K, N >= 32 live in the buffer
    st1 = torch.zeros(2, dtype=torch.int64, device=cuda_device)
    st2 = torch.zeros(2, dtype=torch.int64, device=cuda_device)
    for _ in range(3):
        F.adam_step_capturable(Ws, grads, m1, v1, 1e-2, 0.9, 0.999, 1e-8, st1, pack=[targets.get(id(w)) for w in Ws])
        F.adam_step_capturable(Ws2, grads, m2, v2, 1e-2, 0.9, 0.999, 1e-8, st2)
    for a, b in zip(Ws, Ws2):
        assert torch.equal(a, b)
    ref = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=cuda_device)
    ref_ptr = (ref.data_ptr() + 1023) // 1024 * 1024
    F.repack_weights_into(ref_ptr, Ws, mode)
    torch.cuda.synchronize()
    off, ref_off = ptr - buf.data_ptr(), ref_ptr - ref.data_ptr()
    assert torch.equal(buf[off:off + nbytes], ref[ref_off:ref_off + nbytes])


def test_graphed_step_survives_external_weight_changes(cuda_device):
    """The recorded step reads packed planes that its own Adam keeps current; a weight change from outside (checkpoint
    load, manual copy_) must be picked up before the next replay."""
    import clica_b200
    from clica_b200 import synth
    from clica_b200.graphed import GraphedTrainStep
    if clica_b200.DROPIN_DIR not in sys.path:
        sys.path.insert(0, clica_b200.DROPIN_DIR)
    import losses
    n, B = 10, 512
    f1 = _encoder(n, cuda_device, seed=1)
    f2 = _encoder(n, cuda_device, seed=2)
    g = synth.build_mixing(n, 3, seed=0).to(cuda_device)
    crit = losses.LpSimCLRLoss(p=2, tau=1.0, simclr_compatibility_mode=True)
    z1, z2 = synth.synth_latents(B, n, "sphere", seed=3)
    z1, z2 = z1.to(cuda_device), z2.to(cuda_device)
    s1 = GraphedTrainStep(f1, g, crit, B, n, lr=1e-3, host_io=False)
    s2 = GraphedTrainStep(f2, g, crit, B, n, lr=1e-3, host_io=False)
    assert s1._pack_weights is not None
    s1(z1, z2)
    with torch.no_grad():                                       # f1 <- f2's initial weights, fresh optimizer state
        for a, b in zip(f1.parameters(), f2.parameters()):
            a.copy_(b)
        for st in s1.optimizer.state.values():
            st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
        for grp in s1.optimizer.param_groups:
            grp["_step_state"].zero_()
    for _ in range(3):
        a = s1(z1, z2).clone()
        b = s2(z1, z2).clone()
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (a, b)
    # Adam's first updates are lr * sign(g): an element whose gradient is rounding noise (atomic summation order) may step
    # the other way, so the parameters are compared in the mean, the loss trajectories above element-wise
    for a, b in zip(f1.parameters(), f2.parameters()):
        assert (a - b).abs().mean().item() <= 1e-5
        assert (a - b).abs().max().item() <= 7e-3


@pytest.mark.parametrize("p,d,B", [(2, 10, 1000), (3, 40, 515), (1, 5, 256), (0, 10, 768), (2.5, 16, 300)])
def test_pairs_form_equals_the_rolled_call(p, d, B, cuda_device):
    """lp_infonce_pairs([z1; z2]) -- anchors streamed as their own negatives, both gradient halves in one tensor -- against
    the reference-shaped call lp_infonce(z1, z2, roll(z1, 1, 0)) of main_mlp.py:270-280 (same negative set, other order)."""
    from clica_b200 import functional as F
    g = torch.Generator().manual_seed(int(10 * p) + d + B)
    z1 = torch.randn(B, d, generator=g)
    z1 = z1 / z1.norm(dim=-1, keepdim=True)
    z2 = z1 + 0.05 * torch.randn(B, d, generator=g)
    ab = torch.cat([z1, z2], 0).to(cuda_device).requires_grad_(True)
    a = z1.to(cuda_device).requires_grad_(True)
    b = z2.to(cuda_device).requires_grad_(True)
    m0, li0, pos0, neg0 = F.lp_infonce(a, b, torch.roll(a, 1, 0), float(p), 0.7, 0.5, True)
    m1, li1, pos1, neg1 = F.lp_infonce_pairs(ab, float(p), 0.7, 0.5, True)
    (3.0 * m0).backward()
    (3.0 * m1).backward()
    assert abs(m0.item() - m1.item()) <= 2e-6 * max(1.0, abs(m0.item()))
    assert torch.allclose(li0, li1, rtol=0, atol=5e-6 * max(1.0, li0.abs().max().item()))
    assert abs(pos0.item() - pos1.item()) <= 1e-6 * max(1.0, abs(pos0.item())) and abs(neg0.item() - neg1.item()) <= 2e-6 * max(1.0, abs(neg0.item()))
    scale = max(a.grad.abs().max().item(), b.grad.abs().max().item())
    assert (ab.grad[:B] - a.grad).abs().max().item() <= 2e-5 * scale
    assert (ab.grad[B:] - b.grad).abs().max().item() <= 2e-5 * scale
