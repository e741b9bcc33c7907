"""GPU (-m gpu): the fused Lp-InfoNCE kernels, called through the C ABI, against the oracle and the golden
vectors generated from the reference.

Tolerances (fp32 path vs fp64 truth).  The reference's OWN fp32 run sits <= 5e-6 (relative to
max(1,|loss|)) from its fp64 run (tests/test_oracle_vs_golden.py::test_fp32_reference_noise_floor); the
CUDA path is held to the same band: per-item loss 5e-6 * max(1, max|loss_i|), gradients 2e-5 * max|grad|
(3e-5 for the generic-exponent kernels, which go through ex2/lg2.approx).
"""
import numpy as np
import pytest
import torch

from conftest import golden_loss_cases, load_golden

pytestmark = pytest.mark.gpu

LOSS_TOL = 5e-6
GRAD_TOL = 2e-5


def _run(z1, z2, z3, p, tau, alpha, compat, dev, gl=None, roll=False):
    from clica_b200 import functional as F
    a = torch.tensor(z1, device=dev, requires_grad=True)
    b = torch.tensor(z2, device=dev, requires_grad=True)
    n = torch.roll(a, 1, 0) if roll else torch.tensor(z3, device=dev, requires_grad=True)
    mean, per_item, pos_mean, neg_mean = F.lp_infonce(a, b, n, p, tau, alpha, compat)
    if gl is None:
        mean.backward()
    else:
        (per_item * torch.tensor(gl, device=dev)).sum().backward()
    out = dict(loss_mean=mean.item(), loss_i=per_item.detach().cpu().numpy(), pos_mean=pos_mean.item(),
               neg_mean=neg_mean.item(), g1=a.grad.cpu().numpy(), g2=b.grad.cpu().numpy())
    if not roll:
        out["g3"] = n.grad.cpu().numpy()
    return out


def _check(out, ref, roll, grad_tol=GRAD_TOL):
    scale = max(1.0, float(np.abs(ref["loss_i"]).max()))
    assert np.abs(out["loss_i"] - ref["loss_i"]).max() <= LOSS_TOL * scale
    assert abs(out["loss_mean"] - ref["loss_mean"]) <= LOSS_TOL * scale
    assert abs(out["pos_mean"] - ref["pos_mean"]) <= LOSS_TOL * max(1.0, abs(ref["pos_mean"]))
    assert abs(out["neg_mean"] - ref["neg_mean"]) <= LOSS_TOL * scale
    gmax = max(float(np.abs(ref["g1"]).max()), 1e-30)
    assert np.abs(out["g1"] - ref["g1"]).max() <= grad_tol * gmax
    assert np.abs(out["g2"] - ref["g2"]).max() <= grad_tol * gmax
    if not roll:
        assert np.abs(out["g3"] - ref["g3"]).max() <= grad_tol * gmax


# "*_cpupin" cases pin the CPU oracle only (e.g. p = 5: an integer exponent on the generic ex2/lg2 kernels, whose GPU
# tolerance has not been measured yet); everything else runs on the GPU
@pytest.mark.parametrize("name", [n for n in golden_loss_cases() if n != "nopow_p2" and not n.endswith("_cpupin")])
def test_golden_vectors_from_the_reference(name, cuda_device):
    g = load_golden("lpnce_" + name)
    roll = bool(g["roll"])
    out = _run(g["z1"], g["z2"], None if roll else g["z3"], float(g["p"]), float(g["tau"]), float(g["alpha"]),
               bool(g["compat"]), cuda_device, gl=g["gl"] if "gl" in g else None, roll=roll)
    ref = {k[:-3]: g[k] for k in g if k.endswith("_64")}
    ref["loss_mean"], ref["pos_mean"], ref["neg_mean"] = float(ref["loss_mean"]), float(ref["pos_mean"]), float(ref["neg_mean"])
    generic = float(g["p"]) not in (1.0, 2.0, 3.0, 4.0)          # real exponents go through ex2/lg2.approx
    _check(out, ref, roll, grad_tol=3e-5 if generic else GRAD_TOL)


def test_pow_false_is_refused_not_approximated(cuda_device):
    """pow=False is outside the CUDA path: the C ABI must say so (CLICA_E_UNSUPPORTED), not compute pow=True."""
    from clica_b200 import _lib
    lib = _lib.load()
    a = torch.randn(8, 4, device=cuda_device)
    out = torch.empty(8 * 5 + 3, device=cuda_device)
    ws = torch.empty(1 << 16, dtype=torch.uint8, device=cuda_device)
    rc = lib.clica_lpnce_fwd(a.data_ptr(), 4, a.data_ptr(), 4, a.data_ptr(), 4, 8, 8, 4, 2.0, 1.0, 0.5, 1, 0,
                             out[16:].data_ptr(), out[24:].data_ptr(), out[32:].data_ptr(), out.data_ptr(), out[40:].data_ptr(),
                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == -2 and b"pow=False" in lib.clica_last_error()
    rc = lib.clica_lpnce_fwd(a.data_ptr(), 4, a.data_ptr(), 4, a.data_ptr(), 4, 8, 8, 4, 0.5, 1.0, 0.5, 1, 1,
                             out[16:].data_ptr(), out[24:].data_ptr(), out[32:].data_ptr(), out.data_ptr(), out[40:].data_ptr(),
                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == -2


@pytest.mark.parametrize("p", [1, 2, 3, 4, 2.5])
@pytest.mark.parametrize("B,M,d", [(300, 517, 10), (129, 128, 3), (257, 1031, 40), (64, 2000, 16), (1000, 999, 33),
                                   (130, 257, 128), (64, 100, 64), (96, 201, 100), (40, 70, 256)])
def test_random_shapes_against_c_oracle(p, B, M, d, cuda_device):
    from oracle import c_oracle
    rng = np.random.RandomState(B * 7 + M + d)
    z1 = rng.randn(B, d).astype(np.float32) * 0.7
    z2 = (z1 + 0.05 * rng.randn(B, d)).astype(np.float32)
    z3 = rng.randn(M, d).astype(np.float32) * 0.7
    z3[: min(B, M) // 2] = z1[: min(B, M) // 2]          # include exact self pairs (distance 0)
    tau = 0.6
    out = _run(z1, z2, z3, float(p), tau, 0.5, True, cuda_device)
    ref = c_oracle.lpnce(z1, z2, z3, p, tau, 0.5, include_pos=True)
    _check(out, ref, roll=False, grad_tol=GRAD_TOL if p != 2.5 else 3e-5)


@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("B,d", [(384, 128), (200, 10), (100, 72)])
def test_rolled_negatives_use_the_merged_backward(p, B, d, cuda_device):
    """z3 = torch.roll(z1, 1, 0) (main_mlp.py:272): autograd-detected, single merged backward pass; also the
    feature-split kernels (d > 40: 2 / 4 / 8 lanes per pair) of BASELINE's sweep config (d = 128)."""
    from clica_b200 import functional as F
    from oracle import c_oracle
    rng = np.random.RandomState(B + d + p)
    z1 = (rng.randn(B, d) * 0.4).astype(np.float32)
    z2 = (z1 + 0.05 * rng.randn(B, d)).astype(np.float32)
    a = torch.tensor(z1, device=cuda_device, requires_grad=True)
    b = torch.tensor(z2, device=cuda_device, requires_grad=True)
    n = torch.roll(a, 1, 0)
    assert F.is_row_roll_of(n, a)
    out = _run(z1, z2, None, float(p), 0.9, 0.5, True, cuda_device, roll=True)
    ref = c_oracle.lpnce(z1, z2, np.roll(z1, 1, 0), p, 0.9, 0.5, include_pos=True)
    ref["g1"] = ref["g1"] + np.roll(ref["g3"], -1, 0)
    _check(out, ref, roll=True)


def test_strided_views_and_none_arguments(cuda_device):
    """kitti_masks/solver.py:60-75 feeds mu[::2], mu[1::2]; main_3dident passes None for the first three args."""
    import sys
    import clica_b200
    from oracle import c_oracle
    sys.path.insert(0, clica_b200.DROPIN_DIR)
    import losses
    rng = np.random.RandomState(3)
    mu = torch.tensor(rng.randn(400, 10).astype(np.float32), device=cuda_device, requires_grad=True)
    a, b = mu[::2], mu[1::2]
    crit = losses.LpSimCLRLoss(p=1, tau=1.0, simclr_compatibility_mode=True, pow=True)
    mean, per_item, parts = crit(None, None, None, a, b, torch.roll(a, 1, 0))
    mean.backward()
    mu_np = mu.detach().cpu().numpy()
    ref = c_oracle.lpnce(mu_np[::2], mu_np[1::2], np.roll(mu_np[::2], 1, 0), 1, 1.0)
    assert abs(mean.item() - ref["loss_mean"]) <= LOSS_TOL * max(1.0, abs(ref["loss_mean"]))
    g = np.zeros_like(mu_np)
    g[::2] = ref["g1"] + np.roll(ref["g3"], -1, 0)
    g[1::2] = ref["g2"]
    assert np.abs(mu.grad.cpu().numpy() - g).max() <= GRAD_TOL * np.abs(g).max()
    assert isinstance(parts, list) and len(parts) == 2 and parts[0].dim() == 0 and per_item.shape == (200,)


def test_full_size_properties_config2(cuda_device):
    """BASELINE config 2 size (B=6144, d=10, p=2): size-independent properties + sampled rows vs the oracle."""
    from clica_b200 import functional as F
    from oracle import c_oracle
    B, d, p, tau = 6144, 10, 2.0, 1.0
    gen = torch.Generator(device="cpu").manual_seed(0)
    z1 = torch.randn(B, d, generator=gen)
    z1 = z1 / z1.norm(dim=-1, keepdim=True)
    z2 = z1 + 0.05 * torch.randn(B, d, generator=gen)
    a = z1.to(cuda_device).requires_grad_(True)
    b = z2.to(cuda_device).requires_grad_(True)
    mean, per_item, _, _ = F.lp_infonce(a, b, torch.roll(a, 1, 0), p, tau, 0.5, True)
    mean.backward()
    # (1) sampled anchors against the full negative set, fp64 oracle
    rows = np.arange(0, B, 97)
    z1n, z2n = z1.numpy(), z2.numpy()
    ref = c_oracle.lpnce(z1n[rows], z2n[rows], np.roll(z1n, 1, 0), p, tau, need_grad=False)
    assert np.abs(per_item.detach().cpu().numpy()[rows] - ref["loss_i"]).max() <= LOSS_TOL * max(1.0, np.abs(ref["loss_i"]).max())
    # (2) permuting the negatives leaves every per-item loss unchanged
    perm = torch.randperm(B, generator=gen).to(cuda_device)
    _, per_item_perm, _, _ = F.lp_infonce(a.detach(), b.detach(), torch.roll(a.detach(), 1, 0)[perm], p, tau, 0.5, True)
    assert (per_item_perm - per_item.detach()).abs().max().item() <= 2e-6 * max(1.0, per_item.abs().max().item())
    # (3) translation invariance
    shift = torch.full((1, d), 0.25, device=cuda_device)
    _, per_item_shift, _, _ = F.lp_infonce(a.detach() + shift, b.detach() + shift, torch.roll(a.detach(), 1, 0) + shift, p, tau, 0.5, True)
    assert (per_item_shift - per_item.detach()).abs().max().item() <= 2e-5
    # (4) gradients sum to ~0 over all rows (the loss depends on differences only)
    gsum = (a.grad + b.grad).sum(0).abs().max().item()
    assert gsum <= 1e-5
    # (5) KA1: all outputs identical -> loss_i = ln(B+1), zero gradient
    c = torch.ones(B, d, device=cuda_device, requires_grad=True)
    m2, li2, _, _ = F.lp_infonce(c, c.detach().clone(), c.detach().clone(), 3.0, 0.3, 0.5, True)
    m2.backward()
    assert abs(m2.item() - np.log(B + 1.0)) <= 1e-5 and c.grad.abs().max().item() == 0.0


def test_full_size_config3_sampled_rows(cuda_device):
    """BASELINE config 3 size (B=8192, d=40, p=3): sampled anchors (loss + anchor-role gradient) vs the oracle."""
    from clica_b200 import functional as F
    from oracle import c_oracle
    B, d, p, tau = 8192, 40, 3.0, 1.0
    rng = np.random.RandomState(1)
    z1 = (rng.randn(B, d) * 0.5).astype(np.float32)
    z2 = (z1 + 0.05 * rng.randn(B, d)).astype(np.float32)
    z3 = np.roll(z1, 1, 0).copy()
    a = torch.tensor(z1, device=cuda_device, requires_grad=True)
    b = torch.tensor(z2, device=cuda_device, requires_grad=True)
    n = torch.tensor(z3, device=cuda_device, requires_grad=True)
    mean, per_item, _, _ = F.lp_infonce(a, b, n, p, tau, 0.5, True)
    mean.backward()
    rows = np.arange(5, B, 331)
    ref = c_oracle.lpnce(z1[rows], z2[rows], z3, p, tau, gl=np.full(len(rows), 1.0 / B))
    assert np.abs(per_item.detach().cpu().numpy()[rows] - ref["loss_i"]).max() <= LOSS_TOL * max(1.0, np.abs(ref["loss_i"]).max())
    gmax = np.abs(ref["g1"]).max()
    assert np.abs(a.grad.cpu().numpy()[rows] - ref["g1"]).max() <= GRAD_TOL * gmax
    assert np.abs(b.grad.cpu().numpy()[rows] - ref["g2"]).max() <= GRAD_TOL * gmax
    assert torch.isfinite(n.grad).all()


def test_sharded_backward_equals_single_device(cuda_device):
    """Row-sharded formulation: per-shard fwd + gathered lse + clica_lpnce_bwd_sharded must
    reproduce the single-device gradient of the global mean loss with z3 = roll(z1)."""
    from clica_b200 import functional as F
    from clica_b200 import sharded
    B, d, p, tau, W = 1024, 10, 2.0, 0.7, 4
    rng = np.random.RandomState(11)
    z1 = torch.tensor(rng.randn(B, d).astype(np.float32), device=cuda_device)
    z2 = z1 + 0.05 * torch.tensor(rng.randn(B, d).astype(np.float32), device=cuda_device)
    a = z1.clone().requires_grad_(True)
    b = z2.clone().requires_grad_(True)
    mean, per_item, _, _ = F.lp_infonce(a, b, torch.roll(a, 1, 0), p, tau, 0.5, True)
    mean.backward()
    Bl = B // W
    stat_parts, pos_parts, loss_parts = [], [], []
    for r in range(W):
        li, lse, pos, stat = sharded.local_forward(z1[r * Bl:(r + 1) * Bl], z2[r * Bl:(r + 1) * Bl], z1, p, tau, 0.5, True)
        stat_parts.append(stat.clone()), pos_parts.append(pos.clone()), loss_parts.append(li.clone())
    lse_all = torch.cat(stat_parts)
    assert (torch.cat(loss_parts) - per_item.detach()).abs().max().item() <= 2e-6 * max(1.0, per_item.abs().max().item())
    for r in range(W):
        g1, g2 = sharded.local_backward(z1[r * Bl:(r + 1) * Bl], z2[r * Bl:(r + 1) * Bl], z1, lse_all, pos_parts[r],
                                        r * Bl, p, tau, 0.5, True)
        gmax = a.grad.abs().max().item()
        assert (g1 - a.grad[r * Bl:(r + 1) * Bl]).abs().max().item() <= 2e-5 * gmax
        assert (g2 - b.grad[r * Bl:(r + 1) * Bl]).abs().max().item() <= 2e-5 * gmax


@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("roll", [False, True])
def test_negative_and_signed_upstream_gradients(p, roll, cuda_device):
    """(-loss).backward() and per-item weights of mixed sign: the backward coefficient E = 2 gl (1-alpha)/tau is then
    negative, which the p = 1 gradient (w * sign(t)) must carry through (copysignf(w, t) dropped it)."""
    from clica_b200 import functional as F
    from oracle import c_oracle
    B, M, d, tau = 260, 260 if roll else 333, 10, 0.8
    rng = np.random.RandomState(100 + p + 10 * roll)
    z1 = (rng.randn(B, d) * 0.6).astype(np.float32)
    z2 = (z1 + 0.05 * rng.randn(B, d)).astype(np.float32)
    z3 = np.roll(z1, 1, 0).copy() if roll else (rng.randn(M, d) * 0.6).astype(np.float32)
    gl = rng.randn(B).astype(np.float32) / B                       # mixed signs
    ref_w = c_oracle.lpnce(z1, z2, z3, p, tau, 0.5, include_pos=True, gl=gl)
    ref_m = c_oracle.lpnce(z1, z2, z3, p, tau, 0.5, include_pos=True)
    for mode in ("neg_mean", "weighted"):
        a = torch.tensor(z1, device=cuda_device, requires_grad=True)
        b = torch.tensor(z2, device=cuda_device, requires_grad=True)
        n = torch.roll(a, 1, 0) if roll else torch.tensor(z3, device=cuda_device, requires_grad=True)
        mean, per_item, _, _ = F.lp_infonce(a, b, n, float(p), tau, 0.5, True)
        if mode == "neg_mean":
            (-mean).backward()
            ref, sgn = ref_m, -1.0
        else:
            (per_item * torch.tensor(gl, device=cuda_device)).sum().backward()
            ref, sgn = ref_w, 1.0
        g1 = sgn * ref["g1"]
        if roll:
            g1 = g1 + sgn * np.roll(ref["g3"], -1, 0)
        gmax = float(np.abs(g1).max())
        assert np.abs(a.grad.cpu().numpy() - g1).max() <= GRAD_TOL * gmax, (mode, p, roll)
        assert np.abs(b.grad.cpu().numpy() - sgn * ref["g2"]).max() <= GRAD_TOL * gmax
        if not roll:
            assert np.abs(n.grad.cpu().numpy() - sgn * ref["g3"]).max() <= GRAD_TOL * gmax


@pytest.mark.parametrize("p", [1, 2, 3])
def test_sharded_backward_with_a_negative_scale(p, cuda_device):
    """clica_lpnce_bwd_sharded with g_scale < 0 (the merged role weights w by E_i and E_j, both negative here)."""
    from clica_b200 import sharded
    from oracle import c_oracle
    B, d, tau, W = 512, 10, 0.9, 2
    rng = np.random.RandomState(50 + p)
    z1n = (rng.randn(B, d) * 0.5).astype(np.float32)
    z2n = (z1n + 0.05 * rng.randn(B, d)).astype(np.float32)
    ref = c_oracle.lpnce(z1n, z2n, z1n, p, tau, 0.5, include_pos=True)       # z3 = all anchors (gathered)
    g1_ref = -(ref["g1"] + ref["g3"])
    z1, z2 = torch.tensor(z1n, device=cuda_device), torch.tensor(z2n, device=cuda_device)
    Bl = B // W
    stats, poss = [], []
    for r in range(W):
        _, _, pos, stat = sharded.local_forward(z1[r * Bl:(r + 1) * Bl], z2[r * Bl:(r + 1) * Bl], z1, float(p), tau, 0.5, True)
        stats.append(stat.clone()), poss.append(pos.clone())
    stat_all = torch.cat(stats)
    scale = torch.tensor(-1.0, device=cuda_device)
    gmax = float(np.abs(g1_ref).max())
    for r in range(W):
        g1, g2 = sharded.local_backward(z1[r * Bl:(r + 1) * Bl], z2[r * Bl:(r + 1) * Bl], z1, stat_all, poss[r],
                                        r * Bl, float(p), tau, 0.5, True, g_scale=scale)
        assert np.abs(g1.cpu().numpy() - g1_ref[r * Bl:(r + 1) * Bl]).max() <= GRAD_TOL * gmax
        assert np.abs(g2.cpu().numpy() + ref["g2"][r * Bl:(r + 1) * Bl]).max() <= GRAD_TOL * gmax


def test_integer_exponent_on_the_generic_kernels_p5(cuda_device):
    """`--p 5` is reachable from the CLI (main_mlp.py:111-116, int): it runs on the generic ex2/lg2 kernels."""
    g = load_golden("lpnce_roll_p5_generic_cpupin")
    roll = bool(g["roll"])
    out = _run(g["z1"], g["z2"], None if roll else g["z3"], float(g["p"]), float(g["tau"]), float(g["alpha"]),
               bool(g["compat"]), cuda_device, gl=g["gl"] if "gl" in g else None, roll=roll)
    ref = {k[:-3]: g[k] for k in g if k.endswith("_64")}
    ref["loss_mean"], ref["pos_mean"], ref["neg_mean"] = float(ref["loss_mean"]), float(ref["pos_mean"]), float(ref["neg_mean"])
    _check(out, ref, roll, grad_tol=3e-5)
