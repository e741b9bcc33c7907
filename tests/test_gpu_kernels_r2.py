"""GPU (-m gpu): the fused mixing-net kernel, the one-launch weight packing and the two forms of the p = 2 loss
kernels (dot form on centred data vs subtract-then-square; 2 or 4 owner rows per thread)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("n,L,M", [(10, 3, 12288), (5, 3, 1000), (40, 3, 4099), (16, 1, 77), (7, 4, 513)])
def test_fused_mixing_matches_the_torch_modules(n, L, M, cuda_device):
    """clica_mixing_fwd vs the nn.Sequential shape of invertible_network_utils.py:87-123 run by torch (fp32)."""
    from clica_b200 import functional as F
    from clica_b200 import synth
    g = synth.build_mixing(n, L, seed=n + L).to(cuda_device)
    plan = F.mixing_plan(g)
    assert plan is not None and len(plan[0]) == L and plan[1] == 0.2
    x = torch.randn(M, n, device=cuda_device)
    y = F.mixing_forward(x, *plan)
    ref32 = g(x)                                           # torch: cuBLAS fp32 + elementwise LeakyReLU
    ref64 = x.double()
    for i, W in enumerate(plan[0]):
        ref64 = ref64 @ W.double().t()
        if i != L - 1:
            ref64 = torch.nn.functional.leaky_relu(ref64, 0.2)
    scale = ref64.abs().max().item()
    assert (y.double() - ref64).abs().max().item() <= 2e-6 * scale
    assert (y - ref32).abs().max().item() <= 4e-6 * scale


def test_graphed_step_with_fused_mixing(cuda_device, monkeypatch):
    import copy
    import sys
    import clica_b200
    from clica_b200 import synth
    from clica_b200.graphed import GraphedTrainStep
    sys.path.insert(0, clica_b200.DROPIN_DIR)
    import encoders
    import losses
    n, B = 10, 512
    torch.manual_seed(0)
    f1 = encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(cuda_device)
    f2 = copy.deepcopy(f1)
    g = synth.build_mixing(n, 3, seed=0).to(cuda_device)
    crit = losses.LpSimCLRLoss(p=2, tau=1.0, simclr_compatibility_mode=True)
    z1, z2 = synth.synth_latents(B, n, "sphere", seed=3)
    z1, z2 = z1.to(cuda_device), z2.to(cuda_device)
    monkeypatch.setenv("CLICA_FUSED_MIXING", "1")
    fused = GraphedTrainStep(f1, g, crit, B, n, lr=1e-3, host_io=False)
    assert fused._mix is not None
    monkeypatch.setenv("CLICA_FUSED_MIXING", "0")
    plain = GraphedTrainStep(f2, g, crit, B, n, lr=1e-3, host_io=False)
    assert plain._mix is None and fused.launches_per_replay == plain.launches_per_replay + 1
    for _ in range(3):
        a = fused(z1, z2).clone()
        b = plain(z1, z2).clone()
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), (a, b)


# ---- p = 2: dot form (|a'|^2/2 + |b'|^2/2 - a'.b' on centred rows) against the subtract-then-square form ---------
def _loss_run(z1, z2, z3, p, tau, compat, dev, roll, gl=None):
    from clica_b200 import functional as F
    a = torch.tensor(z1, device=dev, requires_grad=True)
    b = torch.tensor(z2, device=dev, requires_grad=True)
    n = torch.roll(a, 1, 0) if roll else torch.tensor(z3, device=dev, requires_grad=True)
    mean, per_item, pos_mean, neg_mean = F.lp_infonce(a, b, n, p, tau, 0.5, compat)
    if gl is None:
        mean.backward()
    else:
        (per_item * torch.tensor(gl, device=dev)).sum().backward()
    g3 = None if roll else n.grad.cpu().numpy()
    return mean.item(), per_item.detach().cpu().numpy(), a.grad.cpu().numpy(), b.grad.cpu().numpy(), g3


def _data(kind, B, M, d, rng):
    z1 = rng.randn(B, d).astype(np.float32)
    z3 = rng.randn(M, d).astype(np.float32)
    if kind == "sphere":                      # a trained encoder's outputs: unit sphere (dot form everywhere)
        z1 /= np.linalg.norm(z1, axis=1, keepdims=True)
        z3 /= np.linalg.norm(z3, axis=1, keepdims=True)
    elif kind == "init":                      # an untrained encoder: tiny spread around a common offset
        z1 = (0.3 + 0.02 * z1).astype(np.float32)
        z3 = (0.3 + 0.02 * z3).astype(np.float32)
    elif kind == "mixed":                     # a few far rows: some (warp, tile) pairs must leave the dot form
        z1 *= 0.4
        z3 *= 0.4
        z1[::37] *= 9.0
        z3[5::53] *= 9.0
    z2 = (z1 + 0.05 * rng.randn(B, d)).astype(np.float32)
    return z1, z2, z3


@pytest.mark.parametrize("kind", ["sphere", "init", "mixed"])
@pytest.mark.parametrize("B,M,d", [(300, 517, 10), (257, 1031, 40), (129, 128, 3), (1000, 999, 7), (64, 2000, 16)])
@pytest.mark.parametrize("r4", ["0", "1"])
def test_p2_dot_form_against_the_oracle_and_the_subtract_form(kind, B, M, d, r4, cuda_device, monkeypatch):
    from oracle import c_oracle
    rng = np.random.RandomState(B + M + d)
    z1, z2, z3 = _data(kind, B, M, d, rng)
    z3[: min(B, M) // 3] = z1[: min(B, M) // 3]          # exact self pairs (distance 0)
    tau = 0.8
    monkeypatch.setenv("CLICA_LPNCE_R4", r4)
    monkeypatch.setenv("CLICA_LPNCE_DOT", "1")
    monkeypatch.setenv("CLICA_LPNCE_DOT_BWD", "1")        # the backward's dot form is opt-in (slower on B200): test it too
    dot = _loss_run(z1, z2, z3, 2.0, tau, True, cuda_device, False)
    monkeypatch.setenv("CLICA_LPNCE_DOT", "0")
    sub = _loss_run(z1, z2, z3, 2.0, tau, True, cuda_device, False)
    ref = c_oracle.lpnce(z1, z2, z3, 2, tau, 0.5, include_pos=True)
    scale = max(1.0, float(np.abs(ref["loss_i"]).max()))
    gmax = max(float(np.abs(ref["g1"]).max()), float(np.abs(ref["g3"]).max()), 1e-30)
    for out in (dot, sub):
        assert abs(out[0] - ref["loss_mean"]) <= 5e-6 * scale
        assert np.abs(out[1] - ref["loss_i"]).max() <= 5e-6 * scale
        assert np.abs(out[2] - ref["g1"]).max() <= 2e-5 * gmax
        assert np.abs(out[3] - ref["g2"]).max() <= 2e-5 * gmax
        assert np.abs(out[4] - ref["g3"]).max() <= 2e-5 * gmax


@pytest.mark.parametrize("kind", ["sphere", "init"])
def test_p2_dot_form_rolled_negatives_full_size(kind, cuda_device, monkeypatch):
    """BASELINE config 2 size with z3 = roll(z1) (merged backward): dot form vs subtract form vs sampled oracle rows."""
    from oracle import c_oracle
    B, d, tau = 6144, 10, 1.0
    rng = np.random.RandomState(7)
    z1, z2, _ = _data(kind, B, B, d, rng)
    monkeypatch.setenv("CLICA_LPNCE_DOT", "1")
    monkeypatch.setenv("CLICA_LPNCE_DOT_BWD", "1")
    dot = _loss_run(z1, z2, None, 2.0, tau, True, cuda_device, True)
    monkeypatch.setenv("CLICA_LPNCE_DOT", "0")
    sub = _loss_run(z1, z2, None, 2.0, tau, True, cuda_device, True)
    rows = np.arange(3, B, 211)
    z3 = np.roll(z1, 1, 0)
    ref = c_oracle.lpnce(z1[rows], z2[rows], z3, 2, tau, need_grad=False)
    scale = max(1.0, float(np.abs(ref["loss_i"]).max()))
    assert np.abs(dot[1][rows] - ref["loss_i"]).max() <= 5e-6 * scale
    assert np.abs(sub[1][rows] - ref["loss_i"]).max() <= 5e-6 * scale
    gmax = float(np.abs(sub[2]).max())
    assert np.abs(dot[2] - sub[2]).max() <= 2e-5 * gmax and np.abs(dot[3] - sub[3]).max() <= 2e-5 * gmax


def test_loss_workspace_counters_are_left_zero(cuda_device):
    """The in-kernel split merge keeps arrival counters at the head of the workspace and must restore the zeros."""
    from clica_b200 import functional as F
    rng = np.random.RandomState(0)
    for B, d, p in [(700, 10, 2.0), (300, 40, 3.0), (130, 128, 1.0)]:
        z1 = torch.tensor(rng.randn(B, d).astype(np.float32), device=cuda_device, requires_grad=True)
        z2 = (z1.detach() + 0.05).requires_grad_(True)
        for sym in (True, False):
            z3 = torch.roll(z1, 1, 0) if sym else z1.detach().flip(0).clone().requires_grad_(True)
            mean, li, _, _ = F.lp_infonce(z1, z2, z3, p, 1.0, 0.5, True)
            mean.backward()
    torch.cuda.synchronize()
    for key, buf in F._workspaces.items():
        if key[2] in ("lpnce", "lpnce_bwd"):
            assert int(buf[:65536].to(torch.int32).abs().sum().item()) == 0, key


def test_fused_weight_packing_is_bit_identical(cuda_device, monkeypatch):
    """CLICA_PACK_FUSED=1 packs the five hidden weight matrices in one launch: same arithmetic, same planes."""
    from clica_b200 import functional as F
    n, M = 10, 777
    widths = [n, 10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n, n]
    g = torch.Generator().manual_seed(5)
    Ws = [((torch.rand(widths[i + 1], widths[i], generator=g) * 2 - 1) / widths[i] ** 0.5).to(cuda_device) for i in range(7)]
    bs = [((torch.rand(widths[i + 1], generator=g) * 2 - 1) / widths[i] ** 0.5).to(cuda_device) for i in range(7)]
    x = torch.randn(M, n, generator=g).to(cuda_device)
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("CLICA_PACK_FUSED", flag)
        F.invalidate_packed_weights()
        launches0 = F._lib.load().clica_launch_count(6)
        outs.append(F.mlp_forward(x, Ws, bs, slope=0.01, mode=0).clone())
        outs.append(F._lib.load().clica_launch_count(6) - launches0)
    assert torch.equal(outs[0], outs[2])
    assert outs[1] == 5 and outs[3] == 1          # five split launches -> one


# ---- SimCLRLoss (losses.py:162-202): the kernel family in its dot-product-similarity form ----------------------
def _simclr_cases():
    import glob
    from conftest import GOLDEN_DIR
    return sorted(os.path.basename(p)[len("simclr_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "simclr_*.npz")))


@pytest.mark.parametrize("name", _simclr_cases())
def test_simclr_loss_on_the_golden_vectors_from_the_reference(name, cuda_device):
    import sys
    import clica_b200
    from conftest import load_golden
    sys.path.insert(0, clica_b200.DROPIN_DIR)
    import losses
    g = load_golden("simclr_" + name)
    roll = bool(g["roll"])
    a = torch.tensor(g["z1"], device=cuda_device, requires_grad=True)
    b = torch.tensor(g["z2"], device=cuda_device, requires_grad=True)
    n = torch.roll(a, 1, 0) if roll else torch.tensor(g["z3"], device=cuda_device, requires_grad=True)
    crit = losses.SimCLRLoss(normalize=bool(g["normalize"]), tau=float(g["tau"]), alpha=float(g["alpha"]))
    mean, per_item, parts = crit(None, None, None, a, b, n)
    if "gl" in g:
        (per_item * torch.tensor(g["gl"], device=cuda_device)).sum().backward()
    else:
        mean.backward()
    # tolerance: the reference's own fp32 run defines the achievable band (logits of +-1e3 in `indep_large_logits`
    # carry 1e-4 of absolute fp32 rounding); held to max(3x that, the Lp kernels' 5e-6 / 2e-5)
    # loss_i = 2 (alpha * (-pos/tau) + (1 - alpha) * lse) is a difference of two terms of size |pos| / tau: fp32 rounding
    # of the TERMS bounds what any fp32 evaluation order can promise (indep_large_logits: terms ~ 1.4e3, loss ~ 0)
    term_l = float(np.abs((g["z1"].astype(np.float64) * g["z2"]).sum(1)).max()) / float(g["tau"]) if not bool(g["normalize"]) else 1.0 / float(g["tau"])
    scale = max(1.0, float(np.abs(g["loss_i_64"]).max()), term_l)
    ref_err = float(np.abs(g["loss_i_32"] - g["loss_i_64"]).max())
    assert np.abs(per_item.detach().cpu().numpy() - g["loss_i_64"]).max() <= max(5e-6 * scale, 3 * ref_err)
    assert abs(mean.item() - float(g["loss_mean_64"])) <= max(5e-6 * scale, 3 * ref_err)
    assert abs(parts[0].item() - float(g["pos_mean_64"])) <= 5e-6 * max(1.0, abs(float(g["pos_mean_64"])))
    assert abs(parts[1].item() - float(g["neg_mean_64"])) <= max(5e-6 * scale, 3 * ref_err)
    term = float(np.abs(g["z1"]).max()) / (len(g["z1"]) * float(g["tau"]))
    gmax = max(float(np.abs(g["g1_64"]).max()), 1e-30)
    gtol = max(2e-5 * gmax, 3 * float(np.abs(g["g1_32"] - g["g1_64"]).max()), 2e-6 * term)
    assert np.abs(a.grad.cpu().numpy() - g["g1_64"]).max() <= gtol
    assert np.abs(b.grad.cpu().numpy() - g["g2_64"]).max() <= gtol
    if not roll:
        assert np.abs(n.grad.cpu().numpy() - g["g3_64"]).max() <= gtol


@pytest.mark.parametrize("B,M,d,tau", [(300, 517, 10, 1.0), (257, 1031, 40, 0.5), (6144, 6144, 10, 1.0), (130, 257, 128, 2.0)])
def test_simclr_random_shapes_against_the_oracle(B, M, d, tau, cuda_device):
    from clica_b200 import functional as F
    from oracle import simclr_oracle
    rng = np.random.RandomState(B + d)
    z1 = rng.randn(B, d).astype(np.float32)
    z1 /= np.linalg.norm(z1, axis=1, keepdims=True)
    z2 = (z1 + 0.05 * rng.randn(B, d)).astype(np.float32)
    z3 = rng.randn(M, d).astype(np.float32)
    z3 /= np.linalg.norm(z3, axis=1, keepdims=True)
    a = torch.tensor(z1, device=cuda_device, requires_grad=True)
    b = torch.tensor(z2, device=cuda_device, requires_grad=True)
    n = torch.tensor(z3, device=cuda_device, requires_grad=True)
    mean, per_item, pos_mean, neg_mean = F.lp_infonce(a, b, n, 0.0, tau, 0.5, True)
    mean.backward()
    rows = np.arange(0, B, max(1, B // 300))
    ref = simclr_oracle.simclr(z1, z2, z3, tau, 0.5)
    scale = max(1.0, float(np.abs(ref["loss_i"]).max()))
    assert np.abs(per_item.detach().cpu().numpy()[rows] - ref["loss_i"][rows]).max() <= 5e-6 * scale
    assert abs(mean.item() - ref["loss_mean"]) <= 5e-6 * scale
    gmax = float(np.abs(ref["g1"]).max())
    assert np.abs(a.grad.cpu().numpy() - ref["g1"]).max() <= 2e-5 * gmax
    assert np.abs(b.grad.cpu().numpy() - ref["g2"]).max() <= 2e-5 * gmax
    assert np.abs(n.grad.cpu().numpy() - ref["g3"]).max() <= 2e-5 * gmax


def test_p_below_one_runs_through_the_reference_delegation_on_cuda(cuda_device):
    """losses.py:433-442 (p < 1, unreachable from the CLIs): handed to the reference's own torch code on CUDA tensors --
    needs baseline/_ref on the box (placed by build())."""
    import sys
    import warnings
    import clica_b200
    from clica_b200 import vendor
    if vendor.vendored_dir() is None:
        pytest.skip("baseline/_ref is absent")
    sys.path.insert(0, clica_b200.DROPIN_DIR)
    import losses
    rng = np.random.RandomState(2)
    a = torch.tensor(rng.randn(64, 6).astype(np.float32), device=cuda_device, requires_grad=True)
    b = (a.detach() + 0.05).requires_grad_(True)
    crit = losses.LpSimCLRLoss(p=0.5, tau=1.0, simclr_compatibility_mode=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean, per_item, parts = crit(None, None, None, a, b, torch.roll(a, 1, 0))
    mean.backward()
    assert torch.isfinite(mean) and per_item.shape == (64,) and torch.isfinite(a.grad).all()
