"""GPU (-m gpu), opt-in: kernels written after round 1's GPU budget was spent.  They are off by default in the
product path and these tests are skipped unless CLICA_EXPERIMENTAL=1, so that the regular GPU suite only contains
code that has been run on a B200."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CLICA_EXPERIMENTAL", "0") != "1",
                                 reason="experimental kernels: set CLICA_EXPERIMENTAL=1")]


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


# ---- fast forward of the fused loss (CLICA_LPNCE_FAST=1: fixed reference point, underflow fallback in finalize) ----
def _loss_run(z1, z2, z3, p, tau, compat, dev, roll):
    from clica_b200 import functional as F
    a = torch.tensor(z1, device=dev, requires_grad=True)
    b = torch.tensor(z2, device=dev, requires_grad=True)
    n = torch.roll(a, 1, 0) if roll else torch.tensor(z3, device=dev, requires_grad=True)
    mean, per_item, pos_mean, neg_mean = F.lp_infonce(a, b, n, p, tau, 0.5, compat)
    mean.backward()
    return mean.item(), per_item.detach().cpu().numpy(), a.grad.cpu().numpy(), b.grad.cpu().numpy()


def test_fast_forward_on_the_golden_vectors(cuda_device, monkeypatch):
    import numpy as np
    from conftest import golden_loss_cases, load_golden
    monkeypatch.setenv("CLICA_LPNCE_FAST", "1")
    for name in golden_loss_cases():
        g = load_golden("lpnce_" + name)
        if not bool(g["pow"]) or "gl" in g or name.endswith("_cpupin"):
            continue
        roll = bool(g["roll"])
        mean, li, g1, g2 = _loss_run(g["z1"], g["z2"], None if roll else g["z3"], float(g["p"]), float(g["tau"]),
                                     bool(g["compat"]), cuda_device, roll) if float(g["alpha"]) == 0.5 else (None,) * 4
        if mean is None:
            continue
        scale = max(1.0, float(np.abs(g["loss_i_64"]).max()))
        assert np.abs(li - g["loss_i_64"]).max() <= 5e-6 * scale, name      # includes the underflowing rows (fallback)
        gmax = max(float(np.abs(g["g1_64"]).max()), 1e-30)
        tol = 2e-5 if float(g["p"]) in (1.0, 2.0, 3.0, 4.0) else 3e-5
        assert np.abs(g1 - g["g1_64"]).max() <= tol * gmax, name
        assert np.abs(g2 - g["g2_64"]).max() <= tol * gmax, name


@pytest.mark.parametrize("p", [1, 2, 3, 2.5])
@pytest.mark.parametrize("B,M,d", [(300, 517, 10), (257, 1031, 40), (130, 257, 128), (6144, 6144, 10)])
def test_fast_forward_matches_the_default_forward(p, B, M, d, cuda_device, monkeypatch):
    import numpy as np
    rng = np.random.RandomState(B + M + d)
    z1 = rng.randn(B, d).astype(np.float32) * 0.7
    z2 = (z1 + 0.05 * rng.randn(B, d)).astype(np.float32)
    z3 = rng.randn(M, d).astype(np.float32) * 0.7
    monkeypatch.setenv("CLICA_LPNCE_FAST", "0")
    ref = _loss_run(z1, z2, z3, float(p), 0.6, True, cuda_device, False)
    monkeypatch.setenv("CLICA_LPNCE_FAST", "1")
    out = _loss_run(z1, z2, z3, float(p), 0.6, True, cuda_device, False)
    scale = max(1.0, float(np.abs(ref[1]).max()))
    assert abs(out[0] - ref[0]) <= 2e-6 * scale
    assert np.abs(out[1] - ref[1]).max() <= 2e-6 * scale
    gmax = float(np.abs(ref[2]).max())
    assert np.abs(out[2] - ref[2]).max() <= 1e-5 * gmax and np.abs(out[3] - ref[3]).max() <= 1e-5 * gmax


def test_fast_forward_underflow_fallback(cuda_device, monkeypatch):
    """Every logit of every row far below 2^-126 without the positive pair in the denominator (non-compat mode): the
    fast kernel's sums flush to zero and the finalize kernel must recompute the rows; compare with the fp64 oracle."""
    import numpy as np
    from oracle import c_oracle
    rng = np.random.RandomState(11)
    z1 = rng.randn(70, 6).astype(np.float32)
    z2 = (z1 + 0.05 * rng.randn(70, 6)).astype(np.float32)
    z3 = (rng.randn(90, 6) * 2 + 9).astype(np.float32)
    monkeypatch.setenv("CLICA_LPNCE_FAST", "1")
    mean, li, g1, g2 = _loss_run(z1, z2, z3, 2.0, 0.05, False, cuda_device, False)
    ref = c_oracle.lpnce(z1, z2, z3, 2, 0.05, 0.5, include_pos=False)
    scale = max(1.0, float(np.abs(ref["loss_i"]).max()))
    assert np.isfinite(li).all()
    assert np.abs(li - ref["loss_i"]).max() <= 5e-6 * scale
    assert np.abs(g1 - ref["g1"]).max() <= 2e-5 * max(float(np.abs(ref["g1"]).max()), 1e-30)


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
