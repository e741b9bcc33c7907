"""GPU (-m gpu): FusedAdam (host and device step count) and the CUDA-graph training step against eager references.

Reference behaviour: torch.optim.Adam as main_mlp.py:312 builds it, and the train_step body main_mlp.py:258-285.
"""
import copy
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _modules():
    import clica_b200
    if clica_b200.DROPIN_DIR not in sys.path:
        sys.path.insert(0, clica_b200.DROPIN_DIR)
    import encoders
    import losses
    return encoders, losses


def _mixing(n, dev, seed=0):
    from clica_b200 import synth
    return synth.build_mixing(n, 3, seed=seed).to(dev)


def _encoder(n, dev, seed=0):
    encoders, _ = _modules()
    torch.manual_seed(seed)
    return encoders.get_mlp(n, n, [10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n]).to(dev)


def _latents(B, n, dev, seed=1):
    g = torch.Generator().manual_seed(seed)
    z1 = torch.randn(B, n, generator=g)
    z1 = z1 / z1.norm(dim=-1, keepdim=True)
    z2 = z1 + 0.05 * torch.randn(B, n, generator=g)
    return z1, z2


@pytest.mark.parametrize("capturable", [False, True])
def test_fused_adam_matches_torch_adam(cuda_device, capturable):
    from clica_b200.optim import FusedAdam
    torch.manual_seed(0)
    shapes = [(37, 5), (37,), (128, 64), (1,), (1000, 3)]
    p_ref = [torch.randn(s, device=cuda_device).requires_grad_(True) for s in shapes]
    p_our = [p.detach().clone().requires_grad_(True) for p in p_ref]
    o_ref = torch.optim.Adam(p_ref, lr=1e-2)
    o_our = FusedAdam(p_our, lr=1e-2, capturable=capturable)
    for step in range(5):
        for a, b in zip(p_ref, p_our):
            gr = torch.randn_like(a) * (10.0 ** (step - 2))
            a.grad, b.grad = gr.clone(), gr.clone()
        v0 = [p._version for p in p_our]
        o_ref.step()
        o_our.step()
        assert all(p._version > v for p, v in zip(p_our, v0)), "parameters were updated without a version bump"
    for a, b in zip(p_ref, p_our):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (a - b).abs().max().item()
    if capturable:
        assert o_our.device_step_count() == 5


def test_fused_adam_steps_reach_the_encoder_gemms(cuda_device):
    """The tensor-core layers read PACKED copies of the weights: an optimizer step must invalidate them
    (regression: raw-pointer updates that left the version counters untouched kept stale planes alive)."""
    from clica_b200.optim import FusedAdam
    _, losses = _modules()
    n, B = 8, 256
    f_our = _encoder(n, cuda_device)
    f_ref = copy.deepcopy(f_our)
    g = _mixing(n, cuda_device)
    crit = losses.LpSimCLRLoss(p=2, tau=1.0, simclr_compatibility_mode=True)
    o_our = FusedAdam(f_our.parameters(), lr=1e-3)
    o_ref = torch.optim.Adam(f_ref.parameters(), lr=1e-3)
    z1, z2 = (t.to(cuda_device) for t in _latents(B, n, cuda_device))
    seq = {"our": [], "ref": []}
    for name, f, opt in (("our", f_our, o_our), ("ref", f_ref, o_ref)):
        for _ in range(4):
            opt.zero_grad()
            a, b = f(g(z1)), f(g(z2))
            total, _, _ = crit(None, None, None, a, b, torch.roll(a, 1, 0))
            total.backward()
            opt.step()
            seq[name].append(total.item())
    assert seq["our"][0] == pytest.approx(seq["ref"][0], rel=1e-6)
    assert abs(seq["ref"][3] - seq["ref"][0]) > 1e-4, "the test needs a loss that moves"
    for a, b in zip(seq["our"], seq["ref"]):
        assert a == pytest.approx(b, rel=1e-4), (seq["our"], seq["ref"])


@pytest.mark.parametrize("host_io", [False, True])
def test_graphed_step_matches_the_eager_step(cuda_device, host_io):
    from clica_b200.graphed import GraphedTrainStep
    from clica_b200.optim import FusedAdam
    _, losses = _modules()
    n, B, p, tau, lr = 10, 512, 2, 1.0, 1e-3
    f_g = _encoder(n, cuda_device)
    f_e = copy.deepcopy(f_g)
    init = [q.detach().clone() for q in f_g.parameters()]
    g = _mixing(n, cuda_device)
    crit = losses.LpSimCLRLoss(p=p, tau=tau, simclr_compatibility_mode=True)
    step = GraphedTrainStep(f_g, g, crit, B, n, lr=lr, host_io=host_io)
    # construction (warm-up + capture) must leave the model untouched
    for q, q0 in zip(f_g.parameters(), init):
        assert torch.equal(q, q0)
    assert step.optimizer.device_step_count() == 0
    assert step.launches_per_replay >= 12      # chained GEMM launches: 15 at these sizes (26 with CLICA_TC_CHAIN=0)

    opt = FusedAdam(f_e.parameters(), lr=lr)
    got, want = [], []
    for it in range(4):
        z1, z2 = _latents(B, n, cuda_device, seed=10 + it)
        if host_io:
            got.append(step.step_host(z1, z2))
        else:
            out = step(z1.to(cuda_device), z2.to(cuda_device))
            got.append(tuple(float(x) for x in out.cpu()))
        z1d, z2d = z1.to(cuda_device), z2.to(cuda_device)
        opt.zero_grad()
        a, b = f_e(g(z1d)), f_e(g(z2d))
        total, _, parts = crit(None, None, None, a, b, torch.roll(a, 1, 0))
        total.backward()
        opt.step()
        want.append((total.item(), parts[0].item(), parts[1].item()))
    assert step.optimizer.device_step_count() == 4
    for g_, w_ in zip(got, want):
        for x, y in zip(g_, w_):
            assert x == pytest.approx(y, rel=1e-4, abs=1e-6), (got, want)
    assert abs(want[3][0] - want[0][0]) > 1e-5
    # same trajectory => same parameters.  Adam's first steps move every element by ~lr * sign(g), so an element
    # whose gradient is rounding noise (e.g. the last bias: analytically zero) may differ by O(lr); compare the
    # mean over all parameters, in units of lr, and check that the model did move
    diff = torch.cat([(q - r).abs().reshape(-1) for q, r in zip(f_g.parameters(), f_e.parameters())])
    moved = torch.cat([(q - q0).abs().reshape(-1) for q, q0 in zip(f_g.parameters(), init)])
    assert moved.mean().item() > 0.5 * lr
    assert diff.mean().item() <= 0.05 * lr, diff.mean().item() / lr
