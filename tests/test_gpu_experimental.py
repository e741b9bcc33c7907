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
