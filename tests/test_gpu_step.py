"""GPU (-m gpu): the main_mlp.py train_step body on the drop-in modules vs the golden 3-step run of the reference."""
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _digest(t):
    a = t.detach().double().cpu().numpy().ravel()
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])


def test_three_adam_steps_match_the_reference(cuda_device):
    import clica_b200
    sys.path.insert(0, clica_b200.DROPIN_DIR)
    import encoders
    import losses
    g = load_golden("step_n5")
    n, B, p, tau, lr = int(g["n"]), int(g["B"]), int(g["p"]), float(g["tau"]), float(g["lr"])
    mods = []
    for i, W in enumerate(g["g_weights"]):
        lin = torch.nn.Linear(n, n, bias=False)
        with torch.no_grad():
            lin.weight.copy_(torch.tensor(W))
        lin.weight.requires_grad = False
        mods.append(lin)
        if i != 2:
            mods.append(torch.nn.LeakyReLU(0.2))
    gnet = torch.nn.Sequential(*mods).to(cuda_device)
    torch.manual_seed(4321)
    f = encoders.get_mlp(n_in=n, n_out=n, layers=[10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n])
    for k, v in f.state_dict().items():
        assert np.allclose(_digest(v), g["init_digest_" + k], rtol=1e-6), k     # same init as the reference
    f = f.to(cuda_device)
    crit = losses.LpSimCLRLoss(p=p, tau=tau, simclr_compatibility_mode=True)
    opt = torch.optim.Adam(f.parameters(), lr=lr)
    h = lambda z: f(gnet(z))
    for step in range(3):
        z1 = torch.tensor(g["z1"][step], device=cuda_device)
        z2 = torch.tensor(g["z2"][step], device=cuda_device)
        opt.zero_grad()
        a, b = h(z1), h(z2)
        total, _, parts = crit(z1, z2, torch.roll(z1, 1, 0), a, b, torch.roll(a, 1, 0))
        total.backward()
        if step == 0:
            for k, prm in f.named_parameters():
                ref = g["grad0_head_" + k]
                got = prm.grad.detach().cpu().numpy().ravel()[:8]
                scale = np.sqrt(g["grad0_digest_" + k][2] / prm.numel()) + 1e-30      # rms of the reference grad
                ratio = np.abs(got - ref).max() / (scale + 1e-30)
                # measured on B200: <= 1e-4 of the gradient's rms for every parameter (3xTF32 encoder, fused loss); the last
                # bias' gradient is analytically 0
                assert np.abs(got - ref).max() <= 1e-4 * scale + 1e-10, (k, ratio)
        opt.step()
        rec = g["losses"][step]
        assert abs(total.item() - rec[0]) <= 5e-6 * max(1.0, abs(rec[0]))
        assert abs(parts[1].item() - rec[2]) <= 5e-6 * max(1.0, abs(rec[2]))
    for k, v in f.state_dict().items():
        # Adam's first steps move every weight by ~lr regardless of gradient scale; compare aggregate digests
        d_ref, d_got = g["final_digest_" + k], _digest(v)
        assert abs(d_got[1] - d_ref[1]) <= 1e-4 * d_ref[1] + 1e-6, k
