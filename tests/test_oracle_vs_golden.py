"""CPU (-m "not gpu"): pin the oracle restatements against outputs of the reference itself.

The golden ``.npz`` files were produced by tests/golden/make_golden.py importing the unmodified
/root/reference/losses.py and encoders.py.  Tolerances: the fp64 oracle must agree with the fp64
reference run to 1e-11 (same math, different summation order); the fp32 torch port must agree with
the fp32 reference run to a few fp32 ulps.
"""
import numpy as np
import pytest
import torch

from conftest import golden_loss_cases, load_golden
from oracle import c_oracle, mlp_oracle, torch_port


def _close(a, b, rtol, atol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.max(np.abs(a - b) - rtol * np.abs(b)) if a.size else 0.0
    assert err <= atol, f"{what}: max excess error {err:.3e} (rtol={rtol}, atol={atol})"


@pytest.mark.parametrize("name", golden_loss_cases())
def test_c_oracle_matches_reference_fp64(name):
    g = load_golden("lpnce_" + name)
    roll = bool(g["roll"])
    z3 = np.roll(g["z1"], 1, axis=0) if roll else g["z3"]
    out = c_oracle.lpnce(g["z1"], g["z2"], z3, float(g["p"]), float(g["tau"]), float(g["alpha"]),
                         include_pos=bool(g["compat"]), use_pow=bool(g["pow"]),
                         gl=g["gl"] if "gl" in g else None)
    scale = max(1.0, abs(float(g["loss_mean_64"])))
    _close(out["loss_mean"], g["loss_mean_64"], 0, 1e-11 * scale, "loss_mean")
    _close(out["loss_i"], g["loss_i_64"], 1e-12, 1e-11 * scale, "loss_i")
    _close(out["pos_mean"], g["pos_mean_64"], 1e-12, 1e-12, "pos_mean")
    _close(out["neg_mean"], g["neg_mean_64"], 1e-12, 1e-11 * scale, "neg_mean")
    g1 = out["g1"] + (np.roll(out["g3"], -1, axis=0) if roll else 0.0)   # RollBackward
    gmax = max(np.abs(g["g1_64"]).max(), 1e-30)
    _close(g1, g["g1_64"], 1e-10, 1e-12 * gmax + 1e-300, "g1")
    _close(out["g2"], g["g2_64"], 1e-10, 1e-12 * gmax + 1e-300, "g2")
    if not roll:
        _close(out["g3"], g["g3_64"], 1e-10, 1e-12 * gmax + 1e-300, "g3")


@pytest.mark.parametrize("name", golden_loss_cases())
def test_torch_port_matches_reference_fp32(name):
    g = load_golden("lpnce_" + name)
    roll = bool(g["roll"])
    if not bool(g["pow"]) and False:
        pytest.skip()
    torch.set_num_threads(1)
    a = torch.tensor(g["z1"], requires_grad=True)
    b = torch.tensor(g["z2"], requires_grad=True)
    n = torch.roll(a, 1, 0) if roll else torch.tensor(g["z3"], requires_grad=True)
    pexp = float(g["p"])
    pexp = int(pexp) if pexp == int(pexp) else pexp          # the CLI passes ints; 2.5 pins the real-exponent path
    mean, per_item, parts = torch_port.lp_infonce(a, b, n, pexp, float(g["tau"]),
                                                  float(g["alpha"]), bool(g["compat"]),
                                                  bool(g["pow"]))
    if "gl" in g:
        (per_item * torch.tensor(g["gl"])).sum().backward()
    else:
        mean.backward()
    scale = max(1.0, float(np.abs(g["loss_i_32"]).max()))
    _close(per_item.detach().numpy(), g["loss_i_32"], 0, 4e-6 * scale, "loss_i")
    _close(parts[0].item(), g["pos_mean_32"], 1e-5, 1e-7, "pos_mean")
    gmax = float(np.abs(g["g1_32"]).max()) + 1e-30
    _close(a.grad.numpy(), g["g1_32"], 1e-5, 1e-6 * gmax, "g1")
    _close(b.grad.numpy(), g["g2_32"], 1e-5, 1e-6 * gmax, "g2")


def test_fp32_reference_noise_floor():
    """Documents how far the reference's own fp32 run sits from its fp64 run (sets GPU tolerances)."""
    worst = 0.0
    for name in golden_loss_cases():
        g = load_golden("lpnce_" + name)
        scale = max(1.0, float(np.abs(g["loss_i_64"]).max()))
        worst = max(worst, float(np.abs(g["loss_i_32"] - g["loss_i_64"]).max()) / scale)
    assert worst < 5e-6


def test_known_answer_KA1():
    """All outputs identical => loss_i = 2(1-alpha) ln(B+1) for any tau, p."""
    z = np.tile(np.linspace(-1, 1, 7, dtype=np.float32), (200, 1))
    for p in (1, 2, 3):
        out = c_oracle.lpnce(z, z, np.roll(z, 1, 0), p, tau=0.3, alpha=0.5)
        assert np.allclose(out["loss_i"], np.log(201.0), atol=1e-12)
        assert np.abs(out["g1"]).max() == 0.0 and np.abs(out["g3"]).max() == 0.0


def test_invariances_P3():
    rng = np.random.RandomState(0)
    z1 = rng.randn(70, 9).astype(np.float32)
    z2 = (z1 + 0.1 * rng.randn(70, 9)).astype(np.float32)
    z3 = rng.randn(55, 9).astype(np.float32)
    base = c_oracle.lpnce(z1, z2, z3, 3, tau=0.8, need_grad=False)["loss_i"]
    perm = c_oracle.lpnce(z1, z2, z3[rng.permutation(55)], 3, tau=0.8, need_grad=False)["loss_i"]
    assert np.allclose(base, perm, rtol=1e-12, atol=1e-12)
    # loss(p, tau, z) == loss(p, 1, z / tau^(1/p)); inputs are re-rounded to fp32, hence 1e-5
    s = 0.8 ** (1.0 / 3.0)
    resc = c_oracle.lpnce(z1 / s, z2 / s, z3 / s, 3, tau=1.0, need_grad=False)["loss_i"]
    assert np.allclose(base, resc, rtol=1e-5, atol=1e-5)


def test_mlp_oracle_matches_reference():
    g = load_golden("mlp_small")
    keys = [str(k) for k in g["keys"]]
    Ws = [g["param_" + k] for k in keys if k.endswith("weight")]
    bs = [g["param_" + k] for k in keys if k.endswith("bias")]
    y, acts, pre = mlp_oracle.mlp_forward(g["x"], Ws, bs, slope=0.01)
    _close(y, g["y_64"], 1e-12, 1e-13, "y")
    dWs, dbs, dx = mlp_oracle.mlp_backward(g["gy"], Ws, acts, pre, slope=0.01, need_dx=True)
    wk = [k for k in keys if k.endswith("weight")]
    bk = [k for k in keys if k.endswith("bias")]
    for k, dW in zip(wk, dWs):
        _close(dW, g["grad64_" + k], 1e-11, 1e-13, "dW " + k)
    for k, db in zip(bk, dbs):
        _close(db, g["grad64_" + k], 1e-11, 1e-13, "db " + k)
    _close(dx, g["dx_64"], 1e-11, 1e-13, "dx")


def test_port_encoder_structure_and_init_match_reference():
    g = load_golden("mlp_init_n5")
    torch.manual_seed(1234)
    f = torch_port.build_encoder(5)
    sd = f.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [type(m).__name__ for m in f] == [str(t) for t in g["module_types"]]
    for k, v in sd.items():
        assert np.array_equal(v.numpy().ravel()[:8], g["head_" + k]), k
    y = f(torch.tensor(g["x"]))
    _close(y.detach().numpy(), g["y"], 1e-5, 1e-7, "y")


def simclr_cases():
    import glob
    import os
    from conftest import GOLDEN_DIR
    return sorted(os.path.basename(p)[len("simclr_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "simclr_*.npz")))


@pytest.mark.parametrize("name", simclr_cases())
def test_simclr_oracle_matches_reference_fp64(name):
    """oracle/simclr_oracle.py vs losses.SimCLRLoss of the reference (fp64 run), incl. normalize=True via the chain rule."""
    from oracle import simclr_oracle
    g = load_golden("simclr_" + name)
    roll, normalize = bool(g["roll"]), bool(g["normalize"])
    z1, z2 = g["z1"].astype(np.float64), g["z2"].astype(np.float64)
    z3 = np.roll(z1, 1, axis=0) if roll else g["z3"].astype(np.float64)

    def unit(z):
        return z / np.linalg.norm(z, axis=1, keepdims=True)

    def unit_vjp(z, gu):      # gradient through z -> z / |z|
        n = np.linalg.norm(z, axis=1, keepdims=True)
        u = z / n
        return (gu - u * (u * gu).sum(1, keepdims=True)) / n
    a, b, c = (unit(z1), unit(z2), unit(z3)) if normalize else (z1, z2, z3)
    out = simclr_oracle.simclr(a, b, c, float(g["tau"]), float(g["alpha"]), gl=g["gl"] if "gl" in g else None)
    g1, g2, g3 = out["g1"], out["g2"], out["g3"]
    if normalize:
        g1, g2, g3 = unit_vjp(z1, g1), unit_vjp(z2, g2), unit_vjp(z3, g3)
    if roll:
        g1 = g1 + np.roll(g3, -1, axis=0)
    scale = max(1.0, float(np.abs(g["loss_i_64"]).max()))
    _close(out["loss_i"], g["loss_i_64"], 0, 1e-10 * scale, "loss_i")
    _close(out["loss_mean"], g["loss_mean_64"], 0, 1e-10 * scale, "loss_mean")
    _close(out["pos_mean"], g["pos_mean_64"], 1e-12, 1e-10 * scale, "pos_mean")
    _close(out["neg_mean"], g["neg_mean_64"], 1e-12, 1e-10 * scale, "neg_mean")
    # gradients are differences of terms of size |z| / (B tau); when a row's positive takes all the soft-max mass they
    # cancel almost completely (case indep_large_logits), so the absolute tolerance is relative to the TERM size
    gmax = max(np.abs(g["g1_64"]).max(), np.abs(z1).max() / (len(z1) * float(g["tau"])))
    _close(g1, g["g1_64"], 1e-9, 1e-11 * gmax, "g1")
    _close(g2, g["g2_64"], 1e-9, 1e-11 * gmax, "g2")
    if not roll:
        _close(g3, g["g3_64"], 1e-9, 1e-11 * gmax, "g3")
