"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py [--reference /root/reference]

It imports the reference's own ``losses.LpSimCLRLoss`` and ``encoders.get_mlp`` (unmodified, from
``/root/reference``), evaluates them with autograd on small seeded inputs in float32 and float64 and
stores inputs + outputs as ``.npz``.  Nothing here is read at test time except the ``.npz`` files; the
GPU box has no /root/reference.

Fixtures
  lpnce_<name>.npz    loss cases (inputs fp32; reference outputs in fp64 "*_64" and fp32 "*_32")
  mlp_small.npz       get_mlp(4, 3, [12, 20, 12]) weights, input, output and all gradients
  mlp_init_n5.npz     get_mlp(5, 5, [50,250,250,250,250,50]) after torch.manual_seed(1234): a few
                      values of every parameter (pins init scheme + RNG draw order) + fwd/bwd digests
  step_n5.npz         3 Adam steps of the main_mlp.py train_step body (n=5, B=64, p=2): per-step
                      losses and parameter digests
  simclr_<name>.npz   losses.SimCLRLoss cases (losses.py:162-202; `--only simclr` regenerates just these)
"""
import argparse
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _t(a, dtype):
    return torch.tensor(np.asarray(a), dtype=dtype)


def run_loss_case(losses, z1, z2, z3, p, tau, alpha, compat, use_pow, roll, dtype, gl=None):
    """Evaluate the reference loss + autograd grads. roll=True: z3 = torch.roll(z1_rec, 1, 0)."""
    a = _t(z1, dtype).requires_grad_(True)
    b = _t(z2, dtype).requires_grad_(True)
    if roll:
        n = torch.roll(a, 1, 0)
    else:
        n = _t(z3, dtype).requires_grad_(True)
    crit = losses.LpSimCLRLoss(p=p, tau=tau, alpha=alpha, simclr_compatibility_mode=compat,
                               pow=use_pow)
    mean, per_item, parts = crit(None, None, None, a, b, n)
    if gl is None:
        mean.backward()
    else:
        (per_item * _t(gl, dtype)).sum().backward()
    out = dict(loss_mean=mean.item(), loss_i=per_item.detach().numpy(),
               pos_mean=parts[0].item(), neg_mean=parts[1].item(),
               g1=a.grad.numpy(), g2=b.grad.numpy())
    if not roll:
        out["g3"] = n.grad.numpy()
    return out


def loss_cases():
    rng = np.random.RandomState(20261017)

    def pair(B, d, scale=1.0, c=0.05):
        z1 = (rng.randn(B, d) * scale).astype(np.float32)
        z2 = (z1 + c * rng.randn(B, d)).astype(np.float32)
        return z1, z2

    cases = {}
    z1, z2 = pair(64, 10)
    cases["roll_p2_d10"] = dict(z1=z1, z2=z2, p=2, tau=1.0, alpha=0.5, compat=True, pow=True, roll=True)
    z1, z2 = pair(96, 5)
    cases["roll_p1_d5_tau05"] = dict(z1=z1, z2=z2, p=1, tau=0.5, alpha=0.5, compat=True, pow=True, roll=True)
    z1, z2 = pair(80, 40, scale=0.3)
    cases["indep_p3_d40"] = dict(z1=z1, z2=z2, z3=(rng.randn(112, 40) * 0.3).astype(np.float32),
                                 p=3, tau=1.0, alpha=0.5, compat=True, pow=True, roll=False)
    z1, z2 = pair(33, 3)
    cases["indep_p2_d3_logmeanexp"] = dict(z1=z1, z2=z2, z3=rng.randn(47, 3).astype(np.float32),
                                           p=2, tau=0.7, alpha=0.5, compat=False, pow=True, roll=False)
    z1 = np.tile(rng.randn(1, 10).astype(np.float32), (128, 1))
    cases["all_equal_KA1"] = dict(z1=z1, z2=z1.copy(), p=2, tau=1.0, alpha=0.5, compat=True, pow=True, roll=True)
    z1, z2 = pair(64, 10, scale=3.0)
    cases["spread_underflow_p3"] = dict(z1=z1, z2=z2, z3=(rng.randn(64, 10) * 3 + 6).astype(np.float32),
                                        p=3, tau=0.1, alpha=0.5, compat=True, pow=True, roll=False)
    z1, z2 = pair(40, 8, scale=2.0)
    cases["spread_logmeanexp_p2"] = dict(z1=z1, z2=z2, z3=(rng.randn(56, 8) * 2 + 4).astype(np.float32),
                                         p=2, tau=0.05, alpha=0.5, compat=False, pow=True, roll=False)
    z1, z2 = pair(48, 6)
    z2[::3] = z1[::3]                       # exact duplicates: positive distance exactly 0
    z3 = rng.randn(40, 6).astype(np.float32)
    z3[:10] = z1[:10]                       # negatives identical to anchors
    z3[10:14, :3] = z1[10:14, :3]           # partially identical features (sign(0) = 0 per feature)
    cases["duplicates_p1"] = dict(z1=z1, z2=z2, z3=z3, p=1, tau=1.0, alpha=0.5, compat=True, pow=True, roll=False)
    cases["duplicates_p3"] = dict(z1=z1, z2=z2, z3=z3, p=3, tau=1.0, alpha=0.5, compat=True, pow=True, roll=False)
    z1, z2 = pair(50, 7)
    cases["alpha03_p2_weighted"] = dict(z1=z1, z2=z2, p=2, tau=1.3, alpha=0.3, compat=True, pow=True, roll=True,
                                        gl=rng.rand(50).astype(np.float32))
    z1, z2 = pair(40, 6)
    cases["nopow_p2"] = dict(z1=z1, z2=z2, z3=rng.randn(52, 6).astype(np.float32), p=2, tau=1.0, alpha=0.5,
                             compat=True, pow=False, roll=False)
    z1, z2 = pair(37, 13)
    cases["indep_p4_generic"] = dict(z1=z1, z2=z2, z3=rng.randn(29, 13).astype(np.float32), p=4, tau=2.0,
                                     alpha=0.5, compat=True, pow=True, roll=False)
    z1, z2 = pair(1, 10)
    cases["single_row"] = dict(z1=z1, z2=z2, p=2, tau=1.0, alpha=0.5, compat=True, pow=True, roll=True)
    # --- added after the first GPU pass (drawn AFTER the cases above, so those stay bit-identical) ---
    # real (non-integer) exponent and an integer exponent without a dedicated kernel (--p 5 is CLI-reachable)
    z1, z2 = pair(45, 9, scale=0.8)
    cases["indep_p2p5_generic"] = dict(z1=z1, z2=z2, z3=(rng.randn(61, 9) * 0.8).astype(np.float32), p=2.5, tau=0.8,
                                       alpha=0.5, compat=True, pow=True, roll=False)
    z1, z2 = pair(52, 10, scale=0.6)
    cases["roll_p5_generic_cpupin"] = dict(z1=z1, z2=z2, p=5, tau=1.0, alpha=0.5, compat=True, pow=True, roll=True)
    # wide feature vectors (BASELINE config 5 sweeps d = 128; the kernels split a pair over 2 / 4 / 8 lanes)
    z1, z2 = pair(48, 128, scale=0.25)
    cases["roll_p2_d128_wide"] = dict(z1=z1, z2=z2, p=2, tau=1.0, alpha=0.5, compat=True, pow=True, roll=True)
    z1, z2 = pair(36, 200, scale=0.05)
    cases["indep_p1_d200_wide"] = dict(z1=z1, z2=z2, z3=(rng.randn(70, 200) * 0.05).astype(np.float32), p=1, tau=2.0,
                                       alpha=0.5, compat=True, pow=True, roll=False)
    # row / column counts that straddle the kernels' 64-row and 128-column tiles
    z1, z2 = pair(131, 10)
    cases["indep_p3_ragged_tiles"] = dict(z1=z1, z2=z2, z3=rng.randn(259, 10).astype(np.float32), p=3, tau=1.5,
                                          alpha=0.5, compat=True, pow=True, roll=False)
    return cases


def make_loss_fixtures(losses):
    for name, c in loss_cases().items():
        store = dict(z1=c["z1"], z2=c["z2"], p=c["p"], tau=c["tau"], alpha=c["alpha"],
                     compat=int(c["compat"]), pow=int(c["pow"]), roll=int(c["roll"]))
        if not c["roll"]:
            store["z3"] = c["z3"]
        if "gl" in c:
            store["gl"] = c["gl"]
        for tag, dt in (("64", torch.float64), ("32", torch.float32)):
            out = run_loss_case(losses, c["z1"], c["z2"], c.get("z3"), c["p"], c["tau"], c["alpha"],
                                c["compat"], c["pow"], c["roll"], dt, c.get("gl"))
            for k, v in out.items():
                store[f"{k}_{tag}"] = np.asarray(v)
        np.savez_compressed(os.path.join(HERE, f"lpnce_{name}.npz"), **store)
        print("wrote lpnce_%s.npz  loss=%.6f" % (name, store["loss_mean_64"]))


def simclr_cases():
    rng = np.random.RandomState(20261018)

    def pair(B, d, scale=1.0, c=0.05):
        z1 = (rng.randn(B, d) * scale).astype(np.float32)
        z2 = (z1 + c * rng.randn(B, d)).astype(np.float32)
        return z1, z2

    def unit(z):
        return (z / np.linalg.norm(z, axis=1, keepdims=True)).astype(np.float32)

    cases = {}
    z1, z2 = pair(64, 10)
    cases["roll_d10"] = dict(z1=unit(z1), z2=unit(z2), normalize=False, tau=1.0, alpha=0.5, roll=True)
    z1, z2 = pair(50, 7)
    cases["indep_normalize_tau05"] = dict(z1=z1, z2=z2, z3=rng.randn(77, 7).astype(np.float32), normalize=True,
                                          tau=0.5, alpha=0.5, roll=False)
    z1, z2 = pair(40, 16, scale=3.0)
    cases["indep_large_logits"] = dict(z1=z1, z2=z2, z3=(rng.randn(61, 16) * 3).astype(np.float32), normalize=False,
                                       tau=0.1, alpha=0.5, roll=False)
    z1, z2 = pair(48, 10)
    cases["roll_alpha03_weighted"] = dict(z1=unit(z1), z2=unit(z2), normalize=False, tau=0.7, alpha=0.3, roll=True,
                                          gl=(rng.rand(48) - 0.3).astype(np.float32))
    z1, z2 = pair(131, 40, scale=0.4)
    cases["indep_d40_ragged"] = dict(z1=z1, z2=z2, z3=(rng.randn(259, 40) * 0.4).astype(np.float32), normalize=False,
                                     tau=1.0, alpha=0.5, roll=False)
    z1, z2 = pair(36, 128, scale=0.2)
    cases["roll_d128_wide"] = dict(z1=z1, z2=z2, normalize=False, tau=1.0, alpha=0.5, roll=True)
    return cases


def make_simclr_fixtures(losses):
    for name, c in simclr_cases().items():
        store = dict(z1=c["z1"], z2=c["z2"], normalize=int(c["normalize"]), tau=c["tau"], alpha=c["alpha"],
                     roll=int(c["roll"]))
        if not c["roll"]:
            store["z3"] = c["z3"]
        if "gl" in c:
            store["gl"] = c["gl"]
        for tag, dt in (("64", torch.float64), ("32", torch.float32)):
            a = _t(c["z1"], dt).requires_grad_(True)
            b = _t(c["z2"], dt).requires_grad_(True)
            n = torch.roll(a, 1, 0) if c["roll"] else _t(c["z3"], dt).requires_grad_(True)
            crit = losses.SimCLRLoss(normalize=c["normalize"], tau=c["tau"], alpha=c["alpha"])
            mean, per_item, parts = crit(None, None, None, a, b, n)
            if "gl" in c:
                (per_item * _t(c["gl"], dt)).sum().backward()
            else:
                mean.backward()
            out = dict(loss_mean=mean.item(), loss_i=per_item.detach().numpy(), pos_mean=parts[0].item(),
                       neg_mean=parts[1].item(), g1=a.grad.numpy(), g2=b.grad.numpy())
            if not c["roll"]:
                out["g3"] = n.grad.numpy()
            for k, v in out.items():
                store[f"{k}_{tag}"] = np.asarray(v)
        np.savez_compressed(os.path.join(HERE, f"simclr_{name}.npz"), **store)
        print("wrote simclr_%s.npz  loss=%.6f" % (name, store["loss_mean_64"]))


def digest(t):
    a = t.detach().double().numpy().ravel()
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])


def make_mlp_fixtures(encoders):
    # (1) small net, everything stored
    torch.manual_seed(7)
    f = encoders.get_mlp(n_in=4, n_out=3, layers=[12, 20, 12])
    x = torch.randn(19, 4)
    gy = torch.randn(19, 3)
    x.requires_grad_(True)
    y = f(x)
    y.backward(gy)
    store = dict(x=x.detach().numpy(), gy=gy.numpy(), y=y.detach().numpy(), dx=x.grad.numpy(),
                 keys=np.array(list(f.state_dict().keys())))
    for k, v in f.state_dict().items():
        store["param_" + k] = v.numpy()
    for k, prm in f.named_parameters():
        store["grad_" + k] = prm.grad.numpy()
    # fp64 replay of the same net
    f64 = encoders.get_mlp(n_in=4, n_out=3, layers=[12, 20, 12]).double()
    f64.load_state_dict({k: v.double() for k, v in f.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    y64 = f64(x64)
    y64.backward(gy.double())
    store["y_64"] = y64.detach().numpy()
    store["dx_64"] = x64.grad.numpy()
    for k, prm in f64.named_parameters():
        store["grad64_" + k] = prm.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "mlp_small.npz"), **store)
    print("wrote mlp_small.npz; modules:", [type(m).__name__ for m in f])

    # (2) standard n=5 encoder: pins init scheme, RNG draw order, state_dict keys/shapes
    n = 5
    torch.manual_seed(1234)
    f = encoders.get_mlp(n_in=n, n_out=n, layers=[10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n])
    torch.manual_seed(99)
    x = torch.randn(32, n)
    y = f(x)
    y.pow(2).mean().backward()
    sd = f.state_dict()
    store = dict(x=x.numpy(), y=y.detach().numpy(), keys=np.array(list(sd.keys())),
                 shapes=np.array([list(v.shape) + [0] * (2 - v.dim()) for v in sd.values()]),
                 module_types=np.array([type(m).__name__ for m in f]),
                 leaky_slope=np.array([m.negative_slope for m in f if isinstance(m, torch.nn.LeakyReLU)]))
    for k, v in sd.items():
        store["head_" + k] = v.numpy().ravel()[:8]
        store["digest_" + k] = digest(v)
    for k, prm in f.named_parameters():
        store["graddigest_" + k] = digest(prm.grad)
    np.savez_compressed(os.path.join(HERE, "mlp_init_n5.npz"), **store)
    print("wrote mlp_init_n5.npz; n_params =", sum(v.numel() for v in sd.values()))


def make_step_fixture(losses, encoders):
    """Three steps of main_mlp.py:258-285 (unsupervised branch) with the reference loss/encoder."""
    n, B, p, tau, lr = 5, 64, 2, 1.0, 1e-4
    rng = np.random.RandomState(5)
    gW = [np.linalg.qr(rng.randn(n, n))[0].astype(np.float32) * (0.75 + 0.5 * rng.rand(n).astype(np.float32))
          for _ in range(3)]
    g_mods = []
    for i, W in enumerate(gW):
        lin = torch.nn.Linear(n, n, bias=False)
        with torch.no_grad():
            lin.weight.copy_(torch.tensor(W))
        lin.weight.requires_grad = False
        g_mods.append(lin)
        if i != 2:
            g_mods.append(torch.nn.LeakyReLU(0.2))
    g = torch.nn.Sequential(*g_mods)
    torch.manual_seed(4321)
    f = encoders.get_mlp(n_in=n, n_out=n, layers=[10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n])
    init_sd = {k: v.clone() for k, v in f.state_dict().items()}
    crit = losses.LpSimCLRLoss(p=p, tau=tau, simclr_compatibility_mode=True)
    opt = torch.optim.Adam(f.parameters(), lr=lr)
    h = lambda z: f(g(z))
    z1s, z2s, rec = [], [], []
    for step in range(3):
        z1 = rng.uniform(-1, 1, size=(B, n)).astype(np.float32)
        z2 = np.clip(z1 + 0.05 * rng.randn(B, n), -1, 1).astype(np.float32)
        z1s.append(z1), z2s.append(z2)
        a, b = torch.tensor(z1), torch.tensor(z2)
        opt.zero_grad()
        a_rec, b_rec = h(a), h(b)
        n_rec = torch.roll(a_rec, 1, 0)
        total, _, parts = crit(a, b, torch.roll(a, 1, 0), a_rec, b_rec, n_rec)
        total.backward()
        if step == 0:
            grad0 = {k: prm.grad.clone() for k, prm in f.named_parameters()}
        opt.step()
        rec.append([total.item(), parts[0].item(), parts[1].item()])
    store = dict(n=n, B=B, p=p, tau=tau, lr=lr, z1=np.stack(z1s), z2=np.stack(z2s), g_weights=np.stack(gW),
                 losses=np.array(rec), keys=np.array(list(init_sd.keys())))
    for k, v in f.state_dict().items():
        store["final_digest_" + k] = digest(v)
        store["init_digest_" + k] = digest(init_sd[k])
        store["final_head_" + k] = v.numpy().ravel()[:8]
    for k, v in grad0.items():
        store["grad0_digest_" + k] = digest(v)
        store["grad0_head_" + k] = v.numpy().ravel()[:8]
    np.savez_compressed(os.path.join(HERE, "step_n5.npz"), **store)
    print("wrote step_n5.npz  losses:", rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--only", default=None, choices=[None, "simclr"])
    args = ap.parse_args()
    warnings.simplefilter("ignore", SyntaxWarning)
    sys.path.insert(0, args.reference)
    import losses      # the reference's, unmodified
    import encoders    # the reference's, unmodified
    assert os.path.dirname(os.path.abspath(losses.__file__)) == os.path.abspath(args.reference)
    torch.set_num_threads(1)   # deterministic summation order for the fp32 variants
    if args.only == "simclr":
        make_simclr_fixtures(losses)
        return
    make_loss_fixtures(losses)
    make_simclr_fixtures(losses)
    make_mlp_fixtures(encoders)
    make_step_fixture(losses, encoders)


if __name__ == "__main__":
    main()
