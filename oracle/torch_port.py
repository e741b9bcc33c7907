"""oracle/torch_port.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's InfoNCE training step restated with the *same torch formulation the reference
executes* (so that timing it on host cores is a fair stand-in for the reference's CPU path, and so
that its fp32 rounding behaviour is the reference's):

* ``lp_infonce``      <- /root/reference/losses.py:443-477 (p >= 1 branch) and :506-510
* ``build_encoder``   <- /root/reference/encoders.py:36-58 with main_mlp.py:297-309's widths
* ``build_mixing``    <- the *shape* of invertible_network_utils.py:87-123 (n x n bias-free Linear +
                         LeakyReLU(0.2), frozen).  The condition-number search of the reference is
                         setup-only and deliberately not restated.
* ``train_step``      <- /root/reference/main_mlp.py:258-285 (unsupervised branch)

Pinned against the reference's own outputs by tests/test_oracle_vs_golden.py.
"""
import math

import torch
from torch import nn


def lp_infonce(z1_rec, z2_rec, z3_rec, p, tau=1.0, alpha=0.5, compat=True, use_pow=True):
    """Materialised B x M x d formulation (what the reference runs). Returns (mean, per_item, [pos_mean, neg_mean])."""
    diff = z1_rec[:, None, :] - z3_rec[None, :, :]           # [B, M, d]
    neg = torch.norm(diff, p=p, dim=-1)                       # [B, M]
    pos = torch.norm(z1_rec - z2_rec, p=p, dim=-1)            # [B]
    if use_pow:
        neg, pos = neg.pow(p), pos.pow(p)
    part_pos = pos / tau
    if compat:
        logits = torch.cat([neg, pos[:, None]], dim=1) / (-tau)
        part_neg = torch.logsumexp(logits, dim=1)
    else:
        part_neg = torch.logsumexp(neg / (-tau), dim=1) - math.log(neg.shape[1])
    per_item = 2.0 * (alpha * part_pos + (1.0 - alpha) * part_neg)
    return per_item.mean(), per_item, [part_pos.mean(), part_neg.mean()]


def lp_infonce_chunked(z1_rec, z2_rec, z3_rec, p, tau=1.0, alpha=0.5, compat=True, rows=256):
    """Same math, anchors processed in row chunks (rows are independent) to bound memory."""
    outs = []
    for s in range(0, z1_rec.shape[0], rows):
        outs.append(lp_infonce(z1_rec[s:s + rows], z2_rec[s:s + rows], z3_rec, p, tau, alpha,
                               compat)[1])
    per_item = torch.cat(outs)
    return per_item.mean(), per_item


def encoder_widths(n):
    """main_mlp.py:297-309: hidden widths 10n, 50n x4, 10n; in = out = n."""
    return [n, 10 * n, 50 * n, 50 * n, 50 * n, 50 * n, 10 * n, n]


def build_encoder(n, widths=None, slope=0.01):
    """Linear(+LeakyReLU(0.01)) stack, no activation after the last Linear; default nn.Linear init."""
    widths = widths or encoder_widths(n)
    mods = []
    for li in range(len(widths) - 1):
        mods.append(nn.Linear(widths[li], widths[li + 1]))
        if li != len(widths) - 2:
            mods.append(nn.LeakyReLU(slope))
    return nn.Sequential(*mods)


def build_mixing(n, n_layers=3, seed=0):
    """Frozen mixing net g: n_layers bias-free n x n Linear with LeakyReLU(0.2) in between."""
    gen = torch.Generator().manual_seed(seed)
    mods = []
    for li in range(n_layers):
        lin = nn.Linear(n, n, bias=False)
        q, _ = torch.linalg.qr(torch.randn(n, n, generator=gen))
        scale = 0.75 + 0.5 * torch.rand(n, generator=gen)    # well-conditioned, not orthogonal
        with torch.no_grad():
            lin.weight.copy_(q * scale[None, :])
        mods.append(lin)
        if li != n_layers - 1:
            mods.append(nn.LeakyReLU(0.2))
    g = nn.Sequential(*mods)
    for prm in g.parameters():
        prm.requires_grad = False
    return g


def synth_latents(B, n, space="sphere", c_param=0.05, seed=0, dtype=torch.float32):
    """Synthetic (anchor, positive) latents of the benchmark shape.

    sphere: marginal uniform on S^{n-1}, conditional = project(z + c_param*N(0,1))  (spaces.py:134-170)
    real:   marginal N(0,1), conditional z + c_param*N(0,1)
    """
    gen = torch.Generator().manual_seed(seed)
    z = torch.randn(B, n, generator=gen, dtype=dtype)
    if space == "sphere":
        z = z / z.norm(dim=-1, keepdim=True)
    zt = z + c_param * torch.randn(B, n, generator=gen, dtype=dtype)
    if space == "sphere":
        zt = zt / zt.norm(dim=-1, keepdim=True)
    return z, zt


def train_step(f, g, optimizer, z1, z2, p, tau=1.0, alpha=0.5):
    """One unsupervised step; returns (loss, [pos_mean, neg_mean]) as Python floats."""
    optimizer.zero_grad()
    z1_rec = f(g(z1))
    z2_rec = f(g(z2))
    z3_rec = torch.roll(z1_rec, 1, 0)
    total, _, parts = lp_infonce(z1_rec, z2_rec, z3_rec, p, tau, alpha, compat=True)
    total.backward()
    optimizer.step()
    return total.item(), [x.item() for x in parts]


def sampled_train_step(f, g, optimizer, z1, z2, z3_rec_full, rows, p, tau=1.0, alpha=0.5):
    """Bounded sample of one step for CPU timing: the first ``rows`` (anchor, positive) pairs go
    through encoder fwd+bwd and are contrasted against ALL ``M = len(z3_rec_full)`` negatives
    (a pre-encoded leaf that requires grad, so the column-term gradient is computed as in the full
    step).  Per-pair cost equals the full step's; pairs/s = rows / time."""
    optimizer.zero_grad()
    z3_rec_full.grad = None
    z1_rec = f(g(z1[:rows]))
    z2_rec = f(g(z2[:rows]))
    total, _, parts = lp_infonce(z1_rec, z2_rec, z3_rec_full, p, tau, alpha, compat=True)
    total.backward()
    optimizer.step()
    return total.item(), [x.item() for x in parts]
