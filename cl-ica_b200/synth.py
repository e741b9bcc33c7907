"""Synthetic inputs of the benchmark / smoke shape (there are no datasets: cl-ica samples latents on the fly).

* ``build_mixing``  a frozen mixing net g with the SHAPE of the reference's ``construct_invertible_mlp`` output
  (``invertible_network_utils.py:87-123``: bias-free n x n Linear layers, LeakyReLU(0.2) in between).  The
  reference's condition-number rejection search is setup-only and takes minutes; here
  the weights are well-conditioned scaled-orthogonal matrices.  g is forward-only, 3 n^2 MAC/row, left to torch.
* ``synth_latents`` (anchor, positive) latents: sphere = uniform marginal on S^{n-1} + projected Gaussian
  conditional (``spaces.py:134-170``, --c-p 2 --c-param 0.05); real = N(0,1) marginal + Gaussian conditional.
"""
import torch
from torch import nn


def build_mixing(n, n_layers=3, seed=0):
    gen = torch.Generator().manual_seed(seed)
    mods = []
    for li in range(n_layers):
        lin = nn.Linear(n, n, bias=False)
        q, _ = torch.linalg.qr(torch.randn(n, n, generator=gen))
        scale = 0.75 + 0.5 * torch.rand(n, generator=gen)
        with torch.no_grad():
            lin.weight.copy_(q * scale[None, :])
        mods.append(lin)
        if li != n_layers - 1:
            mods.append(nn.LeakyReLU(0.2))
    g = nn.Sequential(*mods)
    for prm in g.parameters():
        prm.requires_grad = False
    return g


def synth_latents(B, n, space="sphere", c_param=0.05, seed=0):
    gen = torch.Generator().manual_seed(seed)
    z = torch.randn(B, n, generator=gen)
    if space == "sphere":
        z = z / z.norm(dim=-1, keepdim=True)
    zt = z + c_param * torch.randn(B, n, generator=gen)
    if space == "sphere":
        zt = zt / zt.norm(dim=-1, keepdim=True)
    return z, zt
