"""Drop-in ``encoders`` module: ``get_mlp`` with the reference's signature, forward/backward on B200 kernels.

``main_mlp.py:297-313`` builds ``f = encoders.get_mlp(n, n, [10n, 50n, 50n, 50n, 50n, 10n],
output_normalization=...)`` and then uses ``f.to(device)``, ``print(f)``, ``f.parameters()``, ``f(x)``,
``f[-1].r`` and ``f.state_dict()``.  The object returned here is an ``nn.Sequential`` subclass holding the
very same ``nn.Linear`` / ``nn.LeakyReLU`` (/ output-normalisation) modules in the same order -- same
parameter init (``nn.Linear`` default, same RNG draw order), same ``state_dict`` keys -- whose ``forward``
sends the Linear+LeakyReLU prefix through ``clica_mlp_fwd`` / ``clica_mlp_bwd`` when the input is a CUDA
fp32 tensor.  CPU inputs (BASELINE config 1) and stacks with BatchNorm/GroupNorm run the torch modules.
``get_flow`` (FrEIA, out of scope) is re-exported from the reference checkout when one is reachable.
"""
from typing import List, Optional

import torch
from torch import nn

import clica_b200
from clica_b200 import functional as _F

from _reference import load_reference_module

__all__ = ["get_mlp", "get_flow", "FusedMLP"]


class FusedMLP(nn.Sequential):
    """``nn.Sequential`` of Linear / LeakyReLU (/ trailing output norm) with a fused CUDA forward+backward."""

    def _plan(self):
        """(linears, slope, n_consumed) for the leading Linear, LeakyReLU, ..., Linear run, or None."""
        mods = list(self)
        linears, slope, i = [], None, 0
        while i < len(mods) and isinstance(mods[i], nn.Linear):
            linears.append(mods[i])
            i += 1
            if i < len(mods) and isinstance(mods[i], nn.LeakyReLU) and i + 1 < len(mods) \
                    and isinstance(mods[i + 1], nn.Linear):
                s = float(mods[i].negative_slope)
                if slope is not None and s != slope:
                    return None
                slope = s
                i += 1
            else:
                break
        if not linears or any(l.bias is None for l in linears):
            return None
        if any(isinstance(m, (nn.Linear, nn.LeakyReLU)) for m in mods[i:]):
            return None      # irregular stack (e.g. BatchNorm in between): leave it to torch
        return linears, (0.01 if slope is None else slope), i

    def forward(self, input):
        x = input
        plan = self._plan() if (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32) else None
        if plan is None:
            return super().forward(x)
        linears, slope, consumed = plan
        if any((not l.weight.is_cuda) or l.weight.dtype != torch.float32 for l in linears):
            raise RuntimeError("FusedMLP: CUDA input but the parameters are not CUDA fp32 (call .to(device))")
        lead = x.shape[:-1]
        y = _F.mlp_forward(x.reshape(-1, x.shape[-1]), [l.weight for l in linears], [l.bias for l in linears],
                           slope=slope)
        y = y.reshape(*lead, y.shape[-1])
        for m in list(self)[consumed:]:
            y = m(y)
        return y


def get_mlp(n_in: int, n_out: int, layers: List[int], layer_normalization: Optional[str] = None,
            output_normalization: Optional[str] = None, output_normalization_kwargs=None):
    """Creates an MLP (reference ``encoders.py:10-85``; same arguments, same module order, same init).

    Args:
        n_in / n_out: input / output width.  layers: hidden widths (``n_out`` is appended to the list the
        caller passed, as the reference does).  layer_normalization: None, "bn" or "gn".
        output_normalization: None, "fixed_sphere", "learnable_sphere", "fixed_box", "learnable_box".
    """
    if len(layers) == 0:
        raise ValueError("get_mlp needs at least one hidden layer (the reference's empty-`layers` path is broken)")
    if layer_normalization not in (None, "bn", "gn"):
        layer_normalization = None    # the reference silently ignores unknown values
    layers.append(n_out)
    mods: List[nn.Module] = []
    width = n_in
    for idx, nxt in enumerate(layers):
        mods.append(nn.Linear(width, nxt))
        if idx != len(layers) - 1:
            if layer_normalization == "bn":
                mods.append(nn.BatchNorm1d(nxt))
            elif layer_normalization == "gn":
                mods.append(nn.GroupNorm(1, nxt))
            mods.append(nn.LeakyReLU())
        width = nxt

    kwargs = {} if output_normalization_kwargs is None else output_normalization_kwargs
    if output_normalization is not None:
        if output_normalization not in ("fixed_sphere", "learnable_sphere", "fixed_box", "learnable_box"):
            raise ValueError("output_normalization")
        import layers as ls      # the reference's layers.py (output norms are O(B*d) torch code, out of scope)
        if output_normalization == "fixed_sphere":
            mods.append(ls.RescaleLayer(fixed_r=True, **kwargs))
        elif output_normalization == "learnable_sphere":
            mods.append(ls.RescaleLayer(init_r=1.0, fixed_r=False))
        elif output_normalization == "fixed_box":
            mods.append(ls.SoftclipLayer(n=n_out, fixed_abs_bound=True, **kwargs))
        else:
            mods.append(ls.SoftclipLayer(n=n_out, fixed_abs_bound=False, **kwargs))
    return FusedMLP(*mods)


def get_flow(*args, **kwargs):
    """FrEIA-based invertible flow (reference ``encoders.py:88-152``): out of scope, forwarded to the reference."""
    ref = load_reference_module("encoders")
    if ref is None:
        raise RuntimeError("get_flow is not part of the B200 hot path and no reference checkout is reachable")
    return ref.get_flow(*args, **kwargs)
