"""Drop-in ``losses`` module: same surface as the reference's losses.py, LpSimCLRLoss on B200 kernels.

``main_mlp.py:4,142-147`` does ``import losses`` and builds ``losses.LpSimCLRLoss(p, tau,
simclr_compatibility_mode=True)`` (or ``losses.SimCLRLoss`` for ``--p 0``).  With this directory ahead
of the reference on ``sys.path`` that import resolves here:

* :class:`LpSimCLRLoss` keeps the reference constructor and call signature (``losses.py:416-431``) and
  return structure ``(mean, per_item, [pos_mean, neg_mean])`` (``losses.py:467-477``); for CUDA fp32
  inputs with ``p >= 1`` and ``pow=True`` it runs the fused CUDA kernel (forward and backward).
* :class:`SimCLRLoss` (``losses.py:162-202``, selected by ``main_mlp.py:146-147`` for ``--p 0``) runs the same
  fused kernels in their dot-product-similarity form (logit = z1.z3 / tau) for CUDA fp32 inputs.
* every other name of the reference module (CLLoss, UniformityLoss, ...) is re-exported from the reference
  checkout when one is reachable -- none of them is on the hot path.
* inputs outside the kernel's domain (CPU tensors = BASELINE config 1 "reference plumbing", p < 1,
  pow=False, non-fp32, d > 320) are handed to the reference's own ``LpSimCLRLoss.loss`` -- explicitly, with
  a one-time warning -- or raise if no reference checkout is reachable.  On a CUDA fp32 input inside the
  domain there is exactly one path and a missing CUDA library raises.
"""
import warnings

import torch

import clica_b200
from clica_b200 import functional as _F

from _reference import load_reference_module

_ref = load_reference_module("losses")
if _ref is not None:
    for _name in dir(_ref):
        if not _name.startswith("__"):
            globals()[_name] = getattr(_ref, _name)
    _Base = _ref.LpSimCLRLoss
else:
    class _Base:   # minimal stand-in so that the class below is importable on a box without the reference
        def __init__(self, p, tau=1.0, alpha=0.5, simclr_compatibility_mode=False, pow=True):
            self.p, self.tau, self.alpha = p, tau, alpha
            self.simclr_compatibility_mode, self.pow = simclr_compatibility_mode, pow

        def __call__(self, z1, z2_con_z1, z3, z1_rec, z2_con_z1_rec, z3_rec):
            return self.loss(z1, z2_con_z1, z3, z1_rec, z2_con_z1_rec, z3_rec)

_MAX_D = 320
_warned = set()

if _ref is not None:
    _SimBase = _ref.SimCLRLoss
else:
    class _SimBase:   # stand-in with the reference constructor (losses.py:172-175)
        def __init__(self, normalize=False, tau=1.0, alpha=0.5):
            self.normalize, self.tau, self.alpha = normalize, tau, alpha

        def __call__(self, z1, z2_con_z1, z3, z1_rec, z2_con_z1_rec, z3_rec):
            return self.loss(z1, z2_con_z1, z3, z1_rec, z2_con_z1_rec, z3_rec)


def _cuda_domain(tensors, max_d=_MAX_D):
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            return "non-tensor input"
        if not t.is_cuda:
            return "CPU tensors"
        if t.dtype != torch.float32:
            return f"dtype {t.dtype}"
        if t.dim() != 2:
            return f"{t.dim()}-D input"
    if tensors[0].shape[1] > max_d:
        return f"feature width {tensors[0].shape[1]} > {max_d}"
    return None


def _delegate(cls_name, why, call):
    if _ref is None:
        raise RuntimeError(f"{cls_name}: {why} is outside the CUDA kernel's domain and no reference "
                           "checkout is reachable (set CLICA_REFERENCE_DIR)")
    if (cls_name, why) not in _warned:
        _warned.add((cls_name, why))
        warnings.warn(f"clica_b200.{cls_name}: {why} -> delegating to the reference's torch code", stacklevel=3)
    return call()


class SimCLRLoss(_SimBase):
    """InfoNCE on dot-product similarities (reference ``losses.py:162-202``), fused on sm_100a.

    Args (identical to the reference): normalize=False, tau=1.0, alpha=0.5.  ``normalize=True`` divides the three
    inputs by their row norms with torch ops (O(B*d), autograd chains through) before the fused kernel."""

    def loss(self, z1, z2_con_z1, z3, z1_rec, z2_con_z1_rec, z3_rec):
        del z1, z2_con_z1, z3
        why = _cuda_domain((z1_rec, z2_con_z1_rec, z3_rec))
        if why is not None:
            return _delegate("SimCLRLoss", why,
                             lambda: super(SimCLRLoss, self).loss(None, None, None, z1_rec, z2_con_z1_rec, z3_rec))
        if self.normalize:
            z1_rec = z1_rec / torch.norm(z1_rec, p=2, dim=-1, keepdim=True)
            z2_con_z1_rec = z2_con_z1_rec / torch.norm(z2_con_z1_rec, p=2, dim=-1, keepdim=True)
            z3_rec = z3_rec / torch.norm(z3_rec, p=2, dim=-1, keepdim=True)
        # p = 0 selects the dot-product form of the kernel family: "distance" -z1.z3, positive "distance" -z1.z2
        mean, per_item, pos_mean, neg_mean = _F.lp_infonce(z1_rec, z2_con_z1_rec, z3_rec, 0.0, float(self.tau),
                                                           float(self.alpha), True)
        return mean, per_item, [pos_mean, neg_mean]


class LpSimCLRLoss(_Base):
    """Extended InfoNCE objective on an Lp norm (reference ``losses.py:405-477``), fused on sm_100a.

    Args (identical to the reference): p, tau=1.0, alpha=0.5, simclr_compatibility_mode=False, pow=True.
    """

    def _outside_domain(self, z1_rec, z2_rec, z3_rec):
        for t in (z1_rec, z2_rec, z3_rec):
            if not isinstance(t, torch.Tensor):
                return "non-tensor input"
            if not t.is_cuda:
                return "CPU tensors"
            if t.dtype != torch.float32:
                return f"dtype {t.dtype}"
            if t.dim() != 2:
                return f"{t.dim()}-D input"
        if self.p < 1.0:
            return "p < 1"
        if not self.pow:
            return "pow=False"
        if z1_rec.shape[1] > _MAX_D:
            return f"feature width {z1_rec.shape[1]} > {_MAX_D}"
        return None

    def loss(self, z1, z2_con_z1, z3, z1_rec, z2_con_z1_rec, z3_rec):
        del z1, z2_con_z1, z3           # unused by the reference as well (losses.py:431); may be None
        why = self._outside_domain(z1_rec, z2_con_z1_rec, z3_rec)
        if why is not None:
            return _delegate("LpSimCLRLoss", why,
                             lambda: super(LpSimCLRLoss, self).loss(None, None, None, z1_rec, z2_con_z1_rec, z3_rec))
        mean, per_item, pos_mean, neg_mean = _F.lp_infonce(
            z1_rec, z2_con_z1_rec, z3_rec, float(self.p), float(self.tau), float(self.alpha),
            bool(self.simclr_compatibility_mode))
        return mean, per_item, [pos_mean, neg_mean]
