"""Device-side latent samplers behind the reference's ``spaces`` surface.

The reference draws its latents with host-bound code: ``NSphereSpace.normal`` syncs on an ``allclose`` assert
(``spaces.py:162-164``), ``NBoxSpace.*`` loops over boolean-mask assignments with one ``.item()`` per round
(``spaces_utils.py:127-142``), ``generalized_normal`` samples its Gamma variates on the CPU (``spaces_utils.py:96-103``).
Once the training step itself takes < 1 ms those dominate a step of the unchanged ``main_mlp.py``.

This module provides ``NRealSpace`` / ``NSphereSpace`` / ``NBoxSpace`` with the reference's constructor and method
signatures (``uniform(size, device)``, ``normal(mean, std, size, device)``, ``laplace(mean, lbd, size, device)``,
``generalized_normal(mean, lbd, p, size, device)``); for a CUDA ``device`` each call is ONE kernel launch
(``clica_sample_latents``: Philox4x32-10, no host synchronisation), for anything else -- CPU devices, tensor-valued
``std``, ``von_mises_fisher`` -- the reference's own method runs.  Same distributions, a different random stream: results
are reproducible under ``torch.manual_seed`` (the seed keys the generator, a per-process call counter is the offset) but
not bit-identical to the reference's draws.

``install()`` registers a module named ``spaces`` that re-exports the reference's names with these three classes in
place, so that the unchanged script's ``import spaces`` picks them up
(``python -m clica_b200.launch --device-samplers -- <main_mlp.py args>``).
"""
import itertools
import sys
import types

import torch

from . import _lib

_calls = itertools.count(1)
SPACE_REAL, SPACE_SPHERE, SPACE_BOX = 0, 1, 2
DIST_UNIFORM, DIST_NORMAL, DIST_LAPLACE, DIST_GNORMAL = 0, 1, 2, 3


def _is_cuda(device):
    try:
        return torch.device(device).type == "cuda"
    except Exception:
        return False


def sample(size, n, space, dist, device, mean=None, scale=1.0, p=2.0, box=(-1.0, 1.0)):
    """``[size, n]`` fp32 samples on the CUDA ``device`` in one launch (see ``clica_sample_latents``)."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty((int(size), int(n)), dtype=torch.float32, device=dev)
    mean_rows, ld_mean, mptr = 0, 0, None
    if mean is not None:
        mean = mean.to(device=dev, dtype=torch.float32)
        if mean.dim() == 1:
            mean = mean.unsqueeze(0)
        if mean.dim() != 2 or mean.shape[1] != n or mean.shape[0] not in (1, size):
            raise ValueError(f"mean must be [n], [1, n] or [size, n]; got {tuple(mean.shape)} for size={size}, n={n}")
        mean = mean.contiguous()
        mean_rows, ld_mean, mptr = mean.shape[0] if mean.shape[0] == 1 else int(size), n, mean.data_ptr()
        if size == 1:
            mean_rows = 1
    seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
    with torch.cuda.device(dev):
        rc = lib.clica_sample_latents(out.data_ptr(), n, int(size), int(n), int(space), int(dist), mptr, ld_mean,
                                      mean_rows, float(scale), float(p), float(box[0]), float(box[1]), seed,
                                      next(_calls), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "clica_sample_latents")
    return out


def _scalar(x):
    return isinstance(x, (int, float))


def _make_classes(ref):
    """Subclasses of the reference's space classes (same constructors) with device-side sampling methods."""

    class NRealSpace(ref.NRealSpace):
        def normal(self, mean, std, size, device="cpu"):
            if _is_cuda(device) and _scalar(std):
                return sample(size, self.n, SPACE_REAL, DIST_NORMAL, device, mean, std)
            return super().normal(mean, std, size, device)

        def laplace(self, mean, lbd, size, device="cpu"):
            if _is_cuda(device) and _scalar(lbd):
                return sample(size, self.n, SPACE_REAL, DIST_LAPLACE, device, mean, lbd)
            return super().laplace(mean, lbd, size, device)

        def generalized_normal(self, mean, lbd, p, size, device=None):
            if device is not None and _is_cuda(device) and _scalar(lbd):
                return sample(size, self.n, SPACE_REAL, DIST_GNORMAL, device, mean, lbd, p)
            return super().generalized_normal(mean, lbd, p, size, device)

    class NSphereSpace(ref.NSphereSpace):
        # the reference projects onto the UNIT sphere whatever `r` is (spaces.py:137,167); so does the kernel
        def uniform(self, size, device="cpu"):
            if _is_cuda(device):
                return sample(size, self.n, SPACE_SPHERE, DIST_UNIFORM, device)
            return super().uniform(size, device)

        def normal(self, mean, std, size, device="cpu"):
            if _is_cuda(device) and _scalar(std):
                return sample(size, self.n, SPACE_SPHERE, DIST_NORMAL, device, mean, std)
            return super().normal(mean, std, size, device)

        def laplace(self, mean, lbd, size, device="cpu"):
            if _is_cuda(device) and _scalar(lbd):
                return sample(size, self.n, SPACE_SPHERE, DIST_LAPLACE, device, mean, lbd)
            return super().laplace(mean, lbd, size, device)

        def generalized_normal(self, mean, lbd, p, size, device="cpu"):
            if _is_cuda(device) and _scalar(lbd):
                return sample(size, self.n, SPACE_SPHERE, DIST_GNORMAL, device, mean, lbd, p)
            return super().generalized_normal(mean, lbd, p, size, device)

    class NBoxSpace(ref.NBoxSpace):
        def uniform(self, size, device="cpu"):
            if _is_cuda(device):
                return sample(size, self.n, SPACE_BOX, DIST_UNIFORM, device, box=(self.min_, self.max_))
            return super().uniform(size, device)

        def normal(self, mean, std, size, device="cpu"):
            if _is_cuda(device) and _scalar(std):
                return sample(size, self.n, SPACE_BOX, DIST_NORMAL, device, mean, std, box=(self.min_, self.max_))
            return super().normal(mean, std, size, device)

        def laplace(self, mean, lbd, size, device="cpu"):
            if _is_cuda(device) and _scalar(lbd):
                return sample(size, self.n, SPACE_BOX, DIST_LAPLACE, device, mean, lbd, box=(self.min_, self.max_))
            return super().laplace(mean, lbd, size, device)

        def generalized_normal(self, mean, lbd, p, size, device=None):
            if device is not None and _is_cuda(device) and _scalar(lbd):
                return sample(size, self.n, SPACE_BOX, DIST_GNORMAL, device, mean, lbd, p, box=(self.min_, self.max_))
            return super().generalized_normal(mean, lbd, p, size, device)

    return NRealSpace, NSphereSpace, NBoxSpace


def build_module(reference_spaces):
    """A module object with the reference ``spaces`` names, the three space classes replaced."""
    mod = types.ModuleType("spaces")
    mod.__dict__.update({k: v for k, v in reference_spaces.__dict__.items() if not k.startswith("__")})
    mod.NRealSpace, mod.NSphereSpace, mod.NBoxSpace = _make_classes(reference_spaces)
    mod.__doc__ = __doc__
    mod.__clica_device_samplers__ = True
    return mod


def install():
    """Make ``import spaces`` resolve to the device-sampler module (the reference's ``spaces`` must be importable,
    i.e. the reference directory is on ``sys.path`` -- ``clica_b200.launch`` arranges that)."""
    cur = sys.modules.get("spaces")
    if cur is not None and getattr(cur, "__clica_device_samplers__", False):
        return cur
    import importlib
    ref = importlib.import_module("spaces")
    if "NSphereSpace" not in ref.__dict__:
        raise RuntimeError("the module named `spaces` on sys.path is not the reference's spaces.py")
    mod = build_module(ref)
    sys.modules["spaces"] = mod
    return mod
