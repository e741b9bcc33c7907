"""clica_b200 -- B200 (sm_100a) implementation of cl-ica's InfoNCE training-step hot path.

Layout (only what the path needs):
  csrc/          hand-written CUDA kernels + the C ABI (include/clica.h) -> lib/libclica_sm100.so
  _lib.py        ctypes binding (fails loudly when the library is missing)
  functional.py  autograd.Function wrappers: lp_infonce, mlp_forward, adam_step
  dropin/        modules named like the reference's (losses.py, encoders.py) so main_mlp.py runs unchanged
  optim.py       FusedAdam (torch.optim.Optimizer API) on clica_adam_step
  graphed.py     GraphedTrainStep: the whole step as one CUDA graph (host batch in, loss scalars out)
  sharded.py     one-process-per-GPU step: batch shards + NCCL all-gather of encoder outputs
  launch.py      runs the reference's byte-identical main_mlp.py against the drop-in modules
"""
import os

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
DROPIN_DIR = os.path.join(PACKAGE_DIR, "dropin")

from . import _lib  # noqa: E402
from ._lib import ClicaError, build  # noqa: E402,F401


def __getattr__(name):
    # functional / optim / sharded import torch; keep `import clica_b200` itself light
    if name in ("functional", "optim", "sharded", "launch", "graphed", "vendor", "samplers", "synth"):
        import importlib
        return importlib.import_module("clica_b200." + name)
    raise AttributeError(name)
