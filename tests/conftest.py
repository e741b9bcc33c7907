"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path.

``-m "not gpu"`` runs on the CPU-only build container (oracle vs golden fixtures, host logic,
C-ABI symbol check, gloo world_size-2 sharding logic).  ``-m gpu`` runs on a B200 and exercises the
CUDA path through the C-ABI library.
"""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def golden_loss_cases():
    return sorted(os.path.basename(p)[len("lpnce_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "lpnce_*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
