"""CPU (-m "not gpu"): the C-ABI library loads without a GPU and exports exactly what include/clica.h declares."""
import ctypes
import os
import re

import clica_b200
from clica_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "clica.h")).read()
    return sorted(set(re.findall(r"CLICA_API[^;(]*?\b(clica_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("clica_lpnce_fwd", "clica_lpnce_bwd", "clica_lpnce_bwd_sharded", "clica_linear_act_fwd",
                 "clica_linear_act_bwd_data", "clica_linear_bwd_weight", "clica_mlp_fwd", "clica_mlp_bwd",
                 "clica_adam_step", "clica_last_error", "clica_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in clica.h but not exported"


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.clica_abi_version() == 1
    assert lib.clica_last_error() is not None


def test_every_entry_point_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "clica.h")).read()
    for cite in ("losses.py:443-477", "losses.py:506-510", "encoders.py:38-48", "main_mlp.py:283"):
        assert cite in text
