"""ctypes loader for oracle/lpnce_oracle.c  (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

``build()`` compiles the C restatement with gcc (-O2 -fopenmp) into ``oracle/liblpnce_oracle.so``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "lpnce_oracle.c")
_SO = os.path.join(_HERE, "liblpnce_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", _SO, _SRC, "-lm"]
        subprocess.run(cmd, check=True, cwd=_HERE)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        fp, dp, ci, cd = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double),
                          ctypes.c_int, ctypes.c_double)
        lib.lpnce_oracle_fwd.argtypes = [fp, ci, fp, ci, fp, ci, ci, ci, ci, cd, cd, cd, ci, ci,
                                         dp, dp, dp, dp]
        lib.lpnce_oracle_fwd.restype = ci
        lib.lpnce_oracle_bwd.argtypes = [fp, ci, fp, ci, fp, ci, ci, ci, ci, cd, cd, cd, ci, ci,
                                         dp, dp, dp, dp, dp, dp]
        lib.lpnce_oracle_bwd.restype = ci
        lib.lpnce_oracle_num_threads.restype = ci
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().lpnce_oracle_num_threads())


def _f32(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    assert a.ndim == 2
    return a


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _dptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def lpnce(z1, z2, z3, p, tau=1.0, alpha=0.5, include_pos=True, use_pow=True, gl=None,
          need_grad=True):
    """float64 loss/grad of the Lp-InfoNCE objective for fp32 inputs z1,z2:[B,d], z3:[M,d].

    Returns a dict with loss_mean, loss_i, lse, pos, pos_mean, neg_mean and (need_grad) g1, g2, g3,
    the gradients of ``sum_i gl[i]*loss_i`` (gl=None -> the mean), g1 holding only the anchor-row
    contribution (add the un-rolled g3 yourself when z3 is a roll of z1).
    """
    lib = _load()
    z1, z2, z3 = _f32(z1), _f32(z2), _f32(z3)
    B, d = z1.shape
    M = z3.shape[0]
    assert z2.shape == (B, d) and z3.shape[1] == d
    loss_i, lse, pos = np.zeros(B), np.zeros(B), np.zeros(B)
    scal = np.zeros(3)
    rc = lib.lpnce_oracle_fwd(_fptr(z1), d, _fptr(z2), d, _fptr(z3), d, B, M, d, float(p),
                              float(tau), float(alpha), int(bool(include_pos)), int(bool(use_pow)),
                              _dptr(loss_i), _dptr(lse), _dptr(pos), _dptr(scal))
    if rc != 0:
        raise ValueError("lpnce_oracle_fwd rejected its arguments")
    out = dict(loss_mean=scal[0], pos_mean=scal[1], neg_mean=scal[2], loss_i=loss_i, lse=lse,
               pos=pos)
    if need_grad:
        g1, g2, g3 = np.zeros((B, d)), np.zeros((B, d)), np.zeros((M, d))
        glc = None if gl is None else np.ascontiguousarray(np.asarray(gl, dtype=np.float64))
        rc = lib.lpnce_oracle_bwd(_fptr(z1), d, _fptr(z2), d, _fptr(z3), d, B, M, d, float(p),
                                  float(tau), float(alpha), int(bool(include_pos)),
                                  int(bool(use_pow)), _dptr(lse), _dptr(pos), _dptr(glc),
                                  _dptr(g1), _dptr(g2), _dptr(g3))
        if rc != 0:
            raise ValueError("lpnce_oracle_bwd rejected its arguments")
        out.update(g1=g1, g2=g2, g3=g3)
    return out
