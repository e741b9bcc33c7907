"""ctypes binding of libclica_sm100.so -- the only door between the Python host and the CUDA kernels.

Every function declared in ``include/clica.h`` is bound here with explicit argtypes.  Loading fails
LOUDLY (RuntimeError) when the library is missing or a symbol is absent: there is no Python/PyTorch
fallback for the CUDA path.  ``build()`` runs the in-tree Makefile (nvcc, sm_100a).
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "lib", "libclica_sm100.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

GEMM_3XTF32, GEMM_TF32, GEMM_FP32 = 0, 1, 3
GEMM_MODES = {"3xtf32": GEMM_3XTF32, "tf32": GEMM_TF32, "fp32": GEMM_FP32}

_c_int, _c_float, _c_size_t, _vp = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p
_i64 = ctypes.c_int64

# name -> (restype, argtypes); mirrors include/clica.h declaration by declaration
SIGNATURES = {
    "clica_abi_version": (_c_int, []),
    "clica_last_error": (ctypes.c_char_p, []),
    "clica_device_info": (_c_int, [_vp, _vp, _vp]),
    "clica_lpnce_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "clica_lpnce_fwd": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _c_int, _c_int, _c_int, _c_int, _c_float,
                                 _c_float, _c_float, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _c_size_t, _vp]),
    "clica_lpnce_bwd_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "clica_lpnce_bwd": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _c_int, _c_int, _c_int, _c_int, _c_float,
                                 _c_float, _c_float, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp, _c_int, _vp,
                                 _c_int, _vp, _c_int, _vp, _c_size_t, _vp]),
    "clica_lpnce_bwd_sharded_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "clica_lpnce_bwd_sharded": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _c_int, _vp, _vp, _c_int, _c_int,
                                         _c_int, _c_int, _c_float, _c_float, _c_float, _c_int, _vp, _vp,
                                         _c_int, _vp, _c_int, _vp, _c_size_t, _vp]),
    "clica_linear_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "clica_linear_act_fwd": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _vp, _c_int, _c_int, _c_int, _c_int,
                                      _c_float, _c_int, _vp, _c_size_t, _vp]),
    "clica_linear_act_bwd_data": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _c_int, _c_float, _vp, _c_int,
                                           _c_int, _c_int, _c_int, _c_int, _vp, _c_size_t, _vp]),
    "clica_linear_bwd_weight": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _c_int, _vp, _c_int, _c_int, _c_int,
                                         _c_int, _vp, _c_size_t, _vp]),
    "clica_mlp_act_floats": (_c_size_t, [_c_int, _c_int, _c_int]),
    "clica_mlp_workspace_bytes": (_c_size_t, [_c_int, _c_int, _vp, _c_int]),
    "clica_mlp_packed_weight_bytes": (_c_size_t, [_c_int, _vp, _c_int]),
    "clica_mlp_pack_weights": (_c_int, [_c_int, _vp, _vp, _c_int, _vp, _c_size_t, _vp]),
    "clica_mlp_fwd": (_c_int, [_c_int, _vp, _vp, _vp, _vp, _c_int, _c_float, _c_int, _vp, _vp, _c_size_t, _vp]),
    "clica_mlp_bwd": (_c_int, [_c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_float, _c_int, _vp, _c_int,
                               _vp, _c_size_t, _vp]),
    "clica_mlp_bwd_range": (_c_int, [_c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_float, _c_int, _vp, _c_int,
                                     _c_int, _c_int, _vp, _c_size_t, _vp]),
    "clica_mixing_fwd": (_c_int, [_vp, _c_int, _vp, _c_int, _c_int, _c_int, _c_float, _vp, _c_int, _vp]),
    "clica_sample_latents": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _c_int, _c_int, _c_float,
                                      _c_float, _c_float, _c_float, ctypes.c_uint64, ctypes.c_uint64, _vp]),
    "clica_tc_set_sm_reserve": (_c_int, [_c_int]),
    "clica_launch_count": (ctypes.c_longlong, [_c_int]),
    "clica_prof_enable": (_c_int, [_c_int]),
    "clica_prof_collect": (_c_int, [_vp, _vp]),
    "clica_adam_step": (_c_int, [_c_int, _vp, _vp, _vp, _vp, _vp, _c_float, _c_float, _c_float, _c_float,
                                 _i64, _c_float, _vp]),
    "clica_adam_step_capturable": (_c_int, [_c_int, _vp, _vp, _vp, _vp, _vp, _c_float, _c_float, _c_float,
                                            _c_float, _vp, _c_float, _vp]),
    "clica_adam_step_capturable_packed": (_c_int, [_c_int, _vp, _vp, _vp, _vp, _vp, _c_float, _c_float, _c_float,
                                                   _c_float, _vp, _c_float, _vp, _vp, _vp, _vp, _vp]),
    "clica_mlp_packed_weight_layout": (_c_int, [_c_int, _vp, _c_int, _vp, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


class ClicaError(RuntimeError):
    """A libclica_sm100 call returned a non-zero status."""


def build(verbose: bool = False, jobs: int = 8) -> str:
    """Compile every CUDA source for sm_100a into ``cl-ica_b200/lib/libclica_sm100.so`` (in-tree)."""
    res = subprocess.run(["make", "-C", CSRC_DIR, f"-j{jobs}"], capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libclica_sm100.so failed:\n" + (res.stdout or "") + (res.stderr or ""))
    return LIB_PATH


def load():
    """Load the shared library (once) and bind every declared symbol."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension is not built (run `python -c 'import "
                "__graft_entry__ as g; g.build()'` or `make -C cl-ica_b200/csrc`). There is no fallback path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise RuntimeError(f"libclica_sm100.so does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.clica_abi_version() != 1:
            raise RuntimeError("libclica_sm100.so ABI version mismatch")
        _lib = lib
        return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().clica_last_error()
        raise ClicaError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def gemm_mode_from_env(default: str = "3xtf32") -> int:
    name = os.environ.get("CLICA_GEMM_MODE", default).lower()
    if name not in GEMM_MODES:
        raise ValueError(f"CLICA_GEMM_MODE must be one of {sorted(GEMM_MODES)}, got {name!r}")
    return GEMM_MODES[name]
