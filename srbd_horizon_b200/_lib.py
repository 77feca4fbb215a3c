"""ctypes binding of csrc/libsddp.so (include/sddp.h).  There is no CPU fallback: if the
library is missing or no CUDA device is present the product path raises."""
from __future__ import annotations

import ctypes
import os
import subprocess

from .config import SddpConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("SDDP_LIB", os.path.join(CSRC, "libsddp.so"))   # SDDP_LIB: A/B builds while tuning
#: the optional fp32 build of the same sources (-DSDDP_F32): float arrays at the ABI, narrower than the reference
LIB_PATH_F32 = os.environ.get("SDDP_LIB_F32", os.path.join(CSRC, "libsddp_f32.so"))

_vp = ctypes.c_void_p
_ip = ctypes.POINTER(ctypes.c_int)

#: every symbol include/sddp.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sddp_abi_version": (ctypes.c_int, []),
    "sddp_real_bytes": (ctypes.c_int, []),
    "sddp_config_size": (ctypes.c_size_t, []),
    "sddp_dims": (ctypes.c_int, [ctypes.c_int, _ip, _ip, _ip]),
    "sddp_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(SddpConfig)]),
    "sddp_create": (ctypes.c_int, [ctypes.POINTER(SddpConfig), ctypes.POINTER(_vp)]),
    "sddp_destroy": (ctypes.c_int, [_vp]),
    "sddp_last_error": (ctypes.c_char_p, [_vp]),
    "sddp_set_config": (ctypes.c_int, [_vp, ctypes.POINTER(SddpConfig)]),
    "sddp_eval_derivatives": (ctypes.c_int, [_vp, ctypes.c_int] + [_vp] * 13 + [_vp]),
    "sddp_solve_batch": (ctypes.c_int, [_vp, ctypes.c_int] + [_vp] * 10 + [_vp]),
    "sddp_backward_pass": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_double, _vp, _vp, _vp, _vp, _vp]),
    "sddp_forward_pass": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int] + [_vp] * 12 + [_vp]),
    "sddp_defects": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sddp_solve_batch_host": (ctypes.c_int, [_vp, ctypes.c_int] + [_vp] * 12),
    "sddp_set_gait_tables": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "sddp_mpc_advance": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    "sddp_plant_step": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_longlong, _vp]),
    "sddp_fp64_peak_tflops": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), _vp]),
    "sddp_set_dispatch_order": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int]),
    "sddp_launch_count": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_longlong)]),
    "sddp_record_doubles": (ctypes.c_longlong, [_vp]),
    "sddp_slab_alloc": (ctypes.c_int, [_vp, ctypes.c_longlong, ctypes.POINTER(_vp)]),
    "sddp_ipc_export": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "sddp_ipc_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "sddp_ipc_close": (ctypes.c_int, [_vp]),
    "sddp_set_result_peers": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(_vp), ctypes.c_longlong]),
}

_libs = {}


def build(force: bool = False) -> str:
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in ("sddp.cu", "sddp_solver.cuh", "sddp_backward_srbd.cuh", "sddp_model.cuh", "Makefile")]
    srcs.append(os.path.join(_HERE, "..", "include", "sddp.h"))
    outs = [os.path.join(CSRC, "libsddp.so"), os.path.join(CSRC, "libsddp_f32.so")]
    stale = any((not os.path.exists(o)) or any(os.path.getmtime(s) > os.path.getmtime(o) for s in srcs) for o in outs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-B", "-j2", "all"])      # the two libraries in parallel
    return LIB_PATH


def lib(dtype: str = "f64") -> ctypes.CDLL:
    """The product library (csrc/libsddp.so, or the A/B build named by SDDP_LIB); dtype="f32": the optional fp32 build."""
    if dtype not in ("f64", "f32"):
        raise ValueError("dtype: 'f64' or 'f32'")
    path = LIB_PATH if dtype == "f64" else LIB_PATH_F32
    if path not in _libs:
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = ctypes.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.sddp_abi_version() != 5:
            raise RuntimeError(f"{os.path.basename(path)} ABI version mismatch")
        if L.sddp_real_bytes() != (8 if dtype == "f64" else 4):
            raise RuntimeError(f"{os.path.basename(path)} is not the {dtype} build")
        if L.sddp_config_size() != ctypes.sizeof(SddpConfig):
            raise RuntimeError("SddpConfig layout mismatch between config.py and include/sddp.h")
        _libs[path] = L
    return _libs[path]


class SddpError(RuntimeError):
    pass


def check(rc: int, handle=None, L=None) -> None:
    if rc != 0:
        msg = (L or lib()).sddp_last_error(handle)
        raise SddpError(f"sddp error {rc}: {msg.decode() if msg else '?'}")
