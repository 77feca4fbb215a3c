"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sddp.h declares, its SddpConfig matches the ctypes mirror and the oracle's struct, and it
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

from srbd_horizon_b200 import _lib
from srbd_horizon_b200.config import MODEL_LIP, MODEL_SRBD, SddpConfig, make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "sddp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sddp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    """libsddp.so exports the whole header."""
    L = _lib.lib()
    declared = _header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(_lib.SYMBOLS) == declared
    assert L.sddp_abi_version() == 5 and L.sddp_config_size() == ctypes.sizeof(SddpConfig)


def test_f32_library_exports_the_same_header():
    """The optional fp32 build: same symbols, float arrays (sddp_real_bytes() == 4), same SddpConfig."""
    L = _lib.lib("f32")
    for name in _header_functions():
        assert hasattr(L, name), name
    assert L.sddp_abi_version() == 5 and L.sddp_real_bytes() == 4 and _lib.lib().sddp_real_bytes() == 8
    assert L.sddp_config_size() == ctypes.sizeof(SddpConfig)
    cfg = make_config(MODEL_SRBD, 50, 0.05)
    assert 0 < L.sddp_workspace_bytes(ctypes.byref(cfg)) < _lib.lib().sddp_workspace_bytes(ctypes.byref(cfg))


def test_inequality_options_are_checked_without_a_gpu():
    """One library serves the inequality extensions (friction cone, bounds) and the model scheduler; bad values are
    rejected by the config check (no GPU needed: sddp_workspace_bytes returns 0 for a config it refuses)."""
    L = _lib.lib()
    cfg = make_config(MODEL_SRBD, 10, 0.05, {"friction_cone_weight": 1.0, "force_bound_weight": 1.0, "lip_tail_start": 4})
    assert L.sddp_workspace_bytes(ctypes.byref(cfg)) > 0
    for name, bad in (("friction_cone_mu", -1.0), ("force_bound_weight", -1.0), ("bound_sharpness", float("inf")), ("lip_tail_start", 11),
                      ("mu_min", 0.0), ("mu_max", float("nan")), ("alpha_0", float("inf")), ("mu0", -1.0)):
        c2 = cfg.copy()
        setattr(c2, name, bad)
        assert L.sddp_workspace_bytes(ctypes.byref(c2)) == 0, name
    lip = make_config(MODEL_LIP, 10, 0.05)
    lip.lip_tail_start = 3
    assert L.sddp_workspace_bytes(ctypes.byref(lip)) == 0


def test_config_layout_matches_header_and_oracle():
    L = _lib.lib()
    assert L.sddp_config_size() == ctypes.sizeof(SddpConfig)
    hdr = open(os.path.join(ROOT, "include", "sddp.h")).read()
    orc = open(os.path.join(ROOT, "oracle", "sddp_oracle.h")).read()
    body = lambda s, name: re.sub(r"/\*.*?\*/", "", re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), s, re.S).group(1), flags=re.S)
    fields = lambda b: re.findall(r"(?:int32_t|double)\s+([^;]+);", b)
    f_h, f_o = fields(body(hdr, "SddpConfig")), fields(body(orc, "OrcConfig"))
    assert f_h == f_o
    flat = [n.strip().split("[")[0] for f in f_h for n in f.split(",")]
    assert flat == [n for n, _ in SddpConfig._fields_]


def test_dims_and_workspace_queries_need_no_gpu():
    L = _lib.lib()
    nx, nu, np_ = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert L.sddp_dims(MODEL_SRBD, nx, nu, np_) == 0 and (nx.value, nu.value, np_.value) == (37, 24, 19)
    assert L.sddp_dims(MODEL_LIP, nx, nu, np_) == 0 and (nx.value, nu.value, np_.value) == (30, 15, 11)
    assert L.sddp_dims(7, nx, nu, np_) == -1
    cfg = make_config(MODEL_SRBD, 50, 0.05)
    assert 0 < L.sddp_workspace_bytes(ctypes.byref(cfg)) < 2 << 30


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = make_config(MODEL_SRBD, 20, 0.05)
    h = ctypes.c_void_p()
    assert _lib.lib().sddp_create(ctypes.byref(cfg), ctypes.byref(h)) == -2
    assert b"no CPU path" in _lib.lib().sddp_last_error(None)
    from srbd_horizon_b200.ddp import BatchedDDP
    with pytest.raises(RuntimeError):
        BatchedDDP(cfg)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "srbd_horizon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'#include[^\n]*oracle|libsddp_oracle|^\s*(from|import)\s+oracle', text, re.M), f
