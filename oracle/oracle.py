"""ctypes wrapper around oracle/libsddp_oracle.so -- TEST INFRASTRUCTURE ONLY.

The checker / reported CPU baseline; never on the product path.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from srbd_horizon_b200.config import DIMS, HIST, SddpConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsddp_oracle.so")
_lib = None
BUILD_FLAGS = "-O3 -march=x86-64-v3"      # of the library in use (oracle/Makefile; native_twin() changes it)


def native_twin() -> str:
    """For the TIMED CPU baseline of bench.py only ("a -O3 -march=native twin for timing"): compile the
    same source for the host CPU of THIS box into oracle/_native/ and use it from now on.  The shipped library is
    x86-64-v3 because it is built in another container; falls back to it (and says so) if gcc is missing or fails."""
    global _LIB_PATH, _lib, BUILD_FLAGS
    out_dir = os.path.join(_HERE, "_native")
    out = os.path.join(out_dir, "libsddp_oracle_native.so")
    try:
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-o", out, os.path.join(_HERE, "sddp_oracle.c"), "-lm", "-lpthread"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=120)
        _LIB_PATH, _lib, BUILD_FLAGS = out, None, "-O3 -march=native (compiled on this box)"
    except Exception:
        pass
    return BUILD_FLAGS

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sddp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsddp_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        cp = ctypes.POINTER(SddpConfig)
        L.orc_dynamics.argtypes = [cp, _dp, _dp, _dp]
        L.orc_dynamics_kind.argtypes = [cp, ctypes.c_int, _dp, _dp, _dp]
        L.orc_cost.argtypes = [cp, ctypes.c_int, _dp, _dp, _dp]
        L.orc_cost.restype = ctypes.c_double
        L.orc_derivs.argtypes = [cp, ctypes.c_int] + [_dp] * 10
        L.orc_solve.argtypes = [cp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp]
        L.orc_solve.restype = ctypes.c_int
        L.orc_solve_batch.argtypes = [cp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dp, ctypes.c_int]
        L.orc_total_cost.argtypes = [cp, _dp, _dp, _dp]
        L.orc_total_cost.restype = ctypes.c_double
        L.orc_backward.argtypes = [cp, _dp, _dp, _dp, _dp, ctypes.c_double, _dp, _dp, _dp]
        L.orc_backward.restype = ctypes.c_int
        L.orc_forward.argtypes = [cp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, ctypes.c_double, ctypes.c_double, _dp, _dp]
        L.orc_forward.restype = ctypes.c_double
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _c(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def dynamics(cfg: SddpConfig, x, u, kind: int = 1):
    """x + dt ode(x, u) of a node of the given kind (3: LIP-style tail node)."""
    nx, nu, _ = DIMS[cfg.model]
    x, u = _c(x, (nx,)), _c(u, (nu,))
    xn = np.empty(nx)
    lib().orc_dynamics_kind(ctypes.byref(cfg), int(kind), _p(x), _p(u), _p(xn))
    return xn


def cost(cfg: SddpConfig, kind: int, x, u, p) -> float:
    nx, nu, np_ = DIMS[cfg.model]
    x, p = _c(x, (nx,)), _c(p, (np_,))
    u = _c(u, (nu,)) if u is not None else np.zeros(nu)
    return lib().orc_cost(ctypes.byref(cfg), kind, _p(x), _p(u), _p(p))


def derivs(cfg: SddpConfig, kind: int, x, u, p):
    nx, nu, np_ = DIMS[cfg.model]
    x, p = _c(x, (nx,)), _c(p, (np_,))
    u = _c(u, (nu,)) if u is not None else np.zeros(nu)
    out = dict(fx=np.zeros((nx, nx)), fu=np.zeros((nx, nu)), lx=np.zeros(nx), lu=np.zeros(nu),
               lxx=np.zeros((nx, nx)), lux=np.zeros((nu, nx)), luu=np.zeros((nu, nu)))
    lib().orc_derivs(ctypes.byref(cfg), kind, _p(x), _p(u), _p(p), _p(out["fx"]), _p(out["fu"]),
                     _p(out["lx"]), _p(out["lu"]), _p(out["lxx"]), _p(out["lux"]), _p(out["luu"]))
    return out


def total_cost(cfg, X, U, params) -> float:
    X, U, params = _c(X), _c(U), _c(params)
    return lib().orc_total_cost(ctypes.byref(cfg), _p(X), _p(U), _p(params))


def backward(cfg, X, U, params, defect, mu):
    nx, nu, _ = DIMS[cfg.model]
    N = cfg.N
    X, U, params, defect = _c(X, (N + 1, nx)), _c(U, (N, nu)), _c(params), _c(defect, (N, nx))
    K = np.zeros((N, nu, nx)); kff = np.zeros((N, nu)); dV = np.zeros(3)
    rc = lib().orc_backward(ctypes.byref(cfg), _p(X), _p(U), _p(params), _p(defect), float(mu), _p(K), _p(kff), _p(dV))
    return rc, K, kff, dV


def forward(cfg, x0, X, U, params, defect, K, kff, alpha, rho):
    nx, nu, _ = DIMS[cfg.model]
    N = cfg.N
    x0, X, U, params, defect, K, kff = map(_c, (x0, X, U, params, defect, K, kff))
    Xn = np.zeros((N + 1, nx)); Un = np.zeros((N, nu))
    J = lib().orc_forward(ctypes.byref(cfg), _p(x0), _p(X), _p(U), _p(params), _p(defect), _p(K), _p(kff),
                          float(alpha), float(rho), _p(Xn), _p(Un))
    return J, Xn, Un


def solve_batch(cfg: SddpConfig, x0, params, X0, U0, nthreads: int = 1):
    """x0[B,nx], params[B,N+1,np], X0[B,N+1,nx] / U0[B,N,nu] warm starts.
    Returns dict(X, U, K, k, hist, iters, status, cost)."""
    nx, nu, np_ = DIMS[cfg.model]
    N = cfg.N
    x0 = _c(x0); B = x0.shape[0]
    assert x0.shape == (B, nx)
    params = _c(params, (B, N + 1, np_))
    X = _c(X0, (B, N + 1, nx)).copy(); U = _c(U0, (B, N, nu)).copy()
    K = np.zeros((B, N, nu, nx)); kff = np.zeros((B, N, nu))
    hist = np.zeros((B, cfg.max_iters, HIST))
    iters = np.zeros(B, dtype=np.int32); status = np.zeros(B, dtype=np.int32); cst = np.zeros(B)
    lib().orc_solve_batch(ctypes.byref(cfg), B, _p(x0), _p(params), _p(X), _p(U), _p(K), _p(kff), _p(hist),
                          iters.ctypes.data_as(_ip), status.ctypes.data_as(_ip), _p(cst), int(nthreads))
    return dict(X=X, U=U, K=K, k=kff, hist=hist, iters=iters, status=status, cost=cst)
