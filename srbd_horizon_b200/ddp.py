"""DDP solver: the host-side mirror of /root/reference/python/ddp.py (class DDPSolver).

Reference surface kept (ddp.py:10-123): `DDPSolver(prb, opts)`, `solve() -> bool`,
`getSolutionDict()`, `setInitialState(x0)`, `set_x_warmstart(x)`, `set_u_warmstart(u)`; the
solution dict holds one `dim x nodes` array per variable plus 'x_opt' / 'u_opt' (ddp.py:125-151).
`opts` accepts the reference's keys (ddp.py:17-35) and the extensions of config.DEFAULT_OPTS.

Where the reference hands CasADi functions to the external `pyddp` module, this class calls the
C ABI of csrc/libsddp.so (include/sddp.h) on torch CUDA tensors.  Additive batch API:
`solve_batch(x0[B,nx], params[B,N+1,np], X0, U0)` on device tensors and `solve_batch_host` on
numpy arrays (host <-> device copies inside the call).

There is no CPU path: constructing a solver without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .config import DIMS, HIST, MODEL_LIP, MODEL_SRBD, STATUS_CONVERGED, Gains, RobotConstants, SddpConfig, make_config


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _np_ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


@dataclass
class BatchResult:
    X: torch.Tensor          # [B, N+1, nx]
    U: torch.Tensor          # [B, N, nu]
    K: Optional[torch.Tensor]    # [B, N, nu, nx] feedback gains of the last backward pass
    k: Optional[torch.Tensor]    # [B, N, nu]
    hist: Optional[torch.Tensor]  # [B, max_iters, 4]: cost, alpha, mu, max|defect|
    iters: torch.Tensor      # [B] int32
    status: torch.Tensor     # [B] int32, 0 = converged
    cost: torch.Tensor       # [B]


class BatchedDDP:
    """Thin owner of one `SddpHandle` (one device, one stream at a time)."""

    def __init__(self, cfg: SddpConfig, device: Optional[torch.device] = None, dtype: str = "f64"):
        """dtype: "f64" (the product library, fp64 like the reference) or "f32" (the optional fp32 build libsddp_f32.so:
        float arrays, float Riccati recursion; narrower than the reference, tolerance in tests/test_gpu_f32.py)."""
        if not torch.cuda.is_available():
            raise RuntimeError("srbd_horizon_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = _lib.lib(dtype)
        self.dtype = dtype
        self.tdtype = torch.float64 if dtype == "f64" else torch.float32
        self.ndtype = np.float64 if dtype == "f64" else np.float32
        self.cfg = cfg.copy()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.nx, self.nu, self.np = DIMS[cfg.model]
        self.N = cfg.N
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.L.sddp_create(ctypes.byref(self.cfg), ctypes.byref(h)), None, self.L)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.sddp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **opts):
        for k, v in opts.items():
            if not hasattr(self.cfg, k):
                raise KeyError(k)
            setattr(self.cfg, k, v)
        _lib.check(self.L.sddp_set_config(self.h, ctypes.byref(self.cfg)), self.h, self.L)

    # -- helpers ------------------------------------------------------------------------------
    def _t(self, a, shape, name):
        t = torch.as_tensor(a, dtype=self.tdtype, device=self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launches(self) -> int:
        n = ctypes.c_longlong()
        self.L.sddp_launch_count(self.h, ctypes.byref(n))
        return n.value

    # -- the solve ----------------------------------------------------------------------------
    # columns of the per-node parameter vector that hold the contact switches (prb.py:159-163 / 372-376)
    _SWITCH_COLS = {MODEL_SRBD: (8, 10, 12, 14), MODEL_LIP: (4, 6, 8, 10)}

    def dispatch_order(self, params, nodes: int = 24):
        """Dispatch order that puts problems with the same contact schedule next to each other (a hash of the switch
        pattern over the first `nodes` nodes), the fastest commanded velocities first inside a schedule.  Scheduling only: co-resident CTAs then stay in step and
        share their instructions in the SM's instruction cache.  Works on a device tensor or a numpy array."""
        cols = list(self._SWITCH_COLS[self.cfg.model])
        n = min(nodes, params.shape[1])
        sl = slice(cols[0], cols[-1] + 1, cols[1] - cols[0])          # the switch columns as a strided view (no copy)
        # Schedules with more swing phases need more iterations (stepping 5.8 on average, up to 14; a flight phase 4.8;
        # stance 4.3 on BASELINE configs[4]): they go first, so that the tail of the batch, when CTAs run out of work,
        # consists of short problems.  Key = -(number of swing entries) * 1e7 + hash of the pattern (< 1e7).
        if isinstance(params, np.ndarray):
            w = (1.0 + np.arange(n * len(cols), dtype=np.float64).reshape(n, len(cols))) ** 2
            sw = params[:, :n, sl].astype(np.float64, copy=False)      # (float32 parameters of the fp32 build: the key needs 27 bits)
            key = np.einsum("bnc,nc->b", sw, w) - 1e7 * (1.0 - sw).sum(axis=(1, 2))
            return self.order_from_keys(key, (params[:, -1, 0:3] ** 2).sum(axis=1))
        w = (1.0 + torch.arange(n * len(cols), dtype=torch.float64, device=params.device).reshape(n, len(cols))) ** 2
        sw = params[:, :n, sl].to(torch.float64)      # (float32 parameters of the fp32 build: the key needs 27 bits)
        return self.order_from_keys((sw * w).sum(dim=(1, 2)) - 1e7 * (1.0 - sw).sum(dim=(1, 2)), (params[:, -1, 0:3] ** 2).sum(dim=1))

    @staticmethod
    def order_from_keys(group, effort):
        """Permutation that sorts by `group` (ascending) and, inside a group, by `effort` (descending).  `effort` is a
        guess of how many iterations a problem needs -- by default the commanded velocity |rdot_ref| (p[0:3] of the last
        node, prb.py:74), which correlates +0.5 with the iteration count inside a schedule (tools/iters_features.py)."""
        if isinstance(group, np.ndarray):
            effort = np.asarray(effort, dtype=np.float64)
            if np.issubdtype(group.dtype, np.integer) and group.size and int(group.max()) - int(group.min()) < 64:
                # few integer groups (the caller's schedule ids): one 16-bit key = group | quantised effort, which numpy sorts
                # with a radix sort (two stable argsorts of 64K keys cost 17 ms, 3.5 % of an end-to-end step; this is 1 ms)
                lo, hi = float(effort.min()), float(effort.max())
                q = ((hi - effort) * (1023.0 / (hi - lo if hi > lo else 1.0))).astype(np.uint16)      # largest effort first
                key = ((group - group.min()).astype(np.uint16) << 10) | q
                return np.argsort(key, kind="stable").astype(np.int32)
            o1 = np.argsort(-effort, kind="stable")
            return o1[np.argsort(np.asarray(group)[o1], kind="stable")].astype(np.int32)
        o1 = torch.argsort(effort, descending=True, stable=True)
        return o1[torch.argsort(group[o1], stable=True)].to(torch.int32)

    def solve(self, x0, params, X0, U0, gains: bool = True, history: bool = True, inplace: bool = False, order=None, gather=None,
              out: Optional[BatchResult] = None) -> BatchResult:
        """x0[B,nx], params[B,N+1,np], warm starts X0[B,N+1,nx], U0[B,N,nu] (device tensors).
        order: None (natural dispatch order), "schedule" (`dispatch_order(params)`) or an int32 device permutation.
        gather: a `parallel.ResultGather` of this solver: the kernel also stores every problem's result record into the
        whole-batch slabs it names (multi-GPU gather); `gather.finish()` afterwards returns the whole-batch views.
        out: a BatchResult of an earlier call with the same shapes: its K, k, hist, iters, status, cost tensors are
        overwritten instead of allocating new ones (K is 7.1 KB per node: a steady-state loop should not allocate it per call)."""
        x0 = torch.as_tensor(x0, dtype=self.tdtype, device=self.device).contiguous()
        B = x0.shape[0]
        N, nx, nu, np_ = self.N, self.nx, self.nu, self.np
        x0 = self._t(x0, (B, nx), "x0")
        params = self._t(params, (B, N + 1, np_), "params")
        X = self._t(X0, (B, N + 1, nx), "X0")
        U = self._t(U0, (B, N, nu), "U0")
        if not inplace:
            X, U = X.clone(), U.clone()
        dev = self.device

        def buf(name, shape, dtype):
            t = getattr(out, name, None) if out is not None else None
            if t is not None and tuple(t.shape) == tuple(shape) and t.dtype == dtype and t.device == dev and t.is_contiguous():
                return t
            return torch.empty(shape, dtype=dtype, device=dev)
        K = buf("K", (B, N, nu, nx), self.tdtype) if gains else None
        k = buf("k", (B, N, nu), self.tdtype) if gains else None
        hist = buf("hist", (B, self.cfg.max_iters, HIST), self.tdtype) if history else None
        iters = buf("iters", (B,), torch.int32)
        status = buf("status", (B,), torch.int32)
        cost = buf("cost", (B,), self.tdtype)
        user_order = order is not None and not isinstance(order, str)
        if isinstance(order, str):
            if order != "schedule":
                raise ValueError("order: None, 'schedule' or a permutation")
            order = self.dispatch_order(params) if B > 1 else None
        if order is not None:
            order = torch.as_tensor(order, dtype=torch.int32, device=dev).contiguous()
            if order.shape != (B,):
                raise ValueError("order must have one entry per problem")
            # a caller's permutation is checked here: the kernel skips entries outside 0..B-1 and a problem that no entry
            # names stays unsolved (status -1), but it cannot tell the caller (include/sddp.h, sddp_set_dispatch_order)
            if user_order and not bool(torch.equal(torch.sort(order.to(torch.int64)).values, torch.arange(B, device=dev))):
                raise ValueError("order must be a permutation of 0..B-1")
        with torch.cuda.device(dev):
            if gather is not None:
                if gather.solver is not self:
                    raise ValueError("gather belongs to another solver")
                gather.arm(B)
            if order is not None:
                _lib.check(self.L.sddp_set_dispatch_order(self.h, _ptr(order), B, 0), self.h, self.L)
            try:
                _lib.check(self.L.sddp_solve_batch(self.h, B, _ptr(x0), _ptr(params), _ptr(X), _ptr(U), _ptr(K), _ptr(k),
                                                   _ptr(hist), _ptr(iters), _ptr(status), _ptr(cost), self._stream()), self.h, self.L)
            finally:
                if order is not None:
                    self.L.sddp_set_dispatch_order(self.h, None, 0, 0)
                if gather is not None:
                    gather.disarm()
        r = BatchResult(X, U, K, k, hist, iters, status, cost)
        r._keepalive = order      # the kernel reads the permutation asynchronously
        return r

    def solve_host(self, x0: np.ndarray, params: np.ndarray, X0: np.ndarray, U0: np.ndarray, gains: bool = False,
                   history: bool = False, out: Optional[Dict[str, np.ndarray]] = None, order=None) -> Dict[str, np.ndarray]:
        """Same solve on HOST numpy buffers through `sddp_solve_batch_host`.  With pinned buffers (inputs and `out`) and no K
        the whole batch is one launch whose CTAs move their own inputs and results over PCIe; otherwise chunked staged
        copies overlap the solves.  `out` may hold preallocated (pinned) X, U, iters, status, cost [, K, k, hist] arrays."""
        B = x0.shape[0]
        N, nx, nu, np_ = self.N, self.nx, self.nu, self.np
        c = lambda a: np.ascontiguousarray(a, dtype=self.ndtype)
        x0, params, X0, U0 = c(x0), c(params), c(X0), c(U0)
        if x0.shape != (B, nx) or params.shape != (B, N + 1, np_) or X0.shape != (B, N + 1, nx) or U0.shape != (B, N, nu):
            raise ValueError("solve_host: bad input shapes")
        out = dict(out) if out else {}
        def buf(name, shape, dtype=None):
            dtype = dtype or self.ndtype
            a = out.get(name)
            if a is None or a.shape != tuple(shape) or a.dtype != dtype or not a.flags.c_contiguous:
                a = np.empty(shape, dtype=dtype)
            out[name] = a
            return a
        X, U = buf("X", (B, N + 1, nx)), buf("U", (B, N, nu))
        K = buf("K", (B, N, nu, nx)) if gains is True else None      # gains="ff": the feed-forward term k only
        k = buf("k", (B, N, nu)) if gains else None
        hist = buf("hist", (B, self.cfg.max_iters, HIST)) if history else None
        iters, status, cost = buf("iters", (B,), np.int32), buf("status", (B,), np.int32), buf("cost", (B,))
        if isinstance(order, str):
            if order != "schedule":
                raise ValueError("order: None, 'schedule' or a permutation")
            order = self.dispatch_order(params) if B > 1 else None
        if order is not None:
            order = np.ascontiguousarray(order, dtype=np.int32)
            if order.shape != (B,):
                raise ValueError("order must have one entry per problem")
        with torch.cuda.device(self.device):
            if order is not None:
                _lib.check(self.L.sddp_set_dispatch_order(self.h, _np_ptr(order), B, 1), self.h, self.L)
            try:
                _lib.check(self.L.sddp_solve_batch_host(self.h, B, _np_ptr(x0), _np_ptr(params), _np_ptr(X0), _np_ptr(U0), _np_ptr(X),
                                                        _np_ptr(U), _np_ptr(K), _np_ptr(k), _np_ptr(hist), _np_ptr(iters), _np_ptr(status),
                                                        _np_ptr(cost)), self.h, self.L)
            finally:
                if order is not None:
                    self.L.sddp_set_dispatch_order(self.h, None, 0, 1)
        out.update(K=K, k=k, hist=hist)
        return out

    # -- stage entry points (used by the stage parity tests) ------------------------------------
    def eval_derivatives(self, kind, x, u, p):
        M = len(kind)
        nx, nu, np_ = self.nx, self.nu, self.np
        dev = self.device
        kind = torch.as_tensor(kind, dtype=torch.int32, device=dev).contiguous()
        x = self._t(x, (M, nx), "x"); u = self._t(u, (M, nu), "u"); p = self._t(p, (M, np_), "p")
        z = lambda *s: torch.empty(s, dtype=self.tdtype, device=dev)
        out = dict(f=z(M, nx), fx=z(M, nx, nx), fu=z(M, nx, nu), l=z(M), lx=z(M, nx), lu=z(M, nu), lxx=z(M, nx, nx),
                   lux=z(M, nu, nx), luu=z(M, nu, nu))
        with torch.cuda.device(dev):
            _lib.check(self.L.sddp_eval_derivatives(self.h, M, _ptr(kind), _ptr(x), _ptr(u), _ptr(p), *[_ptr(out[n]) for n in
                       ("f", "fx", "fu", "l", "lx", "lu", "lxx", "lux", "luu")], self._stream()), self.h, self.L)
        return out

    def backward_pass(self, X, U, params, defect, mu: float):
        B = X.shape[0]
        N, nx, nu, np_ = self.N, self.nx, self.nu, self.np
        X = self._t(X, (B, N + 1, nx), "X"); U = self._t(U, (B, N, nu), "U")
        params = self._t(params, (B, N + 1, np_), "params"); defect = self._t(defect, (B, N, nx), "defect")
        dev = self.device
        K = torch.empty((B, N, nu, nx), dtype=self.tdtype, device=dev); k = torch.empty((B, N, nu), dtype=self.tdtype, device=dev)
        dV = torch.empty((B, 3), dtype=self.tdtype, device=dev); rc = torch.empty(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.L.sddp_backward_pass(self.h, B, _ptr(X), _ptr(U), _ptr(params), _ptr(defect), float(mu), _ptr(K),
                                                 _ptr(k), _ptr(dV), _ptr(rc), self._stream()), self.h, self.L)
        return rc, K, k, dV

    def forward_pass(self, alpha, rho, x0, X, U, params, defect, K, k, trajectories: bool = True):
        B = X.shape[0]
        N, nx, nu, np_ = self.N, self.nx, self.nu, self.np
        dev = self.device
        alpha = torch.as_tensor(alpha, dtype=self.tdtype, device=dev).contiguous()
        rho = torch.as_tensor(rho, dtype=self.tdtype, device=dev).contiguous()
        na = alpha.numel()
        x0 = self._t(x0, (B, nx), "x0"); X = self._t(X, (B, N + 1, nx), "X"); U = self._t(U, (B, N, nu), "U")
        params = self._t(params, (B, N + 1, np_), "params"); defect = self._t(defect, (B, N, nx), "defect")
        K = self._t(K, (B, N, nu, nx), "K"); k = self._t(k, (B, N, nu), "k")
        Jn = torch.empty((B, na), dtype=self.tdtype, device=dev)
        Xn = torch.empty((B, na, N + 1, nx), dtype=self.tdtype, device=dev) if trajectories else None
        Un = torch.empty((B, na, N, nu), dtype=self.tdtype, device=dev) if trajectories else None
        with torch.cuda.device(dev):
            _lib.check(self.L.sddp_forward_pass(self.h, B, na, _ptr(alpha), _ptr(rho), _ptr(x0), _ptr(X), _ptr(U), _ptr(params),
                                                _ptr(defect), _ptr(K), _ptr(k), _ptr(Jn), _ptr(Xn), _ptr(Un), self._stream()), self.h, self.L)
        return Jn, Xn, Un

    def defects(self, X, U, params):
        B = X.shape[0]
        N, nx, nu, np_ = self.N, self.nx, self.nu, self.np
        X = self._t(X, (B, N + 1, nx), "X"); U = self._t(U, (B, N, nu), "U"); params = self._t(params, (B, N + 1, np_), "params")
        D = torch.empty((B, N, nx), dtype=self.tdtype, device=self.device)
        cost = torch.empty(B, dtype=self.tdtype, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.L.sddp_defects(self.h, B, _ptr(X), _ptr(U), _ptr(params), _ptr(D), _ptr(cost), self._stream()), self.h, self.L)
        return D, cost


def fp64_peak_tflops() -> float:
    """Measured FP64 FMA rate of the current device (the roofline denominator of this path)."""
    out = ctypes.c_double()
    _lib.check(_lib.lib().sddp_fp64_peak_tflops(ctypes.byref(out), None))
    return out.value


class DDPSolver:
    """Drop-in for the reference's `ddp.DDPSolver` (ddp.py:10-151) over a `prb.Problem`."""

    def __init__(self, prb, opts: Optional[Dict] = None, device=None, dtype: str = "f64") -> None:
        self.prb = prb
        self.opts = dict(opts or {})
        self.cfg = make_config(prb.model, prb.N, prb.getDt(), self.opts, prb.robot, prb.gains)
        self.max_iters = self.cfg.max_iters
        self.state_var = prb.getState().getVars()
        self.input_var = prb.getInput().getVars()
        self.state_size = sum(v.getDim() for v in self.state_var)
        self.input_size = sum(v.getDim() for v in self.input_var)
        self.param_var = prb.getParameters()
        nx, nu, np_ = DIMS[prb.model]
        if (self.state_size, self.input_size) != (nx, nu):
            raise ValueError("problem variables do not match the model layout")
        self.ddp_solver = BatchedDDP(self.cfg, device, dtype)
        N = prb.N
        self._x0 = np.zeros(nx)
        self._X = np.zeros((N + 1, nx))
        self._U = np.zeros((N, nu))
        self._have_x_ws = False
        self._converged = False
        self.var_solution: Dict[str, np.ndarray] = {}
        self.last: Optional[Dict[str, np.ndarray]] = None

    # ddp.py:122-123
    def setInitialState(self, x0) -> None:
        self._x0 = np.asarray(x0, dtype=np.float64).reshape(-1).copy()

    # ddp.py:113-117 (matrices are dim x nodes, as the examples build them: dsrbd_example.py:63-68)
    def set_u_warmstart(self, u) -> None:
        self._U = np.ascontiguousarray(np.asarray(u, dtype=np.float64).T)

    def set_x_warmstart(self, x) -> None:
        self._X = np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)
        self._have_x_ws = True

    def get_params_value(self) -> np.ndarray:
        """[N+1, np] parameter values, flattened as ddp.py:165-177 does."""
        return self.prb.flat_parameters()

    # ddp.py:96-106
    def solve(self) -> bool:
        params = self.get_params_value()
        if not self._have_x_ws:   # no x warm start given: start every node at the initial state
            self._X = np.tile(self._x0, (self.prb.N + 1, 1))
            self._have_x_ws = True
        r = self.ddp_solver.solve_host(self._x0[None], params[None], self._X[None], self._U[None], gains=True, history=True)
        self.last = {k: (v[0] if v is not None else None) for k, v in r.items()}
        x, u = r["X"][0].T.astype(np.float64), r["U"][0].T.astype(np.float64)      # nx x (N+1), nu x N as pyddp returns them (ddp.py:101)
        self._X, self._U = r["X"][0], r["U"][0]             # the next tick starts from this solution
        self.var_solution = self._createVarSolDict(x, u)
        self.var_solution["x_opt"] = x
        self.var_solution["u_opt"] = u
        self._converged = int(r["status"][0]) == STATUS_CONVERGED
        return self._converged

    def getSolutionDict(self) -> Dict[str, np.ndarray]:
        return self.var_solution

    def is_converged(self) -> bool:
        return self._converged

    # ddp.py:125-151
    def _createVarSolDict(self, x: np.ndarray, u: np.ndarray) -> Dict[str, np.ndarray]:
        out, pos = {}, 0
        for v in self.state_var:
            out[v.getName()] = x[pos:pos + v.getDim(), :]
            pos += v.getDim()
        pos = 0
        for v in self.input_var:
            out[v.getName()] = u[pos:pos + v.getDim(), :]
            pos += v.getDim()
        return out
