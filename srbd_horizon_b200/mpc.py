"""Receding-horizon glue around the solver: what the reference's example loops do between two solves
(dsrbd_example.py:84-131,158-160; dlip_example.py:90-134,161-162)."""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .config import MODEL_SRBD


def shift_back(param) -> None:
    """values of node j -> node j-1 for j = 1..N (dsrbd_example.py:102-106)."""
    v = param.getValues()
    for j in range(1, v.shape[1]):
        param.assign(v[:, j], nodes=j - 1)


def mpc_tick_references(problem, rdot_ref_last: Sequence[float]) -> None:
    """Shift the per-node references one node back and write the newest velocity reference at node N
    (dsrbd_example.py:102-122; the LIP loop shifts rdot_ref only, dlip_example.py:108-125)."""
    N = problem.prb.N
    shift_back(problem.rdot_ref)
    if problem.prb.model == MODEL_SRBD:
        shift_back(problem.w_ref)
        shift_back(problem.oref)
        shift_back(problem.orientation_tracking_gain)
    problem.rdot_ref.assign(list(rdot_ref_last), nodes=N)


def plant_step(ddp_solver, state: np.ndarray, u: np.ndarray) -> np.ndarray:
    """state <- EULER(state, u, dt), SRBD quaternion renormalised (dsrbd_example.py:158-160).
    The Euler step runs on the GPU through sddp_eval_derivatives (f output of stage one)."""
    p = np.zeros((1, ddp_solver.np))
    f = ddp_solver.eval_derivatives([0], np.asarray(state)[None], np.asarray(u)[None], p)["f"][0].cpu().numpy()
    if ddp_solver.cfg.model == MODEL_SRBD:
        f[3:7] /= np.linalg.norm(f[3:7])
    return f


# ------------------------------------------------------------------------------------------------
# Device-side closed loop for a batch of robots (SURVEY.md section 8f, N1 + N2): the tick of the example
# loops -- shift the schedule, write the gait entries of node N, solve, plant step -- without host round trips.
ACTION_STEP, ACTION_STANCE, ACTION_JUMP = 0, 1, 2


class BatchedMPC:
    """B independent MPC loops on one GPU.  Per tick (dsrbd_example.py:84-160):
    x0 <- state; references / contact plan one node back and node N filled from the gait tables
    (`sddp_mpc_advance`, wpg.py:68-101); `sddp_solve_batch` warm-started from the previous solution;
    state <- EULER(state, u_0) with the SRBD quaternion renormalised (`sddp_plant_step`)."""

    def __init__(self, solver, x0, params, U0=None):
        import ctypes

        import torch

        from . import _lib
        from .wpg import gait_tables
        self._ct, self._torch, self._lib = ctypes, torch, _lib
        self.solver = solver
        dev = solver.device
        t = lambda a, dt=solver.tdtype: torch.as_tensor(a, dtype=dt, device=dev).contiguous()
        self.state = t(x0).clone()
        B = self.state.shape[0]
        self.B = B
        self.params = t(params).clone()
        N, nx, nu = solver.N, solver.nx, solver.nu
        self.X = self.state[:, None, :].repeat(1, N + 1, 1).contiguous()
        self.U = t(U0).clone() if U0 is not None else torch.zeros((B, N, nu), dtype=solver.tdtype, device=dev)
        self.step_counter = torch.zeros(B, dtype=torch.int32, device=dev)
        c_init_z = float(solver.cfg.foot[2])
        tabs = [np.ascontiguousarray(a, dtype=np.float64) for a in gait_tables(c_init_z)]
        _lib.check(solver.L.sddp_set_gait_tables(solver.h, *[a.ctypes.data_as(ctypes.c_void_p) for a in tabs]), solver.h, solver.L)
        self.last = None

    def _p(self, t):
        return self._ct.c_void_p(t.data_ptr())

    def advance_schedule(self, actions, rdot_ref_cmd):
        torch = self._torch
        s = self.solver
        a = torch.as_tensor(actions, dtype=torch.int32, device=s.device).contiguous()
        cmd = torch.as_tensor(rdot_ref_cmd, dtype=s.tdtype, device=s.device).contiguous()
        assert a.shape == (self.B,) and cmd.shape == (self.B, 3)
        with torch.cuda.device(s.device):
            self._lib.check(s.L.sddp_mpc_advance(s.h, self.B, self._p(self.params), self._p(a), self._p(self.step_counter),
                                                 self._p(cmd), s._stream()), s.h, s.L)

    def plant_step(self):
        torch = self._torch
        s = self.solver
        with torch.cuda.device(s.device):
            self._lib.check(s.L.sddp_plant_step(s.h, self.B, self._p(self.state), self._p(self.U), s.N * s.nu, s._stream()), s.h, s.L)

    # -- the tick as ONE CUDA graph (launch-bound small fleets: six launches and memsets become one graph launch) --------
    def capture(self, gains: bool = False) -> None:
        """Capture one tick -- schedule advance, solve warm-started in place, plant step -- into a CUDA graph.  Afterwards
        `tick_graph(actions, cmd)` copies the two small inputs into static device buffers and replays the graph; results are
        the same bits as `tick` (same kernels on the same buffers).  The dispatch hint of large fleets is not part of the
        graph (it is a host decision per tick); meant for fleets below one problem per CTA slot, where launches dominate."""
        torch = self._torch
        dev = self.solver.device
        self._g_act = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self._g_cmd = torch.zeros((self.B, 3), dtype=self.solver.tdtype, device=dev)
        keep = [t.clone() for t in (self.state, self.params, self.X, self.U, self.step_counter)]
        restore = lambda: [d.copy_(s) for d, s in zip((self.state, self.params, self.X, self.U, self.step_counter), keep)]

        def body():
            self.advance_schedule(self._g_act, self._g_cmd)
            r = self.solver.solve(self.state, self.params, self.X, self.U, gains=gains, history=False, inplace=True)
            self.plant_step()
            return r

        side = torch.cuda.Stream(device=dev)      # warm-up outside the capture (lazy allocations, module loading)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        restore()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._g_result = body()
        restore()

    def tick_graph(self, actions, rdot_ref_cmd):
        """One closed-loop tick by replaying the captured graph; returns the BatchResult whose tensors the graph overwrites."""
        torch = self._torch
        if getattr(self, "_graph", None) is None:
            raise RuntimeError("call capture() first")
        self._g_act.copy_(torch.as_tensor(actions, dtype=torch.int32), non_blocking=True)
        self._g_cmd.copy_(torch.as_tensor(rdot_ref_cmd, dtype=self.solver.tdtype), non_blocking=True)
        self._graph.replay()
        self.last = self._g_result
        return self._g_result

    def tick(self, actions, rdot_ref_cmd, gains: bool = False):
        """One closed-loop tick for all B robots; returns the BatchResult of the solve (X, U alias the warm start)."""
        self.advance_schedule(actions, rdot_ref_cmd)
        # Dispatch hint (scheduling only): robots that needed many iterations at the previous tick go first, the rest
        # grouped by contact schedule -- consecutive ticks of a receding-horizon loop solve nearly the same problems.
        order = None
        if self.state.shape[0] >= 1024:
            order = self.solver.dispatch_order(self.params)
            if self.last is not None and self.last.iters.shape[0] == self.state.shape[0]:
                order = self.solver.order_from_keys(-self.last.iters.to(self.params.dtype), (self.params[:, -1, 0:3] ** 2).sum(dim=1))
        self.last = self.solver.solve(self.state, self.params, self.X, self.U, gains=gains, history=False, inplace=True, order=order)
        self.plant_step()
        return self.last
