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
