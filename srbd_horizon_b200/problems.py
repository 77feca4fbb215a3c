"""Synthetic problem batches for the BASELINE.json configs.

Every problem is one receding-horizon tick of the reference's MPC loop
(dsrbd_example.py:84-135): an initial state, a contact schedule as `wpg.steps_phase.set`
would have left it in the node parameters, and a reference velocity.

`schedule_params` is the closed form of driving `steps_phase.set` on a fresh problem
(tests/test_problems.py drives the step-by-step scheduler and compares):
  "step"      set("step") N+1+s0 times: node j holds table entry (s0 + j) mod 20;
  "standing"  set("standing") N+1 times: switches 1, c_ref 0 (wpg.py:94-99);
  "jump"      set("standing") N+1 times, then set("jump") 1 + s0 % 8 times: a stance horizon whose
              last 1..8 nodes are a flight phase (switches 0, c_ref untouched, gain 0; wpg.py:89-93).
              (A whole-horizon flight phase would be 2.5 s of free fall -- not a walking problem.)
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .config import DIMS, MODEL_LIP, MODEL_SRBD, RobotConstants
from .wpg import gait_tables

ACTIONS = ("step", "standing", "jump")
ACTION_PROBS = (0.7, 0.2, 0.1)


def schedule_params(model: int, N: int, actions: np.ndarray, s0: np.ndarray, rdot_ref: np.ndarray,
                    robot: Optional[RobotConstants] = None) -> np.ndarray:
    """params[B, N+1, np] for a batch of (action index, phase, reference velocity)."""
    robot = robot or RobotConstants()
    nx, nu, np_ = DIMS[model]
    B = len(actions)
    c_init_z = float(robot.foot[2])
    lc, ls, rc, rs = gait_tables(c_init_z)
    cycle = 20
    j = np.arange(N + 1)
    idx = (np.asarray(s0)[:, None] + j[None, :]) % cycle                      # [B, N+1]
    act = np.asarray(actions)[:, None]
    step = act == 0
    flight = (act == 2) & (j[None, :] > N - 1 - (np.asarray(s0)[:, None] % 8))   # last 1 + s0%8 nodes
    c_left = np.where(step, lc[idx], 0.0)
    c_right = np.where(step, rc[idx], 0.0)
    s_left = np.where(step, ls[idx], np.where(flight, 0.0, 1.0))
    s_right = np.where(step, rs[idx], np.where(flight, 0.0, 1.0))
    p = np.zeros((B, N + 1, np_))
    p[:, 1:, 0:3] = np.asarray(rdot_ref)[:, None, :]                           # prb.py:74 nodes 1..N
    if model == MODEL_SRBD:
        p[:, :, 6] = np.where(flight, 0.0, 1e2)                                # wpg.py:82,91,96
        base = 7
        p[:, :, 18] = 1.0                                                      # oref = identity, prb.py:185-186
    else:
        base = 3
    for i in range(4):
        p[:, :, base + 2 * i] = c_left if i < 2 else c_right
        p[:, :, base + 2 * i + 1] = s_left if i < 2 else s_right
    return p


def nominal(model: int, robot: Optional[RobotConstants] = None) -> Tuple[np.ndarray, np.ndarray]:
    """getInitialState / getStaticInput (prb.py:224-246, 420-441)."""
    robot = robot or RobotConstants()
    if model == MODEL_SRBD:
        x = np.zeros(37); x[0:3] = robot.com; x[6] = 1.0; x[7:19] = robot.foot
        u = np.zeros(24)
        u[5::6] = robot.mass * 9.81 / robot.force_scaling / 4
    else:
        x = np.zeros(30); x[0:3] = robot.com; x[3:15] = robot.foot
        u = np.zeros(15); u[0:2] = robot.com[0:2]
    return x, u


def make_batch(model: int, N: int, B: int, seed: int = 12345, enumerate_schedules: bool = False,
               first: int = 0, x_noise: float = 0.0, robot: Optional[RobotConstants] = None) -> Dict[str, np.ndarray]:
    """Problems `first .. first+B-1` of the seeded family (problem b uses default_rng(seed + b)).

    Returns x0[B,nx], params[B,N+1,np], X0[B,N+1,nx] (x0 repeated, plus N(0, x_noise^2) on r and
    rdot of nodes 1..N for the multiple-shooting config), U0[B,N,nu] (static input), actions, s0.
    """
    robot = robot or RobotConstants()
    nx, nu, np_ = DIMS[model]
    xn, un = nominal(model, robot)
    x0 = np.tile(xn, (B, 1))
    actions = np.zeros(B, dtype=np.int64)
    s0 = np.zeros(B, dtype=np.int64)
    rdot_ref = np.zeros((B, 3))
    noise = np.zeros((B, N, 6))
    ir, ic, ird = (0, 7, 19) if model == MODEL_SRBD else (0, 3, 15)
    for i in range(B):
        b = first + i
        rng = np.random.default_rng(seed + b)
        x0[i, ir:ir + 3] += rng.uniform(-0.02, 0.02, 3)
        axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
        ang = rng.uniform(0.0, 0.1)
        shift = rng.uniform(-0.02, 0.02, (2, 2))
        vel = rng.uniform(-0.1, 0.1, 6)
        ref = rng.uniform(-0.5, 0.5, 2)
        a = rng.choice(3, p=ACTION_PROBS)
        ph = rng.integers(0, 20)
        if x_noise > 0.0:
            noise[i] = rng.normal(0.0, x_noise, (N, 6))
        if model == MODEL_SRBD:
            x0[i, 3:6] = axis * np.sin(ang / 2); x0[i, 6] = np.cos(ang / 2)
            x0[i, 22:25] = vel[3:6]
        for foot in range(2):
            for pt in range(2):
                x0[i, ic + 3 * (2 * foot + pt):ic + 3 * (2 * foot + pt) + 2] += shift[foot]
        x0[i, ird:ird + 3] = vel[0:3]
        rdot_ref[i, 0:2] = ref
        if enumerate_schedules:
            a, ph = (b % 60) // 20, b % 20
        actions[i], s0[i] = a, ph
    params = schedule_params(model, N, actions, s0, rdot_ref, robot)
    X0 = np.repeat(x0[:, None, :], N + 1, axis=1)
    if x_noise > 0.0:
        X0[:, 1:, ir:ir + 3] += noise[:, :, 0:3]
        X0[:, 1:, ird:ird + 3] += noise[:, :, 3:6]
    U0 = np.tile(un, (B, N, 1))
    return dict(x0=x0, params=params, X0=X0, U0=U0, actions=actions, s0=s0)
