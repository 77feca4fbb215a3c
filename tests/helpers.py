"""Shared helpers for the parity tests (seeded problem generators, comparisons)."""
import numpy as np

from srbd_horizon_b200.config import DIMS, MODEL_LIP, MODEL_SRBD, RobotConstants


def relerr(a, b):
    """norm-wise relative error  max|a-b| / max(1e-300, max|b|)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))) if a.size else 0.0


def golden_cases(golden, model=None):
    keys = sorted(k[:-5] for k in golden if k.endswith("_meta"))
    for k in keys:
        m, mode, kind = (int(v) for v in golden[k + "_meta"])
        if model is None or m == model:
            yield k, m, mode, kind


def nominal_state(model, robot=None):
    robot = robot or RobotConstants()
    if model == MODEL_SRBD:
        x = np.zeros(37); x[0:3] = robot.com; x[6] = 1.0; x[7:19] = robot.foot
        u = np.zeros(24)
        for i in range(4):
            u[6 * i + 5] = robot.mass * 9.81 / robot.force_scaling / 4
    else:
        x = np.zeros(30); x[0:3] = robot.com; x[3:15] = robot.foot
        u = np.zeros(15); u[0:2] = robot.com[0:2]
    return x, u


def random_point(rng, model):
    nx, nu, np_ = DIMS[model]
    x, u = nominal_state(model)
    x = x + rng.uniform(-0.05, 0.05, nx)
    u = u + rng.uniform(-0.05, 0.05, nu)
    p = np.zeros(np_)
    if model == MODEL_SRBD:
        x[19:25] = rng.uniform(-0.3, 0.3, 6)
        p[0:3] = rng.uniform(-0.5, 0.5, 3); p[6] = 10.0
        p[7:15:2] = rng.uniform(0, 0.05, 4); p[8:15:2] = rng.integers(0, 2, 4)
        p[15:19] = [0, 0, 0, 1]
    else:
        p[0:3] = rng.uniform(-0.5, 0.5, 3)
        p[3:11:2] = rng.uniform(0, 0.05, 4); p[4:11:2] = rng.integers(0, 2, 4)
    return x, u, p
