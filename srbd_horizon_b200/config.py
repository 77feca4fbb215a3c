"""Problem/solver configuration shared by the host code and the C ABI.

`SddpConfig` mirrors `include/sddp.h:SddpConfig` field for field (ctypes).
Layouts follow the reference's variable creation order:

* SRBD  (/root/reference/python/prb.py:32-68, 224-246)
  x[37] = r[0:3] o[3:7] c0..c3[7:19] rdot[19:22] w[22:25] cdot0..3[25:37]
  u[24] = (cddot_i[3], f_i[3]) for i = 0..3
  p[19] = rdot_ref[0:3] w_ref[3:6] orientation_tracking_gain[6]
          (c_ref_i, cdot_switch_i) for i = 0..3 at [7:15], oref[15:19]
  (parameter flattening order: ddp.py:165-177)
* LIP   (prb.py:264-295, 420-441)
  x[30] = r[0:3] c0..c3[3:15] rdot[15:18] cdot0..3[18:30]
  u[15] = z[0:3] cddot0..3[3:15]
  p[11] = rdot_ref[0:3] (c_ref_i, cdot_switch_i) for i = 0..3

The robot constants (mass, inertia, CoM, foot points) come from an external URDF
in the reference (launch/SRBD_kangaroo_line_feet.launch:9, prb.py:92-95,130-139);
the values below are SYNTHETIC "kangaroo-like" numbers.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

MODEL_SRBD = 0
MODEL_LIP = 1
INERTIA_LITERAL = 0   # prb.py:99, CasADi `*` is element-wise: R o (I/fs) o R^T
INERTIA_ROTATED = 1   # README.md:2 intent: R (I/fs) R^T
HESSIAN_EXACT = 0     # ddp.py:210-214 hands pyddp a scalar L -> exact Hessian
HESSIAN_GN = 1

STATUS_CONVERGED = 0
STATUS_MAX_ITERS = 1
STATUS_LS_FAILED = 2
STATUS_REG_FAILED = 3
STATUS_NAN = 4

HIST = 4  # per-iteration record: cost, alpha, mu, max|defect|

DIMS = {MODEL_SRBD: (37, 24, 19), MODEL_LIP: (30, 15, 11)}


class SddpConfig(ctypes.Structure):
    _fields_ = [
        ("model", ctypes.c_int32),
        ("N", ctypes.c_int32),
        ("inertia_mode", ctypes.c_int32),
        ("hessian_mode", ctypes.c_int32),
        ("multiple_shooting", ctypes.c_int32),
        ("max_iters", ctypes.c_int32),
        ("dense_backward", ctypes.c_int32),
        ("lip_tail_start", ctypes.c_int32),
        ("dt", ctypes.c_double),
        ("mass", ctypes.c_double),
        ("inertia", ctypes.c_double * 9),
        ("com", ctypes.c_double * 3),
        ("foot", ctypes.c_double * 12),
        ("force_scaling", ctypes.c_double),
        ("gravity", ctypes.c_double),
        ("eta2", ctypes.c_double),
        ("r_tracking_gain", ctypes.c_double),
        ("rdot_tracking_gain", ctypes.c_double),
        ("w_tracking_gain", ctypes.c_double),
        ("rel_position_gain", ctypes.c_double),
        ("force_switch_weight", ctypes.c_double),
        ("min_qddot_gain", ctypes.c_double),
        ("min_f_gain", ctypes.c_double),
        ("zmp_tracking_gain", ctypes.c_double),
        ("constraint_weight", ctypes.c_double),
        ("alpha_0", ctypes.c_double),
        ("alpha_converge_threshold", ctypes.c_double),
        ("line_search_decrease_factor", ctypes.c_double),
        ("beta", ctypes.c_double),
        ("cost_reduction_ths", ctypes.c_double),
        ("mu0", ctypes.c_double),
        ("defect_contraction_rate", ctypes.c_double),
        ("mu_min", ctypes.c_double),
        ("mu_max", ctypes.c_double),
        ("mu_factor", ctypes.c_double),
        ("defect_ths", ctypes.c_double),
        ("friction_cone_weight", ctypes.c_double),
        ("friction_cone_mu", ctypes.c_double),
        ("friction_cone_sharpness", ctypes.c_double),
        ("force_bound_weight", ctypes.c_double),
        ("force_bound", ctypes.c_double),
        ("unilateral_weight", ctypes.c_double),
        ("cdot_bound_weight", ctypes.c_double),
        ("cdot_bound", ctypes.c_double),
        ("bound_sharpness", ctypes.c_double),
    ]

    def copy(self) -> "SddpConfig":
        c = SddpConfig()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(self), ctypes.sizeof(SddpConfig))
        return c

    @property
    def dims(self) -> Tuple[int, int, int]:
        return DIMS[self.model]


@dataclass
class RobotConstants:
    """Synthetic stand-in for what the reference reads from the URDF."""
    mass: float = 40.0
    inertia: Sequence[float] = (2.0, 0.02, 0.02,
                                0.02, 1.8, 0.02,
                                0.02, 0.02, 0.5)
    com: Sequence[float] = (0.0, 0.0, 0.88)
    # left_foot_upper, left_foot_lower, right_foot_upper, right_foot_lower
    # (launch/SRBD_kangaroo_line_feet.launch:24-25)
    foot: Sequence[float] = (0.10, 0.10, 0.0,
                             -0.10, 0.10, 0.0,
                             0.10, -0.10, 0.0,
                             -0.10, -0.10, 0.0)
    force_scaling: float = 1000.0   # prb.py:98
    gravity: float = 9.81


@dataclass
class Gains:
    """rospy.get_param defaults of prb.py:142-150, 359-363 and ddp.py:181."""
    r_tracking_gain: float = 1e3
    rdot_tracking_gain: float = 1e4
    w_tracking_gain: float = 1e4
    rel_position_gain: float = 1e4
    force_switch_weight: float = 1e2
    min_qddot_gain: float = 1e0
    min_f_gain: float = 1e-2
    zmp_tracking_gain: float = 1e3
    constraint_weight: float = 1e6


#: Solver option defaults.  The first seven names are the reference's
#: (ddp.py:17-35); pyddp's own defaults for them are unknown, the adapter's
#: shadow values (ddp.py:17-31) are used where it has them.
DEFAULT_OPTS: Dict[str, float] = {
    "max_iters": 100,
    "alpha_0": 1.0,
    "alpha_converge_threshold": 1e-1,      # ddp.py:23
    "line_search_decrease_factor": 0.5,
    "beta": 1e-4,
    "cost_reduction_ths": 1e-6,
    "mu0": 0.0,
    # extensions (not reference option names)
    "multiple_shooting": 1,                # pyddp is a multiple-shooting DDP (README.md:5-6, ddp.py:116-117)
    "defect_contraction_rate": 0.0,        # README.md:6; <=0 means rho = alpha
    "mu_min": 1e-6,
    "mu_max": 1e10,
    "mu_factor": 10.0,
    "defect_ths": 1e-8,
    "inertia_mode": INERTIA_LITERAL,
    "hessian_mode": HESSIAN_EXACT,
    "dense_backward": 0,                   # 1: generic dense Riccati kernel for SRBD (A/B check of the structured one)
    # inequality handling: exponential barrier on the linearised friction cone; 0 = dropped as in the reference
    "friction_cone_weight": 0.0,
    "friction_cone_mu": 0.8,               # prb.py:174
    "friction_cone_sharpness": 6.0,        # ddp.py:182
    # bounds as exponential barriers (ddp.py:204-209); every weight 0 = ignored as in the reference
    "force_bound_weight": 0.0,
    "force_bound": 1.0,                    # isrbd_example.py:200 max_contact_force, in units of force_scaling
    "unilateral_weight": 0.0,              # f_z >= 0 alone (isrbd_example.py:198)
    "cdot_bound_weight": 0.0,
    "cdot_bound": 1.0,                     # contact-point velocity box (isrbd_example.py:195)
    "bound_sharpness": 6.0,                # ddp.py:182
    # model scheduler: first node of the LIP-style tail, 0 = full SRBD (README.md:7, isrbd_example.py:344-353)
    "lip_tail_start": 0,
}


def make_config(model: int, N: int, dt: float, opts: Dict | None = None,
                robot: RobotConstants | None = None, gains: Gains | None = None) -> SddpConfig:
    robot = robot or RobotConstants()
    gains = gains or Gains()
    o = dict(DEFAULT_OPTS)
    if opts:
        unknown = set(opts) - set(o)
        if unknown:
            raise KeyError(f"unknown DDP option(s): {sorted(unknown)}")
        o.update(opts)
    c = SddpConfig()
    c.model = int(model)
    c.N = int(N)
    c.inertia_mode = int(o["inertia_mode"])
    c.hessian_mode = int(o["hessian_mode"])
    c.multiple_shooting = int(o["multiple_shooting"])
    c.max_iters = int(o["max_iters"])
    c.dense_backward = int(o["dense_backward"])
    c.lip_tail_start = int(o["lip_tail_start"])
    c.dt = float(dt)
    c.mass = robot.mass
    c.inertia = (ctypes.c_double * 9)(*robot.inertia)
    c.com = (ctypes.c_double * 3)(*robot.com)
    c.foot = (ctypes.c_double * 12)(*robot.foot)
    c.force_scaling = robot.force_scaling
    c.gravity = robot.gravity
    c.eta2 = 9.81 / 0.88   # prb.py:317
    for name in ("r_tracking_gain", "rdot_tracking_gain", "w_tracking_gain", "rel_position_gain",
                 "force_switch_weight", "min_qddot_gain", "min_f_gain", "zmp_tracking_gain",
                 "constraint_weight"):
        setattr(c, name, float(getattr(gains, name)))
    for name in ("alpha_0", "alpha_converge_threshold", "line_search_decrease_factor", "beta",
                 "cost_reduction_ths", "mu0", "defect_contraction_rate", "mu_min", "mu_max",
                 "mu_factor", "defect_ths", "friction_cone_weight", "friction_cone_mu", "friction_cone_sharpness",
                 "force_bound_weight", "force_bound", "unilateral_weight", "cdot_bound_weight", "cdot_bound", "bound_sharpness"):
        setattr(c, name, float(o[name]))
    if c.N < 1 or c.max_iters < 1:
        raise ValueError("N and max_iters must be >= 1")
    return c
