"""Pin the oracle's model functions to the mpmath goldens transcribed from prb.py/ddp.py
(tests/golden/make_golden.py), plus the known-answer tests of SURVEY.md section 8c (T3, T5)."""
import numpy as np
import pytest

from oracle import oracle as O
from srbd_horizon_b200.config import (HESSIAN_EXACT, HESSIAN_GN, MODEL_LIP, MODEL_SRBD, RobotConstants,
                                      make_config)
from tests.helpers import golden_cases, nominal_state, relerr

TOL = 1e-11   # fp64 formulas vs 80-digit goldens; Hessian entries span 1 .. 1e8


def _cfg(model, mode, hess=HESSIAN_EXACT):
    return make_config(model, 20, 0.05, {"inertia_mode": mode, "hessian_mode": hess})


def test_golden_dynamics_and_cost(golden):
    n = 0
    for k, m, mode, kind in golden_cases(golden):
        cfg = _cfg(m, mode)
        x, u, p = golden[k + "_x"], golden[k + "_u"], golden[k + "_p"]
        if kind != 2:
            assert relerr(O.dynamics(cfg, x, u), golden[k + "_f"]) < 1e-14
        assert abs(O.cost(cfg, kind, x, u, p) - float(golden[k + "_L"])) <= 1e-13 * abs(float(golden[k + "_L"]))
        n += 1
    assert n >= 15


def test_golden_derivatives_exact(golden):
    for k, m, mode, kind in golden_cases(golden):
        cfg = _cfg(m, mode)
        d = O.derivs(cfg, kind, golden[k + "_x"], golden[k + "_u"], golden[k + "_p"])
        names = ["lx", "lxx"] if kind == 2 else ["fx", "fu", "lx", "lu", "lxx", "lux", "luu"]
        for name in names:
            assert relerr(d[name], golden[k + "_" + name]) < TOL, (k, name)


def test_golden_derivatives_gauss_newton(golden):
    for k, m, mode, kind in golden_cases(golden):
        cfg = _cfg(m, mode, HESSIAN_GN)
        d = O.derivs(cfg, kind, golden[k + "_x"], golden[k + "_u"], golden[k + "_p"])
        names = ["lxx"] if kind == 2 else ["lxx", "lux", "luu"]
        for name in names:
            assert relerr(d[name], golden[k + "_gn_" + name]) < TOL, (k, name)


def test_lip_hessian_is_state_independent(golden):
    """T1 prerequisite: every LIP residual is affine, so exact == Gauss-Newton."""
    for k, m, mode, kind in golden_cases(golden, MODEL_LIP):
        for name in (["lxx"] if kind == 2 else ["lxx", "lux", "luu"]):
            assert relerr(golden[k + "_" + name], golden[k + "_gn_" + name]) < 1e-20 + 1e-15


def test_T3_static_fixed_point():
    """f(x_init, u_static) == x_init (prb.py:224-246: fz = m g / fs / 4 per contact point)."""
    for mode in (0, 1):
        cfg = _cfg(MODEL_SRBD, mode)
        x, u = nominal_state(MODEL_SRBD)
        assert np.max(np.abs(O.dynamics(cfg, x, u) - x)) < 1e-15
    cfg = _cfg(MODEL_LIP, 0)
    x, u = nominal_state(MODEL_LIP)
    xn = O.dynamics(cfg, x, u)
    # LIP: rddot = eta2 (r - z) - g with z_z = 0 and r_z = 0.88 => exactly zero (prb.py:317-318, 436-441)
    assert np.max(np.abs(xn - x)) < 1e-14


def test_T5_cost_bookkeeping():
    """L_0 has no trackers, L_N has no input terms and no constraints (ddp.py:216-226, prb.py node ranges)."""
    cfg = _cfg(MODEL_SRBD, 0)
    x, u = nominal_state(MODEL_SRBD)
    p = np.zeros(19); p[6] = 10.0; p[8:15:2] = 1.0; p[18] = 1.0
    x1 = x.copy(); x1[2] += 0.1; x1[19] = 0.3            # tracker violations only
    assert O.cost(cfg, 0, x1, u, p) == pytest.approx(O.cost(cfg, 0, x, u, p), rel=1e-15)
    assert O.cost(cfg, 2, x1, None, p) == pytest.approx(1e3 * 0.01 + 1e4 * 0.09, rel=1e-12)
    x2 = x.copy(); x2[9] += 0.01; x2[25] = 0.2           # constraint violations: c0_z, cdot0_x
    assert O.cost(cfg, 2, x2, None, p) == pytest.approx(0.0, abs=1e-20)
    base = O.cost(cfg, 1, x, u, p)
    assert O.cost(cfg, 1, x2, u, p) - base == pytest.approx(1e6 * (1e-4 + 0.04 + 0.04), rel=1e-9)
    # static cost at the nominal point: only min_f (1e4 |f|^2) is non-zero
    fz = RobotConstants().mass * 9.81 / 1000 / 4
    assert base == pytest.approx(4 * 1e4 * fz * fz, rel=1e-12)


def test_friction_cone_barrier_derivatives_by_finite_differences():
    """Inequality handling (off by default): the exponential barrier on the linearised friction cone adds
    to L, lu and the per-foot 3x3 blocks of luu; gradient and Hessian against central differences of the oracle's own
    cost, and nothing changes when the weight is 0 (the reference drops the cone, prb.py:173-177)."""
    from oracle import oracle as O
    from srbd_horizon_b200.config import MODEL_SRBD, make_config
    from srbd_horizon_b200.problems import nominal
    cfg0 = make_config(MODEL_SRBD, 10, 0.05, {})
    cfg = make_config(MODEL_SRBD, 10, 0.05, {"friction_cone_weight": 2.0, "friction_cone_sharpness": 4.0, "friction_cone_mu": 0.6})
    rng = np.random.default_rng(3)
    x, u = nominal(MODEL_SRBD)
    x = x + 0.01 * rng.standard_normal(37); u = u + 0.02 * rng.standard_normal(24)
    p = np.zeros(19); p[8:15:2] = 1.0; p[18] = 1.0; p[6] = 10.0
    d, d0 = O.derivs(cfg, 1, x, u, p), O.derivs(cfg0, 1, x, u, p)
    assert O.cost(cfg, 1, x, u, p) > O.cost(cfg0, 1, x, u, p)
    assert O.cost(cfg, 2, x, u, p) == O.cost(cfg0, 2, x, u, p)          # terminal node: no input terms
    for name in ("lx", "lxx", "lux"):
        np.testing.assert_array_equal(d[name], d0[name])
    eps = 1e-6
    lu_fd = np.array([(O.cost(cfg, 1, x, u + eps * e, p) - O.cost(cfg, 1, x, u - eps * e, p)) / (2 * eps) for e in np.eye(24)])
    assert np.abs(lu_fd - d["lu"]).max() < 1e-8 * np.abs(d["lu"]).max()
    luu_fd = np.array([(O.derivs(cfg, 1, x, u + eps * e, p)["lu"] - O.derivs(cfg, 1, x, u - eps * e, p)["lu"]) / (2 * eps) for e in np.eye(24)])
    assert np.abs(luu_fd - d["luu"]).max() < 1e-8 * np.abs(d["luu"]).max()
    dd = d["luu"] - d0["luu"]
    mask = np.zeros((24, 24), bool)
    for i in range(4):
        mask[6 * i + 3:6 * i + 6, 6 * i + 3:6 * i + 6] = True
    assert np.all(dd[~mask] == 0) and np.all(np.linalg.eigvalsh(dd) > -1e-9)
