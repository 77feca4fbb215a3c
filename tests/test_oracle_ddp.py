"""Known-answer tests for the oracle's DDP iteration
and a cross-check of the C loop against an independent numpy restatement."""
import numpy as np
import pytest

from oracle import oracle as O
from srbd_horizon_b200.config import DIMS, MODEL_LIP, MODEL_SRBD, make_config
from srbd_horizon_b200.problems import make_batch
from tests import np_ddp
from tests.helpers import relerr

EX_OPTS = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}   # dsrbd_example.py:55-58


def lip_kkt(cfg, x0, params):
    """Dense KKT solve of the LIP problem: every residual is affine and the dynamics are linear
    (prb.py:315-402), so the DDP fixed point is the unique minimiser of an equality-constrained QP."""
    nx, nu, _ = DIMS[MODEL_LIP]
    N = cfg.N
    z0x, z0u = np.zeros(nx), np.zeros(nu)
    f0 = O.dynamics(cfg, z0x, z0u)
    nz = N * nu + N * nx                       # u_0..u_{N-1}, x_1..x_N
    H = np.zeros((nz, nz)); g = np.zeros(nz)
    A = np.zeros((N * nx, nz)); b = np.zeros(N * nx)
    iu = lambda k: slice(k * nu, (k + 1) * nu)
    ix = lambda k: slice(N * nu + (k - 1) * nx, N * nu + k * nx)      # k >= 1
    for k in range(N):
        D = O.derivs(cfg, 0 if k == 0 else 1, z0x, z0u, params[k])
        fx, fu = D["fx"], D["fu"]
        H[iu(k), iu(k)] += D["luu"]; g[iu(k)] += D["lu"]
        if k == 0:
            g[iu(0)] += D["lux"] @ x0
        else:
            H[ix(k), ix(k)] += D["lxx"]; g[ix(k)] += D["lx"]
            H[iu(k), ix(k)] += D["lux"]; H[ix(k), iu(k)] += D["lux"].T
        # x_{k+1} - fx x_k - fu u_k = f0
        rows = slice(k * nx, (k + 1) * nx)
        A[rows, ix(k + 1)] = np.eye(nx); A[rows, iu(k)] = -fu
        b[rows] = f0
        if k == 0:
            b[rows] += fx @ x0
        else:
            A[rows, ix(k)] = -fx
    D = O.derivs(cfg, 2, z0x, None, params[N])
    H[ix(N), ix(N)] += D["lxx"]; g[ix(N)] += D["lx"]
    # null-space free solve of the KKT system in extended precision is overkill; scale and solve
    KKT = np.block([[H, A.T], [A, np.zeros((N * nx, N * nx))]])
    rhs = np.concatenate([-g, b])
    s = 1.0 / np.sqrt(np.maximum(np.abs(np.diag(KKT)), 1.0))
    sol = s * np.linalg.solve(KKT * s[:, None] * s[None, :], s * rhs)
    r = KKT @ sol - rhs
    sol -= s * np.linalg.solve(KKT * s[:, None] * s[None, :], s * r)  # one step of refinement
    U = sol[:N * nu].reshape(N, nu)
    X = np.vstack([x0, sol[N * nu:nz].reshape(N, nx)])
    return X, U


@pytest.mark.parametrize("ms", [0, 1])
def test_T1_T2_lip_is_lqr(ms):
    cfg = make_config(MODEL_LIP, 20, 0.05, dict(EX_OPTS, multiple_shooting=ms))
    b = make_batch(MODEL_LIP, 20, 6)
    r = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=2)
    assert (r["status"] == 0).all()
    assert (r["iters"] <= 2).all()                                   # T2: one full step, one check
    assert (r["hist"][:, 0, 1] == 1.0).all()
    for i in range(6):
        Xk, Uk = lip_kkt(cfg, b["x0"][i], b["params"][i])
        assert relerr(r["X"][i], Xk) < 1e-9
        assert relerr(r["U"][i], Uk) < 1e-9
        assert r["cost"][i] == pytest.approx(O.total_cost(cfg, Xk, Uk, b["params"][i]), rel=1e-10)


def test_T6_model_is_exact_on_lip():
    """Linear-quadratic problem: J(alpha) - J == alpha*D1 + alpha^2*D2 for every alpha (single shooting
    and rho = alpha multiple shooting); fixed-rho: C0 + alpha*D1 + alpha^2*D2."""
    N = 20
    b = make_batch(MODEL_LIP, N, 1, x_noise=0.02)
    x0, params, U = b["x0"][0], b["params"][0], b["U0"][0] + 0.01
    for ms, rate in ((0, 0.0), (1, 0.0), (1, 0.4)):
        cfg = make_config(MODEL_LIP, N, 0.05, dict(multiple_shooting=ms, defect_contraction_rate=rate))
        X = b["X0"][0].copy()
        d = np.zeros((N, 30))
        if ms:
            for k in range(N):
                d[k] = O.dynamics(cfg, X[k], U[k]) - X[k + 1]
        else:
            for k in range(N):
                X[k + 1] = O.dynamics(cfg, X[k], U[k])
        J = O.total_cost(cfg, X, U, params)
        rc, K, kff, dV = O.backward(cfg, X, U, params, d, 0.0)
        assert rc == 0
        for alpha in (1.0, 0.5, 0.25, 0.03):
            rho = rate if rate > 0 else alpha
            Jn, Xn, Un = O.forward(cfg, x0, X, U, params, d, K, kff, alpha, rho)
            model = dV[2] + alpha * dV[0] + alpha * alpha * dV[1]
            assert Jn - J == pytest.approx(model, rel=1e-7, abs=1e-7 * abs(J)), (ms, rate, alpha)
            # T7: the new defects are (1 - rho) d exactly by construction
            for k in range(N):
                dn = O.dynamics(cfg, Xn[k], Un[k]) - Xn[k + 1]
                assert np.max(np.abs(dn - (1 - rho) * d[k])) < 1e-12


@pytest.mark.parametrize("opts", [dict(multiple_shooting=1), dict(multiple_shooting=1, defect_contraction_rate=0.5),
                                  dict(multiple_shooting=1, inertia_mode=1), dict(multiple_shooting=1, hessian_mode=1),
                                  dict(multiple_shooting=0)])
def test_c_loop_matches_numpy_restatement(opts):
    N = 12
    cfg = make_config(MODEL_SRBD, N, 0.05, dict(EX_OPTS, **opts))
    b = make_batch(MODEL_SRBD, N, 3, x_noise=0.01 if opts.get("multiple_shooting") else 0.0)
    r = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"])
    for i in range(3):
        ref = np_ddp.solve(cfg, b["x0"][i], b["params"][i], b["X0"][i], b["U0"][i])
        n = len(ref["hist"])
        assert r["iters"][i] == n and r["status"][i] == ref["status"]
        assert relerr(r["hist"][i, :n, 0], ref["hist"][:, 0]) < 1e-9
        np.testing.assert_array_equal(r["hist"][i, :n, 1], ref["hist"][:, 1])      # same step sizes (T8)
        assert relerr(r["X"][i], ref["X"]) < 1e-9 and relerr(r["U"][i], ref["U"]) < 1e-9
        assert relerr(r["K"][i], ref["K"]) < 1e-8
        # the feed-forward term vanishes at convergence: compare it on the scale of the inputs
        assert np.max(np.abs(r["k"][i] - ref["k"])) < 1e-9 * max(1.0, np.max(np.abs(ref["U"])))


def test_srbd_batch_converges_and_decreases():
    N = 50
    cfg = make_config(MODEL_SRBD, N, 0.05, EX_OPTS)
    b = make_batch(MODEL_SRBD, N, 24)
    r = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=8)
    assert (r["status"] == 0).all()
    assert r["iters"].max() <= 20
    for i in range(24):
        h = r["hist"][i, :r["iters"][i]]
        assert (h[-1, 3] <= cfg.defect_ths)
        assert h[-1, 0] <= h[0, 0]
        # dynamics feasibility of the returned trajectory
        for k in range(0, N, 7):
            assert np.max(np.abs(O.dynamics(cfg, r["X"][i, k], r["U"][i, k]) - r["X"][i, k + 1])) < 1e-7


def test_history_logging_helpers():
    """srbd_horizon_b200.log on an oracle solve (same hist layout as the CUDA path): rows, text table, batch summary."""
    from srbd_horizon_b200 import log
    from srbd_horizon_b200.config import MODEL_LIP, make_config
    from srbd_horizon_b200.problems import make_batch
    cfg = make_config(MODEL_LIP, 10, 0.05, {"max_iters": 20})
    b = make_batch(MODEL_LIP, 10, 3, x_noise=0.01)
    ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=2)
    rows = log.history_rows(ro["hist"], ro["iters"], problem=1)
    assert len(rows) == ro["iters"][1] and rows[0]["cost"] == ro["hist"][1, 0, 0] and rows[0]["alpha"] in (0.0, 1.0)
    assert np.isnan(rows[-1]["cost_change"]) and all(r["cost_change"] <= 1e-9 * abs(r["cost"]) for r in rows[:-1]) or cfg.multiple_shooting
    txt = log.format_history(ro["hist"], ro["iters"], ro["status"], problem=1)
    assert "status: converged" in txt and txt.count("\n") == len(rows) + 1
    s = log.batch_summary(ro["iters"], ro["status"])
    assert s["problems"] == 3 and s["converged"] == 3 and s["max_iters"] == 0 and s["iters_max"] >= s["iters_mean"] >= 1
