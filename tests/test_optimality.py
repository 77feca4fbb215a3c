"""Oracle-independent optimality certificates for SRBD solves ("parity is pinned to the builder's oracle").

pyddp is not available, so nothing can pin the DDP *iteration*; but its *result* can be pinned to the reference's
problem definition alone.  tests/ref_model.py is a numpy transcription of prb.py / ddp.py:179-230 (checked here against
the mpmath golden fixture) that yields J(U) along a rollout and dJ/dU by complex-step differentiation.  At a solution

  * the reduced gradient must vanish: ||dJ/dU||_inf <= 1e-6 ||dJ/dU at the warm start||_inf at the examples' stopping
    threshold (cost_reduction_ths = 1e-6), <= 1e-9 when the solver is run to full convergence;
  * a generic nonlinear least-squares solver (scipy.optimize.least_squares, trust-region reflective, complex-step
    Jacobian) started from the same warm start must arrive at the same inputs to 1e-6 and the same cost to 1e-12.

The CPU tests apply this to the C oracle (so the checker itself is pinned), the GPU tests to the CUDA path."""
import numpy as np
import pytest

from oracle import oracle as O
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.problems import make_batch
from tests.helpers import golden_cases, relerr
from tests.ref_model import SrbdRef

EX_OPTS = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}   # dsrbd_example.py:55-58
TIGHT = dict(EX_OPTS, cost_reduction_ths=1e-14)


def test_ref_model_matches_golden(golden):
    """The numpy transcription reproduces the mpmath fixture: f to 1e-15, L to 1e-15 relative, both inertia modes."""
    n = 0
    for k, model, mode, kind in golden_cases(golden, MODEL_SRBD):
        ref = SrbdRef(20, 0.05, mode)
        x, u, p = golden[k + "_x"], golden[k + "_u"], golden[k + "_p"]
        assert ref.L(x, u, p, kind) == pytest.approx(float(golden[k + "_L"]), rel=1e-14)
        if kind != 2:
            assert np.max(np.abs(ref.f(x, u) - golden[k + "_f"])) < 1e-14
        n += 1
    assert n >= 12


def test_complex_step_gradient_matches_differences():
    N = 3
    b = make_batch(MODEL_SRBD, N, 1, seed=77)
    ref = SrbdRef(N, 0.05)
    U = b["U0"][0] + 0.01 * np.random.default_rng(0).normal(size=(N, 24))
    g = ref.reduced_gradient(b["x0"][0], U, b["params"][0])
    for (k, i) in ((0, 5), (1, 14), (2, 23), (2, 0)):
        h = 1e-6
        Up, Um = U.copy(), U.copy()
        Up[k, i] += h; Um[k, i] -= h
        fd = (ref.total_cost(b["x0"][0], Up, b["params"][0]) - ref.total_cost(b["x0"][0], Um, b["params"][0])) / (2 * h)
        assert g[k, i] == pytest.approx(fd, rel=1e-5, abs=1e-3)


def _certify_stationary(U, b, N, inertia_mode, tol):
    ref = SrbdRef(N, 0.05, inertia_mode)
    for i in range(len(U)):
        g0 = np.abs(ref.reduced_gradient(b["x0"][i], b["U0"][i], b["params"][i])).max()
        gs = np.abs(ref.reduced_gradient(b["x0"][i], U[i], b["params"][i])).max()
        assert gs <= tol * g0, (i, gs, g0)


def _nls(ref, x0, U0, P):
    from scipy.optimize import least_squares
    N = ref.N
    fun = lambda z: ref.total_residuals(x0, z.reshape(N, 24), P)

    def jac(z):
        zc = z.astype(np.complex128)
        J = np.zeros((fun(z).size, z.size))
        for j in range(z.size):
            zc[j] += 1e-30j
            J[:, j] = np.imag(ref.total_residuals(x0, zc.reshape(N, 24), P)) / 1e-30
            zc[j] = z[j]
        return J
    sol = least_squares(fun, U0.reshape(-1), jac=jac, method="trf", x_scale="jac", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=200)
    return sol.x.reshape(N, 24), 2.0 * sol.cost


@pytest.mark.parametrize("opts,tol", [(EX_OPTS, 1e-6), (TIGHT, 1e-9)])
def test_oracle_solution_is_stationary(opts, tol):
    N, B = 10, 3
    cfg = make_config(MODEL_SRBD, N, 0.05, opts)
    b = make_batch(MODEL_SRBD, N, B, enumerate_schedules=True)
    ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=3)
    assert (ro["status"] == 0).all()
    ref = SrbdRef(N, 0.05)
    for i in range(B):      # the returned cost is the reference's cost of the returned inputs (gaps closed)
        assert ro["cost"][i] == pytest.approx(ref.total_cost(b["x0"][i], ro["U"][i], b["params"][i]), rel=1e-9)
    _certify_stationary(ro["U"], b, N, 0, tol)


def test_oracle_equals_generic_nls_solver():
    N, B = 4, 3
    cfg = make_config(MODEL_SRBD, N, 0.05, TIGHT)
    b = make_batch(MODEL_SRBD, N, B)
    ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=3)
    ref = SrbdRef(N, 0.05)
    for i in range(B):
        U, J = _nls(ref, b["x0"][i], b["U0"][i], b["params"][i])
        assert relerr(ro["U"][i], U) < 1e-6, i
        assert ro["cost"][i] == pytest.approx(J, rel=1e-12)


# ---------------------------------------------------------------------------------------------------- CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("opts,tol,mode", [(EX_OPTS, 1e-6, 0), (TIGHT, 1e-9, 0), (TIGHT, 1e-9, 1)])
def test_gpu_solution_is_stationary(opts, tol, mode):
    """Through the C ABI: the CUDA result of dsrbd_example.py's configuration (N = 20) zeroes the reduced gradient of
    the reference's problem, for both inertia modes."""
    from srbd_horizon_b200.ddp import BatchedDDP
    N, B = 20, 3
    cfg = make_config(MODEL_SRBD, N, 0.05, dict(opts, inertia_mode=mode))
    b = make_batch(MODEL_SRBD, N, B, enumerate_schedules=True, first=7)
    r = BatchedDDP(cfg).solve(b["x0"], b["params"], b["X0"], b["U0"])
    assert (r.status.cpu().numpy() == 0).all()
    U, cost = r.U.cpu().numpy(), r.cost.cpu().numpy()
    ref = SrbdRef(N, 0.05, mode)
    for i in range(B):
        assert cost[i] == pytest.approx(ref.total_cost(b["x0"][i], U[i], b["params"][i]), rel=1e-9)
    _certify_stationary(U, b, N, mode, tol)


@pytest.mark.gpu
def test_gpu_equals_generic_nls_solver():
    from srbd_horizon_b200.ddp import BatchedDDP
    N, B = 4, 3
    cfg = make_config(MODEL_SRBD, N, 0.05, TIGHT)
    b = make_batch(MODEL_SRBD, N, B)
    r = BatchedDDP(cfg).solve(b["x0"], b["params"], b["X0"], b["U0"])
    Ug, cost = r.U.cpu().numpy(), r.cost.cpu().numpy()
    ref = SrbdRef(N, 0.05)
    for i in range(B):
        U, J = _nls(ref, b["x0"][i], b["U0"][i], b["params"][i])
        assert relerr(Ug[i], U) < 1e-6, i
        assert cost[i] == pytest.approx(J, rel=1e-12)


# ------------------------------------------------------------------------- extensions: inequality barriers, LIP-style tail
INEQ = dict(friction_cone_weight=5.0, friction_cone_mu=0.7, friction_cone_sharpness=8.0, force_bound_weight=2.0, force_bound=0.15,
            unilateral_weight=3.0, cdot_bound_weight=4.0, cdot_bound=0.4, bound_sharpness=7.0)


@pytest.mark.parametrize("kind", [0, 1, 3])
def test_oracle_derivatives_of_barriers_and_tail_match_ref_model(kind):
    """Oracle lx, lu, lxx, lux, luu, fx, fu with every inequality barrier on and, kind 3, on a node of the LIP-style tail,
    against complex-step / difference derivatives of the numpy transcription."""
    from tests.helpers import random_point
    rng = np.random.default_rng(31 + kind)
    cfg = make_config(MODEL_SRBD, 20, 0.05, dict(INEQ, lip_tail_start=5))
    ref = SrbdRef(20, 0.05, 0, lip_tail_start=5, ineq=INEQ)
    x, u, p = random_point(rng, MODEL_SRBD)
    x[25:37] = rng.uniform(-0.5, 0.5, 12)
    d = O.derivs(cfg, kind, x, u, p)
    assert O.cost(cfg, kind, x, u, p) == pytest.approx(ref.L(x, u, p, kind), rel=1e-13)
    lx, lu = ref.node_gradient(x, u, p, kind)
    assert relerr(d["lx"], lx) < 1e-11 and relerr(d["lu"], lu) < 1e-11
    H = ref.node_hessian(x, u, p, kind)
    assert relerr(d["lxx"], H[:37, :37]) < 1e-6 and relerr(d["luu"], H[37:, 37:]) < 1e-6 and relerr(d["lux"], H[37:, :37]) < 1e-6
    assert np.max(np.abs(O.dynamics(cfg, x, u, kind) - ref.f(x, u, kind))) < 1e-14
    if kind == 3:      # no rotational dynamics on the tail: w is carried over, nothing depends on it through wdot
        assert np.array_equal(O.dynamics(cfg, x, u, kind)[22:25], x[22:25])
        assert np.all(d["fx"][22:25] == np.eye(37)[22:25]) and np.all(d["fu"][22:25] == 0)


@pytest.mark.parametrize("opts", [dict(INEQ), dict(lip_tail_start=4), dict(INEQ, lip_tail_start=6)])
def test_oracle_solution_with_extensions_is_stationary(opts):
    N, B = 10, 2
    cfg = make_config(MODEL_SRBD, N, 0.05, dict(TIGHT, **opts))
    b = make_batch(MODEL_SRBD, N, B, enumerate_schedules=True, first=3)
    ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=2)
    assert (ro["status"] == 0).all()
    ineq = {k: v for k, v in opts.items() if k != "lip_tail_start"}
    ref = SrbdRef(N, 0.05, 0, lip_tail_start=opts.get("lip_tail_start", 0), ineq=ineq)
    for i in range(B):
        assert ro["cost"][i] == pytest.approx(ref.total_cost(b["x0"][i], ro["U"][i], b["params"][i]), rel=1e-9)
        g0 = np.abs(ref.reduced_gradient(b["x0"][i], b["U0"][i], b["params"][i])).max()
        gs = np.abs(ref.reduced_gradient(b["x0"][i], ro["U"][i], b["params"][i])).max()
        assert gs <= 1e-9 * g0, (i, gs, g0)


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [dict(INEQ), dict(lip_tail_start=8), dict(INEQ, lip_tail_start=10)])
def test_gpu_solution_with_extensions_is_stationary(opts):
    """The CUDA results with the inequality barriers and the model scheduler on zero the reduced gradient of the numpy
    transcription of the same extended problem (no oracle involved)."""
    from srbd_horizon_b200.ddp import BatchedDDP
    N, B = 20, 2
    cfg = make_config(MODEL_SRBD, N, 0.05, dict(TIGHT, **opts))
    b = make_batch(MODEL_SRBD, N, B, enumerate_schedules=True, first=3)
    r = BatchedDDP(cfg).solve(b["x0"], b["params"], b["X0"], b["U0"])
    assert (r.status.cpu().numpy() == 0).all()
    U, cost = r.U.cpu().numpy(), r.cost.cpu().numpy()
    ineq = {k: v for k, v in opts.items() if k != "lip_tail_start"}
    ref = SrbdRef(N, 0.05, 0, lip_tail_start=opts.get("lip_tail_start", 0), ineq=ineq)
    for i in range(B):
        assert cost[i] == pytest.approx(ref.total_cost(b["x0"][i], U[i], b["params"][i]), rel=1e-9)
        g0 = np.abs(ref.reduced_gradient(b["x0"][i], b["U0"][i], b["params"][i])).max()
        gs = np.abs(ref.reduced_gradient(b["x0"][i], U[i], b["params"][i])).max()
        assert gs <= 1e-9 * g0, (i, gs, g0)
