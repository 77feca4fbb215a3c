"""The optional fp32 build (libsddp_f32.so, `BatchedDDP(cfg, dtype="f32")`) against the fp64 CPU oracle.

north_star: "the fp64 build matches ... within 1e-9 relative, and an optional fp32 build matches within a stated tolerance".
The fp32 build is NARROWER than the reference (fp64); it keeps float arrays and a float Riccati recursion / derivatives /
rollout, and double only for the trajectory cost, the line-search / convergence logic and the tensor-core accumulators.

STATED TOLERANCE (enforced below, norm-wise max|a - b| / max|b| per problem, against the fp64 oracle on the same inputs):
  trajectories X, inputs U 1e-2   (measured on B200 over the first 6,155 problems of BASELINE configs[4], bench.py --dtype f32:
                                   median 1.1e-6, worst 2.1e-3; over the 60 gait schedules x 2 below: X 2e-4, U 8e-4)
  final cost               1e-5   (measured: worst 7e-7)
  gains K                  2e-2   (measured: median 1e-5, worst 8e-4; conditioning of Quu, up to 1e8)
  status                   identical; iteration counts equal on >= 85 % of the problems
SURVEY.md section 7 predicted ~1e-3 for fp32 because of the 1 .. 1e8 span of the weights (prb.py:202-204, ddp.py:181).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O
from srbd_horizon_b200.config import MODEL_LIP, MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP, DDPSolver
from srbd_horizon_b200.problems import make_batch
from tests.helpers import relerr

EX_OPTS = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}   # dsrbd_example.py:55-58
TOL_X, TOL_U, TOL_COST, TOL_K = 1e-2, 1e-2, 1e-5, 2e-2
# inequality extensions (include/sddp.h): all of them on, tight enough to be active at the solution
INEQ = dict(friction_cone_weight=5.0, friction_cone_mu=0.7, friction_cone_sharpness=8.0, force_bound_weight=2.0, force_bound=0.15,
            unilateral_weight=3.0, cdot_bound_weight=4.0, cdot_bound=0.4, bound_sharpness=7.0)


def _check(cfg, b, B, gains=True):
    s = BatchedDDP(cfg, dtype="f32")
    assert s.L.sddp_real_bytes() == 4
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device="cuda")
    r = s.solve(t(b["x0"]), t(b["params"]), t(b["X0"]), t(b["U0"]), order="schedule" if B > 1 else None)
    assert r.X.dtype == torch.float32 and r.K.dtype == torch.float32
    ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=8)
    X, U, K = (a.double().cpu().numpy() for a in (r.X, r.U, r.K))
    iters, status, cost = r.iters.cpu().numpy(), r.status.cpu().numpy(), r.cost.double().cpu().numpy()
    np.testing.assert_array_equal(status, ro["status"])
    di = np.abs(iters - ro["iters"])
    assert (di == 0).mean() >= 0.85, (di.max(), (di == 0).mean())
    ex = max(relerr(X[i], ro["X"][i]) for i in range(B))
    eu = max(relerr(U[i], ro["U"][i]) for i in range(B))
    ek = max(relerr(K[i], ro["K"][i]) for i in range(B) if iters[i] == ro["iters"][i])
    ec = float(np.max(np.abs(cost - ro["cost"]) / np.abs(ro["cost"])))
    print("fp32 vs fp64 oracle: X %.1e U %.1e K %.1e cost %.1e; iteration counts differ on %d of %d" % (ex, eu, ek, ec, int((di != 0).sum()), B))
    assert ex < TOL_X and eu < TOL_U and ec < TOL_COST and ek < TOL_K, (ex, eu, ek, ec)
    return r, ro


def test_f32_config4_schedules_within_stated_tolerance():
    """BASELINE configs[4] in small: N = 50, two problems of each of the 60 wpg gait schedules."""
    B, N = 120, 50
    cfg = make_config(MODEL_SRBD, N, 0.05, EX_OPTS)
    _check(cfg, make_batch(MODEL_SRBD, N, B, enumerate_schedules=True), B)


@pytest.mark.parametrize("model,N,opts", [
    (MODEL_SRBD, 20, {}),                                   # dsrbd_example.py configuration (configs[1])
    (MODEL_SRBD, 50, {"defect_contraction_rate": 0.5}),     # fixed defect contraction (configs[3])
    (MODEL_SRBD, 20, {"dense_backward": 1}),                # generic dense Riccati kernel
    (MODEL_SRBD, 20, {"lip_tail_start": 10}),               # model scheduler
    (MODEL_LIP, 20, {}),                                    # dlip_example.py configuration (configs[0])
    (MODEL_SRBD, 20, dict(INEQ)),                           # every inequality barrier on (the SrbdT<true> instantiation)
])
def test_f32_solve_within_stated_tolerance(model, N, opts):
    B = 24
    cfg = make_config(model, N, 0.05, dict(EX_OPTS, **opts))
    _check(cfg, make_batch(model, N, B, seed=3), B)


def test_f32_host_path_equals_device_path():
    """sddp_solve_batch_host of the fp32 library (float host buffers, staged and host-direct) == its device entry point, bit for bit."""
    B, N = 64, 20
    cfg = make_config(MODEL_SRBD, N, 0.05, EX_OPTS)
    b = make_batch(MODEL_SRBD, N, B, seed=5)
    s = BatchedDDP(cfg, dtype="f32")
    t = lambda a: torch.as_tensor(a, dtype=torch.float32, device="cuda")
    r = s.solve(t(b["x0"]), t(b["params"]), t(b["X0"]), t(b["U0"]))
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    h = s.solve_host(f(b["x0"]), f(b["params"]), f(b["X0"]), f(b["U0"]), gains=True)
    assert h["X"].dtype == np.float32
    np.testing.assert_array_equal(h["X"], r.X.cpu().numpy())
    np.testing.assert_array_equal(h["U"], r.U.cpu().numpy())
    np.testing.assert_array_equal(h["K"], r.K.cpu().numpy())
    np.testing.assert_array_equal(h["iters"], r.iters.cpu().numpy())
    pin = lambda a: torch.from_numpy(f(a)).pin_memory().numpy()
    out = {k: torch.empty(v.shape, dtype=torch.from_numpy(v).dtype).pin_memory().numpy() for k, v in h.items() if k in ("X", "U", "iters", "status", "cost")}
    d = s.solve_host(pin(b["x0"]), pin(b["params"]), pin(b["X0"]), pin(b["U0"]), out=out)      # host-direct: one launch
    np.testing.assert_array_equal(d["X"], h["X"])
    np.testing.assert_array_equal(d["U"], h["U"])
    np.testing.assert_array_equal(d["cost"], h["cost"])


def test_f32_batched_closed_loop():
    """BatchedMPC (device-side schedule shift, solve, plant step) on the fp32 build: ten ticks of 32 robots stay within the
    stated tolerance of the same loop on the fp64 build."""
    from srbd_horizon_b200.mpc import BatchedMPC
    B, N = 32, 20
    cfg = make_config(MODEL_SRBD, N, 0.05, EX_OPTS)
    b = make_batch(MODEL_SRBD, N, B, seed=11)
    loops = [BatchedMPC(BatchedDDP(cfg, dtype=dt), b["x0"], b["params"], b["U0"]) for dt in ("f64", "f32")]
    act = np.zeros(B, dtype=np.int32)
    cmd = np.tile([0.3, 0.0, 0.0], (B, 1))
    for _ in range(10):
        for m in loops:
            r = m.tick(act, cmd)
            assert bool((r.status == 0).all())
    s64, s32 = loops[0].state.cpu().numpy(), loops[1].state.double().cpu().numpy()
    assert loops[1].state.dtype == torch.float32
    assert relerr(s32, s64) < TOL_X, relerr(s32, s64)


def test_f32_drop_in_solver():
    """The reference-facing class on the fp32 build: DDPSolver(prb, opts, dtype="f32") returns float64 dicts like pyddp."""
    from srbd_horizon_b200 import prb as P
    prob = P.SRBDProblem(); prob.createSRBDProblem(20, 1.0)
    sol64, sol32 = DDPSolver(prob.prb, dict(EX_OPTS)), DDPSolver(prob.prb, dict(EX_OPTS), dtype="f32")
    for s in (sol64, sol32):
        s.setInitialState(prob.getInitialState())
        s.set_u_warmstart(np.tile(prob.getStaticInput()[:, None], (1, 20)))
        assert s.solve()
    a, b = sol64.getSolutionDict(), sol32.getSolutionDict()
    assert b["x_opt"].dtype == np.float64 and b["x_opt"].shape == a["x_opt"].shape
    assert relerr(b["x_opt"], a["x_opt"]) < TOL_X and relerr(b["u_opt"], a["u_opt"]) < TOL_U
