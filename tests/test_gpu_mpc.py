"""BASELINE configs[0] and configs[1]: the reference's example loops (dlip_example.py, dsrbd_example.py) driven through
the drop-in DDPSolver on the GPU, tick by tick against the CPU oracle solving the same problem from the same warm start."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O
from srbd_horizon_b200 import prb as P
from srbd_horizon_b200 import wpg
from srbd_horizon_b200.config import MODEL_LIP, MODEL_SRBD
from srbd_horizon_b200.ddp import DDPSolver
from srbd_horizon_b200.mpc import mpc_tick_references, plant_step
from tests.helpers import relerr

OPTS = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}     # dsrbd_example.py:55-58


def _loop(model, ticks, extra=None, stats=None):
    import time
    ns, T = 20, 1.0                                                             # dsrbd_example.py:30-31
    if model == MODEL_SRBD:
        prob = P.SRBDProblem(); prob.createSRBDProblem(ns, T)
        w_ref, otg = prob.w_ref, prob.orientation_tracking_gain
    else:
        prob = P.LIPProblem(); prob.createLIPProblem(ns, T)
        dummy = P.SRBDProblem(); dummy.createSRBDProblem(ns, T)                 # dlip_example.py:33-36, 85
        w_ref, otg = dummy.w_ref, dummy.orientation_tracking_gain
    solver = DDPSolver(prob.prb, dict(OPTS, **(extra or {})))
    gen = wpg.steps_phase(getattr(prob, "f", None), prob.c, prob.cdot, float(prob.initial_foot_position[0][2]), prob.c_ref, w_ref,
                          otg, prob.cdot_switch, ns, number_of_legs=2, contact_model=prob.contact_model)
    state = prob.getInitialState()
    solver.set_u_warmstart(np.tile(prob.getStaticInput()[:, None], (1, ns)))
    cfg = solver.cfg
    Xw = np.tile(state, (ns + 1, 1)); Uw = np.tile(prob.getStaticInput(), (ns, 1))
    worst = 0.0
    for tick in range(ticks):
        solver.setInitialState(state)
        walking = tick >= 5
        mpc_tick_references(prob, [0.5 if walking else 0.0, 0.0, 0.0])          # dsrbd_example.py:102-122
        gen.set("step" if walking else "standing")                              # :126-131
        params = solver.get_params_value()
        t0 = time.perf_counter()
        ok = solver.solve()                                                     # :135
        if stats is not None:
            stats.setdefault("ms", []).append(1e3 * (time.perf_counter() - t0))
            stats.setdefault("iters", []).append(int(solver.last["iters"]))
        ro = O.solve_batch(cfg, state[None], params[None], Xw[None], Uw[None])
        sol = solver.getSolutionDict()
        assert ok == (ro["status"][0] == 0)
        assert solver.last["iters"] == ro["iters"][0]
        ex, eu = relerr(sol["x_opt"].T, ro["X"][0]), relerr(sol["u_opt"].T, ro["U"][0])
        worst = max(worst, ex, eu)
        assert ex < 1e-9 and eu < 1e-9, (tick, ex, eu)
        # the solution dict has the reference's keys and shapes (ddp.py:125-151)
        assert sol["r"].shape == (3, ns + 1) and sol["c3"].shape == (3, ns + 1) and sol["u_opt"].shape[1] == ns
        if model == MODEL_SRBD:
            assert sol["o"].shape == (4, ns + 1) and sol["f2"].shape == (3, ns) and sol["cddot0"].shape == (3, ns)
        else:
            assert sol["z"].shape == (3, ns)
        Xw, Uw = ro["X"][0], ro["U"][0]          # both sides warm start from their own previous solution
        state = plant_step(solver.ddp_solver, state, sol["u_opt"][:, 0])        # :158-160
        ref = O.dynamics(cfg, solver._x0, sol["u_opt"][:, 0])
        if model == MODEL_SRBD:
            ref[3:7] /= np.linalg.norm(ref[3:7])
        assert np.max(np.abs(state - ref)) < 1e-13
    return worst


def test_config0_dlip_example_closed_loop():
    assert _loop(MODEL_LIP, 30) < 1e-9


def test_config1_dsrbd_example_closed_loop():
    assert _loop(MODEL_SRBD, 30) < 1e-9


def test_model_scheduler_closed_loop():
    """SURVEY 8f N4 (README.md:7, isrbd_example.py:344-353): the dsrbd_example loop with SRBD on nodes 0..9 and the LIP-style
    model on the tail, tick by tick against the oracle at 1e-9.  The tail drops the rotational dynamics (no rigid-body pack,
    no curvature entries on those nodes); the Riccati recursion itself is the same per node, so a tick costs what its
    iteration count says: the test prints both loops' iteration counts and solve times and checks that the mixed horizon
    does not need more iterations than the full model over the loop."""
    full, mixed = {}, {}
    assert _loop(MODEL_SRBD, 30, stats=full) < 1e-9
    assert _loop(MODEL_SRBD, 30, extra={"lip_tail_start": 10}, stats=mixed) < 1e-9
    med = lambda v: float(np.median(v))
    print("closed loop, 30 ticks: full SRBD %d iterations, median %.3f ms per tick; SRBD + LIP-style tail %d iterations, median %.3f ms"
          % (sum(full["iters"]), med(full["ms"]), sum(mixed["iters"]), med(mixed["ms"])))
    assert sum(mixed["iters"]) <= sum(full["iters"]) + 3
