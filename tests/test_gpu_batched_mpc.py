"""Device-side closed loop: sddp_mpc_advance == the host gait scheduler (itself pinned to the
reference's wpg.py) bit for bit, sddp_plant_step == the example's Euler + renormalise, and a batch of closed loops
equals the per-robot DDPSolver loops."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O
from srbd_horizon_b200 import prb as P
from srbd_horizon_b200 import wpg
from srbd_horizon_b200.config import DIMS, MODEL_LIP, MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP, DDPSolver
from srbd_horizon_b200.mpc import BatchedMPC, mpc_tick_references, plant_step
from srbd_horizon_b200.problems import make_batch, nominal

OPTS = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}
NAMES = {0: "step", 1: "standing", 2: "jump"}


def _host_problem(model, ns):
    if model == MODEL_SRBD:
        prob = P.SRBDProblem(); prob.createSRBDProblem(ns, ns * 0.05)
        w_ref, otg = prob.w_ref, prob.orientation_tracking_gain
    else:
        prob = P.LIPProblem(); prob.createLIPProblem(ns, ns * 0.05)
        d = P.SRBDProblem(); d.createSRBDProblem(ns, ns * 0.05)
        w_ref, otg = d.w_ref, d.orientation_tracking_gain
    gen = wpg.steps_phase(None, prob.c, prob.cdot, float(prob.initial_foot_position[0][2]), prob.c_ref, w_ref, otg,
                          prob.cdot_switch, ns, number_of_legs=2, contact_model=2)
    return prob, gen


@pytest.mark.parametrize("model", [MODEL_SRBD, MODEL_LIP])
def test_device_schedule_matches_host_scheduler(model):
    ns, B, ticks = 20, 5, 47
    rng = np.random.default_rng(3)
    cfg = make_config(model, ns, 0.05, OPTS)
    s = BatchedDDP(cfg)
    hosts = [_host_problem(model, ns) for _ in range(B)]
    p0 = np.stack([h[0].prb.flat_parameters() for h in hosts])
    x0, u0 = nominal(model)
    mpc = BatchedMPC(s, np.tile(x0, (B, 1)), p0)
    for t in range(ticks):
        acts = rng.choice(3, size=B, p=[0.6, 0.25, 0.15])
        cmd = np.concatenate([rng.uniform(-0.5, 0.5, (B, 2)), np.zeros((B, 1))], axis=1)
        mpc.advance_schedule(acts, cmd)
        for b, (prob, gen) in enumerate(hosts):
            mpc_tick_references(prob, cmd[b])
            gen.set(NAMES[int(acts[b])])
        ref = np.stack([h[0].prb.flat_parameters() for h in hosts])
        np.testing.assert_array_equal(mpc.params.cpu().numpy(), ref, err_msg=f"tick {t}")
    np.testing.assert_array_equal(mpc.step_counter.cpu().numpy(), np.full(B, ticks))


def test_batched_closed_loop_equals_single_robot_loops():
    ns, B, ticks = 20, 4, 12
    cfg = make_config(MODEL_SRBD, ns, 0.05, OPTS)
    s = BatchedDDP(cfg)
    hosts = [_host_problem(MODEL_SRBD, ns) for _ in range(B)]
    solvers = [DDPSolver(h[0].prb, dict(OPTS)) for h in hosts]
    x0, u0 = nominal(MODEL_SRBD)
    rng = np.random.default_rng(11)
    states = [x0 + np.concatenate([rng.uniform(-0.01, 0.01, 3), np.zeros(34)]) for _ in range(B)]
    for st in states:
        st[3:7] = [0, 0, 0, 1]
    U0 = np.tile(u0, (B, ns, 1))
    mpc = BatchedMPC(s, np.stack(states), np.stack([h[0].prb.flat_parameters() for h in hosts]), U0)
    for sv in solvers:
        sv.set_u_warmstart(U0[0].T)
    acts_seq = rng.choice(3, size=(ticks, B), p=[0.7, 0.2, 0.1])
    for t in range(ticks):
        cmd = np.tile([0.3, 0.0, 0.0], (B, 1))
        r = mpc.tick(acts_seq[t], cmd)
        Xb, Ub = r.X.cpu().numpy(), r.U.cpu().numpy()
        for b in range(B):
            prob, gen = hosts[b]
            solvers[b].setInitialState(states[b])
            mpc_tick_references(prob, cmd[b]); gen.set(NAMES[int(acts_seq[t, b])])
            solvers[b].solve()
            sol = solvers[b].getSolutionDict()
            np.testing.assert_array_equal(sol["x_opt"].T, Xb[b], err_msg=f"tick {t} robot {b}")
            np.testing.assert_array_equal(sol["u_opt"].T, Ub[b])
            states[b] = plant_step(solvers[b].ddp_solver, states[b], sol["u_opt"][:, 0])
        np.testing.assert_allclose(mpc.state.cpu().numpy(), np.stack(states), rtol=0, atol=1e-15)


def test_tick_as_cuda_graph_equals_tick():
    """BatchedMPC.capture(): the whole tick (schedule advance, solve, plant step) as one CUDA graph; replays give the same
    bits as the launch-by-launch tick, and a graph tick is one graph launch instead of six launches / memsets."""
    import time
    ns, B, ticks = 20, 6, 8
    cfg = make_config(MODEL_SRBD, ns, 0.05, OPTS)
    b = make_batch(MODEL_SRBD, ns, B, seed=21)
    loops = [BatchedMPC(BatchedDDP(cfg), b["x0"], b["params"], b["U0"]) for _ in range(2)]
    loops[1].capture()
    rng = np.random.default_rng(2)
    acts = rng.choice(3, size=(ticks, B), p=[0.7, 0.2, 0.1]).astype(np.int32)
    cmd = np.tile([0.3, 0.0, 0.0], (B, 1))
    for t in range(ticks):
        r0 = loops[0].tick(acts[t], cmd)
        r1 = loops[1].tick_graph(acts[t], cmd)
        for f in ("X", "U", "iters", "status", "cost"):
            assert torch.equal(getattr(r0, f), getattr(r1, f)), (t, f)
        assert torch.equal(loops[0].state, loops[1].state) and torch.equal(loops[0].params, loops[1].params)
        assert torch.equal(loops[0].step_counter, loops[1].step_counter)
    ms = []
    for m, fn in ((loops[0], loops[0].tick), (loops[1], loops[1].tick_graph)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(50):
            fn(acts[t % ticks], cmd)
        torch.cuda.synchronize()
        ms.append(1e3 * (time.perf_counter() - t0) / 50)
    print("closed-loop tick of %d robots: %.3f ms launch by launch, %.3f ms as one CUDA graph" % (B, ms[0], ms[1]))


def test_plant_step_matches_example():
    for model in (MODEL_SRBD, MODEL_LIP):
        cfg = make_config(model, 20, 0.05, OPTS)
        s = BatchedDDP(cfg)
        nx, nu, np_ = DIMS[model]
        rng = np.random.default_rng(5)
        x0, u0 = nominal(model)
        B = 7
        X = np.tile(x0, (B, 1)) + rng.uniform(-0.02, 0.02, (B, nx))
        U = np.tile(u0, (B, 20, 1)) + rng.uniform(-0.02, 0.02, (B, 20, nu))
        mpc = BatchedMPC(s, X, np.zeros((B, 21, np_)), U)
        mpc.plant_step()
        out = mpc.state.cpu().numpy()
        for b in range(B):
            ref = O.dynamics(cfg, X[b], U[b, 0])
            if model == MODEL_SRBD:
                ref[3:7] /= np.linalg.norm(ref[3:7])
            assert np.max(np.abs(out[b] - ref)) < 1e-14


def test_large_fleet_ticks_use_the_dispatch_hint_and_change_nothing():
    """B >= 1024 robots: BatchedMPC.tick dispatches the solve by the previous tick's iteration counts / the contact
    schedule; the closed loop must be bit-identical to the same ticks solved in index order."""
    from srbd_horizon_b200.problems import make_batch
    ns, B, ticks = 12, 1500, 3
    cfg = make_config(MODEL_SRBD, ns, 0.05, OPTS)
    b = make_batch(MODEL_SRBD, ns, B, enumerate_schedules=True)
    rng = np.random.default_rng(5)
    acts = rng.choice(3, size=(ticks, B), p=[0.7, 0.2, 0.1])
    cmd = np.tile([0.3, 0.0, 0.0], (B, 1))
    runs = []
    for hinted in (True, False):
        s = BatchedDDP(cfg)
        mpc = BatchedMPC(s, b["x0"], b["params"], b["U0"])
        if not hinted:      # same ticks, index order: call the pieces of tick() by hand
            for t in range(ticks):
                mpc.advance_schedule(acts[t], cmd)
                mpc.last = s.solve(mpc.state, mpc.params, mpc.X, mpc.U, gains=False, history=False, inplace=True)
                mpc.plant_step()
        else:
            for t in range(ticks):
                mpc.tick(acts[t], cmd)
        runs.append((mpc.state.clone(), mpc.X.clone(), mpc.U.clone(), mpc.last.iters.clone()))
    for a, c in zip(runs[0], runs[1]):
        assert torch.equal(a, c)
