"""The batched schedule generator is the closed form of driving the step-by-step gait scheduler
(which tests/test_wpg.py pins to the reference's wpg.py)."""
import numpy as np
import pytest

from srbd_horizon_b200 import prb as P
from srbd_horizon_b200 import wpg
from srbd_horizon_b200.config import DIMS, MODEL_LIP, MODEL_SRBD
from srbd_horizon_b200.problems import make_batch, schedule_params


def _drive(model, N, action, s0, rdot_ref):
    if model == MODEL_SRBD:
        pr = P.SRBDProblem(); pr.createSRBDProblem(N, N * 0.05)
        w_ref, otg = pr.w_ref, pr.orientation_tracking_gain
    else:
        pr = P.LIPProblem(); pr.createLIPProblem(N, N * 0.05)
        dummy = P.SRBDProblem(); dummy.createSRBDProblem(N, N * 0.05)      # dlip_example.py:85
        w_ref, otg = dummy.w_ref, dummy.orientation_tracking_gain
    gen = wpg.steps_phase(None, pr.c, pr.cdot, float(pr.initial_foot_position[0][2]), pr.c_ref, w_ref, otg,
                          pr.cdot_switch, N, number_of_legs=2, contact_model=2)
    def tick(a):
        # dsrbd_example.py:102-106 shifts the gain back one node before wpg.set writes node N
        v = otg.getValues()
        for j in range(1, N + 1):
            otg.assign(v[:, j], nodes=j - 1)
        gen.set(a)

    if action == 0:
        for _ in range(N + 1 + s0):
            tick("step")
    else:
        for _ in range(N + 1):
            tick("standing")
        if action == 2:
            for _ in range(1 + s0 % 8):
                tick("jump")
    pr.rdot_ref.assign(rdot_ref, nodes=range(1, N + 1))
    return pr.prb.flat_parameters()


@pytest.mark.parametrize("model", [MODEL_SRBD, MODEL_LIP])
@pytest.mark.parametrize("N", [20, 50])
def test_schedule_closed_form(model, N):
    cases = [(0, 0), (0, 7), (0, 19), (1, 3), (2, 0), (2, 5), (2, 15)]
    actions = np.array([c[0] for c in cases]); s0 = np.array([c[1] for c in cases])
    ref = np.tile([0.3, -0.2, 0.0], (len(cases), 1))
    p = schedule_params(model, N, actions, s0, ref)
    for i, (a, s) in enumerate(cases):
        np.testing.assert_array_equal(p[i], _drive(model, N, a, s, ref[i]), err_msg=str((a, s)))


def test_batch_is_seeded_and_sliceable():
    a = make_batch(MODEL_SRBD, 50, 12)
    b = make_batch(MODEL_SRBD, 50, 4, first=8)
    for k in ("x0", "params", "X0", "U0"):
        np.testing.assert_array_equal(a[k][8:], b[k])
    nx, nu, np_ = DIMS[MODEL_SRBD]
    assert a["x0"].shape == (12, nx) and a["params"].shape == (12, 51, np_)
    assert np.allclose(np.linalg.norm(a["x0"][:, 3:7], axis=1), 1.0)
    e = make_batch(MODEL_SRBD, 20, 120, enumerate_schedules=True)
    assert len({(int(x), int(y)) for x, y in zip(e["actions"], e["s0"])}) == 60
    m = make_batch(MODEL_SRBD, 20, 3, x_noise=0.01)
    assert np.abs(m["X0"][:, 1:, 0:3] - m["x0"][:, None, 0:3]).max() > 0
    np.testing.assert_array_equal(m["X0"][:, 0], m["x0"])


def test_dispatch_order_from_keys_groups_and_sorts():
    """BatchedDDP.order_from_keys (host version, no GPU): groups ascending, inside a group the effort key descending,
    stable, always a permutation; the torch version gives the same permutation."""
    import torch
    from srbd_horizon_b200.ddp import BatchedDDP
    rng = np.random.default_rng(0)
    g = rng.integers(0, 7, size=500)
    e = rng.random(500)
    gf = g.astype(np.float64)                      # general keys (the hash of dispatch_order): two stable sorts, exact
    o = BatchedDDP.order_from_keys(gf, e)
    assert o.dtype == np.int32 and sorted(o.tolist()) == list(range(500))
    assert (np.diff(g[o]) >= 0).all()
    for k in range(7):
        assert (np.diff(e[o][g[o] == k]) <= 0).all()
    ot = BatchedDDP.order_from_keys(torch.as_tensor(gf), torch.as_tensor(e))
    assert np.array_equal(ot.numpy(), o)
    # few integer groups (the caller's schedule ids): one radix sort of a 16-bit key, the effort quantised to 10 bits
    oi = BatchedDDP.order_from_keys(g, e)
    assert oi.dtype == np.int32 and sorted(oi.tolist()) == list(range(500))
    assert (np.diff(g[oi]) >= 0).all()
    quantum = (e.max() - e.min()) / 1023.0
    for k in range(7):
        assert (np.diff(e[oi][g[oi] == k]) <= quantum * 1.001).all()
