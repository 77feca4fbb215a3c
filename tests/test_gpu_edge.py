"""Edge cases of the CUDA path against the oracle: shortest and long horizons, tiny batches, iteration caps,
non-finite inputs, and BASELINE configs[3] (multiple shooting with a fixed defect contraction rate) at full size."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O
from srbd_horizon_b200.config import DIMS, MODEL_LIP, MODEL_SRBD, STATUS_MAX_ITERS, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch
from tests.helpers import relerr

EX = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}
cpu = lambda t: t.detach().cpu().numpy()


@pytest.mark.parametrize("model", [MODEL_SRBD, MODEL_LIP])
@pytest.mark.parametrize("N", [1, 2, 3, 7, 64])
def test_horizon_lengths(model, N):
    B = 3
    cfg = make_config(model, N, 0.05, EX)
    b = make_batch(model, N, B, x_noise=0.01)
    r = BatchedDDP(cfg).solve(b["x0"], b["params"], b["X0"], b["U0"])
    ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"])
    np.testing.assert_array_equal(cpu(r.iters), ro["iters"])
    np.testing.assert_array_equal(cpu(r.status), ro["status"])
    assert relerr(cpu(r.X), ro["X"]) < 1e-9 and relerr(cpu(r.U), ro["U"]) < 1e-9
    assert relerr(cpu(r.cost), ro["cost"]) < 1e-9


def test_iteration_cap_and_alpha_threshold():
    """max_iters = 1 / 2 stops with MAX_ITERS exactly like the oracle; the adapter's default threshold 1e-1 (ddp.py:23)
    gives four candidate step sizes -- one wave."""
    for opts in ({"max_iters": 1}, {"max_iters": 2}, {"max_iters": 3, "alpha_converge_threshold": 1e-1, "beta": 1e-4}):
        cfg = make_config(MODEL_SRBD, 20, 0.05, opts)
        b = make_batch(MODEL_SRBD, 20, 4, x_noise=0.01)
        r = BatchedDDP(cfg).solve(b["x0"], b["params"], b["X0"], b["U0"])
        ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"])
        np.testing.assert_array_equal(cpu(r.status), ro["status"])
        np.testing.assert_array_equal(cpu(r.iters), ro["iters"])
        assert (ro["status"] == STATUS_MAX_ITERS).any()
        assert relerr(cpu(r.X), ro["X"]) < 1e-9 and relerr(cpu(r.hist)[..., 0], ro["hist"][..., 0]) < 1e-9


def test_non_finite_problem_does_not_poison_the_batch():
    """A NaN initial state fails its own problem (status != 0) and leaves its neighbours bit-identical."""
    cfg = make_config(MODEL_SRBD, 20, 0.05, EX)
    b = make_batch(MODEL_SRBD, 20, 6)
    s = BatchedDDP(cfg)
    good = s.solve(b["x0"], b["params"], b["X0"], b["U0"])
    x0 = b["x0"].copy(); X0 = b["X0"].copy()
    x0[2, 0] = np.nan; X0[2, :, 0] = np.nan
    bad = s.solve(x0, b["params"], X0, b["U0"])
    st = cpu(bad.status)
    assert st[2] != 0
    keep = [0, 1, 3, 4, 5]
    assert (st[keep] == 0).all()
    np.testing.assert_array_equal(cpu(bad.X)[keep], cpu(good.X)[keep])
    ro = O.solve_batch(cfg, x0, b["params"], X0, b["U0"])
    assert ro["status"][2] == st[2]


def test_line_search_beyond_the_first_wave():
    """A poor warm start (10 cm / 0.1 m/s noise on the state guess) needs step sizes down to 1/32: the later waves of
    four candidates must pick the same alpha as the oracle's sequential backtracking (T8)."""
    cfg = make_config(MODEL_SRBD, 20, 0.05, EX)
    b = make_batch(MODEL_SRBD, 20, 16, x_noise=0.1)
    r = BatchedDDP(cfg).solve(b["x0"], b["params"], b["X0"], b["U0"])
    ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=4)
    h, ho = cpu(r.hist), ro["hist"]
    np.testing.assert_array_equal(cpu(r.iters), ro["iters"])
    np.testing.assert_array_equal(h[..., 1], ho[..., 1])
    small = ho[..., 1][(ho[..., 1] > 0) & (ho[..., 1] < 1)]
    assert small.size >= 5 and small.min() <= 1.0 / 16, "the case must exercise several waves"
    assert relerr(cpu(r.X), ro["X"]) < 1e-9 and relerr(cpu(r.U), ro["U"]) < 1e-9


def test_config3_fixed_contraction_full_size_properties():
    """BASELINE configs[3]: multiple shooting, fixed defect contraction rate, SRBD, batch 16K, N = 50.  Without an
    oracle at this size: after k accepted steps every defect is (1 - rho)^k times its initial value, and the first
    512 problems agree with the oracle."""
    B, N, rho = 16384, 50, 0.5
    cfg = make_config(MODEL_SRBD, N, 0.05, dict(EX, defect_contraction_rate=rho))
    b = make_batch(MODEL_SRBD, N, B, x_noise=0.01)
    s = BatchedDDP(cfg)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda")
    x0, p, X0, U0 = t(b["x0"]), t(b["params"]), t(b["X0"]), t(b["U0"])
    X0[:, 0] = x0
    D0, _ = s.defects(X0, U0, p)
    r = s.solve(x0, p, X0, U0, gains=False)
    iters, status = cpu(r.iters), cpu(r.status)
    assert (status == 0).mean() > 0.99
    hist = cpu(r.hist)
    d0 = cpu(D0.abs().amax(dim=(1, 2)))
    for i in range(0, B, 257):
        n = iters[i]
        steps = np.cumsum(hist[i, :n, 1] > 0)
        np.testing.assert_allclose(hist[i, :n, 3], d0[i] * (1 - rho) ** steps, rtol=1e-12, atol=1e-300)
    D, J = s.defects(r.X, r.U, p)
    conv = torch.as_tensor(status == 0, device="cuda")
    assert float(D.abs().amax(dim=(1, 2))[conv].max()) <= cfg.defect_ths * 1.0001
    ro = O.solve_batch(cfg, b["x0"][:512], b["params"][:512], b["X0"][:512], b["U0"][:512], nthreads=16)
    np.testing.assert_array_equal(iters[:512], ro["iters"])
    assert relerr(cpu(r.X)[:512], ro["X"]) < 1e-9 and relerr(cpu(r.U)[:512], ro["U"]) < 1e-9


def test_dispatch_order_changes_nothing_but_the_schedule():
    """sddp_set_dispatch_order: any permutation (random, or grouped by contact schedule) gives bit-identical results,
    on the device entry point and on the chunked host entry point; a non-permutation is rejected."""
    import ctypes
    from srbd_horizon_b200 import _lib
    B, N = 700, 20
    cfg = make_config(MODEL_SRBD, N, 0.05, EX)
    b = make_batch(MODEL_SRBD, N, B, enumerate_schedules=True)
    s = BatchedDDP(cfg)
    r0 = s.solve(b["x0"], b["params"], b["X0"], b["U0"])
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3)).to(torch.int32)
    for order in (perm, "schedule"):
        r = s.solve(b["x0"], b["params"], b["X0"], b["U0"], order=order)
        for f in ("X", "U", "K", "k", "cost", "iters", "status", "hist"):
            assert torch.equal(getattr(r, f), getattr(r0, f)), (f, order if isinstance(order, str) else "perm")
    o = s.dispatch_order(torch.as_tensor(b["params"], device="cuda"))
    assert sorted(o.tolist()) == list(range(B))
    assert np.array_equal(cpu(o), s.dispatch_order(b["params"]))          # device and host versions agree
    os.environ.pop("SDDP_HOST_CHUNK", None)
    rh0 = s.solve_host(b["x0"], b["params"], b["X0"], b["U0"])
    for order in (perm.numpy(), "schedule"):
        rh = s.solve_host(b["x0"], b["params"], b["X0"], b["U0"], order=order)
        for f in ("X", "U", "cost", "iters", "status"):
            np.testing.assert_array_equal(rh[f], rh0[f])
            np.testing.assert_array_equal(rh[f], cpu(getattr(r0, f)))
    os.environ["SDDP_HOST_CHUNK"] = "256"           # three chunks: the permutation is split into chunk-local orders
    try:
        s2 = BatchedDDP(cfg)
        rh2 = s2.solve_host(b["x0"], b["params"], b["X0"], b["U0"], order=perm.numpy())
    finally:
        os.environ.pop("SDDP_HOST_CHUNK", None)
    for f in ("X", "U", "cost", "iters", "status"):
        np.testing.assert_array_equal(rh2[f], rh0[f])
    bad = np.zeros(B, dtype=np.int32)
    rc = s.L.sddp_set_dispatch_order(s.h, bad.ctypes.data_as(ctypes.c_void_p), B, 1)
    assert rc == -1 and b"permutation" in s.L.sddp_last_error(s.h)


def test_device_dispatch_order_is_validated_by_the_wrapper():
    """a device permutation reaches the kernel unchecked at the C ABI (documented in include/sddp.h: entries
    outside 0..B-1 are skipped, a problem nobody names keeps status -1); `BatchedDDP.solve(order=<tensor>)` checks it."""
    import ctypes
    B, N = 40, 10
    cfg = make_config(MODEL_SRBD, N, 0.05, EX)
    b = make_batch(MODEL_SRBD, N, B)
    s = BatchedDDP(cfg)
    dup = torch.arange(B, dtype=torch.int32); dup[3] = 4
    with pytest.raises(ValueError):
        s.solve(b["x0"], b["params"], b["X0"], b["U0"], order=dup)
    with pytest.raises(ValueError):
        s.solve(b["x0"], b["params"], b["X0"], b["U0"], order=torch.arange(B, dtype=torch.int32) + 1)
    # at the C ABI: out-of-range entries are skipped, the unnamed problem is left with status -1, nothing else is touched
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda")
    x0, p, X, U = t(b["x0"]), t(b["params"]), t(b["X0"]).clone(), t(b["U0"]).clone()
    it = torch.zeros(B, dtype=torch.int32, device="cuda"); st = torch.zeros(B, dtype=torch.int32, device="cuda"); c = torch.zeros(B, dtype=torch.float64, device="cuda")
    bad = torch.arange(B, dtype=torch.int32, device="cuda"); bad[7] = B + 5
    vp = lambda a: ctypes.c_void_p(a.data_ptr())
    assert s.L.sddp_set_dispatch_order(s.h, vp(bad), B, 0) == 0
    assert s.L.sddp_solve_batch(s.h, B, vp(x0), vp(p), vp(X), vp(U), None, None, None, vp(it), vp(st), vp(c), None) == 0
    s.L.sddp_set_dispatch_order(s.h, None, 0, 0)
    torch.cuda.synchronize()
    assert int(st[7]) == -1 and bool((st[torch.arange(B) != 7] == 0).all())


def test_solve_reuses_the_output_buffers_it_is_given():
    """BatchedDDP.solve(out=previous result): K, k, hist, iters, status, cost are overwritten in place (a steady-state loop
    allocates nothing per call), results unchanged."""
    B, N = 40, 20
    cfg = make_config(MODEL_SRBD, N, 0.05, EX)
    b = make_batch(MODEL_SRBD, N, B, seed=9)
    s = BatchedDDP(cfg)
    r0 = s.solve(b["x0"], b["params"], b["X0"], b["U0"])
    ref = {f: getattr(r0, f).clone() for f in ("X", "U", "K", "k", "hist", "iters", "status", "cost")}
    ptrs = {f: getattr(r0, f).data_ptr() for f in ("K", "k", "hist", "iters", "status", "cost")}
    r0.K.zero_(); r0.cost.zero_(); r0.iters.zero_()
    r1 = s.solve(b["x0"], b["params"], b["X0"], b["U0"], out=r0)
    for f, p in ptrs.items():
        assert getattr(r1, f).data_ptr() == p, f
    for f, v in ref.items():
        assert torch.equal(getattr(r1, f), v), f
    r2 = s.solve(b["x0"][:8], b["params"][:8], b["X0"][:8], b["U0"][:8], out=r1)      # other shapes: fresh buffers
    assert r2.K.data_ptr() != ptrs["K"] and torch.equal(r2.X, ref["X"][:8])
