import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch
from tests.helpers import relerr
EX = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}
for N, opts in ((50, {"defect_contraction_rate": 0.5}), (20, {"lip_tail_start": 10}), (50, {}), (20, {})):
    cfg = make_config(MODEL_SRBD, N, 0.05, dict(EX, **opts))
    b = make_batch(MODEL_SRBD, N, 24, x_noise=0.01)
    r = BatchedDDP(cfg).solve(b["x0"], b["params"], b["X0"], b["U0"])
    ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=8)
    it, st = r.iters.cpu().numpy(), r.status.cpu().numpy()
    X, U, h = r.X.cpu().numpy(), r.U.cpu().numpy(), r.hist.cpu().numpy()
    print(N, opts, "iters eq", (it == ro["iters"]).all(), "status eq", (st == ro["status"]).all())
    for i in range(24):
        n = min(it[i], ro["iters"][i])
        e = (relerr(X[i], ro["X"][i]), relerr(U[i], ro["U"][i]), relerr(h[i, :n, 0], ro["hist"][i, :n, 0]))
        if max(e) > 1e-10 or it[i] != ro["iters"][i]:
            print("  prob", i, "iters", it[i], ro["iters"][i], "relerr X U cost", e, "alpha", h[i, :n, 1], ro["hist"][i, :n, 1])
