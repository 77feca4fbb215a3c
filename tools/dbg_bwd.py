import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch
from tests.helpers import relerr
EX = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}
for N, opts in ((20, {"lip_tail_start": 10}), (20, {})):
    cfg = make_config(MODEL_SRBD, N, 0.05, dict(EX, **opts))
    b = make_batch(MODEL_SRBD, N, 4, x_noise=0.01)
    s = BatchedDDP(cfg)
    X = b["X0"].copy(); X[:, 0] = b["x0"]
    D, _ = s.defects(X, b["U0"], b["params"])
    rc, K, k, dV = s.backward_pass(X, b["U0"], b["params"], D, 0.0)
    K, k, dV, D = (t.cpu().numpy() for t in (K, k, dV, D))
    for i in range(2):
        rc_o, K_o, k_o, dV_o = O.backward(cfg, X[i], b["U0"][i], b["params"][i], D[i], 0.0)
        print(N, opts, "prob", i, "dV", dV[i], dV_o)
        for kk in range(N - 1, -1, -1):
            print("   node", kk, "relerr K %.2e k %.2e" % (relerr(K[i, kk], K_o[kk]), np.abs(k[i, kk] - k_o[kk]).max() / max(1e-300, np.abs(k_o).max())), "max|K| %.2e" % np.abs(K_o[kk]).max())
