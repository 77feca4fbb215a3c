"""Per-problem parity report on BASELINE configs[4] (120 problems, 60 schedules, N = 50): the structured kernel in its
latency variant (B < slots), its throughput variant (the 120 problems padded to a full grid) and the dense kernel,
each against the CPU oracle.  Developer tool (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import oracle as O
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch
from tests.helpers import relerr

EX = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}
B, N = 120, 50
b = make_batch(MODEL_SRBD, N, B, enumerate_schedules=True)
cfg = make_config(MODEL_SRBD, N, 0.05, EX)
ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=8)
t = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda")


def report(name, r, n=B):
    it = r.iters.cpu().numpy()[:n]
    X, U, K, k, h = (getattr(r, f).cpu().numpy()[:n] for f in ("X", "U", "K", "k", "hist"))
    print(f"== {name}: iters eq {(it == ro['iters']).all()} status eq {(r.status.cpu().numpy()[:n] == ro['status']).all()}")
    rows = []
    for i in range(n):
        m = min(it[i], ro["iters"][i])
        eK = relerr(K[i], ro["K"][i])
        eKn = max(relerr(K[i, j], ro["K"][i, j]) for j in range(N))
        rows.append((eK, eKn, relerr(X[i], ro["X"][i]), relerr(U[i], ro["U"][i]), relerr(h[i, :m, 0], ro["hist"][i, :m, 0]),
                     np.max(np.abs(k[i] - ro["k"][i])), i, int(it[i])))
    rows.sort(reverse=True)
    for r_ in rows[:6]:
        print("   K %.2e (worst node %.2e)  X %.2e U %.2e hist %.2e  k abs %.2e  prob %d iters %d" % r_)
    print("   K > 1e-8:", sum(r_[0] > 1e-8 for r_ in rows), " X/U > 1e-9:", sum(max(r_[2], r_[3]) > 1e-9 for r_ in rows))


s = BatchedDDP(cfg)
report("structured, latency variant (B=120)", s.solve(t(b["x0"]), t(b["params"]), t(b["X0"]), t(b["U0"]), order="schedule"))
rep = 6
tile = lambda a: t(np.concatenate([a] * rep, axis=0))
report("structured, throughput variant (B=720)", s.solve(tile(b["x0"]), tile(b["params"]), tile(b["X0"]), tile(b["U0"])))
sd = BatchedDDP(make_config(MODEL_SRBD, N, 0.05, dict(EX, dense_backward=1)))
report("dense", sd.solve(t(b["x0"]), t(b["params"]), t(b["X0"]), t(b["U0"])))
