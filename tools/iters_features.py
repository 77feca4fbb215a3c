#!/usr/bin/env python
"""Developer tool: which input features predict the DDP iteration count of a problem (for the dispatch order)?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch, nominal
from srbd_horizon_b200.config import RobotConstants
B = 16384
cfg = make_config(MODEL_SRBD, 50, 0.05, {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3})
b = make_batch(MODEL_SRBD, 50, B, enumerate_schedules=True)
r = BatchedDDP(cfg).solve(b["x0"], b["params"], b["X0"], b["U0"], gains=False, history=True)
it = r.iters.cpu().numpy().astype(float)
xn, un = nominal(MODEL_SRBD, RobotConstants())
dx = b["x0"] - xn
feats = {"|dr|": np.linalg.norm(dx[:, 0:3], axis=1), "|do|": np.linalg.norm(dx[:, 3:6], axis=1), "|dc|": np.linalg.norm(dx[:, 7:19], axis=1),
         "|rdot0|": np.linalg.norm(dx[:, 19:22], axis=1), "|w0|": np.linalg.norm(dx[:, 22:25], axis=1),
         "|rdot_ref|": np.linalg.norm(b["params"][:, -1, 0:3], axis=1), "cost0": r.hist[:, 0, 0].cpu().numpy(), "defect0": r.hist[:, 0, 3].cpu().numpy()}
sched = b["actions"] * 20 + b["s0"]
print("overall std of iters", it.std())
res = it - np.array([it[sched == s].mean() for s in range(60)])[sched]
print("std within schedule groups", res.std())
for k, v in feats.items():
    c_all = np.corrcoef(v, it)[0, 1]
    vr = v - np.array([v[sched == s].mean() for s in range(60)])[sched]
    c_in = np.corrcoef(vr, res)[0, 1]
    print(f"{k:12s} corr with iters {c_all:+.3f}   within schedule {c_in:+.3f}")
