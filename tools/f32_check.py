#!/usr/bin/env python
"""fp32 build against the fp64 build on the same problems (developer tool; the enforced tolerance is in tests/test_gpu_f32.py)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srbd_horizon_b200.config import MODEL_LIP, MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=240)
ap.add_argument("--N", type=int, default=50)
ap.add_argument("--model", default="srbd")
ap.add_argument("--opts", default="{}")
a = ap.parse_args()
model = MODEL_SRBD if a.model == "srbd" else MODEL_LIP
opts = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}
opts.update(eval(a.opts))
cfg = make_config(model, a.N, 0.05, opts)
b = make_batch(model, a.N, a.batch, enumerate_schedules=True)
res = {}
for dt in ("f64", "f32"):
    s = BatchedDDP(cfg, dtype=dt)
    t = lambda v: torch.as_tensor(v, dtype=s.tdtype, device="cuda")
    r = s.solve(t(b["x0"]), t(b["params"]), t(b["X0"]), t(b["U0"]), order="schedule")
    torch.cuda.synchronize()
    res[dt] = {k: getattr(r, k).double().cpu().numpy() for k in ("X", "U", "K", "k", "cost", "iters", "status", "hist")}
r64, r32 = res["f64"], res["f32"]
rel = lambda x, y: np.max(np.abs(x - y).reshape(len(x), -1), axis=1) / np.maximum(1e-300, np.max(np.abs(y).reshape(len(y), -1), axis=1))
print("status f64", np.bincount(r64["status"].astype(int) + 1), "f32", np.bincount(r32["status"].astype(int) + 1))
print("iters  f64 mean %.3f max %d | f32 mean %.3f max %d | differ on %d of %d" % (r64["iters"].mean(), r64["iters"].max(), r32["iters"].mean(),
      r32["iters"].max(), int((r64["iters"] != r32["iters"]).sum()), a.batch))
for k in ("X", "U", "K", "cost"):
    e = rel(r32[k], r64[k])
    print("%-5s rel err: median %.2e  p90 %.2e  max %.2e (problem %d)" % (k, np.median(e), np.quantile(e, 0.9), e.max(), int(e.argmax())))
w = int(rel(r32["U"], r64["U"]).argmax())
n = int(max(r64["iters"][w], r32["iters"][w]))
print("worst problem %d: iters %d / %d, status %d / %d" % (w, r64["iters"][w], r32["iters"][w], r64["status"][w], r32["status"][w]))
print(np.array2string(np.hstack([r64["hist"][w, :n], r32["hist"][w, :n]]), precision=6, max_line_width=200))
