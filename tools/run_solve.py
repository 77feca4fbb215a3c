#!/usr/bin/env python
"""Small driver for profiling: solves one synthetic SRBD batch (used under ncu)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1184)
ap.add_argument("--N", type=int, default=50)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--no-gains", action="store_true")
ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="f32: the optional fp32 build (libsddp_f32.so)")
ap.add_argument("--order", default="schedule", choices=["schedule", "index"], help="dispatch order (see sddp_set_dispatch_order)")
a = ap.parse_args()
cfg = make_config(MODEL_SRBD, a.N, 0.05, {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3})
b = make_batch(MODEL_SRBD, a.N, a.batch, enumerate_schedules=True)
s = BatchedDDP(cfg, dtype=a.dtype)
t = lambda v: torch.as_tensor(v, dtype=s.tdtype, device="cuda")
x0, p, X0, U0 = t(b["x0"]), t(b["params"]), t(b["X0"]), t(b["U0"])
times = []
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = s.solve(x0, p, X0, U0, gains=not a.no_gains, history=False, order=None if a.order == "index" else "schedule")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    times.append(ms)
    it = r.iters.double()
    print(f"B={a.batch} N={a.N} ms={ms:.3f} mean_iters={it.mean().item():.3f} max_iters={int(it.max().item())} "
          f"converged={(r.status == 0).double().mean().item():.4f} us_per_node_iter_of_slowest={1e3 * ms / (it.max().item() * a.N):.2f}")
if len(times) > 2:
    ts = sorted(times[1:])
    print(f"SUMMARY B={a.batch} median_ms={ts[len(ts) // 2]:.3f} min_ms={ts[0]:.3f} (first rep excluded, {len(ts)} reps)")
