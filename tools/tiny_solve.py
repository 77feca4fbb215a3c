#!/usr/bin/env python
"""Tiny solve for compute-sanitizer runs (memcheck / racecheck): B problems, short horizon, both models."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srbd_horizon_b200.config import MODEL_LIP, MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch

dtype = sys.argv[1] if len(sys.argv) > 1 else "f64"      # f32: the optional fp32 build
for model, N, opts in ((MODEL_SRBD, 6, {}), (MODEL_SRBD, 5, {"mu0": 1e-3, "defect_contraction_rate": 0.5}), (MODEL_LIP, 6, {}),
                       (MODEL_SRBD, 6, {"lip_tail_start": 3, "friction_cone_weight": 1.0, "force_bound_weight": 1.0, "force_bound": 0.2})):
    cfg = make_config(model, N, 0.05, dict({"max_iters": 6, "alpha_converge_threshold": 1e-3, "beta": 1e-3}, **opts))
    b = make_batch(model, N, 3, x_noise=0.05)
    s = BatchedDDP(cfg, dtype=dtype)
    r = s.solve(b["x0"], b["params"], b["X0"], b["U0"])
    torch.cuda.synchronize()
    print(model, N, opts, r.iters.tolist(), r.status.tolist())
