"""Experiment: how much faster per Riccati node-iteration is the solve kernel when every co-resident CTA runs the same
problem (natural lockstep: shared instruction stream in the SM's instruction caches) than on the mixed batch?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch
N = 50
cfg = make_config(MODEL_SRBD, N, 0.05, {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3})
s = BatchedDDP(cfg)
t = lambda v: torch.as_tensor(v, dtype=torch.float64, device="cuda")
for B in (592, 4736):
    b = make_batch(MODEL_SRBD, N, B, enumerate_schedules=True)
    for name, sel in (("mixed", np.arange(B)), ("identical", np.zeros(B, dtype=int)), ("identical#7", np.full(B, 7)), ("groups of 4", (np.arange(B) // 4) * 4 % B)):
        x0, p, X0, U0 = (t(b[k][sel]) for k in ("x0", "params", "X0", "U0"))
        ms = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = s.solve(x0, p, X0, U0, gains=True, history=False, order="schedule" if name == "mixed" else None); e1.record()
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        it = r.iters.double()
        ni = float(it.sum()) * N
        slots = min(B, 592)
        print(f"B={B} {name:12s} ms={min(ms):8.3f} mean_iters={it.mean().item():.3f} max_iters={int(it.max())} cycles/node-iter/CTA={min(ms)*1e-3*1.965e9*slots/ni:8.0f}")
