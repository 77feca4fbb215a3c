#!/usr/bin/env python
"""Per-phase cycle counts of one CTA (developer tool).  Build the profiling variant first:
   nvcc ... -DSDDP_PROFILE -o build_ab/libsddp_prof.so srbd_horizon_b200/csrc/sddp.cu
   SDDP_LIB=$PWD/build_ab/libsddp_prof.so python tools/phase_timer.py [--batch B]"""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srbd_horizon_b200 import _lib
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=1); ap.add_argument("--N", type=int, default=50)
a = ap.parse_args()
cfg = make_config(MODEL_SRBD, a.N, 0.05, {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3})
b = make_batch(MODEL_SRBD, a.N, a.batch, enumerate_schedules=True)
s = BatchedDDP(cfg)
L = _lib.lib()
out = (ctypes.c_longlong * 32)()
names = {0: "packs", 1: "backward total", 2: "forward wave", 3: "accept/copy", 8: "bwd: wait+sync (top)", 9: "bwd: expand", 10: "bwd: c1 (Quu, gap)",
         11: "bwd: d1 || c2,c3", 12: "bwd: d2 (RHS subst)", 13: "bwd: f,g (syrk, K)", 14: "  d1 alone (thread 0, since c1 end)", 15: "  c2 alone (thread 32)", 16: "  c2+c3 (thread 32)", 18: "  c2+c3+e phase 1 (thread 32)", 17: "  c2+c3+e (thread 32)", 20: "  expand: zero fill + sync", 21: "  expand: z-block + sync", 22: "  expand: affine + sync"}
for rep in range(2):
    L.sddp_debug_profile(out, 1)
    r = s.solve(b["x0"], b["params"], b["X0"], b["U0"], gains=False, history=False)
    torch.cuda.synchronize()
    L.sddp_debug_profile(out, 0)
its = int(r.iters[0].item()) if a.batch == 1 else None
print(f"batch={a.batch} N={a.N} iters(problem 0)={its}; cycles of CTA 0 (all problems it solved):")
tot = sum(out[i] for i in (0, 1, 2, 3))
for i, n in names.items():
    print(f"  {n:26s} {out[i]:>12d}  {100.0 * out[i] / max(tot, 1):5.1f}%" + (f"   {out[i] / (its * a.N):8.0f} cyc/node-iter" if its else ""))
