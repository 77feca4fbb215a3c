#!/usr/bin/env python
"""Per-warp timeline of one backward-pass node of CTA 0 (developer tool).  Build the stamped variant first:
   nvcc ... -DSDDP_STAMP -o build_ab/libsddp_stamp.so srbd_horizon_b200/csrc/sddp.cu
   SDDP_LIB=$PWD/build_ab/libsddp_stamp.so python tools/stamp_timeline.py [--batch B]"""
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
out = (ctypes.c_longlong * 64)()
for rep in range(2):
    s.solve(b["x0"], b["params"], b["X0"], b["U0"], gains=False, history=False)
    torch.cuda.synchronize()
L.sddp_debug_stamps(out)
names = ["node top", "after top sync", "c1 done", "after c1 sync", "w0: d1 done | w1-3: c2 done", "w1-3: c3 done", "w1-3: e done",
         "end of d phase", "after d sync", "f,g done", "after f,g sync", "h done", "fwd: node top (after sync)", "fwd: u^ done", "fwd: after sync", "fwd: integrate / cost done"]
t0 = out[1 * 4 + 0]
print(f"batch={a.batch}: cycles since warp 0 passed the top barrier of node 10 (last iteration of the last problem of CTA 0)")
print(f"{'':30s}" + "".join(f"{'warp ' + str(w):>10s}" for w in range(4)))
for i, n in enumerate(names):
    tb = out[12 * 4 + 0] if i >= 12 else t0      # forward-pass stamps are relative to their own node top
    print(f"{n:30s}" + "".join(f"{out[i * 4 + w] - tb:10d}" for w in range(4)))
