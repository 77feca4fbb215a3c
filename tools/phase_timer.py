#!/usr/bin/env python
"""Cycle budget of one persistent CTA (developer tool).  Build the profiling variant first:
   nvcc ... -DSDDP_PROFILE -o build_ab/libsddp_prof.so srbd_horizon_b200/csrc/sddp.cu
   SDDP_LIB=$PWD/build_ab/libsddp_prof.so python tools/phase_timer.py [--batch B]
Thread 0 of CTA 0 reads clock64() at every phase boundary and adds the time since the previous boundary to that
phase's slot, from kernel entry to exit: the slots are a PARTITION of CTA 0's lifetime, so they sum to its total
cycles, and divided by the Riccati node-iterations CTA 0 processed they give the per-node-iteration budget.
(A clock read right after a block barrier is taken when thread 0 arrives, so waiting at a barrier is charged to the
phase after it; sums over adjacent phases are exact.)"""
import argparse, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srbd_horizon_b200 import _lib
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.problems import make_batch

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=1); ap.add_argument("--N", type=int, default=50)
ap.add_argument("--order", default="schedule")
ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
a = ap.parse_args()
cfg = make_config(MODEL_SRBD, a.N, 0.05, {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3})
b = make_batch(MODEL_SRBD, a.N, a.batch, enumerate_schedules=True)
s = BatchedDDP(cfg, dtype=a.dtype)
L = _lib.lib(a.dtype)
out = (ctypes.c_longlong * 32)()
t = lambda v: torch.as_tensor(v, dtype=s.tdtype, device="cuda")
x0, p, X0, U0 = t(b["x0"]), t(b["params"]), t(b["X0"]), t(b["U0"])
# partition slots, in program order
part = [(7, "queue pop (+ tail wait)"), (4, "init: x0, hist, defects + cost"), (0, "node packs (thread per node)"),
        (19, "bwd: terminal node + prologue"), (8, "bwd: node top"), (9, "bwd: (empty)"), (10, "bwd: c1 Quu, gap shift + sync"),
        (11, "bwd: d1 LDL^T || c2 c3 e + sync"), (12, "bwd: h Wn = Es B + sync"), (13, "bwd: f g syrk, gains + sync"), (1, "bwd: epilogue (dV)"),
        (20, "expand (terminal): zero fill"), (21, "expand (terminal): z-block"), (22, "expand (terminal): affine"),
        (5, "line search: set-up"), (23, "fwd: prologue + prefetch 0"), (24, "fwd: node top wait + sync"), (25, "fwd: u^ = U + a k + K dx + sync"),
        (26, "fwd: integrate / cost"), (27, "fwd: terminal + cost reduce"), (2, "fwd: multi-candidate waves"), (3, "accept: copy X U, scale d"),
        (6, "solve epilogue")]
side = [(14, "d1 alone (warp 0, since c1 end)"), (15, "c2 alone (thread 32)"), (16, "c2+c3 (thread 32)"), (18, "c2+c3+e pass 1 (thread 32)"), (17, "c2+c3+e (thread 32)")]
for rep in range(2):
    L.sddp_debug_profile(out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = s.solve(x0, p, X0, U0, gains=True, history=False, order="schedule" if a.order == "schedule" else None)
    e1.record()
    torch.cuda.synchronize()
    L.sddp_debug_profile(out, 0)
ms = e0.elapsed_time(e1)
probs, iters, waves = out[31], out[30], out[29]
ni = max(iters * a.N, 1)
tot = sum(out[i] for i, _ in part)
slots = min(a.batch, 148 * 4)
mean_it = r.iters.double().mean().item()
print(f"batch={a.batch} N={a.N} kernel(+launch) {ms:.2f} ms; mean iters/solve {mean_it:.3f}; whole-kernel budget = ms x clock x slots / node-iterations "
      f"= {ms * 1e-3 * 1.965e9 * slots / (a.batch * mean_it * a.N):.0f} cycles per node-iteration at 1965 MHz")
print(f"CTA 0: {probs} problems, {iters} DDP iterations ({iters * a.N} node-iterations), {waves} forward waves; lifetime {tot} cycles = {tot / 1.965e6:.2f} ms at 1965 MHz")
print(f"{'phase':42s} {'cycles':>13s} {'share':>7s} {'per node-iteration':>19s}")
for i, n in part:
    print(f"  {n:40s} {out[i]:>13d} {100.0 * out[i] / max(tot, 1):6.1f}% {out[i] / ni:>19.0f}")
print(f"  {'TOTAL (partition)':40s} {tot:>13d} {100.0:6.1f}% {tot / ni:>19.0f}")
print("side measurements (not part of the partition; accumulated since the last boundary):")
for i, n in side:
    print(f"  {n:40s} {out[i]:>13d} {'':7s} {out[i] / ni:>19.0f}")
