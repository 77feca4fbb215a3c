#!/usr/bin/env python
"""bench.py -- SRBD DDP solves/sec at batch 64K on 1/2/4/8 B200 (BASELINE.json metric).

A "step" is one solve of the whole synthetic batch (BASELINE configs[4]: 65,536 SRBD problems,
N = 50, every wpg.py gait schedule, sharded contiguously over the ranks).  One process per GPU
(torchrun for N > 1); NCCL is used only to gather the result trajectories / costs / status.

  python bench.py [--gpus N] [--steps K] [--warmup W]            CUDA path (this repo)
  python bench.py --impl reference [...]                          CPU oracle on the host cores

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from srbd_horizon_b200.config import DIMS, MODEL_SRBD, make_config  # noqa: E402
from srbd_horizon_b200.problems import make_batch  # noqa: E402

METRIC = "srbd_ddp_solves_per_sec"
UNIT = "solves/s"
N_HORIZON = 50
DT = 0.05
EX_OPTS = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}   # dsrbd_example.py:55-58
# SURVEY.md section 8d / BASELINE.md section 4: dense-equivalent algorithmic work per Riccati node
F_NODE = 4 * 37 ** 3 + 8 * 37 ** 2 * 24 + 6 * 37 * 24 ** 2 + 24 ** 3 // 3 + 2 * 37 * 24     # 599,716 FLOP
BYTES_NODE = 15720
# measured DRAM traffic of solve_kernel per Riccati node-iteration: dram__bytes_read+write of one `ncu --set full`
# capture (profiles/r1_solve_kernel_full_raw.csv: 26.55 GB for 4736 problems x 4.963 iterations x 50 nodes)
TRAFFIC_NODE = 26.55e9 / (4736 * 4.963 * 50)
HBM_PEAK_FALLBACK = 6650.0


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": HBM_PEAK_FALLBACK}, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (pynvml, 100 ms period)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_oracle_rate(cfg, batch, seconds: float, threads: int):
    """Times the CPU oracle (oracle/, the CPU restatement of the reference algorithm) on a bounded sample of the
    same workload with `threads` host threads.  Returns (solves/s, problems solved, mean iterations)."""
    from oracle import oracle as O
    n = max(2 * threads, 8)
    sl = lambda k: (batch["x0"][:k], batch["params"][:k], batch["X0"][:k], batch["U0"][:k])
    t0 = time.perf_counter()
    r = O.solve_batch(cfg, *sl(n), nthreads=threads)
    t1 = time.perf_counter() - t0
    target = int(min(len(batch["x0"]), max(n, n * seconds / max(t1, 1e-3))))
    if target > n:
        t0 = time.perf_counter()
        r = O.solve_batch(cfg, *sl(target), nthreads=threads)
        t1 = time.perf_counter() - t0
        n = target
    return n / t1, n, float(r["iters"].mean())


def run_reference(args, rank, world):
    """--impl reference: the reference's DDP core (pyddp + CasADi) is not in /root/reference and cannot run
    here, so this arm times the oracle port (kind "port") on all host threads, each step a bounded sample."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cfg = make_config(MODEL_SRBD, N_HORIZON, DT, EX_OPTS)
    sample = max(2 * threads, min(args.ref_sample, 64 * threads))
    batch = make_batch(MODEL_SRBD, N_HORIZON, sample, enumerate_schedules=True)
    from oracle import oracle as O
    O.lib()
    times, iters = [], []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = O.solve_batch(cfg, batch["x0"], batch["params"], batch["X0"], batch["U0"], nthreads=threads)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt); iters.append(float(r["iters"].mean()))
    t = sum(times) / len(times)
    value = sample / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, sample_note=f"each step = {sample} problems of the same seeded family"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} of 65536 problems per step, {threads} host threads, CPU oracle (C, -O3)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_iters": sum(iters) / len(iters), "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sample_note=None):
    c = {"workload": "BASELINE configs[4]: SRBD DDP batch %d (N=%d, dt=%.2f, nx=37 nu=24 np=19), all 60 wpg gait schedules "
                     "round-robin, seeds default_rng(12345+b), multiple shooting from X=x0 repeated / U=static input" % (args.batch, N_HORIZON, DT),
         "batch": args.batch, "horizon": N_HORIZON, "opts": EX_OPTS, "sharding": "contiguous batch slices per rank, results all-gathered (NCCL)",
         "dispatch": "problems dispatched grouped by contact schedule (device path: hash of the switch pattern of the parameters; host path: the (action, phase) the caller assigned; inside a schedule the largest commanded velocity first), recomputed inside every timed step; results do not depend on it",
         "cache": "inputs larger than L2 (%.1f GB per step), no L2 flush" % (args.batch * (51 * 37 + 50 * 24 + 51 * 19 + 37) * 8 / 1e9)}
    if sample_note:
        c["sample"] = sample_note
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="total problems over all ranks")
    ap.add_argument("--no-gains", action="store_true", help="do not materialise K[B,N,nu,nx] (gains stay in the workspace)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work spent on the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=512)
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from srbd_horizon_b200.ddp import BatchedDDP, DDPSolver, fp64_peak_tflops
    from srbd_horizon_b200.parallel import gather_results, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    nx, nu, np_ = DIMS[MODEL_SRBD]
    lo, hi = shard_range(args.batch, rank, world)
    Bl = hi - lo
    cfg = make_config(MODEL_SRBD, N_HORIZON, DT, EX_OPTS)
    batch = make_batch(MODEL_SRBD, N_HORIZON, Bl, first=lo, enumerate_schedules=True)
    solver = BatchedDDP(cfg, dev)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    x0, params, X0, U0 = t(batch["x0"]), t(batch["params"]), t(batch["X0"]), t(batch["U0"])
    gains = not args.no_gains

    def step():
        r = solver.solve(x0, params, X0, U0, gains=gains, history=False, order="schedule")
        if world > 1:
            gather_results(r, world, args.batch)
        return r

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        r = step()
    sync_all()
    launches0 = solver.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        sync_all()
        t_wall0 = time.perf_counter()
        for s in range(args.steps):
            ev[s][0].record()
            Xc, Uc = X0.clone(), U0.clone()
            kev[s][0].record()
            r = solver.solve(x0, params, Xc, Uc, gains=gains, history=False, inplace=True, order="schedule")
            kev[s][1].record()
            if world > 1:      # (solving in two pieces to overlap the gather with the second solve was measured slower:
                gathered = gather_results(r, world, args.batch)      #  parallel.solve_sharded, 99.5 vs 92.1 ms at 8 GPUs)
            ev[s][1].record()
        sync_all()
        t_wall = time.perf_counter() - t_wall0
    launches = solver.launches - launches0
    ms_dev = sum(a.elapsed_time(b) for a, b in ev)           # device time of the K steps on this rank
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    tt = torch.tensor([ms_dev, ms_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_dev, ms_kernel = float(tt[0]), float(tt[1])
    if world > 1:      # every rank holds the whole gathered batch
        agg = torch.stack([gathered["iters"].sum().to(torch.float64), (gathered["status"] == 0).sum().to(torch.float64)])
    else:
        agg = torch.stack([r.iters.sum().to(torch.float64), (r.status == 0).sum().to(torch.float64)])
    mean_iters = float(agg[0]) / args.batch
    conv_frac = float(agg[1]) / args.batch
    ms_per_step = ms_dev / args.steps
    value = args.batch / (ms_per_step * 1e-3)

    # ---- end to end through the public host API (pinned host buffers, copies inside the timed region)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hx0, hp, hX, hU = pin(batch["x0"]), pin(batch["params"]), pin(batch["X0"]), pin(batch["U0"])
    e2e_steps = max(1, min(args.steps, 3))
    pin_out = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory().numpy()
    hout = {"X": pin_out((Bl, N_HORIZON + 1, nx)), "U": pin_out((Bl, N_HORIZON, nu)), "cost": pin_out((Bl,)),
            "iters": pin_out((Bl,), torch.int32), "status": pin_out((Bl,), torch.int32)}
    # On the host the caller groups the problems by the gait schedule it assigned them (action, phase): cheaper than
    # hashing 500 MB of parameters; recomputed inside every timed step.
    sched_order = lambda: solver.order_from_keys(batch["actions"] * 20 + batch["s0"], (hp.numpy()[:, -1, 0:3] ** 2).sum(axis=1))
    solver.solve_host(hx0.numpy(), hp.numpy(), hX.numpy(), hU.numpy(), out=hout, order=sched_order())     # warm-up (allocates the staging buffer)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        rh = solver.solve_host(hx0.numpy(), hp.numpy(), hX.numpy(), hU.numpy(), out=hout, order=sched_order())
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = args.batch / float(t_e2e[0])
    h2d = (hx0.numel() + hp.numel() + hX.numel() + hU.numel()) * 8
    d2h = (hX.numel() + hU.numel() + Bl) * 8 + 2 * Bl * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (solve_kernel<Srbd>, one launch per step)
    peaks, peak_src = measured_peaks()
    p64 = fp64_peak_tflops()
    units = (args.batch / world) * mean_iters * N_HORIZON            # Riccati node-iterations per launch (per rank)
    t_k = ms_kernel * 1e-3
    ach_tf = units * F_NODE / t_k / 1e12
    ach_gb = units * BYTES_NODE / t_k / 1e9
    roofline = {"bound": "fp64", "achieved": ach_tf, "peak": p64, "unit": "TFLOP/s", "frac": ach_tf / p64 if p64 else None,
                "traffic": units * TRAFFIC_NODE,
                "traffic_note": "bytes per launch = node-iterations x %.0f B measured with ncu --set full at B=4736 (profiles/README.md)" % TRAFFIC_NODE,
                "note": "dense-equivalent algorithmic FLOP (599,716 per Riccati node-iteration, SURVEY 8d) / CUDA-event duration of "
                        "solve_kernel; peak = FP64 FMA rate measured in this run (sddp_fp64_peak_tflops)",
                "kernel": "solve_kernel<Srbd>", "kernel_ms": ms_kernel, "node_iterations_per_launch": units,
                "hbm": {"achieved": ach_gb, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": ach_gb / peaks.get("hbm_gbs", HBM_PEAK_FALLBACK),
                        "peak_source": peak_src, "bytes_per_node_iteration": BYTES_NODE}}

    threads = os.cpu_count() or 1
    cpu_rate, cpu_n, cpu_iters = cpu_oracle_rate(cfg, batch, args.cpu_seconds, threads)
    cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": "port", "mean_iters": cpu_iters,
                    "sample": f"first {cpu_n} of {args.batch} problems, {threads} host threads, CPU oracle (oracle/sddp_oracle.c, gcc -O3)"}

    latency = None
    if not args.no_latency:
        latency = single_solve_latency(dev)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "BatchedDDP.solve_host -> sddp_solve_batch_host (pinned host buffers in and out, 16K-problem chunks: copies overlap solves; X, U, cost, iters, status read back)"},
        "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "mean_iters": mean_iters, "converged_frac": conv_frac, "ddp_iterations_per_sec": value * mean_iters,
        "gains_materialised": gains, "wall_s_timed_region": t_wall, "latency": latency,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def single_solve_latency(dev):
    """p50 single-solve latency, BASELINE configs[1]: one SRBD problem, N=20 (dsrbd_example.py), closed-loop MPC ticks;
    GPU through the reference-facing DDPSolver.solve() (host buffers in and out) next to the CPU oracle."""
    import torch
    from oracle import oracle as O
    from srbd_horizon_b200 import prb as P, wpg
    from srbd_horizon_b200.ddp import DDPSolver
    from srbd_horizon_b200.mpc import mpc_tick_references, plant_step

    ns = 20
    srbd = P.SRBDProblem(); srbd.createSRBDProblem(ns, 1.0)
    solver = DDPSolver(srbd.prb, dict(EX_OPTS), device=dev)
    gen = wpg.steps_phase(srbd.f, srbd.c, srbd.cdot, float(srbd.initial_foot_position[0][2]), srbd.c_ref, srbd.w_ref,
                          srbd.orientation_tracking_gain, srbd.cdot_switch, ns, number_of_legs=2, contact_model=2)
    state = srbd.getInitialState()
    solver.set_u_warmstart(np.tile(srbd.getStaticInput()[:, None], (1, ns)))
    gpu_ms, cpu_ms = [], []
    cfg = solver.cfg
    Xc = np.tile(state, (ns + 1, 1)); Uc = np.tile(srbd.getStaticInput(), (ns, 1))
    for tick in range(140):
        solver.setInitialState(state)
        mpc_tick_references(srbd, [0.5, 0.0, 0.0] if tick >= 10 else [0.0, 0.0, 0.0])
        gen.set("step" if tick >= 10 else "standing")
        params = solver.get_params_value()
        t0 = time.perf_counter()
        solver.solve()
        gpu_ms.append(1e3 * (time.perf_counter() - t0))
        t0 = time.perf_counter()
        ro = O.solve_batch(cfg, state[None], params[None], Xc[None], Uc[None], nthreads=1)
        cpu_ms.append(1e3 * (time.perf_counter() - t0))
        Xc, Uc = ro["X"][0], ro["U"][0]
        state = plant_step(solver.ddp_solver, state, solver.getSolutionDict()["u_opt"][:, 0])
    g, c = sorted(gpu_ms[40:]), sorted(cpu_ms[40:])
    q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]
    return {"workload": "BASELINE configs[1]: single SRBD problem, N=20, 100 closed-loop MPC ticks (walking) after 40 warm-up ticks",
            "gpu_p50_ms": q(g, 0.5), "gpu_p99_ms": q(g, 0.99), "cpu_oracle_p50_ms": q(c, 0.5), "cpu_oracle_p99_ms": q(c, 0.99),
            "gpu_api": "DDPSolver.solve() (host numpy in/out, H2D + kernel + D2H)", "cpu_threads": 1}


if __name__ == "__main__":
    main()
