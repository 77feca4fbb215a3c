#!/usr/bin/env python
"""bench.py -- SRBD DDP solves/sec at batch 64K on 1/2/4/8 B200 (BASELINE.json metric).

A "step" is one solve of the whole synthetic batch (BASELINE configs[4]: 65,536 SRBD problems,
N = 50, every wpg.py gait schedule, sharded contiguously over the ranks).  One process per GPU
(torchrun for N > 1); NCCL is used only to gather the result trajectories / costs / status.

  python bench.py [--gpus N] [--steps K] [--warmup W]            CUDA path (this repo), BASELINE configs[4]
  python bench.py --config {0,1,2,3,4} [...]                      the other BASELINE configs (0/1: single-problem closed loop)
  python bench.py --impl reference [...]                          CPU oracle on the host cores
  python bench.py --dtype f32 [...]                               the optional fp32 build (NARROWER than the reference's fp64;
                                                                  stated tolerance in tests/test_gpu_f32.py) -- not the headline

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
# capture (profiles/r2_solve_kernel_full_raw.csv: 11.25 + 13.46 GB for 4736 problems x 4.963 iterations x 50 nodes)
TRAFFIC_NODE = 24.71e9 / (4736 * 4.963 * 50)
HBM_PEAK_FALLBACK = 6650.0
FP64_NOMINAL_TFLOPS = 148 * 4 * 16 * 2 * 1.965e9 / 1e12      # 148 SMs x 64 FP64 FMA/clk x 2 x 1965 MHz = 37.2 (no FP64 entry in MEASURED_PEAKS.json)

# BASELINE.json configs -> workload.  Batched configs share one code path; 0 / 1 are the reference's own single-problem
# closed loops (dlip_example.py / dsrbd_example.py), timed as latency and reported as solves/s of one problem.
CONFIGS = {
    0: {"name": "configs[0] dlip_example.py: single discrete-LIP DDP walking problem (N=20, closed loop)", "single": "lip"},
    1: {"name": "configs[1] dsrbd_example.py: single SRBD kangaroo line-feet DDP solve (N=20, closed loop)", "single": "srbd"},
    2: {"name": "configs[2] batched SRBD DDP, 4096 random initial states / contact schedules, N=50", "batch": 4096, "opts": {}, "x_noise": 0.0, "enumerate": False},
    3: {"name": "configs[3] multiple-shooting DDP with fixed defect contraction rate 0.5, SRBD, batch 16K, N=50, warm start perturbed (x_noise 0.01)",
        "batch": 16384, "opts": {"defect_contraction_rate": 0.5}, "x_noise": 0.01, "enumerate": False},
    4: {"name": "configs[4] SRBD DDP batch 64K sweep over wpg.py gait schedules", "batch": 65536, "opts": {}, "x_noise": 0.0, "enumerate": True},
}


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
    same workload with `threads` host threads.  Returns (solves/s, problems solved, mean iterations, the oracle's results)."""
    from oracle import oracle as O
    O.native_twin()
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
    return n / t1, n, float(r["iters"].mean()), r


def oracle_flags():
    from oracle import oracle as O
    return O.BUILD_FLAGS


def run_reference(args, rank, world):
    """--impl reference: the reference's DDP core (pyddp + CasADi) is not in /root/reference and cannot run
    here, so this arm times the oracle port (kind "port") on all host threads, each step a bounded sample."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    spec = CONFIGS[args.config]
    if "single" in spec:
        print(json.dumps(single_problem_line(args, None, reference_only=True)), flush=True)
        return
    cfg = make_config(MODEL_SRBD, N_HORIZON, DT, dict(EX_OPTS, **spec["opts"]))
    sample = max(2 * threads, min(args.ref_sample, 64 * threads))
    batch = make_batch(MODEL_SRBD, N_HORIZON, sample, enumerate_schedules=spec["enumerate"], x_noise=spec["x_noise"])
    from oracle import oracle as O
    O.native_twin()
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
                         "sample": f"{sample} of {args.batch} problems per step, {threads} host threads, CPU oracle (C, {O.BUILD_FLAGS})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_iters": sum(iters) / len(iters), "ddp_iterations_per_sec": value * sum(iters) / len(iters), "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sample_note=None):
    spec = CONFIGS[args.config]
    opts = dict(EX_OPTS, **spec.get("opts", {}))
    c = {"workload": "BASELINE %s: SRBD DDP batch %d (N=%d, dt=%.2f, nx=37 nu=24 np=19), %s, seeds default_rng(12345+b), multiple shooting from "
                     "X=x0 repeated / U=static input" % (spec["name"], args.batch, N_HORIZON, DT,
                                                          "all 60 wpg gait schedules round-robin" if spec["enumerate"] else "gait schedules drawn at random (70/20/10 % step/standing/jump)"),
         "batch": args.batch, "horizon": N_HORIZON, "opts": opts, "sharding": "contiguous batch slices per rank, results gathered on every rank (%s)" % args.gather,
         "dispatch": "problems dispatched grouped by contact schedule (device path: hash of the switch pattern of the parameters; host path: the (action, phase) the caller assigned; inside a schedule the largest commanded velocity first), recomputed inside every timed step; results do not depend on it",
         "cache": "inputs larger than L2 (%.1f GB per step), no L2 flush" % (args.batch * (51 * 37 + 50 * 24 + 51 * 19 + 37) * (4 if getattr(args, "dtype", "f64") == "f32" else 8) / 1e9),
         "repeat": "every step re-solves the same batch from the same warm start (X, U cloned inside the timed region)",
         "gains": "K[B,N,24,37] %s is written to HBM every step into the buffer of the previous step (%.1f GB for the whole batch)" % (getattr(args, "dtype", "f64"), args.batch * N_HORIZON * 24 * 37 * (4 if getattr(args, "dtype", "f64") == "f32" else 8) / 1e9) if not args.no_gains
                  else "gains stay in the per-CTA workspace (--no-gains)"}
    if sample_note:
        c["sample"] = sample_note
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS), help="BASELINE.json configs index (default 4: the headline metric)")
    ap.add_argument("--batch", type=int, default=None, help="total problems over all ranks (default: the config's)")
    ap.add_argument("--no-gains", action="store_true", help="do not materialise K[B,N,nu,nx] (gains stay in the workspace)")
    ap.add_argument("--gather", default="auto", choices=["auto", "push", "nccl"],
                    help="multi-GPU result gather: push = every CTA stores its finished problem's result record into every peer's slab over "
                         "NVLink from inside solve_kernel (auto: used when peer access works), nccl = one all-gather of the packed slab after the kernel")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work spent on the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=512)
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="f32: the optional fp32 build libsddp_f32.so (batched configs only); narrower than the reference, labelled so in the line")
    args = ap.parse_args()
    if args.dtype == "f32" and ("single" in CONFIGS[args.config] or args.impl == "reference"):
        raise SystemExit("--dtype f32 applies to the batched CUDA configs (2, 3, 4)")
    spec = CONFIGS[args.config]
    if args.batch is None:
        args.batch = spec.get("batch", 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from srbd_horizon_b200.ddp import BatchedDDP, fp64_peak_tflops
    from srbd_horizon_b200.parallel import ResultGather, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    if "single" in spec:      # configs[0] / [1]: one problem, one GPU (replicas only: nothing to shard)
        if rank == 0:
            print(json.dumps(single_problem_line(args, dev)), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    nx, nu, np_ = DIMS[MODEL_SRBD]
    lo, hi = shard_range(args.batch, rank, world)
    Bl = hi - lo
    cfg = make_config(MODEL_SRBD, N_HORIZON, DT, dict(EX_OPTS, **spec["opts"]))
    batch = make_batch(MODEL_SRBD, N_HORIZON, Bl, first=lo, enumerate_schedules=spec["enumerate"], x_noise=spec["x_noise"])
    solver = BatchedDDP(cfg, dev, dtype=args.dtype)
    esz = 8 if args.dtype == "f64" else 4
    t = lambda a: torch.as_tensor(a, dtype=solver.tdtype, device=dev)
    x0, params, X0, U0 = t(batch["x0"]), t(batch["params"]), t(batch["X0"]), t(batch["U0"])
    gains = not args.no_gains
    gather = ResultGather(solver, args.batch, rank, world, mode=args.gather) if world > 1 else None
    args.gather = gather.mode if gather else "single GPU"

    res = [None]      # K, k, iters, status, cost of the previous step are overwritten (no allocation of K inside a step)

    def step():
        Xc, Uc = X0.clone(), U0.clone()
        res[0] = solver.solve(x0, params, Xc, Uc, gains=gains, history=False, inplace=True, order="schedule", gather=gather, out=res[0])
        return res[0]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    t_w0 = time.perf_counter()
    for _ in range(args.warmup):
        r = step()
        if gather:
            gather.finish()
    sync_all()
    # Clock ramp: at 4 - 8 GPUs a step is 60 - 120 ms, so W = 3 warm-up steps are a fraction of a second and a GPU that
    # boosts late shows up in the max-over-ranks time of the K timed steps (seen at 4 GPUs: 97 / 113 / 140 ms for the same
    # step).  More untimed steps, the same number on every rank, until about 1.5 s of warm-up have run.
    t_w = torch.tensor([time.perf_counter() - t_w0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_w, op=dist.ReduceOp.MAX)
    t_w = float(t_w[0])
    extra_warmup = int(min(30, max(0, np.ceil((1.5 - t_w) / max(t_w / max(args.warmup, 1), 1e-3))))) if args.warmup > 0 else 0
    for _ in range(extra_warmup):
        r = step()
        if gather:
            gather.finish()
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
            r = res[0] = solver.solve(x0, params, Xc, Uc, gains=gains, history=False, inplace=True, order="schedule", gather=gather, out=res[0])
            kev[s][1].record()
            if gather:      # push: a 4-byte all-reduce orders the peers' stores; nccl: one all-gather of the packed slab
                gathered = gather.finish()
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
        # the gathered slab must hold this rank's own results bit for bit (and, through the all-reduce, everyone's)
        own_ok = bool(torch.equal(gathered["X"][lo:hi], r.X) and torch.equal(gathered["U"][lo:hi], r.U) and torch.equal(gathered["cost"][lo:hi], r.cost))
        chk = torch.stack([gathered["cost"].sum(), gathered["X"].sum()])
        chk_max, chk_min = chk.clone(), chk.clone()
        dist.all_reduce(chk_max, op=dist.ReduceOp.MAX); dist.all_reduce(chk_min, op=dist.ReduceOp.MIN)
        gather_ok = own_ok and bool(torch.equal(chk_max, chk_min))
    else:
        agg = torch.stack([r.iters.sum().to(torch.float64), (r.status == 0).sum().to(torch.float64)])
        gather_ok = None
    mean_iters = float(agg[0]) / args.batch
    conv_frac = float(agg[1]) / args.batch
    ms_per_step = ms_dev / args.steps
    value = args.batch / (ms_per_step * 1e-3)

    # ---- end to end through the public host API (pinned host buffers, copies inside the timed region)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=solver.ndtype)).pin_memory()
    hx0, hp, hX, hU = pin(batch["x0"]), pin(batch["params"]), pin(batch["X0"]), pin(batch["U0"])
    e2e_steps = max(1, min(args.steps, 3))
    pin_out = lambda shape, dt=solver.tdtype: torch.empty(shape, dtype=dt).pin_memory().numpy()
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
    h2d = (hx0.numel() + hp.numel() + hX.numel() + hU.numel()) * esz
    d2h = (hX.numel() + hU.numel() + Bl) * esz + 2 * Bl * 4

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
                "fp64_peak_source": "in-run microbenchmark (sddp_fp64_peak_tflops: 8 independent DFMA chains per thread); MEASURED_PEAKS.json has no FP64 entry",
                "fp64_peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": ach_tf / FP64_NOMINAL_TFLOPS,
                "traffic": units * TRAFFIC_NODE,
                "traffic_note": "bytes per launch = node-iterations x %.0f B measured with ncu --set full at B=4736 (profiles/README.md)" % TRAFFIC_NODE,
                "note": "dense-equivalent algorithmic FLOP (599,716 per Riccati node-iteration, SURVEY 8d) / CUDA-event duration of "
                        "solve_kernel; the structured kernel executes about a third of them (profiles/README.md), so this is a "
                        "work-rate figure, not pipe utilisation",
                "kernel": "solve_kernel<Srbd>", "kernel_ms": ms_kernel, "node_iterations_per_launch": units,
                "hbm": {"achieved": ach_gb, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": ach_gb / peaks.get("hbm_gbs", HBM_PEAK_FALLBACK),
                        "peak_source": peak_src, "bytes_per_node_iteration": BYTES_NODE}}

    threads = os.cpu_count() or 1
    cpu_rate, cpu_n, cpu_iters, ro = cpu_oracle_rate(cfg, batch, args.cpu_seconds, threads)
    cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": "port", "mean_iters": cpu_iters,
                    "sample": f"first {cpu_n} of {args.batch} problems, {threads} host threads, CPU oracle (oracle/sddp_oracle.c, gcc {oracle_flags()})"}
    # parity of the timed GPU results against the oracle on that sample (outside every timed region)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    gX, gU, gc = r.X[:cpu_n].cpu().numpy(), r.U[:cpu_n].cpu().numpy(), r.cost[:cpu_n].cpu().numpy()
    per = np.array([max(rel(gX[i], ro["X"][i]), rel(gU[i], ro["U"][i])) for i in range(cpu_n)])
    parity = {"problems": cpu_n, "parity_max_rel_err": float(per.max()), "parity_median_rel_err": float(np.median(per)),
              "frac_above_1e-9": float((per > 1e-9).mean()),
              "cost_max_rel_err": float(np.max(np.abs(gc - ro["cost"]) / np.abs(ro["cost"]))),
              "iters_equal": bool((r.iters[:cpu_n].cpu().numpy() == ro["iters"]).all()), "status_equal": bool((r.status[:cpu_n].cpu().numpy() == ro["status"]).all()),
              "host_path_equals_device_path": bool(np.array_equal(rh["X"], r.X.cpu().numpy()) and np.array_equal(rh["U"], r.U.cpu().numpy())),
              "note": "max over the sample of max|X - X_oracle| / max|X_oracle| and the same for U, per problem"}

    if args.dtype == "f32":
        # the optional fp32 build: same dense-equivalent work against the nominal FP32 FMA rate (no measured FP32 peak);
        # its results are NOT at the reference's precision -- `parity` above shows the distance to the fp64 oracle
        f32_nominal = 2.0 * FP64_NOMINAL_TFLOPS
        roofline.update({"bound": "fp32", "peak": f32_nominal, "frac": ach_tf / f32_nominal, "fp32_peak_source": "nominal: 148 SMs x 128 FP32 FMA/clk x 2 x 1965 MHz",
                         "kernel": "solve_kernel<Srbd> of libsddp_f32.so", "traffic": None,
                         "hbm": dict(roofline["hbm"], achieved=ach_gb / 2, frac=ach_gb / 2 / peaks.get("hbm_gbs", HBM_PEAK_FALLBACK), bytes_per_node_iteration=BYTES_NODE // 2)})
        parity["note"] += "; fp32 build: the stated tolerance is 1e-2 on X and U (tests/test_gpu_f32.py; measured worst 2e-3, median 1e-6), not the 1e-9 of the fp64 build"
    latency = None
    if not args.no_latency and args.dtype == "f64":
        latency = single_solve_latency(dev)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": workload_config(args),
        "ddp_iterations_per_sec": value * mean_iters, "mean_iters": mean_iters, "converged_frac": conv_frac,
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "BatchedDDP.solve_host -> sddp_solve_batch_host (pinned host buffers in and out, at least four chunks per call: copies overlap solves; X, U, cost, iters, status read back)"},
        "gpu_launches": launches,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity, "parity_max_rel_err": parity["parity_max_rel_err"],
        "gather_ms_per_step": ms_per_step - ms_kernel if world > 1 else None, "gather_ok": gather_ok,
        "gains_materialised": gains, "wall_s_timed_region": t_wall, "latency": latency, "extra_untimed_warmup_steps": extra_warmup,
    }
    if args.dtype == "f32":
        line["precision_note"] = ("optional fp32 build (north_star): float storage and Riccati recursion, NARROWER than the reference's fp64; "
                                  "not comparable with the fp64 headline line")
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def closed_loop(kind, dev, ticks, skip, with_cpu=True):
    """The reference's own use (dsrbd_example.py / dlip_example.py): one problem, closed-loop MPC ticks through the
    reference-facing DDPSolver.solve() (host numpy in and out: H2D + kernel + D2H), next to the single-thread oracle.
    Returns per-tick GPU ms, CPU ms and iteration counts after `skip` warm-up ticks."""
    from oracle import oracle as O
    O.native_twin()
    from srbd_horizon_b200 import prb as P, wpg
    from srbd_horizon_b200.ddp import DDPSolver
    from srbd_horizon_b200.mpc import mpc_tick_references, plant_step

    ns = 20
    if kind == "srbd":
        pr = P.SRBDProblem(); pr.createSRBDProblem(ns, 1.0)
    else:
        pr = P.LIPProblem(); pr.createLIPProblem(ns, 1.0)
    gpu_ms, cpu_ms, iters = [], [], []
    solver = DDPSolver(pr.prb, dict(EX_OPTS), device=dev) if dev is not None else None
    cfgm = solver.cfg if solver else make_config(pr.prb.model, ns, pr.prb.getDt(), EX_OPTS, pr.prb.robot, pr.prb.gains)
    aux = pr
    if kind != "srbd":      # dlip_example.py:33-36,85 hands the SRBD problem's w_ref / gain parameters to the gait scheduler
        aux = P.SRBDProblem(); aux.createSRBDProblem(ns, 1.0)
    gen = wpg.steps_phase(getattr(pr, "f", None), pr.c, pr.cdot, float(pr.initial_foot_position[0][2]), pr.c_ref, aux.w_ref,
                          aux.orientation_tracking_gain, pr.cdot_switch, ns, number_of_legs=2, contact_model=2)
    state = pr.getInitialState()
    ustat = pr.getStaticInput()
    if solver:
        solver.set_u_warmstart(np.tile(ustat[:, None], (1, ns)))
    Xc = np.tile(state, (ns + 1, 1)); Uc = np.tile(ustat, (ns, 1))
    for tick in range(ticks):
        if solver:
            solver.setInitialState(state)
        mpc_tick_references(pr, [0.5, 0.0, 0.0] if tick >= 10 else [0.0, 0.0, 0.0])
        gen.set("step" if tick >= 10 else "standing")
        params = pr.prb.flat_parameters()
        if solver:
            t0 = time.perf_counter()
            solver.solve()
            gpu_ms.append(1e3 * (time.perf_counter() - t0))
        if with_cpu or not solver:
            t0 = time.perf_counter()
            ro = O.solve_batch(cfgm, state[None], params[None], Xc[None], Uc[None], nthreads=1)
            cpu_ms.append(1e3 * (time.perf_counter() - t0))
            Xc, Uc = ro["X"][0], ro["U"][0]
            iters.append(int(ro["iters"][0]))
        if solver:
            state = plant_step(solver.ddp_solver, state, solver.getSolutionDict()["u_opt"][:, 0])
        else:      # CPU arm alone: the oracle's Euler step, SRBD quaternion renormalised (dsrbd_example.py:158-160)
            state = O.dynamics(cfgm, state, Uc[0], 0)
            if kind == "srbd":
                state[3:7] /= np.linalg.norm(state[3:7])
    return gpu_ms[skip:], cpu_ms[skip:], iters[skip:]


def single_solve_latency(dev):
    """p50 single-solve latency, BASELINE configs[1]: one SRBD problem, N=20 (dsrbd_example.py), closed-loop MPC ticks;
    GPU through the reference-facing DDPSolver.solve() (host buffers in and out) next to the CPU oracle."""
    gpu_ms, cpu_ms, _ = closed_loop("srbd", dev, 140, 40)
    g, c = sorted(gpu_ms), sorted(cpu_ms)
    q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]
    return {"workload": "BASELINE configs[1]: single SRBD problem, N=20, 100 closed-loop MPC ticks (walking) after 40 warm-up ticks",
            "gpu_p50_ms": q(g, 0.5), "gpu_p99_ms": q(g, 0.99), "cpu_oracle_p50_ms": q(c, 0.5), "cpu_oracle_p99_ms": q(c, 0.99),
            "gpu_api": "DDPSolver.solve() (host numpy in/out, H2D + kernel + D2H)", "cpu_threads": 1}


def single_problem_line(args, dev, reference_only=False):
    """BASELINE configs[0] / [1]: the JSON line of a single-problem closed loop.  A "step" is 20 MPC ticks; value = ticks
    solved per second of solver time (1 / mean latency)."""
    spec = CONFIGS[args.config]
    ticks = 20 * (args.warmup + args.steps)
    gpu_ms, cpu_ms, iters = closed_loop(spec["single"], None if reference_only else dev, ticks + 10, 10 + 20 * args.warmup)
    q = lambda v, f: sorted(v)[min(len(v) - 1, int(f * len(v)))]
    cpu_rate = 1e3 / (sum(cpu_ms) / len(cpu_ms))
    cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"the same {len(cpu_ms)} closed-loop ticks, CPU oracle, one thread",
                    "p50_ms": q(cpu_ms, 0.5)}
    cfgd = {"workload": "BASELINE " + spec["name"], "batch": 1, "horizon": 20, "opts": EX_OPTS,
            "cache": "closed loop: every tick is a new problem (state, shifted schedule); working set far below L2, stated",
            "step": "20 MPC ticks"}
    metric = METRIC if spec["single"] == "srbd" else "lip_ddp_solves_per_sec"
    if reference_only:
        return {"impl": "reference", "metric": metric, "value": cpu_rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 20 * 1e3 / cpu_rate, "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfgd, "cpu_baseline": cpu_baseline, "e2e": {"value": cpu_rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "mean_iters": sum(iters) / len(iters), "gpu_launches": 0}
    rate = 1e3 / (sum(gpu_ms) / len(gpu_ms))
    nx, nu, np_ = (37, 24, 19) if spec["single"] == "srbd" else (30, 15, 11)
    h2d = (nx + 21 * np_ + 21 * nx + 20 * nu) * 8
    d2h = (21 * nx + 20 * nu + 1) * 8 + 8
    return {"metric": metric, "value": rate, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 20 * 1e3 / rate,
            "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfgd,
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 20 * h2d, "d2h_bytes_per_step": 20 * d2h,
                    "api": "DDPSolver.solve() (host numpy in and out): value and e2e are the same measurement for the single-problem configs"},
            "gpu_launches": len(gpu_ms), "latency": {"gpu_p50_ms": q(gpu_ms, 0.5), "gpu_p99_ms": q(gpu_ms, 0.99), "cpu_oracle_p50_ms": q(cpu_ms, 0.5)},
            "roofline": {"bound": "latency", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None, "traffic": None,
                         "note": "one problem occupies one of 148 SMs; the time is a dependent chain (launch + copies + 20 nodes x iterations), no throughput roofline applies"},
            "cpu_baseline": cpu_baseline, "mean_iters": sum(iters) / len(iters)}


if __name__ == "__main__":
    main()
