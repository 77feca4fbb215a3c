"""Multi-GPU sharding of a problem batch.

The path shards naturally: every MPC problem is independent (the reference runs one problem per
process, dsrbd_example.py:82-185).  One process per GPU solves a contiguous slice of the batch.  The
results (trajectories, cost, iterations, status: one packed record per problem, include/sddp.h) reach every
GPU in one of two ways (`ResultGather`):

  push   the solve kernel itself stores every finished problem's record into the whole-batch slab of
         every GPU over NVLink (peer memory mapped through CUDA IPC) while the rest of the batch is
         still being solved; the only collective is a 4-byte all-reduce that orders "all kernels done";
  nccl   the kernel fills the own slab only and ONE in-place all-gather of the packed slab follows.

Gains are not gathered by default (MPC consumes u_0 / K_0 only; K is 7.1 KB per node).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of B problems owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return (B * rank) // world, (B * (rank + 1)) // world


def _gather(t: torch.Tensor, world: int, B: int, sizes=None, async_op: bool = False):
    """all-gather along dim 0 of per-rank shards made by `shard_range` (pads uneven shards).
    With async_op the collective is only enqueued: returns (padded output, work handle, sizes)."""
    if sizes is None:
        sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    m = max(sizes)
    if t.shape[0] != m:
        pad = torch.zeros((m - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], dim=0)
    out = torch.empty((world * m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if async_op:
        return out, dist.all_gather_into_tensor(out, t.contiguous(), async_op=True), sizes
    dist.all_gather_into_tensor(out, t.contiguous())
    if all(s == m for s in sizes):
        return out
    return torch.cat([out[r * m:r * m + sizes[r]] for r in range(world)], dim=0)


def gather_results(r, world: int, B: Optional[int] = None, gains: str = "none") -> Dict[str, torch.Tensor]:
    """Gather a `BatchResult` shard from every rank into whole-batch tensors on every rank.
    gains: "none" | "first" (K_0, k_0 only) | "all"."""
    if B is None:
        t = torch.tensor([r.X.shape[0]], device=r.X.device)
        dist.all_reduce(t)
        B = int(t.item())
    out = {name: _gather(getattr(r, name), world, B) for name in ("X", "U", "cost", "iters", "status")}
    if gains != "none" and r.K is not None:
        if gains == "first":
            out["K"], out["k"] = _gather(r.K[:, 0].contiguous(), world, B), _gather(r.k[:, 0].contiguous(), world, B)
        else:
            out["K"], out["k"] = _gather(r.K, world, B), _gather(r.k, world, B)
    return out


_FIELDS = ("X", "U", "cost", "iters", "status")


def solve_sharded(solve, x0, params, X0, U0, world: int, B: int, rank: int, chunks: int = 2) -> Dict[str, torch.Tensor]:
    """Solve this rank's shard of a batch of B problems in `chunks` pieces and all-gather the results of piece i
    while piece i+1 is being solved (the collectives run on NCCL's stream; there is no other communication on the
    path).  `solve(x0, params, X0, U0)` returns an object with the fields X, U, cost, iters, status for the problems
    it is given.  Returns whole-batch tensors in the original problem order on every rank."""
    n = x0.shape[0]
    if n != shard_range(B, rank, world)[1] - shard_range(B, rank, world)[0]:
        raise ValueError("shard size does not match shard_range(B, rank, world)")
    shard = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    chunks = max(1, min([chunks] + [s for s in shard if s > 0]))
    pending, last = [], None
    for c in range(chunks):
        lo, hi = shard_range(n, c, chunks)
        last = solve(x0[lo:hi], params[lo:hi], X0[lo:hi], U0[lo:hi])
        sizes = [shard_range(s, c, chunks)[1] - shard_range(s, c, chunks)[0] for s in shard]
        pending.append({f: _gather(getattr(last, f), world, B, sizes=sizes, async_op=True) for f in _FIELDS})
    out = {}
    for f in _FIELDS:
        parts = [[] for _ in range(world)]
        for g in pending:
            buf, work, sizes = g[f]
            work.wait()
            m = max(sizes)
            for r in range(world):
                parts[r].append(buf[r * m:r * m + sizes[r]])
        out[f] = torch.cat([p for r in range(world) for p in parts[r]], dim=0)
    return out


class _DevArray:
    """__cuda_array_interface__ view of library-owned device memory, so that torch can alias it without a copy."""

    def __init__(self, ptr: int, shape, owner, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner


def record_layout(N: int, nx: int, nu: int) -> Dict[str, Tuple[int, int]]:
    """Offsets (in doubles) of the fields of a result record (include/sddp.h): X | U | cost | iters | status."""
    xs, us = (N + 1) * nx, N * nu
    return {"X": (0, xs), "U": (xs, us), "cost": (xs + us, 1), "iters": (xs + us + 1, 1), "status": (xs + us + 2, 1), "size": (0, xs + us + 3)}


def slab_views(slab: torch.Tensor, N: int, nx: int, nu: int) -> Dict[str, torch.Tensor]:
    """Whole-batch X[B,N+1,nx], U[B,N,nu], cost[B] as strided views of a slab[B, record]; iters / status converted to int32."""
    lay = record_layout(N, nx, nu)
    B = slab.shape[0]
    f = lambda name: slab[:, lay[name][0]:lay[name][0] + lay[name][1]]
    return {"X": f("X").unflatten(1, (N + 1, nx)), "U": f("U").unflatten(1, (N, nu)), "cost": f("cost").reshape(B),
            "iters": f("iters").reshape(B).to(torch.int32), "status": f("status").reshape(B).to(torch.int32)}


class ResultGather:
    """Whole-batch result slab of one rank and the way the other ranks' records get into it.

    g = ResultGather(solver, B, rank, world)         # allocates the slab, maps the peers' slabs (mode "push")
    r = solver.solve(..., gather=g)                  # the kernel stores the records (own shard: problems lo..hi-1)
    out = g.finish()                                 # orders the ranks; dict of whole-batch views X, U, cost, iters, status
    The views stay valid until the next `solve(gather=g)` on any rank (which first waits for every rank to have
    reached it: consumers are stream-ordered before it).
    """

    def __init__(self, solver, B: int, rank: int, world: int, mode: str = "auto", group=None):
        if world > 8:
            raise ValueError("result peers are the GPUs of one NVLink box (<= 8)")
        self.solver, self.B, self.rank, self.world, self.group = solver, int(B), rank, world, group
        self.lo, self.hi = shard_range(B, rank, world)
        self.L = solver.L
        self.rec = int(self.L.sddp_record_doubles(solver.h))
        ptr = ctypes.c_void_p()
        with torch.cuda.device(solver.device):
            _lib.check(self.L.sddp_slab_alloc(solver.h, self.B, ctypes.byref(ptr)), solver.h, self.L)
        self.ptr = ptr.value
        self.slab = torch.as_tensor(_DevArray(self.ptr, (self.B, self.rec), self, "<f8" if solver.dtype == "f64" else "<f4"), device=solver.device)
        self.peer_ptrs = []
        # the ordering collective runs on the stream with NCCL; with a host backend (gloo: the tests) the device is drained first
        self._on_stream = world > 1 and dist.get_backend(group) == "nccl"
        self._flag = torch.zeros(1, dtype=torch.int32, device=solver.device if self._on_stream else "cpu")
        self.mode = "nccl" if mode == "nccl" or world == 1 else self._map_peers(mode)
        ptrs = self.peer_ptrs if self.mode == "push" else [self.ptr]
        self._arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        self._armed = False

    def _map_peers(self, mode: str) -> str:
        """Exchange the IPC handles of the slabs and map every peer's slab; returns the mode that works on every rank."""
        hbuf = ctypes.create_string_buffer(64)
        ok = self.L.sddp_ipc_export(ctypes.c_void_p(self.ptr), hbuf) == 0
        allh = [None] * self.world
        dist.all_gather_object(allh, bytes(hbuf.raw), group=self.group)
        ptrs = []
        if ok:
            with torch.cuda.device(self.solver.device):
                for r in range(self.world):
                    if r == self.rank:
                        ptrs.append(self.ptr)
                        continue
                    p = ctypes.c_void_p()
                    if self.L.sddp_ipc_open(allh[r], ctypes.byref(p)) != 0:
                        ok = False
                        break
                    ptrs.append(p.value)
        oks = [None] * self.world
        dist.all_gather_object(oks, bool(ok), group=self.group)
        if all(oks):
            self.peer_ptrs = ptrs
            return "push"
        for r, p in enumerate(ptrs):
            if r != self.rank:
                self.L.sddp_ipc_close(ctypes.c_void_p(p))
        if mode == "push":
            raise RuntimeError("ResultGather(mode='push'): peer slabs cannot be mapped (%s)" % self.L.sddp_last_error(None).decode())
        return "nccl"

    def _order_ranks(self) -> None:
        if not self._on_stream:
            torch.cuda.synchronize(self.solver.device)
        dist.all_reduce(self._flag, group=self.group)

    def arm(self, n_local: int) -> None:
        """Called by `BatchedDDP.solve(gather=...)` before the launch."""
        if n_local != self.hi - self.lo:
            raise ValueError("the solve must cover this rank's shard (shard_range)")
        if self.mode == "push" and self.world > 1:
            # nobody may store into a slab whose previous contents are still being read: every rank's consumers are
            # stream-ordered before this 4-byte all-reduce
            self._order_ranks()
        _lib.check(self.L.sddp_set_result_peers(self.solver.h, len(self._arr), self._arr, self.lo), self.solver.h, self.L)
        self._armed = True

    def disarm(self) -> None:
        """After the launch: later solves of the handle (without `gather=`) do not store records."""
        self.L.sddp_set_result_peers(self.solver.h, 0, None, 0)

    def finish(self) -> Dict[str, torch.Tensor]:
        if self.world > 1:
            if self.mode == "push":
                self._order_ranks()      # all kernels (and with them their peer stores) are complete
            elif (self.hi - self.lo) * self.world == self.B:
                dist.all_gather_into_tensor(self.slab, self.slab[self.lo:self.hi], group=self.group)      # in place, one collective
            else:
                for r in range(self.world):
                    lo, hi = shard_range(self.B, r, self.world)
                    dist.broadcast(self.slab[lo:hi], src=r, group=self.group)
        self._armed = False
        return slab_views(self.slab, self.solver.N, self.solver.nx, self.solver.nu)

    def close(self) -> None:
        if getattr(self, "solver", None) is None:
            return
        s = self.solver
        if getattr(s, "h", None):
            with torch.cuda.device(s.device):
                torch.cuda.synchronize(s.device)
                self.L.sddp_set_result_peers(s.h, 0, None, 0)
                for r, p in enumerate(self.peer_ptrs):
                    if r != self.rank:
                        self.L.sddp_ipc_close(ctypes.c_void_p(p))
                self.slab = None
                self.L.sddp_slab_alloc(s.h, 0, ctypes.byref(ctypes.c_void_p()))
        self.peer_ptrs, self.solver = [], None
