"""Multi-GPU sharding of a problem batch.

The path shards naturally: every MPC problem is independent (the reference runs one problem per
process, dsrbd_example.py:82-185).  One process per GPU solves a contiguous slice of the batch; the
only collective is the NCCL all-gather of the result slabs (trajectories, cost, iterations, status).
Gains are not gathered by default (MPC consumes u_0 / K_0 only; K is 7.1 KB per node).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


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
