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


def _gather(t: torch.Tensor, world: int, B: int) -> torch.Tensor:
    """all-gather along dim 0 of per-rank shards made by `shard_range` (pads uneven shards)."""
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    m = max(sizes)
    if t.shape[0] != m:
        pad = torch.zeros((m - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], dim=0)
    out = torch.empty((world * m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
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
