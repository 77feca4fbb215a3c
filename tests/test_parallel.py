"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch and all-gather result slabs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from srbd_horizon_b200.parallel import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_batch():
    for B in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_range(B, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [hi - lo for lo, hi in pieces]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    from srbd_horizon_b200.ddp import BatchResult
    from srbd_horizon_b200.parallel import gather_results, shard_range, solve_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(B, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.float64)
    X = idx[:, None, None] * torch.ones(1, 3, 5, dtype=torch.float64)
    U = -idx[:, None, None] * torch.ones(1, 2, 4, dtype=torch.float64)
    K = idx[:, None, None, None] * torch.ones(1, 2, 4, 5, dtype=torch.float64)
    r = BatchResult(X=X, U=U, K=K, k=U.clone(), hist=None, iters=idx.to(torch.int32), status=torch.zeros(hi - lo, dtype=torch.int32), cost=idx * 2)
    g = gather_results(r, world, B, gains="first")
    ok = (g["X"].shape == (B, 3, 5) and torch.equal(g["X"][:, 0, 0], torch.arange(B, dtype=torch.float64))
          and torch.equal(g["cost"], 2 * torch.arange(B, dtype=torch.float64))
          and torch.equal(g["iters"], torch.arange(B, dtype=torch.int32)) and g["K"].shape == (B, 4, 5)
          and torch.equal(g["K"][:, 0, 0], torch.arange(B, dtype=torch.float64)))
    g2 = gather_results(r, world)     # batch size inferred by all-reduce
    ok = ok and g2["U"].shape == (B, 2, 4)
    # chunked solve with the gathers of one piece overlapping the next solve: original problem order, any chunk count
    def fake_solve(x0c, pc, Xc, Uc):
        i = x0c[:, 0]
        return BatchResult(X=i[:, None, None] * torch.ones(1, 3, 5, dtype=torch.float64), U=-i[:, None, None] * torch.ones(1, 2, 4, dtype=torch.float64),
                           K=None, k=None, hist=None, iters=i.to(torch.int32), status=torch.zeros(i.shape[0], dtype=torch.int32), cost=2 * i)
    x0 = idx[:, None].clone()
    for chunks in (1, 2, 3):
        g3 = solve_sharded(fake_solve, x0, x0, x0, x0, world, B, rank, chunks=chunks)
        ok = ok and (torch.equal(g3["X"][:, 0, 0], torch.arange(B, dtype=torch.float64)) and g3["U"].shape == (B, 2, 4)
                     and torch.equal(g3["iters"], torch.arange(B, dtype=torch.int32)) and torch.equal(g3["cost"], 2 * torch.arange(B, dtype=torch.float64)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_gather_results_gloo_world2(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_record_layout_and_slab_views():
    """The packed result record (include/sddp.h: X | U | cost | iters | status) and its whole-batch views."""
    from srbd_horizon_b200.parallel import record_layout, slab_views
    N, nx, nu, B = 5, 37, 24, 4
    lay = record_layout(N, nx, nu)
    assert lay["size"][1] == (N + 1) * nx + N * nu + 3 and lay["U"][0] == (N + 1) * nx and lay["status"][0] == lay["size"][1] - 1
    slab = torch.arange(B * lay["size"][1], dtype=torch.float64).reshape(B, -1).clone()
    slab[:, lay["iters"][0]] = torch.tensor([3, 4, 5, 6], dtype=torch.float64)
    slab[:, lay["status"][0]] = torch.tensor([0, 1, 0, 2], dtype=torch.float64)
    v = slab_views(slab, N, nx, nu)
    assert v["X"].shape == (B, N + 1, nx) and v["U"].shape == (B, N, nu) and v["cost"].shape == (B,)
    assert v["X"].data_ptr() == slab.data_ptr()                       # views, not copies
    assert v["U"][2, 1, 3] == slab[2, lay["U"][0] + nu + 3] and v["X"][1, 2, 5] == slab[1, 2 * nx + 5]
    assert v["iters"].tolist() == [3, 4, 5, 6] and v["status"].tolist() == [0, 1, 0, 2] and v["iters"].dtype == torch.int32
