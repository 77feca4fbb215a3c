"""Result records and the in-kernel multi-GPU gather (include/sddp.h: sddp_set_result_peers), on ONE GPU:
several slabs of the same device stand in for the peers, and two processes sharing cuda:0 exercise the CUDA IPC
mapping and the rank ordering of `parallel.ResultGather` (gloo for the control plane: NCCL refuses two ranks on one
GPU).  The 2..8 GPU run of the same code is bench.py under torchrun."""
import ctypes
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import torch.distributed as dist
import torch.multiprocessing as mp

from srbd_horizon_b200 import _lib
from srbd_horizon_b200.config import MODEL_SRBD, make_config
from srbd_horizon_b200.ddp import BatchedDDP
from srbd_horizon_b200.parallel import ResultGather, shard_range, slab_views
from srbd_horizon_b200.problems import make_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX_OPTS = {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3}
N = 20


def _problem(B, first=0):
    b = make_batch(MODEL_SRBD, N, B, first=first, enumerate_schedules=True)
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda")
    return t(b["x0"]), t(b["params"]), t(b["X0"]), t(b["U0"])


def test_records_land_in_every_slab_at_the_shard_offset():
    """Two slabs (as two peers would own them), records of a 7-problem shard that starts at record 5 of 16."""
    cfg = make_config(MODEL_SRBD, N, 0.05, EX_OPTS)
    s = BatchedDDP(cfg)
    L = s.L
    rec = L.sddp_record_doubles(s.h)
    assert rec == (N + 1) * 37 + N * 24 + 3
    total, first, B = 16, 5, 7
    slabs = [torch.full((total, rec), -7.0, dtype=torch.float64, device="cuda") for _ in range(2)]
    arr = (ctypes.c_void_p * 2)(*[t.data_ptr() for t in slabs])
    x0, p, X0, U0 = _problem(B, first=first)
    _lib.check(L.sddp_set_result_peers(s.h, 2, arr, first), s.h, L)
    r = s.solve(x0, p, X0, U0, order="schedule")
    _lib.check(L.sddp_set_result_peers(s.h, 0, None, 0), s.h, L)
    r2 = s.solve(x0, p, X0, U0)                       # switched off again: nothing is stored
    torch.cuda.synchronize()
    for t in slabs:
        v = slab_views(t, N, 37, 24)
        assert torch.equal(v["X"][first:first + B], r.X) and torch.equal(v["U"][first:first + B], r.U)
        assert torch.equal(v["cost"][first:first + B], r.cost)
        assert torch.equal(v["iters"][first:first + B], r.iters) and torch.equal(v["status"][first:first + B], r.status)
        assert bool((t[:first] == -7.0).all()) and bool((t[first + B:] == -7.0).all())      # nothing outside the shard
    assert torch.equal(r2.X, r.X)
    # argument checks
    assert L.sddp_set_result_peers(s.h, 9, arr, 0) == -1 and L.sddp_set_result_peers(s.h, 1, None, 0) == -1
    assert L.sddp_set_result_peers(s.h, 1, arr, -1) == -1


def test_result_gather_single_rank_and_host_path_untouched():
    cfg = make_config(MODEL_SRBD, N, 0.05, EX_OPTS)
    s = BatchedDDP(cfg)
    B = 9
    x0, p, X0, U0 = _problem(B)
    g = ResultGather(s, B, 0, 1)
    r = s.solve(x0, p, X0, U0, gather=g)
    out = g.finish()
    assert torch.equal(out["X"], r.X) and torch.equal(out["U"], r.U) and torch.equal(out["iters"], r.iters)
    # the host entry point never stores records (its chunks use chunk-local problem indices)
    g.slab.fill_(-1.0)
    g.arm(B)
    c = lambda t: t.cpu().numpy()
    rh = s.solve_host(c(x0), c(p), c(X0), c(U0))
    g.disarm()
    torch.cuda.synchronize()
    assert bool((g.slab == -1.0).all()) and np.array_equal(rh["X"], c(r.X))
    with pytest.raises(ValueError):
        s.solve(x0[:3], p[:3], X0[:3], U0[:3], gather=g)
    g.close()


def _rank(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        cfg = make_config(MODEL_SRBD, N, 0.05, EX_OPTS)
        s = BatchedDDP(cfg)
        lo, hi = shard_range(B, rank, world)
        x0, p, X0, U0 = _problem(hi - lo, first=lo)
        g = ResultGather(s, B, rank, world, mode="push")
        res = []
        for rep in range(2):                              # second pass: the slab is overwritten only after everyone read it
            r = s.solve(x0, p, X0, U0, gather=g, order="schedule")
            out = g.finish()
            res.append((out["X"].cpu().numpy().copy(), out["U"].cpu().numpy().copy(), out["cost"].cpu().numpy().copy(),
                        out["iters"].cpu().numpy().copy(), out["status"].cpu().numpy().copy()))
        q.put((rank, g.mode, res))
        dist.barrier()
        g.close()
    except Exception as e:      # noqa: BLE001
        q.put((rank, "error: %r" % (e,), None))
    finally:
        dist.destroy_process_group()


def test_two_processes_push_records_into_each_other_over_ipc():
    """Two ranks on cuda:0, each solves half of 10 problems; the kernel of each stores its records into both slabs
    (the peer's through a CUDA IPC mapping).  Every rank must end up with the whole batch, equal to one process solving it."""
    B, world = 10, 2
    cfg = make_config(MODEL_SRBD, N, 0.05, EX_OPTS)
    ref = BatchedDDP(cfg).solve(*_problem(B))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank, args=(r, world, port, B, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    for rank, mode, res in got:
        assert mode == "push", (rank, mode)
        for X, U, cost, iters, status in res:
            np.testing.assert_array_equal(X, ref.X.cpu().numpy())
            np.testing.assert_array_equal(U, ref.U.cpu().numpy())
            np.testing.assert_array_equal(cost, ref.cost.cpu().numpy())
            np.testing.assert_array_equal(iters, ref.iters.cpu().numpy())
            np.testing.assert_array_equal(status, ref.status.cpu().numpy())
