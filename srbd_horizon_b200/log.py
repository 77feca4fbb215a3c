"""Logging of DDP internals (the reference's TODO list, README.md:7: "logging/plotting DDP internals").

The solver records, per problem and iteration, the cost before the step, the accepted step size (0: none), the
regularisation and the largest defect (`hist[B, max_iters, 4]`, include/sddp.h).  These helpers turn that tensor into
rows / text; they work on the result of `BatchedDDP.solve` (torch tensors), of `solve_host` (numpy) and on
`DDPSolver.last`."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

STATUS_NAMES = {0: "converged", 1: "max_iters", 2: "line_search_failed", 3: "regularisation_failed", 4: "non_finite"}


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def history_rows(hist, iters, problem: int = 0) -> List[Dict[str, float]]:
    """One dict per DDP iteration of one problem: iteration, cost, alpha, mu, max_defect, cost_change (to the next row)."""
    h, it = _np(hist), _np(iters)
    if h.ndim == 2:
        h, it = h[None], np.atleast_1d(it)
    n = int(it[problem])
    rows = []
    for i in range(n):
        c = float(h[problem, i, 0])
        nxt = float(h[problem, i + 1, 0]) if i + 1 < n else float("nan")
        rows.append({"iteration": i, "cost": c, "alpha": float(h[problem, i, 1]), "mu": float(h[problem, i, 2]),
                     "max_defect": float(h[problem, i, 3]), "cost_change": nxt - c})
    return rows


def format_history(hist, iters, status=None, problem: int = 0) -> str:
    """The rows of `history_rows` as an aligned text table (what one would print from the MPC loop)."""
    rows = history_rows(hist, iters, problem)
    out = ["iter          cost     alpha        mu   max|defect|   cost change"]
    for r in rows:
        out.append(f"{r['iteration']:4d}  {r['cost']:12.6e}  {r['alpha']:8.2e}  {r['mu']:8.2e}  {r['max_defect']:12.3e}  {r['cost_change']:12.3e}")
    if status is not None:
        st = int(np.atleast_1d(_np(status))[problem])
        out.append(f"status: {STATUS_NAMES.get(st, st)} after {len(rows)} iteration(s)")
    return "\n".join(out)


def batch_summary(iters, status) -> Dict[str, float]:
    """Iteration-count statistics (iters_*) and the number of problems per final status of a batch."""
    it, st = _np(iters).astype(np.float64), _np(status)
    out = {"problems": int(it.size), "iters_mean": float(it.mean()) if it.size else 0.0, "iters_max": int(it.max()) if it.size else 0,
           "iters_p50": float(np.median(it)) if it.size else 0.0, "iters_p99": float(np.percentile(it, 99)) if it.size else 0.0}
    for code, name in STATUS_NAMES.items():
        out[name] = int((st == code).sum())
    return out
