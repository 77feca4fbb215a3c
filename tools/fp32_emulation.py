#!/usr/bin/env python
"""What would an fp32 build of the DDP iteration deliver?  (north_star: "optional fp32 build within a stated tolerance";
SURVEY.md section 7 step 9 predicts ~1e-3 because of the 1 ... 1e8 weight span.)

CPU experiment, no GPU: the Riccati recursion, the gains and the rollout of tests/np_ddp.py with every intermediate
rounded to float32 (derivatives and dynamics are evaluated by the fp64 oracle at the float32 iterate and rounded: the
best case for an fp32 kernel), on BASELINE configs[1] / [4]-like problems, against the fp64 oracle solve.

  python tools/fp32_emulation.py [N] [problems]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from srbd_horizon_b200.config import DIMS, MODEL_SRBD, make_config
from srbd_horizon_b200.problems import make_batch

f32 = np.float32
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cfg = make_config(MODEL_SRBD, N, 0.05, {"max_iters": 100, "alpha_converge_threshold": 1e-12, "beta": 1e-3})
nx, nu, _ = DIMS[MODEL_SRBD]
kind = lambda k: 0 if k == 0 else (2 if k == N else 1)
r32 = lambda a: np.asarray(a, dtype=f32)


def cost(X, U, P):
    return sum(O.cost(cfg, kind(k), X[k].astype(np.float64), U[k].astype(np.float64) if k < N else None, P[k]) for k in range(N + 1))


def backward32(X, U, P, d, mu):
    t = O.derivs(cfg, 2, X[N].astype(np.float64), None, P[N])
    Vx, Vxx = r32(t["lx"]), r32(t["lxx"])
    K = np.zeros((N, nu, nx), f32); kff = np.zeros((N, nu), f32)
    D1 = f32(0); D2 = f32(0); worst_cond = 0.0
    for k in range(N - 1, -1, -1):
        D = {n: r32(v) for n, v in O.derivs(cfg, kind(k), X[k].astype(np.float64), U[k].astype(np.float64), P[k]).items() if isinstance(v, np.ndarray)}
        fx, fu = D["fx"], D["fu"]
        vp = Vx + Vxx @ d[k]
        Qx = D["lx"] + fx.T @ vp; Qu = D["lu"] + fu.T @ vp
        Qxx = D["lxx"] + fx.T @ Vxx @ fx; Qux = D["lux"] + fu.T @ Vxx @ fx; Quu = D["luu"] + fu.T @ Vxx @ fu
        Quu = f32(0.5) * (Quu + Quu.T) + f32(mu) * np.eye(nu, dtype=f32)
        worst_cond = max(worst_cond, float(np.linalg.cond(Quu.astype(np.float64))))
        try:
            L = np.linalg.cholesky(Quu)          # float32 LAPACK
        except np.linalg.LinAlgError:
            return k + 1, K, kff, None, worst_cond
        sol = lambda b: np.linalg.solve(L.T, np.linalg.solve(L, b)).astype(f32)
        kk = -sol(Qu); Kk = -sol(Qux)
        K[k], kff[k] = Kk, kk
        D1 += Qu @ kk; D2 += f32(0.5) * (kk @ Quu @ kk)
        Vx = Qx + Kk.T @ (Quu @ kk) + Kk.T @ Qu + Qux.T @ kk
        Vxx = Qxx + Kk.T @ Quu @ Kk + Kk.T @ Qux + Qux.T @ Kk
        Vxx = f32(0.5) * (Vxx + Vxx.T)
    return 0, K, kff, (float(D1), float(D2)), worst_cond


def solve32(x0, P, X0, U0):
    X, U = r32(X0).copy(), r32(U0).copy(); X[0] = r32(x0)
    d = np.stack([r32(O.dynamics(cfg, X[k].astype(np.float64), U[k].astype(np.float64), kind(k))) - X[k + 1] for k in range(N)])
    J = cost(X, U, P); mu = 0.0; conds = []
    for it in range(cfg.max_iters):
        while True:
            rc, K, kff, dV, wc = backward32(X, U, P, d, mu)
            conds.append(wc)
            if rc == 0:
                break
            mu = max(mu * cfg.mu_factor, cfg.mu_min)
            if mu > cfg.mu_max:
                return X, U, it, "reg_failed", max(conds)
        alpha, acc = 1.0, False
        while alpha >= cfg.alpha_converge_threshold:
            Xn, Un = X.copy(), U.copy()
            for k in range(N):
                Un[k] = U[k] + f32(alpha) * kff[k] + K[k] @ (Xn[k] - X[k])
                Xn[k + 1] = r32(O.dynamics(cfg, Xn[k].astype(np.float64), Un[k].astype(np.float64), kind(k))) - f32(1 - alpha) * d[k]
            Jn = cost(Xn, Un, P)
            dJm = alpha * dV[0] + alpha * alpha * dV[1]
            if np.isfinite(Jn) and Jn - J <= dJm + (1 - cfg.beta) * abs(dJm):
                acc = True
                break
            alpha *= cfg.line_search_decrease_factor
        if not acc:
            mu = max(mu * cfg.mu_factor, cfg.mu_min)
            if mu > cfg.mu_max:
                return X, U, it + 1, "ls_failed", max(conds)
            continue
        dJ = J - Jn
        X, U, J = Xn, Un, Jn
        d = d * f32(1 - alpha)
        mu = mu / cfg.mu_factor
        mu = 0.0 if mu < cfg.mu_min else mu
        if dJ <= cfg.cost_reduction_ths * (1 + abs(J)) and np.abs(d).max() <= cfg.defect_ths:
            return X, U, it + 1, "converged", max(conds)
    return X, U, cfg.max_iters, "max_iters", max(conds)


b = make_batch(MODEL_SRBD, N, B, enumerate_schedules=True)
ro = O.solve_batch(cfg, b["x0"], b["params"], b["X0"], b["U0"], nthreads=8)
rel = lambda a, r: float(np.abs(a - r).max() / np.abs(r).max())
print(f"fp32 emulation of the DDP iteration, SRBD N={N}, {B} problems (eps_f32 = 6e-8)")
for i in range(B):
    X, U, it, st, wc = solve32(b["x0"][i], b["params"][i], b["X0"][i], b["U0"][i])
    print(f"  problem {i:2d}: fp64 iters {ro['iters'][i]:3d} | fp32 {st:10s} iters {it:3d}  max cond(Quu) {wc:8.1e}  rel err X {rel(X, ro['X'][i]):.1e}  U {rel(U, ro['U'][i]):.1e}"
          f"  cost {abs(cost(X, U, b['params'][i]) - ro['cost'][i]) / abs(ro['cost'][i]):.1e}")
