"""Independent numpy restatement of the DDP iteration documented in oracle/sddp_oracle.c.

Built only on the oracle's model primitives (dynamics / cost / derivs, themselves pinned to the
mpmath goldens) and numpy.linalg; used to cross-check the C loop's index arithmetic."""
import numpy as np

from oracle import oracle as O
from srbd_horizon_b200.config import DIMS


def kind(k, N):
    return 0 if k == 0 else (2 if k == N else 1)


def backward(cfg, X, U, params, d, mu):
    nx, nu, _ = DIMS[cfg.model]
    N = cfg.N
    fixed = cfg.defect_contraction_rate > 0
    rho_b = cfg.defect_contraction_rate if fixed else 1.0
    t = O.derivs(cfg, 2, X[N], None, params[N])
    Vx, Vxx = t["lx"].copy(), t["lxx"].copy()
    y = Vx.copy()
    K = np.zeros((N, nu, nx)); kff = np.zeros((N, nu))
    tot = acc1 = acc2 = 0.0
    for k in range(N - 1, -1, -1):
        D = O.derivs(cfg, kind(k, N), X[k], U[k], params[k])
        fx, fu = D["fx"], D["fu"]
        c = rho_b * d[k]
        s = Vxx @ c
        vp = Vx + s
        tot += Vx @ c + 0.5 * c @ s
        ys = y + s if fixed else y
        if fixed:
            acc1 += y @ c + 0.5 * c @ s
        Qx = D["lx"] + fx.T @ vp
        Qu = D["lu"] + fu.T @ vp
        Qxx = D["lxx"] + fx.T @ Vxx @ fx
        Qux = D["lux"] + fu.T @ Vxx @ fx
        Quu = D["luu"] + fu.T @ Vxx @ fu
        Quu = 0.5 * (Quu + Quu.T)
        try:
            L = np.linalg.cholesky(Quu + mu * np.eye(nu))
        except np.linalg.LinAlgError:
            return k + 1, K, kff, None
        sol = lambda b: np.linalg.solve(L.T, np.linalg.solve(L, b))
        kk = -sol(Qu); Kk = -sol(Qux)
        K[k], kff[k] = Kk, kk
        tot += Qu @ kk + 0.5 * kk @ Quu @ kk
        acc2 += 0.5 * kk @ Quu @ kk
        quy = D["lu"] + fu.T @ ys
        qxy = D["lx"] + fx.T @ ys
        if not fixed:
            acc1 += quy @ kk + y @ d[k]
        y = qxy + Kk.T @ quy
        Vx = Qx + Kk.T @ (Quu @ kk) + Kk.T @ Qu + Qux.T @ kk
        Vxx = Qxx + Kk.T @ Quu @ Kk + Kk.T @ Qux + Qux.T @ Kk
        Vxx = 0.5 * (Vxx + Vxx.T)
    dV = np.array([tot - acc1 - acc2, acc2, acc1]) if fixed else np.array([acc1, tot - acc1, 0.0])
    return 0, K, kff, dV


def forward(cfg, x0, X, U, params, d, K, kff, alpha, rho):
    N = cfg.N
    Xn = np.zeros_like(X); Un = np.zeros_like(U)
    Xn[0] = x0
    J = 0.0
    for k in range(N):
        Un[k] = U[k] + alpha * kff[k] + K[k] @ (Xn[k] - X[k])
        J += O.cost(cfg, kind(k, N), Xn[k], Un[k], params[k])
        Xn[k + 1] = O.dynamics(cfg, Xn[k], Un[k]) - (1.0 - rho) * d[k]
    J += O.cost(cfg, 2, Xn[N], None, params[N])
    return J, Xn, Un


def solve(cfg, x0, params, X0, U0):
    nx, nu, _ = DIMS[cfg.model]
    N = cfg.N
    X, U = X0.copy(), U0.copy()
    X[0] = x0
    d = np.zeros((N, nx))
    if not cfg.multiple_shooting:
        for k in range(N):
            X[k + 1] = O.dynamics(cfg, X[k], U[k])
    else:
        for k in range(N):
            d[k] = O.dynamics(cfg, X[k], U[k]) - X[k + 1]
    J = O.total_cost(cfg, X, U, params)
    mu = cfg.mu0
    fixed = cfg.defect_contraction_rate > 0
    hist = []
    status = 1
    for it in range(cfg.max_iters):
        while True:
            rc, K, kff, dV = backward(cfg, X, U, params, d, mu)
            if rc == 0:
                break
            mu = max(mu * cfg.mu_factor, cfg.mu_min)
            if mu > cfg.mu_max:
                return dict(X=X, U=U, K=K, k=kff, hist=np.array(hist), status=3, cost=J)
        dmax = np.abs(d).max()
        model = lambda a: dV[2] + a * dV[0] + a * a * dV[1]
        rec = [J, 0.0, mu, dmax]
        if abs(model(cfg.alpha_0)) <= 1e-3 * cfg.cost_reduction_ths * (1 + abs(J)) and dmax <= cfg.defect_ths:
            hist.append(rec); status = 0
            break
        a = cfg.alpha_0
        ok = False
        while a >= cfg.alpha_converge_threshold:
            rho = cfg.defect_contraction_rate if fixed else a
            Jn, Xn, Un = forward(cfg, x0, X, U, params, d, K, kff, a, rho)
            m = model(a)
            if np.isfinite(Jn) and Jn - J <= m + (1 - cfg.beta) * abs(m):
                ok = True
                break
            a *= cfg.line_search_decrease_factor
        if ok:
            X, U = Xn, Un
            d = d * (1.0 - rho)
            dJ, J = J - Jn, Jn
            rec = [J, a, mu, np.abs(d).max()]
            hist.append(rec)
            mu = mu / cfg.mu_factor
            if mu < cfg.mu_min:
                mu = 0.0
            mu = max(mu, cfg.mu0)
            if dJ <= cfg.cost_reduction_ths * (1 + abs(J)) and rec[3] <= cfg.defect_ths:
                status = 0
                break
        else:
            hist.append(rec)
            mu = max(mu * cfg.mu_factor, cfg.mu_min)
            if mu > cfg.mu_max:
                status = 2
                break
    return dict(X=X, U=U, K=K, k=kff, hist=np.array(hist), status=status, cost=J)
