"""Oracle-independent SRBD problem in numpy (test infrastructure).

A third transcription of /root/reference/python/prb.py:97-109, 166-204 and of how ddp.py:179-230 assembles
f_k, L_k, L_N -- after the mpmath one behind the golden fixture (tests/golden/make_golden.py) and the C oracle
(oracle/sddp_oracle.c) -- written so that every operation also works on complex128.  That gives the total cost of a
trajectory as a function of the inputs alone (single-shooting rollout), its gradient to machine precision by the
complex-step method and, from that, optimality certificates that share nothing with the DDP iteration under test:
`reduced_gradient` must vanish at a solution, and a generic NLP solver started from the same point must arrive at the
same inputs (tests/test_optimality.py).

The Horizon helpers prb.py calls are not in the reference tree; they are restated as in make_golden.py ([EXTERNAL]).
Robot constants: the synthetic set of srbd_horizon_b200/config.py.
"""
import numpy as np

from srbd_horizon_b200.config import Gains, RobotConstants

NX, NU, NP = 37, 24, 19
CW = 1e6          # ddp.py:181 constraint_weight


def _cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def _to_rot(q):   # [EXTERNAL] horizon.utils.utils.toRot, quaternion (x, y, z, w), no normalisation
    qi, qj, qk, qr = q
    return np.array([[1 - 2 * (qj * qj + qk * qk), 2 * (qi * qj - qk * qr), 2 * (qi * qk + qj * qr)],
                     [2 * (qi * qj + qk * qr), 1 - 2 * (qi * qi + qk * qk), 2 * (qj * qk - qi * qr)],
                     [2 * (qi * qk - qj * qr), 2 * (qj * qk + qi * qr), 1 - 2 * (qi * qi + qj * qj)]])


def _quat_prod(q, p):   # [EXTERNAL] horizon.utils.utils.quaterion_product, scalar last
    v = q[3] * p[0:3] + p[3] * q[0:3] + _cross(q[0:3], p[0:3])
    return np.array([v[0], v[1], v[2], q[3] * p[3] - (q[0] * p[0] + q[1] * p[1] + q[2] * p[2])])


class SrbdRef:
    def __init__(self, N, dt, inertia_mode=0, robot=None, gains=None, lip_tail_start=0, ineq=None):
        """lip_tail_start: first node of the LIP-style tail (include/sddp.h; isrbd_example.py:344-353), 0 = none.
        ineq: dict of the inequality options of include/sddp.h (friction_cone_*, force_bound*, unilateral_weight,
        cdot_bound*, bound_sharpness); missing weights are 0 (= the reference, which drops every inequality)."""
        self.N, self.dt, self.inertia_mode = N, dt, inertia_mode
        self.tail0 = lip_tail_start
        self.ineq = dict(friction_cone_weight=0.0, friction_cone_mu=0.8, friction_cone_sharpness=6.0, force_bound_weight=0.0,
                         force_bound=1.0, unilateral_weight=0.0, cdot_bound_weight=0.0, cdot_bound=1.0, bound_sharpness=6.0)
        self.ineq.update(ineq or {})
        self.robot = robot or RobotConstants()
        self.gains = gains or Gains()
        self.fs = self.robot.force_scaling
        self.I = np.array(self.robot.inertia, dtype=np.float64).reshape(3, 3)
        self.foot = np.array(self.robot.foot, dtype=np.float64).reshape(4, 3)
        self.com = np.array(self.robot.com, dtype=np.float64)

    # prb.py:92-106
    def qddot(self, x, u):
        r, o, w = x[0:3], x[3:7], x[22:25]
        c = [x[7 + 3 * i:10 + 3 * i] for i in range(4)]
        f = [u[6 * i + 3:6 * i + 6] for i in range(4)]
        R = _to_rot(o)
        if self.inertia_mode == 0:      # prb.py:99 as written: CasADi `*` is element-wise
            Iw = R * (self.I / self.fs) * R.T
        else:                           # README.md:2 intent
            Iw = R @ (self.I / self.fs) @ R.T
        fsum = f[0] + f[1] + f[2] + f[3]
        tau = sum(_cross(c[i] - r, f[i]) for i in range(4))
        rddot = fsum / (self.robot.mass / self.fs) + np.array([0.0, 0.0, -self.robot.gravity])      # [EXTERNAL] kin_dyn.fSRBD
        wdot = np.linalg.solve(Iw, tau - _cross(w, Iw @ w))
        return rddot, wdot

    def kind(self, k):
        """0: node 0, 1: nodes 1..N-1, 2: node N, 3: a node 1..N-1 of the LIP-style tail"""
        if k == 0:
            return 0
        if k == self.N:
            return 2
        return 3 if (self.tail0 > 0 and k >= self.tail0) else 1

    # prb.py:107-109, ddp.py:228-230 (explicit Euler); tail nodes: no rotational dynamics (wdot = 0)
    def f(self, x, u, kind=1):
        rddot, wdot = self.qddot(x, u)
        if kind == 3:
            wdot = 0.0 * wdot
        w, o = x[22:25], x[3:7]
        qdot = _quat_prod(np.array([w[0] / 2, w[1] / 2, w[2] / 2, 0.0 * w[0]]), o)      # [EXTERNAL] LOCAL_WORLD_ALIGNED
        xd = np.concatenate([x[19:22], qdot, x[25:37], rddot, wdot] + [u[6 * i:6 * i + 3] for i in range(4)])
        return x + self.dt * xd

    # prb.py:166-204 as ddp.py:179-226 stacks them: residuals, then sqrt(1e6) x equality constraints, so that
    # L = sum(res^2).  kind 0: node 0, 1: nodes 1..N-1, 2: node N (trackers only, no constraints, no inputs)
    def residuals(self, x, u, p, kind):
        g = self.gains
        sq = np.sqrt
        r, o, rdot, w = x[0:3], x[3:7], x[19:22], x[22:25]
        c = [x[7 + 3 * i:10 + 3 * i] for i in range(4)]
        cdot = [x[25 + 3 * i:28 + 3 * i] for i in range(4)]
        res = []
        if kind >= 1:      # nodes 1..N (prb.py:184-199)
            res.append(sq(g.r_tracking_gain) * (r[2:3] - self.com[2]))
            qe = _quat_prod(o, p[15:19])
            res.append(p[6] * np.array([qe[0], qe[1], qe[2], qe[3] - 1]))
            res.append(sq(g.rdot_tracking_gain) * (rdot - p[0:3]))
            res.append(sq(g.w_tracking_gain) * (w - p[3:6]))
            d1, d2 = -(self.foot[0] - self.foot[2]), -(self.foot[1] - self.foot[3])
            for a, b, d in ((0, 2, d1), (1, 3, d2)):
                res.append(sq(g.rel_position_gain) * (-c[a][0:2] + c[b][0:2] - d[0:2]))
        if kind != 2:      # nodes 0..N-1 (prb.py:200-204, constraints :166-181 with weight 1e6)
            rddot, wdot = self.qddot(x, u)
            if kind == 3:      # LIP-style tail (isrbd_example.py:352-353): wdot = 0; lip_zero_angular_momentum, lip_com_height
                wdot = 0.0 * wdot
                res += [sq(CW) * w, sq(CW) * (r[2:3] - self.com[2])]
            res += [sq(g.min_qddot_gain) * rddot, sq(g.min_qddot_gain) * wdot]
            for i in range(4):
                cddot, f, sw = u[6 * i:6 * i + 3], u[6 * i + 3:6 * i + 6], p[8 + 2 * i]
                res.append(sq(g.min_qddot_gain) * cddot)
                res.append(self.fs * sq(g.min_f_gain) * f)
                res.append(self.fs * sq(g.force_switch_weight) * (1 - sw) * f)
                res.append(sq(CW) * (c[i][2:3] - p[7 + 2 * i]))
                res.append(sq(CW) * sw * cdot[i][0:2])
            for a, b in ((0, 1), (2, 3)):
                res.append(sq(CW) * (cdot[a][0:2] - cdot[b][0:2]))
        return np.concatenate(res)

    def barriers(self, x, u, kind):
        """Inequality terms of nodes 0..N-1 (extensions; ddp.py:197-209 sketches them, prb.py:173-177 builds the cone)."""
        q = self.ineq
        if kind == 2 or not any(q[k] for k in ("friction_cone_weight", "force_bound_weight", "unilateral_weight", "cdot_bound_weight")):
            return 0.0
        cost = 0.0
        kb = q["bound_sharpness"]
        for i in range(4):
            f, cdot = u[6 * i + 3:6 * i + 6], x[25 + 3 * i:28 + 3 * i]
            if q["friction_cone_weight"]:
                mu, kc = q["friction_cone_mu"], q["friction_cone_sharpness"]
                for gr in (f[0] - mu * f[2], -f[0] - mu * f[2], f[1] - mu * f[2], -f[1] - mu * f[2], -f[2]):
                    cost = cost + q["friction_cone_weight"] * np.exp(kc * gr)
            if q["force_bound_weight"]:
                cost = cost + q["force_bound_weight"] * np.sum(np.exp(kb * (f - q["force_bound"])) + np.exp(kb * (-q["force_bound"] - f)))
            if q["unilateral_weight"]:
                cost = cost + q["unilateral_weight"] * np.exp(-kb * f[2])
            if q["cdot_bound_weight"]:
                cost = cost + q["cdot_bound_weight"] * np.sum(np.exp(kb * (cdot - q["cdot_bound"])) + np.exp(kb * (-q["cdot_bound"] - cdot)))
        return cost

    def L(self, x, u, p, kind):
        return np.sum(self.residuals(x, u, p, kind) ** 2) + self.barriers(x, u, kind)

    def rollout(self, x0, U):
        X = [np.asarray(x0, dtype=U.dtype)]
        for k in range(self.N):
            X.append(self.f(X[k], U[k], self.kind(k)))
        return X

    def total_cost(self, x0, U, P):
        """J(U) = sum_k L_k(x_k, u_k, p_k) + L_N(x_N, p_N) along the rollout from x0 (single shooting)."""
        X = self.rollout(x0, U)
        J = self.L(X[self.N], None, P[self.N], 2)
        for k in range(self.N):
            J = J + self.L(X[k], U[k], P[k], self.kind(k))
        return J

    def total_residuals(self, x0, U, P):
        """All residuals of the trajectory rolled out from x0: J(U) = |total_residuals|^2 (without inequality barriers)."""
        X = self.rollout(x0, U)
        return np.concatenate([self.residuals(X[k], U[k], P[k], self.kind(k)) for k in range(self.N)]
                              + [self.residuals(X[self.N], None, P[self.N], 2)])

    def node_gradient(self, x, u, p, kind, h=1e-30):
        """(lx, lu) of one node by complex step."""
        z = np.concatenate([x, u]).astype(np.complex128)
        g = np.zeros(NX + NU)
        for i in range(NX + NU):
            z[i] += 1j * h
            g[i] = np.imag(self.L(z[:NX], z[NX:], p, kind)) / h
            z[i] -= 1j * h
        return g[:NX], g[NX:]

    def node_hessian(self, x, u, p, kind, h=1e-6):
        """Hessian of L of one node with respect to [x; u]: central differences of the complex-step gradient."""
        n = NX + NU
        z = np.concatenate([x, u]).astype(np.float64)
        H = np.zeros((n, n))
        for j in range(n):
            zp, zm = z.copy(), z.copy()
            zp[j] += h; zm[j] -= h
            gp = np.concatenate(self.node_gradient(zp[:NX], zp[NX:], p, kind))
            gm = np.concatenate(self.node_gradient(zm[:NX], zm[NX:], p, kind))
            H[:, j] = (gp - gm) / (2 * h)
        return 0.5 * (H + H.T)

    def reduced_gradient(self, x0, U, P, h=1e-30):
        """dJ/dU [N, nu] by the complex-step method (exact to rounding; no subtractive cancellation)."""
        U = np.asarray(U, dtype=np.float64)
        g = np.zeros_like(U)
        Uc = U.astype(np.complex128)
        for k in range(self.N):
            for i in range(NU):
                Uc[k, i] += 1j * h
                g[k, i] = np.imag(self.total_cost(x0, Uc, P)) / h
                Uc[k, i] = U[k, i]
        return g
