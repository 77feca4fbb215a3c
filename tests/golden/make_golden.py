#!/usr/bin/env python
"""Generate tests/golden/model_golden.npz -- golden vectors for the model functions.

This is a LITERAL transcription of the reference's problem definitions
(/root/reference/python/prb.py and the way /root/reference/python/ddp.py:179-230
turns them into f_k, L_k, L_N) into mpmath arithmetic at 80 significant digits.
Only VALUES are transcribed (ode, residual vectors); every derivative in the
fixture is obtained from those values by high-precision central differences
(step 1e-20 / 1e-15 at 80 digits => error << 1e-25), so the fixture shares no
derivative code with the oracle or with the CUDA kernels.

The Horizon helpers prb.py calls are not in the reference tree; they are restated
from their published definitions and marked [EXTERNAL]:
  utils.toRot, utils.quaterion_product, kin_dyn.fSRBD,
  utils.double_integrator_with_floating_base(LOCAL_WORLD_ALIGNED), utils.double_integrator.

Robot constants are the synthetic set of srbd_horizon_b200/config.py (the
reference reads them from an external URDF).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from srbd_horizon_b200.config import Gains, RobotConstants  # noqa: E402  (plain data only)

mp.mp.dps = 80
ROBOT = RobotConstants()
GAINS = Gains()
FS = mp.mpf(ROBOT.force_scaling)
CW = mp.mpf(GAINS.constraint_weight)   # ddp.py:181


def M(rows):
    return mp.matrix(rows)


def skew(v):
    return M([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def cross(a, b):
    return M([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def toRot(q):  # [EXTERNAL] horizon.utils.utils.toRot, quaternion (x,y,z,w), no normalisation
    qi, qj, qk, qr = q
    return M([[1 - 2 * (qj * qj + qk * qk), 2 * (qi * qj - qk * qr), 2 * (qi * qk + qj * qr)],
              [2 * (qi * qj + qk * qr), 1 - 2 * (qi * qi + qk * qk), 2 * (qj * qk - qi * qr)],
              [2 * (qi * qk - qj * qr), 2 * (qj * qk + qi * qr), 1 - 2 * (qi * qi + qj * qj)]])


def quaterion_product(q, p):  # [EXTERNAL] horizon.utils.utils.quaterion_product, scalar last
    qv, pv = M(q[0:3]), M(p[0:3])
    v = q[3] * pv + p[3] * qv + cross(qv, pv)
    return [v[0], v[1], v[2], q[3] * p[3] - (qv.T * pv)[0]]


def fSRBD(m, I, f, r, c, w):  # [EXTERNAL] horizon.utils.kin_dyn.fSRBD
    fsum = M([0, 0, 0])
    tau = M([0, 0, 0])
    for i in range(len(f)):
        fsum += f[i]
        tau += cross(c[i] - r, f[i])
    rddot = fsum / m + M([0, 0, -mp.mpf(ROBOT.gravity)])
    wdot = mp.inverse(I) * (tau - cross(w, I * w))
    return rddot, wdot


def elementwise(A, B):
    return M([[A[i, j] * B[i, j] for j in range(3)] for i in range(3)])


def srbd_split(x, u):
    r, o = M(x[0:3]), list(x[3:7])
    c = [M(x[7 + 3 * i:10 + 3 * i]) for i in range(4)]
    rdot, w = M(x[19:22]), M(x[22:25])
    cdot = [M(x[25 + 3 * i:28 + 3 * i]) for i in range(4)]
    cddot = [M(u[6 * i:6 * i + 3]) for i in range(4)] if u is not None else None
    f = [M(u[6 * i + 3:6 * i + 6]) for i in range(4)] if u is not None else None
    return r, o, c, rdot, w, cdot, cddot, f


def srbd_qddot(x, u, inertia_mode):
    """prb.py:92-106: rddot, wdot and the aggregate qddot = [rddot; wdot; cddot_i]"""
    r, o, c, rdot, w, cdot, cddot, f = srbd_split(x, u)
    I = M([[mp.mpf(ROBOT.inertia[3 * i + j]) for j in range(3)] for i in range(3)])
    w_R_b = toRot(o)
    if inertia_mode == 0:   # prb.py:99 as written: CasADi `*` on SX is element-wise
        Iw = elementwise(elementwise(w_R_b, I / FS), w_R_b.T)
    else:                   # README.md:2 intent
        Iw = w_R_b * (I / FS) * w_R_b.T
    rddot, wdot = fSRBD(mp.mpf(ROBOT.mass) / FS, Iw, f, r, c, w)
    return rddot, wdot, cddot


def srbd_ode(x, u, inertia_mode):
    """prb.py:107-109 with [EXTERNAL] double_integrator_with_floating_base(LOCAL_WORLD_ALIGNED):
    xdot = [rdot; quat_prod([w/2, 0], o); cdot_i; qddot]"""
    r, o, c, rdot, w, cdot, cddot, f = srbd_split(x, u)
    rddot, wdot, _ = srbd_qddot(x, u, inertia_mode)
    qdot = quaterion_product([w[0] / 2, w[1] / 2, w[2] / 2, mp.mpf(0)], o)
    out = list(rdot) + qdot
    for i in range(4):
        out += list(cdot[i])
    out += list(rddot) + list(wdot)
    for i in range(4):
        out += list(cddot[i])
    return out


def srbd_residuals(x, u, p, kind, inertia_mode):
    """(weight-included residual list, constraint list) active at a node of `kind`
    0: node 0 (range(0,ns) terms only), 1: 1..N-1 (all), 2: node N (range(1,ns+1) terms, no constraints)."""
    r, o, c, rdot, w, cdot, cddot, f = srbd_split(x, u)
    rdot_ref, w_ref, otg = M(p[0:3]), M(p[3:6]), p[6]
    c_ref = [p[7 + 2 * i] for i in range(4)]
    sw = [p[8 + 2 * i] for i in range(4)]
    oref = list(p[15:19])
    foot = [M([mp.mpf(v) for v in ROBOT.foot[3 * i:3 * i + 3]]) for i in range(4)]
    com = [mp.mpf(v) for v in ROBOT.com]
    res, con = [], []
    sq = mp.sqrt
    if kind >= 1:  # nodes=range(1, ns+1), prb.py:184-199
        res.append(sq(GAINS.r_tracking_gain) * (r[2] - com[2]))
        qe = quaterion_product(o, oref)
        res += [otg * qe[0], otg * qe[1], otg * qe[2], otg * (qe[3] - 1)]
        res += list(sq(GAINS.rdot_tracking_gain) * (rdot - rdot_ref))
        res += list(sq(GAINS.w_tracking_gain) * (w - w_ref))
        d1 = -(foot[0] - foot[2])
        d2 = -(foot[1] - foot[3])
        g = sq(GAINS.rel_position_gain)
        res.append(g * (-c[0][1] + c[2][1] - d1[1]))
        res.append(g * (-c[0][0] + c[2][0] - d1[0]))
        res.append(g * (-c[1][1] + c[3][1] - d2[1]))
        res.append(g * (-c[1][0] + c[3][0] - d2[0]))
    if kind <= 1:  # nodes=range(0, ns), prb.py:200-204
        rddot, wdot, _ = srbd_qddot(x, u, inertia_mode)
        g = sq(GAINS.min_qddot_gain)
        res += list(g * rddot) + list(g * wdot)
        for i in range(4):
            res += list(g * cddot[i])
        for i in range(4):
            res += list(FS * sq(GAINS.min_f_gain) * f[i])
            res += list(FS * sq(GAINS.force_switch_weight) * (1 - sw[i]) * f[i])
        # equality constraints (prb.py:166-181), summed into L_k only (ddp.py:191-196, 216-226)
        con += [cdot[0][0] - cdot[1][0], cdot[0][1] - cdot[1][1]]
        con += [cdot[2][0] - cdot[3][0], cdot[2][1] - cdot[3][1]]
        for i in range(4):
            con.append(c[i][2] - c_ref[i])
            con += [sw[i] * cdot[i][0], sw[i] * cdot[i][1]]
    return res, con


def lip_split(x, u):
    r = M(x[0:3])
    c = [M(x[3 + 3 * i:6 + 3 * i]) for i in range(4)]
    rdot = M(x[15:18])
    cdot = [M(x[18 + 3 * i:21 + 3 * i]) for i in range(4)]
    z = M(u[0:3]) if u is not None else None
    cddot = [M(u[3 + 3 * i:6 + 3 * i]) for i in range(4)] if u is not None else None
    return r, c, rdot, cdot, z, cddot


ETA2 = mp.mpf("9.81") / mp.mpf("0.88")


def lip_rddot(r, z):
    return ETA2 * (r - z) - M([0, 0, mp.mpf("9.81")])   # prb.py:317-318


def lip_ode(x, u):
    r, c, rdot, cdot, z, cddot = lip_split(x, u)
    out = list(rdot)
    for i in range(4):
        out += list(cdot[i])
    out += list(lip_rddot(r, z))
    for i in range(4):
        out += list(cddot[i])
    return out


def lip_residuals(x, u, p, kind):
    r, c, rdot, cdot, z, cddot = lip_split(x, u)
    rdot_ref = M(p[0:3])
    c_ref = [p[3 + 2 * i] for i in range(4)]
    sw = [p[4 + 2 * i] for i in range(4)]
    foot = [M([mp.mpf(v) for v in ROBOT.foot[3 * i:3 * i + 3]]) for i in range(4)]
    com = [mp.mpf(v) for v in ROBOT.com]
    res, con = [], []
    sq = mp.sqrt
    csum = c[0] + c[1] + c[2] + c[3]
    if kind >= 1:  # prb.py:390-392, 394-401
        res.append(sq(GAINS.r_tracking_gain) * (r[2] - com[2]))
        res += [sq(GAINS.r_tracking_gain) * (r[k] - csum[k] * mp.mpf("0.25")) for k in range(2)]
        res += list(sq(GAINS.rdot_tracking_gain) * (rdot - rdot_ref))
        d1 = -(foot[0] - foot[2])
        d2 = -(foot[1] - foot[3])
        g = sq(GAINS.rel_position_gain)
        res.append(g * (-c[0][1] + c[2][1] - d1[1]))
        res.append(g * (-c[0][0] + c[2][0] - d1[0]))
        res.append(g * (-c[1][1] + c[3][1] - d2[1]))
        res.append(g * (-c[1][0] + c[3][0] - d2[0]))
    if kind <= 1:  # prb.py:393, 402, 379-387
        res += list(sq(GAINS.zmp_tracking_gain) * (z - csum * mp.mpf("0.25")))
        g = sq(GAINS.min_qddot_gain)
        res += list(g * lip_rddot(r, z))
        for i in range(4):
            res += list(g * cddot[i])
        con += [cdot[0][0] - cdot[1][0], cdot[0][1] - cdot[1][1]]
        con += [cdot[2][0] - cdot[3][0], cdot[2][1] - cdot[3][1]]
        for i in range(4):
            con.append(c[i][2] - c_ref[i])
            con += [sw[i] * cdot[i][0], sw[i] * cdot[i][1]]
    return res, con


# ----------------------------------------------------------------------------- assembly as in ddp.py
class Model:
    def __init__(self, name, inertia_mode=0):
        self.name = name
        self.inertia_mode = inertia_mode
        self.nx, self.nu, self.np = (37, 24, 19) if name == "srbd" else (30, 15, 11)

    def ode(self, x, u):
        return srbd_ode(x, u, self.inertia_mode) if self.name == "srbd" else lip_ode(x, u)

    def residuals(self, x, u, p, kind):
        if self.name == "srbd":
            return srbd_residuals(x, u, p, kind, self.inertia_mode)
        return lip_residuals(x, u, p, kind)

    def f(self, x, u, dt):   # ddp.py:228-230 (integrators.EULER)
        xd = self.ode(x, u)
        return [x[i] + dt * xd[i] for i in range(self.nx)]

    def L(self, x, u, p, kind):   # ddp.py:179-226
        res, con = self.residuals(x, u, p, kind)
        return sum(v * v for v in res) + CW * sum(v * v for v in con)

    def stacked(self, x, u, p, kind):   # sqrt-weighted residual stack, for the Gauss-Newton Hessian
        res, con = self.residuals(x, u, p, kind)
        return list(res) + [mp.sqrt(CW) * v for v in con]


def jac_fd(fun, z, h=mp.mpf("1e-25")):
    cols = []
    for i in range(len(z)):
        zp, zm = list(z), list(z)
        zp[i] += h
        zm[i] -= h
        a, b = fun(zp), fun(zm)
        cols.append([(a[k] - b[k]) / (2 * h) for k in range(len(a))])
    return np.array([[float(cols[j][i]) for j in range(len(z))] for i in range(len(cols[0]))])


def grad_hess_fd(fun, z, h=mp.mpf("1e-18")):
    n = len(z)
    f0 = fun(z)
    g = np.zeros(n)
    H = np.zeros((n, n))
    fp, fm = [], []
    for i in range(n):
        zp, zm = list(z), list(z)
        zp[i] += h
        zm[i] -= h
        fp.append(fun(zp))
        fm.append(fun(zm))
        g[i] = float((fp[i] - fm[i]) / (2 * h))
        H[i, i] = float((fp[i] - 2 * f0 + fm[i]) / (h * h))
    for i in range(n):
        for j in range(i + 1, n):
            zpp, zmm = list(z), list(z)
            zpp[i] += h; zpp[j] += h
            zmm[i] -= h; zmm[j] -= h
            v = (fun(zpp) - fp[i] - fp[j] + 2 * f0 - fm[i] - fm[j] + fun(zmm)) / (2 * h * h)
            H[i, j] = H[j, i] = float(v)
    return float(f0), g, H


def sample_point(rng, model):
    foot = np.array(ROBOT.foot)
    if model.name == "srbd":
        x = np.zeros(37)
        x[0:3] = np.array(ROBOT.com) + rng.uniform(-0.05, 0.05, 3)
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        ang = rng.uniform(0.05, 0.4)
        x[3:6] = ax * np.sin(ang / 2); x[6] = np.cos(ang / 2)
        x[3:7] *= rng.uniform(0.95, 1.05)     # Euler steps leave the quaternion un-normalised
        x[7:19] = foot + rng.uniform(-0.05, 0.05, 12)
        x[19:25] = rng.uniform(-0.5, 0.5, 6)
        x[25:37] = rng.uniform(-0.3, 0.3, 12)
        u = np.zeros(24)
        for i in range(4):
            u[6 * i:6 * i + 3] = rng.uniform(-1, 1, 3)
            u[6 * i + 3:6 * i + 6] = np.array([0, 0, ROBOT.mass * 9.81 / 1000 / 4]) + rng.uniform(-0.05, 0.05, 3)
        p = np.zeros(19)
        p[0:3] = rng.uniform(-0.5, 0.5, 3)
        p[3:6] = rng.uniform(-0.2, 0.2, 3)
        p[6] = rng.choice([10.0, 100.0, 3.7])
        for i in range(4):
            p[7 + 2 * i] = rng.uniform(0, 0.05)
            p[8 + 2 * i] = rng.choice([0.0, 1.0, 0.3])
        q = rng.normal(size=4) * 0.1 + np.array([0, 0, 0, 1.0])
        p[15:19] = q / np.linalg.norm(q)
    else:
        x = np.zeros(30)
        x[0:3] = np.array(ROBOT.com) + rng.uniform(-0.05, 0.05, 3)
        x[3:15] = foot + rng.uniform(-0.05, 0.05, 12)
        x[15:18] = rng.uniform(-0.5, 0.5, 3)
        x[18:30] = rng.uniform(-0.3, 0.3, 12)
        u = np.concatenate([x[0:3] * [1, 1, 0] + rng.uniform(-0.05, 0.05, 3), rng.uniform(-1, 1, 12)])
        p = np.zeros(11)
        p[0:3] = rng.uniform(-0.5, 0.5, 3)
        for i in range(4):
            p[3 + 2 * i] = rng.uniform(0, 0.05)
            p[4 + 2 * i] = rng.choice([0.0, 1.0, 0.3])
    return x, u, p


def main():
    rng = np.random.default_rng(20261018)
    dt = 0.05
    out = {}
    cases = [("srbd", 0), ("srbd", 1), ("lip", 0)]
    idx = 0
    for name, mode in cases:
        model = Model(name, mode)
        nx, nu = model.nx, model.nu
        for kind in (0, 1, 2):
            for rep in range(2 if name == "srbd" else 1):
                x, u, p = sample_point(rng, model)
                xm = [mp.mpf(float(v)) for v in x]
                um = [mp.mpf(float(v)) for v in u]
                pm = [mp.mpf(float(v)) for v in p]
                zm = xm + um
                dtm = mp.mpf(dt)
                key = f"case{idx:02d}"
                idx += 1
                out[key + "_meta"] = np.array([0 if name == "srbd" else 1, mode, kind], dtype=np.int64)
                out[key + "_x"], out[key + "_u"], out[key + "_p"] = x, u, p
                if kind != 2:
                    fz = lambda z: model.f(z[:nx], z[nx:], dtm)
                    out[key + "_f"] = np.array([float(v) for v in fz(zm)])
                    Jf = jac_fd(fz, zm)
                    out[key + "_fx"], out[key + "_fu"] = Jf[:, :nx], Jf[:, nx:]
                Lz = lambda z: model.L(z[:nx], z[nx:], pm, kind)
                L0, g, H = grad_hess_fd(Lz, zm)
                out[key + "_L"] = np.array(L0)
                out[key + "_lx"], out[key + "_lu"] = g[:nx], g[nx:]
                out[key + "_lxx"], out[key + "_lux"], out[key + "_luu"] = H[:nx, :nx], H[nx:, :nx], H[nx:, nx:]
                Jr = jac_fd(lambda z: model.stacked(z[:nx], z[nx:], pm, kind), zm)
                Hgn = 2.0 * Jr.T @ Jr
                out[key + "_gn_lxx"], out[key + "_gn_lux"], out[key + "_gn_luu"] = Hgn[:nx, :nx], Hgn[nx:, :nx], Hgn[nx:, nx:]
                print(key, name, "inertia_mode", mode, "kind", kind, "L =", L0, flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
