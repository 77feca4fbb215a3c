/*
 * sddp_oracle.c -- CPU oracle (plain C, fp64) for the srbd_horizon DDP hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see sddp_oracle.h.  PARITY UNPINNED for the DDP
 * iteration (pyddp's source is absent); the model functions are pinned to
 * sympy golden vectors transcribed from the reference's prb.py.
 *
 * What is restated from the reference (file:line under /root/reference/python):
 *   state / input / parameter layouts ........ prb.py:32-68, 71-72, 143, 159-163, 185 (SRBD)
 *                                              prb.py:264-295, 298, 372-376          (LIP)
 *   SRBD ode .................................. prb.py:97-109   (fSRBD, toRot, quaternion kinematics
 *                                              with a world-aligned angular velocity)
 *   LIP ode ................................... prb.py:315-329
 *   residuals and their node ranges ........... prb.py:184-204 (SRBD), prb.py:390-402 (LIP)
 *   equality constraints (penalised by 1e6) ... prb.py:166-181, 379-387 with ddp.py:181,191-196
 *   L_k = sum ||res||^2 + 1e6 sum ||g||^2 ..... ddp.py:179-214 (no 1/2 factor, no inequality terms)
 *   L_N = sum ||res||^2 (no constraints) ...... ddp.py:216-226
 *   f_k = x + dt * ode(x,u) ................... ddp.py:228-230
 *   option names .............................. ddp.py:14-35
 *
 * What the oracle defines because nothing upstream fixes it (the DDP iteration):
 *
 *   defects      d_k = f(X_k,U_k) - X_{k+1}          (zero in single shooting after the initial rollout)
 *   backward     c_k = d_k (rho = alpha mode) or rho*d_k (fixed contraction rate)
 *                v+  = Vx' + Vxx' c_k
 *                Qx = lx + fx^T v+,  Qu = lu + fu^T v+
 *                Qxx = lxx + fx^T Vxx' fx, Qux = lux + fu^T Vxx' fx, Quu = luu + fu^T Vxx' fu
 *                (iLQR: no second-order dynamics terms)
 *                Quu_r = Quu + mu I = L L^T (Cholesky; failure -> mu = max(mu*mu_factor, mu_min), restart)
 *                k = -Quu_r^-1 Qu,  K = -Quu_r^-1 Qux
 *                Vx  = Qx + K^T Quu k + K^T Qu + Qux^T k
 *                Vxx = sym(Qxx + K^T Quu K + K^T Qux + Qux^T K)
 *   model of the cost change for step alpha:  dJ(alpha) = C0 + alpha*D1 + alpha^2*D2
 *                tot = sum_k [Qu.k + 1/2 k^T Quu k] + sum_k [Vx'.c + 1/2 c^T Vxx' c]   (= dJ(1))
 *                y-recursion  y_N = l_Nx,  y_k = (lx + fx^T(y'+s)) + K^T (lu + fu^T(y'+s))
 *                  rho = alpha : s = 0        D1 = sum_k (lu + fu^T y').k + y'.d_k   (exact first-order
 *                                             change along the alpha=1 step), D2 = tot - D1, C0 = 0
 *                  fixed rho   : s = Vxx' c   C0 = sum_k y'.c + 1/2 c^T Vxx' c  (feedback-only response
 *                                             to the gaps), D2 = 1/2 sum k^T Quu k, D1 = tot - C0 - D2
 *                (with d = 0 and mu = 0 this is the classic  alpha*sum Qu.k + alpha^2/2 sum k^T Quu k)
 *   forward      x^_0 = x0, u^_k = U_k + alpha k_k + K_k (x^_k - X_k),
 *                x^_{k+1} = f(x^_k, u^_k) - (1 - rho) d_k,    rho = alpha or the fixed rate
 *   line search  alpha_j = alpha_0 * factor^j while alpha_j >= alpha_converge_threshold; accept the
 *                first (largest) alpha with   J(alpha) - J <= dJ(alpha) + (1 - beta) |dJ(alpha)|
 *                (for dJ < 0 this is the Armijo test  J(alpha) - J <= beta dJ(alpha))
 *   on accept    X,U <- trial, d <- (1 - rho) d, mu <- mu / mu_factor (below mu_min -> 0, floor mu0);
 *                converged when  J - J_new <= cost_reduction_ths * (1 + |J_new|)  and max|d| <= defect_ths
 *   before the   converged when |dJ(alpha_0)| <= cost_reduction_ths * (1 + |J|) * 1e-3 and max|d| <= defect_ths
 *   line search  (stationary point: nothing left to gain)
 *   on failure   mu <- max(mu*mu_factor, mu_min); mu > mu_max -> status LS_FAILED
 *   bad input    a non-finite initial gap (multiple shooting) -> status NAN before the first iteration
 */
#include "sddp_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define NXM 37
#define NUM 24
#define NPM 19
#define NZ 34 /* variables the rigid-body angular acceleration depends on: r,o,c0..3,w,f0..3 */

/* ---- SRBD layout (prb.py:32-68, 224-246) ---- */
enum { SX_R = 0, SX_O = 3, SX_C = 7, SX_RD = 19, SX_W = 22, SX_CD = 25 };
#define SU_CDD(i) (6 * (i))
#define SU_F(i) (6 * (i) + 3)
enum { SP_RDREF = 0, SP_WREF = 3, SP_OTG = 6, SP_OREF = 15 };
#define SP_CREF(i) (7 + 2 * (i))
#define SP_SW(i) (8 + 2 * (i))
/* ---- LIP layout (prb.py:264-295, 420-441) ---- */
enum { LX_R = 0, LX_C = 3, LX_RD = 15, LX_CD = 18 };
enum { LU_Z = 0 };
#define LU_CDD(i) (3 + 3 * (i))
enum { LP_RDREF = 0 };
#define LP_CREF(i) (3 + 2 * (i))
#define LP_SW(i) (4 + 2 * (i))

void orc_dims(int model, int *nx, int *nu, int *np) {
    if (model == 0) { *nx = 37; *nu = 24; *np = 19; }
    else            { *nx = 30; *nu = 15; *np = 11; }
}

/* ------------------------------------------------------------------ small helpers */
static void skew(const double v[3], double S[9]) {
    S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}
static void cross(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static void mat3mul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j];
        C[3 * i + j] = s;
    }
}
static void mat3mulT(const double A[9], const double B[9], double C[9]) { /* A * B^T */
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * j + k];
        C[3 * i + j] = s;
    }
}
static void mat3vec(const double A[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; i++) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void inv3(const double A[9], double M[9]) {
    double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    double id = 1.0 / det;
    M[0] = c00 * id; M[1] = (A[2] * A[7] - A[1] * A[8]) * id; M[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    M[3] = c01 * id; M[4] = (A[0] * A[8] - A[2] * A[6]) * id; M[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    M[6] = c02 * id; M[7] = (A[1] * A[6] - A[0] * A[7]) * id; M[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

/* ------------------------------------------------------------------ rotation from a (non-normalised)
 * quaternion (x,y,z,w), horizon utils.toRot as used at prb.py:97.  R_ij = delta_ij + q^T Qt_ij q. */
static double QT[3][3][4][4];
static int QT_ready = 0;
static void qt_add(int i, int j, int a, int b, double v) {
    if (a == b) QT[i][j][a][a] += v;
    else { QT[i][j][a][b] += 0.5 * v; QT[i][j][b][a] += 0.5 * v; }
}
static void qt_init(void) {
    if (QT_ready) return;
    memset(QT, 0, sizeof QT);
    enum { x = 0, y = 1, z = 2, w = 3 };
    qt_add(0, 0, y, y, -2); qt_add(0, 0, z, z, -2);
    qt_add(0, 1, x, y, 2);  qt_add(0, 1, z, w, -2);
    qt_add(0, 2, x, z, 2);  qt_add(0, 2, y, w, 2);
    qt_add(1, 0, x, y, 2);  qt_add(1, 0, z, w, 2);
    qt_add(1, 1, x, x, -2); qt_add(1, 1, z, z, -2);
    qt_add(1, 2, y, z, 2);  qt_add(1, 2, x, w, -2);
    qt_add(2, 0, x, z, 2);  qt_add(2, 0, y, w, -2);
    qt_add(2, 1, y, z, 2);  qt_add(2, 1, x, w, 2);
    qt_add(2, 2, x, x, -2); qt_add(2, 2, y, y, -2);
    QT_ready = 1;
}
static void rot_all(const double q[4], double R[9], double Ra[4][9], double Rab[4][4][9]) {
    qt_init();
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = (i == j) ? 1.0 : 0.0;
        for (int a = 0; a < 4; a++) {
            double g = 0;
            for (int b = 0; b < 4; b++) { g += QT[i][j][a][b] * q[b]; Rab[a][b][3 * i + j] = 2.0 * QT[i][j][a][b]; }
            Ra[a][3 * i + j] = 2.0 * g;
            s += q[a] * g;
        }
        R[3 * i + j] = s;
    }
}

/* Inertia as the reference builds it (prb.py:99): J(o) and its o-derivatives.
 * literal: element-wise  R o (I/fs) o R^T  (CasADi `*` on SX is element-wise)
 * rotated: R (I/fs) R^T  (the README's intent) */
static void inertia_all(const OrcConfig *c, const double q[4], double J[9], double Ja[4][9], double Jab[4][4][9]) {
    double R[9], Ra[4][9], Rab[4][4][9], Ib[9];
    rot_all(q, R, Ra, Rab);
    for (int i = 0; i < 9; i++) Ib[i] = c->inertia[i] / c->force_scaling;
    if (c->inertia_mode == 0) {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
            int ij = 3 * i + j, ji = 3 * j + i;
            J[ij] = R[ij] * R[ji] * Ib[ij];
            for (int a = 0; a < 4; a++) {
                Ja[a][ij] = (Ra[a][ij] * R[ji] + R[ij] * Ra[a][ji]) * Ib[ij];
                for (int b = 0; b < 4; b++)
                    Jab[a][b][ij] = (Rab[a][b][ij] * R[ji] + Ra[a][ij] * Ra[b][ji] + Ra[b][ij] * Ra[a][ji] + R[ij] * Rab[a][b][ji]) * Ib[ij];
            }
        }
    } else {
        double T[9], T2[9], T3[9];
        mat3mul(R, Ib, T); mat3mulT(T, R, J);
        for (int a = 0; a < 4; a++) {
            mat3mul(Ra[a], Ib, T2); mat3mulT(T2, R, T3);
            double T4[9]; mat3mulT(T, Ra[a], T4);
            for (int i = 0; i < 9; i++) Ja[a][i] = T3[i] + T4[i];
            for (int b = 0; b < 4; b++) {
                double A1[9], A2[9], A3[9], A4[9], t[9];
                mat3mul(Rab[a][b], Ib, t); mat3mulT(t, R, A1);
                mat3mulT(T2, Ra[b], A2);                          /* Ra I Rb^T */
                mat3mul(Ra[b], Ib, t); mat3mulT(t, Ra[a], A3);    /* Rb I Ra^T */
                mat3mulT(T, Rab[a][b], A4);                       /* R I Rab^T */
                for (int i = 0; i < 9; i++) Jab[a][b][i] = A1[i] + A2[i] + A3[i] + A4[i];
            }
        }
    }
}

/* ------------------------------------------------------------------ SRBD rigid body accelerations
 * rddot = sum f_i / (m/fs) + [0,0,-g]          (kin_dyn.fSRBD, prb.py:99)
 * wdot  = J^-1 ( sum (c_i - r) x f_i  -  w x J w )
 * z = [r(3) o(4) c0..c3(12) w(3) f0..f3(12)] are the variables wdot depends on. */
typedef struct {
    double wd[3], rdd[3];
    double Jac[3][NZ];
    double Hc[NZ][NZ]; /* Hessian of lambda^T wdot at lambda = wdot (curvature term of ||wdot||^2) */
} RB;
enum { Z_R = 0, Z_O = 3, Z_C = 7, Z_W = 19, Z_F = 22 };

/* tail: a node of the LIP-style tail (include/sddp.h lip_tail_start; isrbd_example.py:344-353): no rotational dynamics,
 * wdot = 0 identically, so its Jacobian and curvature vanish; rddot is unchanged. */
static void srbd_rb(const OrcConfig *c, const double *x, const double *u, int order, int tail, RB *rb) {
    const double *r = x + SX_R, *o = x + SX_O, *w = x + SX_W;
    double J[9], Ja[4][9], Jab[4][4][9], M[9];
    inertia_all(c, o, J, Ja, Jab);
    inv3(J, M);
    double tau[3] = {0, 0, 0}, fsum[3] = {0, 0, 0};
    for (int i = 0; i < 4; i++) {
        const double *ci = x + SX_C + 3 * i, *fi = u + SU_F(i);
        double d[3] = {ci[0] - r[0], ci[1] - r[1], ci[2] - r[2]}, t[3];
        cross(d, fi, t);
        for (int k = 0; k < 3; k++) { tau[k] += t[k]; fsum[k] += fi[k]; }
    }
    double Jw[3], wJw[3], h[3];
    mat3vec(J, w, Jw); cross(w, Jw, wJw);
    for (int k = 0; k < 3; k++) h[k] = tau[k] - wJw[k];
    mat3vec(M, h, rb->wd);
    double ms = c->mass / c->force_scaling;
    rb->rdd[0] = fsum[0] / ms; rb->rdd[1] = fsum[1] / ms; rb->rdd[2] = fsum[2] / ms - c->gravity;
    if (tail) {
        rb->wd[0] = rb->wd[1] = rb->wd[2] = 0.0;
        if (order >= 1) memset(rb->Jac, 0, sizeof rb->Jac);
        if (order >= 2) memset(rb->Hc, 0, sizeof rb->Hc);
        return;
    }
    if (order < 1) return;

    /* first derivatives: Jac[:,p] = M (dh/dp - J_p wd) */
    double dh[3][NZ];
    memset(dh, 0, sizeof dh);
    double S[9];
    for (int i = 0; i < 4; i++) {
        const double *ci = x + SX_C + 3 * i, *fi = u + SU_F(i);
        skew(fi, S);
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
            dh[a][Z_R + b] += S[3 * a + b];           /* d tau / d r   = sum skew(f_i) */
            dh[a][Z_C + 3 * i + b] = -S[3 * a + b];   /* d tau / d c_i = -skew(f_i)    */
        }
        double d[3] = {ci[0] - r[0], ci[1] - r[1], ci[2] - r[2]};
        skew(d, S);
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) dh[a][Z_F + 3 * i + b] = S[3 * a + b];
    }
    {   /* d(-w x Jw)/dw = skew(Jw) - skew(w) J */
        double SJw[9], Sw[9], SwJ[9];
        skew(Jw, SJw); skew(w, Sw); mat3mul(Sw, J, SwJ);
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) dh[a][Z_W + b] = SJw[3 * a + b] - SwJ[3 * a + b];
    }
    for (int a = 0; a < 4; a++) { /* d/do_a: -w x (J_a w) - J_a wd */
        double Jaw[3], t[3], Jawd[3];
        mat3vec(Ja[a], w, Jaw); cross(w, Jaw, t); mat3vec(Ja[a], rb->wd, Jawd);
        for (int k = 0; k < 3; k++) dh[k][Z_O + a] = -t[k] - Jawd[k];
    }
    for (int p = 0; p < NZ; p++) {
        double col[3] = {dh[0][p], dh[1][p], dh[2][p]}, out[3];
        mat3vec(M, col, out);
        for (int k = 0; k < 3; k++) rb->Jac[k][p] = out[k];
    }
    if (order < 2) return;

    /* curvature: phi = lambda^T wdot, lambda = wd fixed, nu = M lambda.
     * phi_pq = nu^T ( h_pq - J_p wd_q - J_q wd_p - J_pq wd ) */
    double nu[3];
    mat3vec(M, rb->wd, nu);
    memset(rb->Hc, 0, sizeof rb->Hc);
    double Sn[9];
    skew(nu, Sn);
    for (int i = 0; i < 4; i++)
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
            rb->Hc[Z_C + 3 * i + a][Z_F + 3 * i + b] += -Sn[3 * a + b];
            rb->Hc[Z_F + 3 * i + b][Z_C + 3 * i + a] += -Sn[3 * a + b];
            rb->Hc[Z_R + a][Z_F + 3 * i + b] += Sn[3 * a + b];
            rb->Hc[Z_F + 3 * i + b][Z_R + a] += Sn[3 * a + b];
        }
    {   /* (w,w): skew(nu) J - J skew(nu) */
        double A[9], Bm[9];
        mat3mul(Sn, J, A); mat3mul(J, Sn, Bm);
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) rb->Hc[Z_W + a][Z_W + b] += A[3 * a + b] - Bm[3 * a + b];
    }
    for (int a = 0; a < 4; a++) {
        /* (w,o_a): (skew(nu) J_a - J_a skew(nu)) w */
        double A[9], Bm[9], D[9], v[3];
        mat3mul(Sn, Ja[a], A); mat3mul(Ja[a], Sn, Bm);
        for (int i = 0; i < 9; i++) D[i] = A[i] - Bm[i];
        mat3vec(D, w, v);
        for (int k = 0; k < 3; k++) { rb->Hc[Z_W + k][Z_O + a] += v[k]; rb->Hc[Z_O + a][Z_W + k] += v[k]; }
        /* -(J_a nu) . Jac[:,q]  on row o_a and, symmetrically, column o_a */
        double Jan[3];
        mat3vec(Ja[a], nu, Jan);
        for (int q = 0; q < NZ; q++) {
            double s = Jan[0] * rb->Jac[0][q] + Jan[1] * rb->Jac[1][q] + Jan[2] * rb->Jac[2][q];
            rb->Hc[Z_O + a][q] -= s;
            rb->Hc[q][Z_O + a] -= s;
        }
        for (int b = 0; b < 4; b++) {
            /* (o_a,o_b): w^T skew(nu) J_ab w  -  nu^T J_ab wd */
            double t[3], t2[3], t3[3];
            mat3vec(Jab[a][b], w, t); mat3vec(Sn, t, t2);
            mat3vec(Jab[a][b], rb->wd, t3);
            rb->Hc[Z_O + a][Z_O + b] += dot3(w, t2) - dot3(nu, t3);
        }
    }
}

/* ------------------------------------------------------------------ dynamics */
static void srbd_ode(const OrcConfig *c, int tail, const double *x, const double *u, double *xd) {
    RB rb;
    srbd_rb(c, x, u, 0, tail, &rb);
    const double *o = x + SX_O, *w = x + SX_W;
    for (int k = 0; k < 3; k++) xd[SX_R + k] = x[SX_RD + k];
    /* odot = quat_prod([w/2, 0], o)  (LOCAL_WORLD_ALIGNED, prb.py:107-108) */
    double wxo[3];
    cross(w, o, wxo);
    for (int k = 0; k < 3; k++) xd[SX_O + k] = 0.5 * (o[3] * w[k] + wxo[k]);
    xd[SX_O + 3] = -0.5 * dot3(w, o);
    for (int k = 0; k < 12; k++) xd[SX_C + k] = x[SX_CD + k];
    for (int k = 0; k < 3; k++) { xd[SX_RD + k] = rb.rdd[k]; xd[SX_W + k] = rb.wd[k]; }
    for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) xd[SX_CD + 3 * i + k] = u[SU_CDD(i) + k];
}
static void lip_ode(const OrcConfig *c, const double *x, const double *u, double *xd) {
    for (int k = 0; k < 3; k++) xd[LX_R + k] = x[LX_RD + k];
    for (int k = 0; k < 12; k++) xd[LX_C + k] = x[LX_CD + k];
    for (int k = 0; k < 3; k++) xd[LX_RD + k] = c->eta2 * (x[LX_R + k] - u[LU_Z + k]);
    xd[LX_RD + 2] -= c->gravity;
    for (int k = 0; k < 12; k++) xd[LX_CD + k] = u[LU_CDD(0) + k];
}
void orc_dynamics_kind(const OrcConfig *c, int kind, const double *x, const double *u, double *xn) {
    int nx, nu, np; orc_dims(c->model, &nx, &nu, &np);
    double xd[NXM];
    if (c->model == 0) srbd_ode(c, kind == ORC_NODE_TAIL, x, u, xd); else lip_ode(c, x, u, xd);
    for (int i = 0; i < nx; i++) xn[i] = x[i] + c->dt * xd[i];
}
void orc_dynamics(const OrcConfig *c, const double *x, const double *u, double *xn) { orc_dynamics_kind(c, ORC_NODE_MID, x, u, xn); }

/* ------------------------------------------------------------------ cost accumulator
 * every residual of prb.py except min_qddot's wdot rows is affine in z=[x;u]:
 *   res = sum_i coef_i z[idx_i] + const ;  cost += wgt * res^2                       */
typedef struct {
    int nx, nu, derivs;
    double cost;
    double *lx, *lu, *lxx, *lux, *luu;
} Acc;
static void acc_affine(Acc *a, double wgt, int n, const int *idx, const double *coef, double res) {
    a->cost += wgt * res * res;
    if (!a->derivs) return;
    int nx = a->nx;
    for (int i = 0; i < n; i++) {
        double g = 2.0 * wgt * res * coef[i];
        if (idx[i] < nx) a->lx[idx[i]] += g; else a->lu[idx[i] - nx] += g;
        for (int j = 0; j < n; j++) {
            double hh = 2.0 * wgt * coef[i] * coef[j];
            int zi = idx[i], zj = idx[j];
            if (zi < nx && zj < nx) a->lxx[zi * nx + zj] += hh;
            else if (zi >= nx && zj >= nx) a->luu[(zi - nx) * a->nu + (zj - nx)] += hh;
            else if (zi >= nx && zj < nx) a->lux[(zi - nx) * nx + zj] += hh;
        }
    }
}
static void acc_single(Acc *a, double wgt, int idx, double res) { /* res = z[idx] + const */
    double one = 1.0;
    acc_affine(a, wgt, 1, &idx, &one, res);
}

/* exponential barrier on one component (ddp.py:204-209): cost += wgt * exp(kap * g),  g = sgn * z[idx] + const */
static void acc_barrier(Acc *a, double wgt, double kap, int idx, double sgn, double g) {
    double e = wgt * exp(kap * g);
    a->cost += e;
    if (!a->derivs) return;
    int nx = a->nx;
    if (idx < nx) { a->lx[idx] += kap * sgn * e; a->lxx[idx * nx + idx] += kap * kap * e; }
    else { a->lu[idx - nx] += kap * sgn * e; a->luu[(idx - nx) * a->nu + (idx - nx)] += kap * kap * e; }
}

static void srbd_cost(const OrcConfig *c, int kind, const double *x, const double *u, const double *p, Acc *a) {
    const int nx = 37;
    int track = (kind != ORC_NODE_FIRST), input = (kind != ORC_NODE_TERM), tail = (kind == ORC_NODE_TAIL);
    if (track) {
        /* rz_tracking, prb.py:184 */
        acc_single(a, c->r_tracking_gain, SX_R + 2, x[SX_R + 2] - c->com[2]);
        /* o_tracking_xyz / o_tracking_w, prb.py:185-189: otg * (quat_prod(o, oref) - [0,0,0,1]) */
        const double *o = x + SX_O, *q = p + SP_OREF;
        double otg = p[SP_OTG];
        double E[4][4] = {{q[3], q[2], -q[1], q[0]},
                          {-q[2], q[3], q[0], q[1]},
                          {q[1], -q[0], q[3], q[2]},
                          {-q[0], -q[1], -q[2], q[3]}};
        for (int i = 0; i < 4; i++) {
            int idx[4] = {SX_O, SX_O + 1, SX_O + 2, SX_O + 3};
            double res = E[i][0] * o[0] + E[i][1] * o[1] + E[i][2] * o[2] + E[i][3] * o[3] - (i == 3 ? 1.0 : 0.0);
            acc_affine(a, otg * otg, 4, idx, E[i], res);
        }
        /* rdot_tracking, w_tracking, prb.py:190-191 */
        for (int k = 0; k < 3; k++) {
            acc_single(a, c->rdot_tracking_gain, SX_RD + k, x[SX_RD + k] - p[SP_RDREF + k]);
            acc_single(a, c->w_tracking_gain, SX_W + k, x[SX_W + k] - p[SP_WREF + k]);
        }
        /* rel_pos_{y,x}_1_4 (c0,c2) and _3_6 (c1,c3), prb.py:153-154, 192-199 */
        for (int pair = 0; pair < 2; pair++) for (int ax = 0; ax < 2; ax++) {
            int ia = SX_C + 3 * pair + ax, ib = SX_C + 3 * (pair + 2) + ax;
            double d = -(c->foot[3 * pair + ax] - c->foot[3 * (pair + 2) + ax]);
            int idx[2] = {ia, ib};
            double coef[2] = {-1.0, 1.0};
            acc_affine(a, c->rel_position_gain, 2, idx, coef, -x[ia] + x[ib] - d);
        }
    }
    if (input) {
        double fs = c->force_scaling;
        /* min_qddot = [rddot; wdot; cddot_i], prb.py:104-106,200 */
        RB rb;
        srbd_rb(c, x, u, a->derivs ? (c->hessian_mode == 0 ? 2 : 1) : 0, tail, &rb);
        double gq = c->min_qddot_gain, ms = c->mass / fs;
        for (int k = 0; k < 3; k++) {
            int idx[4]; double coef[4];
            for (int i = 0; i < 4; i++) { idx[i] = nx + SU_F(i) + k; coef[i] = 1.0 / ms; }
            acc_affine(a, gq, 4, idx, coef, rb.rdd[k]);
        }
        for (int k = 0; k < 3; k++) a->cost += gq * rb.wd[k] * rb.wd[k];
        if (a->derivs) {
            int zmap[NZ];
            for (int i = 0; i < 3; i++) { zmap[Z_R + i] = SX_R + i; zmap[Z_W + i] = SX_W + i; }
            for (int i = 0; i < 4; i++) zmap[Z_O + i] = SX_O + i;
            for (int i = 0; i < 12; i++) zmap[Z_C + i] = SX_C + i;
            for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) zmap[Z_F + 3 * i + k] = nx + SU_F(i) + k;
            for (int pi = 0; pi < NZ; pi++) {
                double g = 0;
                for (int k = 0; k < 3; k++) g += rb.Jac[k][pi] * rb.wd[k];
                g *= 2.0 * gq;
                int zi = zmap[pi];
                if (zi < nx) a->lx[zi] += g; else a->lu[zi - nx] += g;
                for (int qi = 0; qi < NZ; qi++) {
                    double hh = 0;
                    for (int k = 0; k < 3; k++) hh += rb.Jac[k][pi] * rb.Jac[k][qi];
                    if (c->hessian_mode == 0) hh += rb.Hc[pi][qi];
                    hh *= 2.0 * gq;
                    int zj = zmap[qi];
                    if (zi < nx && zj < nx) a->lxx[zi * nx + zj] += hh;
                    else if (zi >= nx && zj >= nx) a->luu[(zi - nx) * a->nu + (zj - nx)] += hh;
                    else if (zi >= nx && zj < nx) a->lux[(zi - nx) * nx + zj] += hh;
                }
            }
        }
        for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) {
            acc_single(a, gq, nx + SU_CDD(i) + k, u[SU_CDD(i) + k]);
            /* min_f_i and f_i_active, prb.py:201-204 */
            double sw = p[SP_SW(i)];
            acc_single(a, fs * fs * c->min_f_gain, nx + SU_F(i) + k, u[SU_F(i) + k]);
            acc_single(a, fs * fs * c->force_switch_weight * (1.0 - sw) * (1.0 - sw), nx + SU_F(i) + k, u[SU_F(i) + k]);
        }
        /* Friction cone as an exponential barrier (extension, off by default): the reference builds the linearised cone
         * (prb.py:173-177) and drops it; its adapter sketches  cost += exp_parameter * exp(g)  for g <= 0
         * (ddp.py:197-203).  Here  L += w sum_rows exp(kappa g_row(f_i)),  g = A f,
         * A = [1 0 -mu; -1 0 -mu; 0 1 -mu; 0 -1 -mu; 0 0 -1]  (stance frame = world frame, prb.py:175). */
        if (c->friction_cone_weight != 0.0) {
            const double w = c->friction_cone_weight, kap = c->friction_cone_sharpness, fm = c->friction_cone_mu;
            const double A[5][3] = {{1, 0, -fm}, {-1, 0, -fm}, {0, 1, -fm}, {0, -1, -fm}, {0, 0, -1}};
            for (int i = 0; i < 4; i++) {
                const double *f = u + SU_F(i);
                for (int r = 0; r < 5; r++) {
                    double g = A[r][0] * f[0] + A[r][1] * f[1] + A[r][2] * f[2];
                    double e = w * exp(kap * g);
                    a->cost += e;
                    if (a->derivs) {
                        for (int k = 0; k < 3; k++) {
                            a->lu[SU_F(i) + k] += kap * e * A[r][k];
                            for (int l = 0; l < 3; l++)
                                a->luu[(SU_F(i) + k) * a->nu + SU_F(i) + l] += kap * kap * e * A[r][k] * A[r][l];
                        }
                    }
                }
            }
        }
        /* Bounds as exponential barriers (extension, off by default; ddp.py:204-209 sketches them for variable bounds):
         * force box (isrbd_example.py:200), unilaterality f_z >= 0 (isrbd_example.py:198), contact-point velocity box (:195). */
        {
            const double kb = c->bound_sharpness;
            for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) {
                if (c->force_bound_weight != 0.0) {
                    acc_barrier(a, c->force_bound_weight, kb, nx + SU_F(i) + k, 1.0, u[SU_F(i) + k] - c->force_bound);
                    acc_barrier(a, c->force_bound_weight, kb, nx + SU_F(i) + k, -1.0, -c->force_bound - u[SU_F(i) + k]);
                }
                if (c->cdot_bound_weight != 0.0) {
                    acc_barrier(a, c->cdot_bound_weight, kb, SX_CD + 3 * i + k, 1.0, x[SX_CD + 3 * i + k] - c->cdot_bound);
                    acc_barrier(a, c->cdot_bound_weight, kb, SX_CD + 3 * i + k, -1.0, -c->cdot_bound - x[SX_CD + 3 * i + k]);
                }
            }
            if (c->unilateral_weight != 0.0)
                for (int i = 0; i < 4; i++) acc_barrier(a, c->unilateral_weight, kb, nx + SU_F(i) + 2, -1.0, -u[SU_F(i) + 2]);
        }
        /* equality constraints, weight 1e6 (ddp.py:181,191-196); prb.py:166-181 */
        double cw = c->constraint_weight;
        if (tail) {     /* LIP-style tail: lip_zero_angular_momentum (w) and lip_com_height (r_z - com_z), isrbd_example.py:352-353 */
            for (int k = 0; k < 3; k++) acc_single(a, cw, SX_W + k, x[SX_W + k]);
            acc_single(a, cw, SX_R + 2, x[SX_R + 2] - c->com[2]);
        }
        for (int leg = 0; leg < 2; leg++) for (int ax = 0; ax < 2; ax++) { /* relative_vel_left_1 / right_3 */
            int ia = SX_CD + 3 * (2 * leg) + ax, ib = SX_CD + 3 * (2 * leg + 1) + ax;
            int idx[2] = {ia, ib};
            double coef[2] = {1.0, -1.0};
            acc_affine(a, cw, 2, idx, coef, x[ia] - x[ib]);
        }
        for (int i = 0; i < 4; i++) {
            acc_single(a, cw, SX_C + 3 * i + 2, x[SX_C + 3 * i + 2] - p[SP_CREF(i)]);      /* cz_tracking */
            double sw = p[SP_SW(i)];
            for (int ax = 0; ax < 2; ax++) {                                               /* cdotxy_tracking */
                int id = SX_CD + 3 * i + ax;
                acc_affine(a, cw, 1, &id, &sw, sw * x[id]);
            }
        }
    }
}

static void lip_cost(const OrcConfig *c, int kind, const double *x, const double *u, const double *p, Acc *a) {
    const int nx = 30;
    int track = (kind != ORC_NODE_FIRST), input = (kind != ORC_NODE_TERM);
    if (track) {
        acc_single(a, c->r_tracking_gain, LX_R + 2, x[LX_R + 2] - c->com[2]);                 /* prb.py:390 */
        for (int ax = 0; ax < 2; ax++) {                                                      /* rxy_tracking :391 */
            int idx[5] = {LX_R + ax, LX_C + ax, LX_C + 3 + ax, LX_C + 6 + ax, LX_C + 9 + ax};
            double coef[5] = {1.0, -0.25, -0.25, -0.25, -0.25};
            double res = x[idx[0]] - 0.25 * (x[idx[1]] + x[idx[2]] + x[idx[3]] + x[idx[4]]);
            acc_affine(a, c->r_tracking_gain, 5, idx, coef, res);
        }
        for (int k = 0; k < 3; k++)                                                           /* rdot_tracking :392 */
            acc_single(a, c->rdot_tracking_gain, LX_RD + k, x[LX_RD + k] - p[LP_RDREF + k]);
        for (int pair = 0; pair < 2; pair++) for (int ax = 0; ax < 2; ax++) {                 /* rel_pos :394-401 */
            int ia = LX_C + 3 * pair + ax, ib = LX_C + 3 * (pair + 2) + ax;
            double d = -(c->foot[3 * pair + ax] - c->foot[3 * (pair + 2) + ax]);
            int idx[2] = {ia, ib};
            double coef[2] = {-1.0, 1.0};
            acc_affine(a, c->rel_position_gain, 2, idx, coef, -x[ia] + x[ib] - d);
        }
    }
    if (input) {
        for (int k = 0; k < 3; k++) {                                                         /* zmp_tracking :393 */
            int idx[5] = {nx + LU_Z + k, LX_C + k, LX_C + 3 + k, LX_C + 6 + k, LX_C + 9 + k};
            double coef[5] = {1.0, -0.25, -0.25, -0.25, -0.25};
            double res = u[LU_Z + k] - 0.25 * (x[idx[1]] + x[idx[2]] + x[idx[3]] + x[idx[4]]);
            acc_affine(a, c->zmp_tracking_gain, 5, idx, coef, res);
        }
        for (int k = 0; k < 3; k++) {                                                         /* min_qddot: rddot :402 */
            int idx[2] = {LX_R + k, nx + LU_Z + k};
            double coef[2] = {c->eta2, -c->eta2};
            double res = c->eta2 * (x[LX_R + k] - u[LU_Z + k]) - (k == 2 ? c->gravity : 0.0);
            acc_affine(a, c->min_qddot_gain, 2, idx, coef, res);
        }
        for (int k = 0; k < 12; k++) acc_single(a, c->min_qddot_gain, nx + LU_CDD(0) + k, u[LU_CDD(0) + k]);
        double cw = c->constraint_weight;                                                     /* :379-387 */
        for (int leg = 0; leg < 2; leg++) for (int ax = 0; ax < 2; ax++) {
            int ia = LX_CD + 3 * (2 * leg) + ax, ib = LX_CD + 3 * (2 * leg + 1) + ax;
            int idx[2] = {ia, ib};
            double coef[2] = {1.0, -1.0};
            acc_affine(a, cw, 2, idx, coef, x[ia] - x[ib]);
        }
        for (int i = 0; i < 4; i++) {
            acc_single(a, cw, LX_C + 3 * i + 2, x[LX_C + 3 * i + 2] - p[LP_CREF(i)]);
            double sw = p[LP_SW(i)];
            for (int ax = 0; ax < 2; ax++) {
                int id = LX_CD + 3 * i + ax;
                acc_affine(a, cw, 1, &id, &sw, sw * x[id]);
            }
        }
    }
}

double orc_cost(const OrcConfig *c, int kind, const double *x, const double *u, const double *p) {
    Acc a;
    int np;
    orc_dims(c->model, &a.nx, &a.nu, &np);
    a.derivs = 0; a.cost = 0;
    a.lx = a.lu = a.lxx = a.lux = a.luu = 0;
    if (c->model == 0) srbd_cost(c, kind, x, u, p, &a); else lip_cost(c, kind, x, u, p, &a);
    return a.cost;
}

void orc_derivs(const OrcConfig *c, int kind, const double *x, const double *u, const double *p,
                double *fx, double *fu, double *lx, double *lu, double *lxx, double *lux, double *luu) {
    int nx, nu, np;
    orc_dims(c->model, &nx, &nu, &np);
    Acc a;
    a.nx = nx; a.nu = nu; a.derivs = 1; a.cost = 0;
    a.lx = lx; a.lu = lu; a.lxx = lxx; a.lux = lux; a.luu = luu;
    memset(lx, 0, sizeof(double) * nx); memset(lu, 0, sizeof(double) * nu);
    memset(lxx, 0, sizeof(double) * nx * nx); memset(lux, 0, sizeof(double) * nu * nx);
    memset(luu, 0, sizeof(double) * nu * nu);
    if (c->model == 0) srbd_cost(c, kind, x, u, p, &a); else lip_cost(c, kind, x, u, p, &a);
    if (kind == ORC_NODE_TERM || !fx) return;

    double dt = c->dt;
    memset(fx, 0, sizeof(double) * nx * nx); memset(fu, 0, sizeof(double) * nx * nu);
    for (int i = 0; i < nx; i++) fx[i * nx + i] = 1.0;
    if (c->model == 1) {
        for (int k = 0; k < 3; k++) {
            fx[(LX_R + k) * nx + LX_RD + k] += dt;
            fx[(LX_RD + k) * nx + LX_R + k] += dt * c->eta2;
            fu[(LX_RD + k) * nu + LU_Z + k] += -dt * c->eta2;
        }
        for (int k = 0; k < 12; k++) {
            fx[(LX_C + k) * nx + LX_CD + k] += dt;
            fu[(LX_CD + k) * nu + LU_CDD(0) + k] += dt;
        }
        return;
    }
    RB rb;
    srbd_rb(c, x, u, 1, kind == ORC_NODE_TAIL, &rb);
    const double *o = x + SX_O, *w = x + SX_W;
    for (int k = 0; k < 3; k++) fx[(SX_R + k) * nx + SX_RD + k] += dt;
    for (int k = 0; k < 12; k++) fx[(SX_C + k) * nx + SX_CD + k] += dt;
    {   /* odot_v = (o_w w + w x o_v)/2,  odot_w = -(w . o_v)/2 */
        double Sw[9], So[9];
        skew(w, Sw); skew(o, So);
        for (int a_ = 0; a_ < 3; a_++) {
            for (int b = 0; b < 3; b++) {
                fx[(SX_O + a_) * nx + SX_O + b] += dt * 0.5 * Sw[3 * a_ + b];
                fx[(SX_O + a_) * nx + SX_W + b] += dt * 0.5 * ((a_ == b ? o[3] : 0.0) - So[3 * a_ + b]);
            }
            fx[(SX_O + a_) * nx + SX_O + 3] += dt * 0.5 * w[a_];
            fx[(SX_O + 3) * nx + SX_O + a_] += -dt * 0.5 * w[a_];
            fx[(SX_O + 3) * nx + SX_W + a_] += -dt * 0.5 * o[a_];
        }
    }
    double ms = c->mass / c->force_scaling;
    for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) {
        fu[(SX_RD + k) * nu + SU_F(i) + k] += dt / ms;
        fu[(SX_CD + 3 * i + k) * nu + SU_CDD(i) + k] += dt;
    }
    for (int a_ = 0; a_ < 3; a_++) {
        for (int b = 0; b < 3; b++) {
            fx[(SX_W + a_) * nx + SX_R + b] += dt * rb.Jac[a_][Z_R + b];
            fx[(SX_W + a_) * nx + SX_W + b] += dt * rb.Jac[a_][Z_W + b];
        }
        for (int b = 0; b < 4; b++) fx[(SX_W + a_) * nx + SX_O + b] += dt * rb.Jac[a_][Z_O + b];
        for (int b = 0; b < 12; b++) fx[(SX_W + a_) * nx + SX_C + b] += dt * rb.Jac[a_][Z_C + b];
        for (int i = 0; i < 4; i++) for (int b = 0; b < 3; b++)
            fu[(SX_W + a_) * nu + SU_F(i) + b] += dt * rb.Jac[a_][Z_F + 3 * i + b];
    }
}

/* ------------------------------------------------------------------ DDP */
static int node_kind_c(const OrcConfig *c, int k) {
    if (k == 0) return ORC_NODE_FIRST;
    if (k == c->N) return ORC_NODE_TERM;
    return (c->model == 0 && c->lip_tail_start > 0 && k >= c->lip_tail_start) ? ORC_NODE_TAIL : ORC_NODE_MID;
}
#define node_kind(k, N) node_kind_c(c, (k))

double orc_total_cost(const OrcConfig *c, const double *X, const double *U, const double *params) {
    int nx, nu, np; orc_dims(c->model, &nx, &nu, &np);
    double J = 0;
    for (int k = 0; k < c->N; k++) J += orc_cost(c, node_kind(k, c->N), X + k * nx, U + k * nu, params + k * np);
    J += orc_cost(c, ORC_NODE_TERM, X + c->N * nx, 0, params + c->N * np);
    return J;
}

/* Cholesky of the n x n SPD matrix A (row-major, lower factor in place); 0 on success */
static int chol(double *A, int n) {
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
        if (!(d > 0.0) || !isfinite(d)) return j + 1;
        d = sqrt(d);
        A[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = s / d;
        }
    }
    return 0;
}
static void chol_solve(const double *L, int n, double *b) { /* b <- (L L^T)^-1 b */
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L[i * n + k] * b[k];
        b[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= L[k * n + i] * b[k];
        b[i] = s / L[i * n + i];
    }
}

int orc_backward(const OrcConfig *c, const double *X, const double *U, const double *params,
                 const double *defect, double mu, double *K, double *kff, double *dV) {
    int nx, nu, np; orc_dims(c->model, &nx, &nu, &np);
    const int N = c->N;
    const int fixed_rho = (c->defect_contraction_rate > 0.0);
    const double rho_b = fixed_rho ? c->defect_contraction_rate : 1.0;
    double Vx[NXM], Vxx[NXM * NXM], y[NXM];
    double fx[NXM * NXM], fu[NXM * NUM], lx[NXM], lu[NUM], lxx[NXM * NXM], lux[NUM * NXM], luu[NUM * NUM];
    double Qx[NXM], Qu[NUM], Qxx[NXM * NXM], Qux[NUM * NXM], Quu[NUM * NUM], L[NUM * NUM];
    double T[NXM * NXM], Tu[NXM * NUM], vp[NXM], cg[NXM], s[NXM], ys[NXM], col[NUM];
    double tot = 0, acc1 = 0, acc2 = 0; /* acc1: D1 (rho=alpha) or C0 (fixed); acc2: 1/2 sum k^T Quu k */

    orc_derivs(c, ORC_NODE_TERM, X + N * nx, 0, params + N * np, 0, 0, Vx, lu, Vxx, lux, luu);
    memcpy(y, Vx, sizeof(double) * nx);
    for (int k = N - 1; k >= 0; k--) {
        orc_derivs(c, node_kind(k, N), X + k * nx, U + k * nu, params + k * np, fx, fu, lx, lu, lxx, lux, luu);
        for (int i = 0; i < nx; i++) cg[i] = rho_b * defect[k * nx + i];
        /* s = Vxx' c, v+ = Vx' + s, gap terms of the model */
        double g1 = 0, g2 = 0, yg = 0;
        for (int i = 0; i < nx; i++) {
            double t = 0;
            for (int j = 0; j < nx; j++) t += Vxx[i * nx + j] * cg[j];
            s[i] = t; vp[i] = Vx[i] + t;
            g1 += Vx[i] * cg[i]; g2 += cg[i] * t; yg += y[i] * cg[i];
        }
        tot += g1 + 0.5 * g2;
        for (int i = 0; i < nx; i++) ys[i] = fixed_rho ? y[i] + s[i] : y[i];
        if (fixed_rho) acc1 += yg + 0.5 * g2;
        /* T = Vxx' fx, Tu = Vxx' fu   (loops ordered i-l-j: unit-stride inner loop, same summation order over l) */
        for (int i = 0; i < nx; i++) {
            for (int j = 0; j < nx; j++) T[i * nx + j] = 0.0;
            for (int j = 0; j < nu; j++) Tu[i * nu + j] = 0.0;
            for (int l = 0; l < nx; l++) {
                const double v = Vxx[i * nx + l];
                for (int j = 0; j < nx; j++) T[i * nx + j] += v * fx[l * nx + j];
                for (int j = 0; j < nu; j++) Tu[i * nu + j] += v * fu[l * nu + j];
            }
        }
        for (int i = 0; i < nx; i++) {
            double t = lx[i];
            for (int l = 0; l < nx; l++) t += fx[l * nx + i] * vp[l];
            Qx[i] = t;
            for (int j = 0; j < nx; j++) Qxx[i * nx + j] = lxx[i * nx + j];
            for (int l = 0; l < nx; l++) {
                const double v = fx[l * nx + i];
                for (int j = 0; j < nx; j++) Qxx[i * nx + j] += v * T[l * nx + j];
            }
        }
        for (int i = 0; i < nu; i++) {
            double t = lu[i];
            for (int l = 0; l < nx; l++) t += fu[l * nu + i] * vp[l];
            Qu[i] = t;
            for (int j = 0; j < nx; j++) Qux[i * nx + j] = lux[i * nx + j];
            for (int j = 0; j < nu; j++) Quu[i * nu + j] = luu[i * nu + j];
            for (int l = 0; l < nx; l++) {
                const double v = fu[l * nu + i];
                for (int j = 0; j < nx; j++) Qux[i * nx + j] += v * T[l * nx + j];
                for (int j = 0; j < nu; j++) Quu[i * nu + j] += v * Tu[l * nu + j];
            }
        }
        for (int i = 0; i < nu; i++) for (int j = 0; j < i; j++) { /* symmetrise Quu */
            double m_ = 0.5 * (Quu[i * nu + j] + Quu[j * nu + i]);
            Quu[i * nu + j] = Quu[j * nu + i] = m_;
        }
        memcpy(L, Quu, sizeof(double) * nu * nu);
        for (int i = 0; i < nu; i++) L[i * nu + i] += mu;
        if (chol(L, nu)) return k + 1;
        double *Kk = K + (size_t)k * nu * nx, *kk = kff + (size_t)k * nu;
        for (int i = 0; i < nu; i++) col[i] = -Qu[i];
        chol_solve(L, nu, col);
        memcpy(kk, col, sizeof(double) * nu);
        for (int j = 0; j < nx; j++) {
            for (int i = 0; i < nu; i++) col[i] = -Qux[i * nx + j];
            chol_solve(L, nu, col);
            for (int i = 0; i < nu; i++) Kk[i * nx + j] = col[i];
        }
        /* model terms */
        double Quuk[NUM], quk = 0, kQk = 0;
        for (int i = 0; i < nu; i++) {
            double t = 0;
            for (int j = 0; j < nu; j++) t += Quu[i * nu + j] * kk[j];
            Quuk[i] = t; quk += Qu[i] * kk[i]; kQk += kk[i] * t;
        }
        tot += quk + 0.5 * kQk;
        acc2 += 0.5 * kQk;
        /* y recursion: qu_y = lu + fu^T ys,  qx_y = lx + fx^T ys */
        double quy[NUM], qxy[NXM];
        for (int i = 0; i < nu; i++) { double t = lu[i]; for (int l = 0; l < nx; l++) t += fu[l * nu + i] * ys[l]; quy[i] = t; }
        for (int i = 0; i < nx; i++) { double t = lx[i]; for (int l = 0; l < nx; l++) t += fx[l * nx + i] * ys[l]; qxy[i] = t; }
        if (!fixed_rho) { double t = yg; for (int i = 0; i < nu; i++) t += quy[i] * kk[i]; acc1 += t; }
        for (int i = 0; i < nx; i++) { double t = qxy[i]; for (int l = 0; l < nu; l++) t += Kk[l * nx + i] * quy[l]; y[i] = t; }
        /* value function */
        for (int i = 0; i < nx; i++) {
            double t = Qx[i];
            for (int l = 0; l < nu; l++) t += Kk[l * nx + i] * (Quuk[l] + Qu[l]) + Qux[l * nx + i] * kk[l];
            Vx[i] = t;
        }
        /* Tu(reuse as QuuK: nu x nx) */
        double *QuuK = Tu;
        for (int i = 0; i < nu; i++) {
            for (int j = 0; j < nx; j++) QuuK[i * nx + j] = 0.0;
            for (int l = 0; l < nu; l++) {
                const double v = Quu[i * nu + l];
                for (int j = 0; j < nx; j++) QuuK[i * nx + j] += v * Kk[l * nx + j];
            }
        }
        for (int i = 0; i < nx; i++) {
            for (int j = 0; j < nx; j++) T[i * nx + j] = Qxx[i * nx + j];
            for (int l = 0; l < nu; l++) {
                const double ki = Kk[l * nx + i], qi = Qux[l * nx + i];
                for (int j = 0; j < nx; j++) T[i * nx + j] += ki * (QuuK[l * nx + j] + Qux[l * nx + j]) + qi * Kk[l * nx + j];
            }
        }
        for (int i = 0; i < nx; i++) for (int j = 0; j < nx; j++) Vxx[i * nx + j] = 0.5 * (T[i * nx + j] + T[j * nx + i]);
    }
    if (fixed_rho) { dV[2] = acc1; dV[1] = acc2; dV[0] = tot - acc1 - acc2; }
    else           { dV[2] = 0.0;  dV[0] = acc1; dV[1] = tot - acc1; }
    return 0;
}

double orc_forward(const OrcConfig *c, const double *x0, const double *X, const double *U, const double *params,
                   const double *defect, const double *K, const double *kff,
                   double alpha, double rho, double *Xn, double *Un) {
    int nx, nu, np; orc_dims(c->model, &nx, &nu, &np);
    const int N = c->N;
    double J = 0;
    memcpy(Xn, x0, sizeof(double) * nx);
    for (int k = 0; k < N; k++) {
        const double *Kk = K + (size_t)k * nu * nx, *kk = kff + (size_t)k * nu;
        double *xk = Xn + k * nx, *uk = Un + k * nu;
        for (int i = 0; i < nu; i++) {
            double t = 0;
            for (int j = 0; j < nx; j++) t += Kk[i * nx + j] * (xk[j] - X[k * nx + j]);
            uk[i] = U[k * nu + i] + alpha * kk[i] + t;
        }
        J += orc_cost(c, node_kind(k, N), xk, uk, params + k * np);
        orc_dynamics_kind(c, node_kind(k, N), xk, uk, Xn + (k + 1) * nx);
        for (int i = 0; i < nx; i++) Xn[(k + 1) * nx + i] -= (1.0 - rho) * defect[k * nx + i];
    }
    J += orc_cost(c, ORC_NODE_TERM, Xn + N * nx, 0, params + N * np);
    return J;
}

static double max_abs(const double *v, int n) {
    double m = 0;
    for (int i = 0; i < n; i++) { double a = fabs(v[i]); if (a > m || a != a) m = a; }
    return m;
}

int orc_solve(const OrcConfig *c, const double *x0, const double *params,
              double *X, double *U, double *K, double *kff, double *hist, int *iters, double *cost) {
    int nx, nu, np; orc_dims(c->model, &nx, &nu, &np);
    const int N = c->N;
    double *Xn = (double *)malloc(sizeof(double) * (N + 1) * nx);
    double *Un = (double *)malloc(sizeof(double) * N * nu);
    double *d = (double *)calloc((size_t)N * nx, sizeof(double));
    memset(hist, 0, sizeof(double) * c->max_iters * ORC_HIST);
    memcpy(X, x0, sizeof(double) * nx);
    if (!c->multiple_shooting) {
        for (int k = 0; k < N; k++) orc_dynamics_kind(c, node_kind(k, N), X + k * nx, U + k * nu, X + (k + 1) * nx);
    } else {
        for (int k = 0; k < N; k++) {
            orc_dynamics_kind(c, node_kind(k, N), X + k * nx, U + k * nu, d + k * nx);
            for (int i = 0; i < nx; i++) d[k * nx + i] -= X[(k + 1) * nx + i];
        }
    }
    double J = orc_total_cost(c, X, U, params);
    double mu = c->mu0;
    int status = ORC_MAX_ITERS, it = 0;
    const int fixed_rho = (c->defect_contraction_rate > 0.0);
    int bad_start = 0;     /* a non-finite initial gap (NaN / inf in the warm start or x0): status NAN, no iteration */
    for (int i = 0; i < N * nx; i++) if (!(fabs(d[i]) <= 1.79e308)) bad_start = 1;
    if (bad_start) status = ORC_NAN;
    for (it = 0; it < (bad_start ? 0 : c->max_iters); it++) {
        double dV[3];
        int reg_fail = 0;
        while (orc_backward(c, X, U, params, d, mu, K, kff, dV)) {
            mu = fmax(mu * c->mu_factor, c->mu_min);
            if (mu > c->mu_max) { reg_fail = 1; break; }
        }
        double dmax = max_abs(d, N * nx);
        double *h = hist + it * ORC_HIST;
        h[0] = J; h[1] = 0.0; h[2] = mu; h[3] = dmax;
        if (reg_fail) { status = ORC_REG_FAILED; it++; break; }
        if (!isfinite(dV[0]) || !isfinite(dV[1]) || !isfinite(J)) { status = ORC_NAN; it++; break; }
        double a0 = c->alpha_0;
        double dJ0 = dV[2] + a0 * dV[0] + a0 * a0 * dV[1];
        if (fabs(dJ0) <= 1e-3 * c->cost_reduction_ths * (1.0 + fabs(J)) && dmax <= c->defect_ths) {
            status = ORC_OK; it++; break;
        }
        int accepted = 0;
        double Jn = J, alpha;
        for (alpha = a0; alpha >= c->alpha_converge_threshold; alpha *= c->line_search_decrease_factor) {
            double rho = fixed_rho ? c->defect_contraction_rate : alpha;
            Jn = orc_forward(c, x0, X, U, params, d, K, kff, alpha, rho, Xn, Un);
            double dJm = dV[2] + alpha * dV[0] + alpha * alpha * dV[1];
            if (isfinite(Jn) && Jn - J <= dJm + (1.0 - c->beta) * fabs(dJm)) { accepted = 1; break; }
        }
        if (accepted) {
            double rho = fixed_rho ? c->defect_contraction_rate : alpha;
            memcpy(X, Xn, sizeof(double) * (N + 1) * nx);
            memcpy(U, Un, sizeof(double) * N * nu);
            for (int i = 0; i < N * nx; i++) d[i] *= (1.0 - rho);
            dmax = max_abs(d, N * nx);
            double dJ = J - Jn;
            J = Jn;
            h[0] = J; h[1] = alpha; h[3] = dmax;
            mu = mu / c->mu_factor;
            if (mu < c->mu_min) mu = 0.0;
            if (mu < c->mu0) mu = c->mu0;
            if (dJ <= c->cost_reduction_ths * (1.0 + fabs(J)) && dmax <= c->defect_ths) { status = ORC_OK; it++; break; }
        } else {
            mu = fmax(mu * c->mu_factor, c->mu_min);
            if (mu > c->mu_max) { status = ORC_LS_FAILED; it++; break; }
        }
    }
    *iters = it;
    *cost = J;
    free(Xn); free(Un); free(d);
    return status;
}

typedef struct {
    const OrcConfig *c; int B; const double *x0, *params;
    double *X, *U, *K, *kff, *hist; int *iters, *status; double *cost;
    int next; pthread_mutex_t lock;
} BatchJob;

static void *batch_worker(void *arg) {
    BatchJob *j = (BatchJob *)arg;
    const OrcConfig *c = j->c;
    int nx, nu, np; orc_dims(c->model, &nx, &nu, &np);
    const int N = c->N;
    for (;;) {
        pthread_mutex_lock(&j->lock);
        int b = j->next++;
        pthread_mutex_unlock(&j->lock);
        if (b >= j->B) break;
        j->status[b] = orc_solve(c, j->x0 + (size_t)b * nx, j->params + (size_t)b * (N + 1) * np,
                                 j->X + (size_t)b * (N + 1) * nx, j->U + (size_t)b * N * nu,
                                 j->K + (size_t)b * N * nu * nx, j->kff + (size_t)b * N * nu,
                                 j->hist + (size_t)b * c->max_iters * ORC_HIST, j->iters + b, j->cost + b);
    }
    return 0;
}

void orc_solve_batch(const OrcConfig *c, int B, const double *x0, const double *params,
                     double *X, double *U, double *K, double *kff,
                     double *hist, int *iters, int *status, double *cost, int nthreads) {
    qt_init();
    BatchJob j = {c, B, x0, params, X, U, K, kff, hist, iters, status, cost, 0, PTHREAD_MUTEX_INITIALIZER};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], 0, batch_worker, &j);
    batch_worker(&j);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], 0);
}
