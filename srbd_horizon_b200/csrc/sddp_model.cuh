// sddp_model.cuh -- analytic dynamics / cost / derivatives of the two srbd_horizon problems
// as sm_100a device code (north_star stage one).
//
// What is implemented (file:line under /root/reference/python):
//   SRBD ode    prb.py:97-109    rdot, quaternion kinematics (world-aligned w), cdot_i,
//                                rddot = sum f_i/(m/fs) - g e_z,
//                                wdot  = J(o)^-1 (sum (c_i - r) x f_i - w x J(o) w),
//                                J = R o (I/fs) o R^T element-wise as written at prb.py:99 ("literal")
//                                or R (I/fs) R^T ("rotated", README.md:2), cddot_i
//   LIP ode     prb.py:315-329
//   L_k, L_N    ddp.py:179-226 over the residuals prb.py:184-204 / 390-402 and the equality
//               constraints prb.py:166-181 / 379-387 (weight 1e6, ddp.py:181)
//   f_k         ddp.py:228-230 explicit Euler
//
// Organisation (one model = one struct of static device functions):
//   accel      the non-elementwise part of the ode (SRBD: wdot, rddot), uniform per node
//   xdot_i     i-th component of the ode given accel        -> lanes own state components
//   cost_lane  this lane's share of L (lanes own residuals) -> warp-shuffle sum
//   pack       single-thread "rigid-body pack": wdot, its Jacobian wrt the 34 variables it
//              depends on, and the sparse curvature blocks of lambda^T wdot -- evaluated for all
//              nodes of a problem in parallel, one thread per node
//   expand     block-parallel: writes lx,lu,lxx,lux,luu of a node into the Q buffers from the pack
//   expand_f   block-parallel: dense fx, fu (generic path / stage-one kernel)
#pragma once
#include <math.h>
#include "../../include/sddp.h"

// Scalar type of the build.  `real` is the storage type of every array (the ABI's sddp_real) and the arithmetic type of the
// Riccati recursion, the derivatives and the rollout: double in the product library libsddp.so, float in the optional
// -DSDDP_F32 build (libsddp_f32.so; north_star: "an optional fp32 build matches within a stated tolerance").  What stays
// `double` in both builds, spelled out where it happens: the cost of a trajectory (cost_lane, the line-search sums and the
// acceptance / convergence tests: they resolve 1e-6 relative), the expected-decrease accumulators of the backward pass and
// the accumulators of the tensor-core products (mma.sync m8n8k4 f64: there is no fp32 tensor-core path that keeps 24 bits).
typedef sddp_real real;
#ifdef SDDP_F32
typedef float2 real2;
__device__ __forceinline__ real2 make_real2(real a, real b) { return make_float2(a, b); }
#else
typedef double2 real2;
__device__ __forceinline__ real2 make_real2(real a, real b) { return make_double2(a, b); }
#endif
constexpr int RV = 16 / (int)sizeof(real);      // elements per 16-byte vector (cp.async.cg 16, bulk copies)

// Phase timer (developer builds only, -DSDDP_PROFILE): thread 0 of CTA 0 accumulates clock64() deltas per phase
// into g_prof[]; read back with sddp_debug_profile().  Phases are delimited by block barriers.
#ifdef SDDP_PROFILE
__device__ long long g_prof[32];
__device__ long long g_prof_last;
#define PROF_DECL
#define PROF(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long t_ = clock64(); g_prof[i] += t_ - g_prof_last; g_prof_last = t_; } } while (0)
#define PROF_RESET do { if (threadIdx.x == 0 && blockIdx.x == 0) g_prof_last = clock64(); } while (0)
#define PROF_T(i, thr) do { if (threadIdx.x == (thr) && blockIdx.x == 0) { g_prof[i] += clock64() - g_prof_last; } } while (0)
#define PROF_INC(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_prof[i] += 1; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_RESET
#define PROF_T(i, thr)
#define PROF_INC(i)
#endif



// Per-warp timeline of one backward node of CTA 0 (developer builds only, -DSDDP_STAMP): lane 0 of every warp stores
// clock64() at each stamp (no read-modify-write, so the probes do not stall the warps); sddp_debug_stamps() reads them.
#ifdef SDDP_STAMP
__device__ long long g_stamp[16 * 4];
#define STAMP(id) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && k == 10) g_stamp[(id) * 4 + (threadIdx.x >> 5)] = clock64(); } while (0)
#else
#define STAMP(id)
#endif

// The inequality barriers (include/sddp.h: friction cone, force / velocity boxes, unilaterality) are off unless one of
// their weights is > 0.  Their exp() code sits in a second instantiation of the SRBD model (SrbdT<true>) and of every
// kernel that uses it: merely compiled into the default kernels it costs them 14 % (register pressure: ptxas then
// serialises the shared-memory loads of the factorisation through one register quad), so the default instantiation
// (SrbdT<false>) has none of it.  One library, the variant is picked from the configuration at sddp_create.
#define SDDP_CONE_ON(c) (HAS_INEQ && (c).ineq != 0)

struct DevCfg {
    int model, N, inertia_mode, hessian_mode, ms, max_iters;
    real dt, mscaled, inv_ms, Ib[9], com[3], foot[12], fs, g, eta2;
    real w_r, w_rdot, w_w, w_rel, w_fsw, gq, w_minf, w_zmp, cw;
    real w_cone, cone_mu, cone_k;   // friction-cone barrier (include/sddp.h); w_cone = 0: off
    real w_fb, fb, w_uni, w_cdb, cdb, kb;   // force box, unilaterality, contact-point velocity box (include/sddp.h); weights 0: off
    int ineq;                         // any inequality barrier on
    int lip_tail;                     // first node of the LIP-style tail, 0 = off (include/sddp.h lip_tail_start)
    real drel[2][2];   // drel[pair][axis] = -(foot[pair][axis] - foot[pair+2][axis])   prb.py:153-154
    double alpha0, alpha_min, ls_factor, beta, cost_ths, mu0, rho_fixed, mu_min, mu_max, mu_factor, defect_ths;   // solver scalars: double in both builds
    const unsigned long long* ztab;   // SRBD: descriptors of the 595 upper-triangular entries of the wdot Hessian block
};

// z-block descriptor bit layout (built on the host, sddp.cu:build_ztab)
enum { ZT_ROUNDS = 5, ZT_THREADS = 128, ZT_NXX_NUX = 517, ZT_LAZY_THREADS = 96 };   // 595 entries: 253 xx + 264 ux first, then 78 uu
// Quu work table of the structured backward pass: ztab[ZT_C1OFF + tid] = four 16-bit descriptors type | i1 << 2 | i2 << 6
enum { ZT_C1OFF = 640, ZT_COFF = 768, ZT_CROUNDS = 2, ZT_AOFF = 768 + 2 * 96, ZT_NAFF = 44, ZT_TOTAL = 768 + 2 * 96 + 48 };   // ZT_AOFF: dst | value index << 12 of the affine Hessian entries (Srbd::apply_rec)   // ZT_COFF: the xx / ux entries that have a curvature term
#define C1_TYPE(d) (int)((d) & 3)
#define C1_I1(d) (int)(((d) >> 2) & 15)
#define C1_I2(d) (int)(((d) >> 6) & 15)
#define ZT_VALID(d) ((d) & 1ull)
#define ZT_KIND(d) (int)(((d) >> 1) & 3)
#define ZT_DA(d) (int)(((d) >> 3) & 63)
#define ZT_DB(d) (int)(((d) >> 9) & 63)
#define ZT_PI(d) (int)(((d) >> 15) & 63)
#define ZT_QI(d) (int)(((d) >> 21) & 63)
#define ZT_HSIGN(d) (int)(((d) >> 27) & 3)
#define ZT_HOFF(d) (int)(((d) >> 29) & 511)
#define ZT_DST1(d) (int)(((d) >> 38) & 4095)
#define ZT_DST2(d) (int)(((d) >> 50) & 4095)
enum { ZT_QUX_OFF = 37 * 37 + 1, ZT_LDUX = 44 };   // layout of SmemSrbd (static_assert there)

enum { NODE_FIRST = 0, NODE_MID = 1, NODE_TERM = 2, NODE_TAIL = 3 };   // TAIL: a MID node of the LIP-style tail (no rotational dynamics)

#define SDDP_DEV __device__ __forceinline__
#ifndef SDDP_E_UNROLL
#define SDDP_E_UNROLL 1
#endif
#ifndef SDDP_NOINLINE
#define SDDP_NOINLINE
#endif

namespace m3 {
SDDP_DEV void cross(const real* a, const real* b, real* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
SDDP_DEV real dot(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
SDDP_DEV void mv(const real* A, const real* v, real* o) {
    o[0] = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
    o[1] = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
    o[2] = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
}
SDDP_DEV void mm(const real* A, const real* B, real* C) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
SDDP_DEV void mmT(const real* A, const real* B, real* C) {   // A * B^T
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
SDDP_DEV void inv(const real* A, real* M) {
    real c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    real id = real(1.0) / (A[0] * c00 + A[1] * c01 + A[2] * c02);
    M[0] = c00 * id; M[1] = (A[2] * A[7] - A[1] * A[8]) * id; M[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    M[3] = c01 * id; M[4] = (A[0] * A[8] - A[2] * A[6]) * id; M[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    M[6] = c02 * id; M[7] = (A[1] * A[6] - A[0] * A[7]) * id; M[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}
// skew(v)[a][b]
SDDP_DEV real skew_ab(const real* v, int a, int b) {
    if (a == b) return 0.0;
    int k = 3 - a - b;                       // the remaining axis
    real s = ((b - a + 3) % 3 == 1) ? -real(1.0) : real(1.0);   // (0,1)->-v2, (1,2)->-v0, (2,0)->-v1
    return s * v[k];
}
}  // namespace m3

// rotation of a (non-normalised) quaternion (x,y,z,w): horizon utils.toRot as used at prb.py:97
SDDP_DEV void quat_R(const real* q, real* R) {
    real x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = real(1.0) - real(2.0) * (y * y + z * z); R[1] = real(2.0) * (x * y - z * w); R[2] = real(2.0) * (x * z + y * w);
    R[3] = real(2.0) * (x * y + z * w); R[4] = real(1.0) - real(2.0) * (x * x + z * z); R[5] = real(2.0) * (y * z - x * w);
    R[6] = real(2.0) * (x * z - y * w); R[7] = real(2.0) * (y * z + x * w); R[8] = real(1.0) - real(2.0) * (x * x + y * y);
}
// dR/dq_a (linear in q, so d2R/dq_a dq_b = quat_dR(e_b, a))
SDDP_DEV void quat_dR(const real* q, int a, real* D) {
    real x = real(2.0) * q[0], y = real(2.0) * q[1], z = real(2.0) * q[2], w = real(2.0) * q[3];
    switch (a) {
    case 0: D[0] = 0; D[1] = y; D[2] = z; D[3] = y; D[4] = -2 * x; D[5] = -w; D[6] = z; D[7] = w; D[8] = -2 * x; break;
    case 1: D[0] = -2 * y; D[1] = x; D[2] = w; D[3] = x; D[4] = 0; D[5] = z; D[6] = -w; D[7] = z; D[8] = -2 * y; break;
    case 2: D[0] = -2 * z; D[1] = -w; D[2] = x; D[3] = w; D[4] = -2 * z; D[5] = y; D[6] = x; D[7] = y; D[8] = 0; break;
    default: D[0] = 0; D[1] = -z; D[2] = y; D[3] = z; D[4] = 0; D[5] = -x; D[6] = -y; D[7] = x; D[8] = 0; break;
    }
}

// =====================================================================================  SRBD
template <bool INEQ>
struct SrbdT {
    static constexpr bool HAS_INEQ = INEQ;
    static constexpr int NX = 37, NU = 24, NP = 19, NACC = 6, PACK = 256, NZ = 34;
    // x: r[0:3] o[3:7] c_i[7+3i] rdot[19:22] w[22:25] cdot_i[25+3i];  u: cddot_i[6i] f_i[6i+3]
    // p: rdot_ref[0:3] w_ref[3:6] otg[6] (c_ref_i, sw_i)[7+2i, 8+2i] oref[15:19]
    enum { XR = 0, XO = 3, XC = 7, XRD = 19, XW = 22, XCD = 25 };
    enum { ZR = 0, ZO = 3, ZC = 7, ZW = 19, ZF = 22 };
    enum { PK_WD = 0, PK_RDD = 3, PK_NU = 6, PK_HWW = 9, PK_JAC = 18, PK_HO = 120 };
    // indices of the affine Hessian values (aff_value) that the table at ZT_AOFF / the compact curvature descriptors name
    enum { AF_CD = 0, AF_RELP = 12, AF_RELN = 13, AF_NCW = 14, AF_CW = 15, AF_RDOT = 16, AF_RZ = 17, AF_WW = 18, AF_OO = 19, AF_N = 29 };

    SDDP_DEV static void inertia(const DevCfg& c, const real* R, real* J) {
        if (c.inertia_mode == 0) {
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) J[3 * i + j] = R[3 * i + j] * R[3 * j + i] * c.Ib[3 * i + j];
        } else {
            real T[9];
            m3::mm(R, c.Ib, T);
            m3::mmT(T, R, J);
        }
    }
    // d/dq_a of the inertia given R, Ra
    SDDP_DEV static void inertia_d(const DevCfg& c, const real* R, const real* Ra, real* Ja) {
        if (c.inertia_mode == 0) {
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++)
                    Ja[3 * i + j] = (Ra[3 * i + j] * R[3 * j + i] + R[3 * i + j] * Ra[3 * j + i]) * c.Ib[3 * i + j];
        } else {
            real T[9], A[9], B[9];
            m3::mm(Ra, c.Ib, T); m3::mmT(T, R, A);
            m3::mm(R, c.Ib, T);  m3::mmT(T, Ra, B);
#pragma unroll
            for (int i = 0; i < 9; i++) Ja[i] = A[i] + B[i];
        }
    }
    // d2/dq_a dq_b
    SDDP_DEV static void inertia_dd(const DevCfg& c, const real* R, const real* Ra, const real* Rb, const real* Rab, real* Jab) {
        if (c.inertia_mode == 0) {
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    int ij = 3 * i + j, ji = 3 * j + i;
                    Jab[ij] = (Rab[ij] * R[ji] + Ra[ij] * Rb[ji] + Rb[ij] * Ra[ji] + R[ij] * Rab[ji]) * c.Ib[ij];
                }
        } else {
            real T[9], A[9];
            m3::mm(Rab, c.Ib, T); m3::mmT(T, R, Jab);
            m3::mm(Ra, c.Ib, T);  m3::mmT(T, Rb, A);
#pragma unroll
            for (int i = 0; i < 9; i++) Jab[i] += A[i];
            m3::mm(Rb, c.Ib, T);  m3::mmT(T, Ra, A);
#pragma unroll
            for (int i = 0; i < 9; i++) Jab[i] += A[i];
            m3::mm(R, c.Ib, T);   m3::mmT(T, Rab, A);
#pragma unroll
            for (int i = 0; i < 9; i++) Jab[i] += A[i];
        }
    }

    // acc[0:3] = wdot, acc[3:6] = rddot      (kin_dyn.fSRBD as called at prb.py:99)
    // The part of the accelerations that depends on the state only: pre[0:9] = I_w^-1, pre[9:12] = w x I_w w.
    // (The forward pass computes it for x^_k on an idle warp while u^_k is still being formed.)
    static constexpr int NPRE = 12;
    SDDP_DEV static void accel_pre(const DevCfg& c, const real* x, real* pre) {
        real R[9], J[9];
        quat_R(x + XO, R);
        inertia(c, R, J);
        m3::inv(J, pre);
        real Jw[3];
        m3::mv(J, x + XW, Jw);
        m3::cross(x + XW, Jw, pre + 9);
    }
    SDDP_DEV static void accel_post(const DevCfg& c, const real* x, const real* u, const real* pre, real* acc, bool tail = false) {
        real tau[3] = {0, 0, 0}, fsum[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const real* ci = x + XC + 3 * i;
            const real* fi = u + 6 * i + 3;
            real d[3] = {ci[0] - x[0], ci[1] - x[1], ci[2] - x[2]}, t[3];
            m3::cross(d, fi, t);
            tau[0] += t[0]; tau[1] += t[1]; tau[2] += t[2];
            fsum[0] += fi[0]; fsum[1] += fi[1]; fsum[2] += fi[2];
        }
        real h[3] = {tau[0] - pre[9], tau[1] - pre[10], tau[2] - pre[11]};
        m3::mv(pre, h, acc);
        if (tail) { acc[0] = 0.0; acc[1] = 0.0; acc[2] = 0.0; }      // LIP-style tail: no rotational dynamics
        acc[3] = fsum[0] * c.inv_ms; acc[4] = fsum[1] * c.inv_ms; acc[5] = fsum[2] * c.inv_ms - c.g;
    }
    SDDP_DEV static void accel(const DevCfg& c, const real* x, const real* u, real* acc, bool tail = false) {
        real pre[NPRE];
        accel_pre(c, x, pre);
        accel_post(c, x, u, pre, acc, tail);
    }
    // Branch free: lanes of a warp own different components, and a divergent branch per component kind costs more
    // than the work.  Every component but the quaternion rate is a copy from one of the arrays.
    SDDP_DEV static real xdot_i(const DevCfg& c, int i, const real* x, const real* u, const real* acc) {
        const real* src = x;                                                              // rdot, cdot_j
        int idx = (i < 3) ? XRD + i : XCD + (i - XC);
        if (i >= XRD) { src = acc; idx = (i < XW) ? i - XRD + 3 : i - XW; }                 // rddot, wdot
        if (i >= XCD) { src = u; const int e = i - XCD; idx = e + 3 * (e / 3); }            // cddot_j = u[6 j + k]
        const real v = src[idx];
        // odot = quat_prod([w/2, 0], o)  (world-aligned angular velocity, prb.py:107-108)
        const real* o = x + XO;
        const real* w = x + XW;
        const int a = min(max(i - XO, 0), 3), a3 = min(a, 2);
        const int b2 = a3 == 2 ? 0 : a3 + 1, d = a3 == 0 ? 2 : a3 - 1;
        const real qv = (a == 3) ? -real(0.5) * (w[0] * o[0] + w[1] * o[1] + w[2] * o[2]) : real(0.5) * (o[3] * w[a3] + (w[b2] * o[d] - w[d] * o[b2]));
        return (i >= XO && i < XC) ? qv : v;
    }

    // quaternion error rows: quat_prod(o, oref) = E(oref) o        (prb.py:187)
    SDDP_DEV static real E_row(const real* q, int i, int a) {
        // E = [[q3, q2, -q1, q0], [-q2, q3, q0, q1], [q1, -q0, q3, q2], [-q0, -q1, -q2, q3]]
        if (i == a) return q[3];
        if (a == 3) return q[i];
        if (i == 3) return -q[a];
        return -m3::skew_ab(q, i, a);
    }

    // Inequality barriers of one foot (include/sddp.h, extensions, off by default): exponential barrier on the linearised
    // friction cone, on the force box |f_k| <= fb and on f_z >= 0.  Value, gradient (3) and Hessian (xx, xy, xz, yy, yz, zz)
    // with respect to the foot's force f.
    SDDP_DEV static void cone_terms(const DevCfg& c, const real* f, double& val, double* g, double* h) {
        val = 0.0;
        g[0] = g[1] = g[2] = 0.0;
        h[0] = h[1] = h[2] = h[3] = h[4] = h[5] = 0.0;
        if (c.w_cone != 0.0) {
            const double k = c.cone_k, m = c.cone_mu, kz = k * m * (double)f[2];
            const double e1 = c.w_cone * exp(k * (double)f[0] - kz), e2 = c.w_cone * exp(-k * (double)f[0] - kz);
            const double e3 = c.w_cone * exp(k * (double)f[1] - kz), e4 = c.w_cone * exp(-k * (double)f[1] - kz), e5 = c.w_cone * exp(-k * (double)f[2]);
            const double s4 = (e1 + e2) + (e3 + e4);
            val = s4 + e5;
            g[0] = k * (e1 - e2); g[1] = k * (e3 - e4); g[2] = -k * (m * s4 + e5);
            h[0] = k * k * (e1 + e2); h[2] = -k * k * m * (e1 - e2);
            h[3] = k * k * (e3 + e4); h[4] = -k * k * m * (e3 - e4); h[5] = k * k * (m * m * s4 + e5);
        }
        if (c.w_fb != 0.0) {
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const double ep = c.w_fb * exp(c.kb * ((double)f[a] - c.fb)), em = c.w_fb * exp(c.kb * (-c.fb - (double)f[a]));
                val += ep + em;
                g[a] += c.kb * (ep - em);
                h[a == 0 ? 0 : (a == 1 ? 3 : 5)] += c.kb * c.kb * (ep + em);
            }
        }
        if (c.w_uni != 0.0) {
            const double e = c.w_uni * exp(-c.kb * (double)f[2]);
            val += e; g[2] -= c.kb * e; h[5] += c.kb * c.kb * e;
        }
    }
    // Barrier on the contact-point velocity box |v| <= cdb: value, first and second derivative.
    __device__ __noinline__ static void cdot_box_cold(const DevCfg& c, double v, double& val, double& g, double& h) {
        const double ep = c.w_cdb * exp(c.kb * (v - c.cdb)), em = c.w_cdb * exp(c.kb * (-c.cdb - v));
        val = ep + em; g = c.kb * (ep - em); h = c.kb * c.kb * (ep + em);
    }
    // Out-of-line copy for the call sites inside hot loops (forward-pass cost, node expansion): with the barrier off they
    // cost one predicated call instead of five inlined exp().
#ifdef SDDP_CONE_INLINE
    SDDP_DEV static void cone_terms_cold(const DevCfg& c, const real* f, double& val, double* g, double* h) { cone_terms(c, f, val, g, h); }
#else
    __device__ __noinline__ static void cone_terms_cold(const DevCfg& c, const real* f, double& val, double* g, double* h) {
        cone_terms(c, f, val, g, h);
    }
#endif
    SDDP_DEV static int cone_hidx(int a, int b) {      // index of (a, b) in (xx, xy, xz, yy, yz, zz)
        const int lo = min(a, b), hi = max(a, b);
        return lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
    }

    // parts: 1 = terms indexed by the state (8: those of the contact-point velocities, split off for load balance),
    //        2 = terms indexed by the input, 4 = rddot / wdot terms (need accel)
    SDDP_DEV static double cost_lane(const DevCfg& c, int kind, int lane, const real* x, const real* u, const real* p, const real* acc,
                                     int parts = 15) {
        double s = 0.0;
        if (parts & 9) {
            const bool track = kind != NODE_FIRST, input = kind != NODE_TERM;
            for (int i = lane; i < NX; i += 32) {
                if (!(parts & (i < XCD ? 1 : 8))) continue;
                if (i == 2 || (i >= XRD && i < XCD)) {                     // rz / rdot / w tracking (prb.py:184,190-191)
                    if (track) {
                        const double ref = (i == 2) ? c.com[2] : (double)p[i - XRD];            // rdot_ref = p[0:3], w_ref = p[3:6]
                        const double wg = (i == 2) ? c.w_r : (i < XW ? c.w_rdot : c.w_w);
                        const double r = (double)x[i] - ref;
                        s += wg * r * r;
                    }
                    if (kind == NODE_TAIL && (i == 2 || i >= XW)) {        // lip_com_height, lip_zero_angular_momentum (isrbd_example.py:352-353)
                        const double r = (i == 2) ? (double)x[2] - c.com[2] : (double)x[i];
                        s += c.cw * r * r;
                    }
                } else if (i >= XO && i < XC) {                            // otg * (quat_prod(o, oref) - [0,0,0,1])  (prb.py:185-189)
                    if (track) {
                        const real* q = p + 15;
                        const real* o = x + XO;
                        const int a = i - XO;
                        double r;
                        if (a == 3) r = (double)o[3] * (double)q[3] - ((double)o[0] * (double)q[0] + (double)o[1] * (double)q[1] + (double)o[2] * (double)q[2]) - 1.0;
                        else {
                            const int b2 = a == 2 ? 0 : a + 1, d = a == 0 ? 2 : a - 1;
                            r = (double)o[3] * (double)q[a] + (double)q[3] * (double)o[a] + ((double)o[b2] * (double)q[d] - (double)o[d] * (double)q[b2]);
                        }
                        s += (double)p[6] * (double)p[6] * r * r;
                    }
                } else if (i >= XC) {                                      // contact points and their velocities
                    const bool vel = i >= XCD;
                    const int e = i - (vel ? XCD : XC), foot = e / 3, ax = e - 3 * foot;
                    if (SDDP_CONE_ON(c) && c.w_cdb != 0.0 && vel && input) {      // contact-point velocity box (extension)
                        double val, g_, h_;
                        cdot_box_cold(c, (double)x[i], val, g_, h_);
                        s += val;
                    }
                    if (!vel && ax == 2) {                                 // cz_tracking (prb.py:180)
                        if (input) { const double r = (double)x[i] - (double)p[7 + 2 * foot]; s += c.cw * r * r; }
                    } else if (ax < 2) {
                        if (vel && input) { const double r = (double)p[8 + 2 * foot] * (double)x[i]; s += c.cw * r * r; }      // cdotxy_tracking (:181)
                        // pair terms: rel_pos (c_j, c_j+2), j < 2 (prb.py:192-199) / relative_vel (cdot_0,1), (cdot_2,3) (:166-170)
                        const bool pair = vel ? ((foot & 1) == 0 && input) : (foot < 2 && track);
                        if (pair) {
                            const double r = (double)x[i + (vel ? 3 : 6)] - (double)x[i] - (vel ? 0.0 : c.drel[foot][ax]);
                            s += (vel ? c.cw : c.w_rel) * r * r;
                        }
                    }
                }
            }
        }
        if (kind != NODE_TERM) {        // nodes 0..N-1: prb.py:200-204
            if ((parts & 2) && lane < NU) {
                int i = lane / 6, r = lane % 6;
                double v = (double)u[lane];
                if (r < 3) s += c.gq * v * v;
                else { double a = 1.0 - (double)p[8 + 2 * i]; s += (c.w_minf + c.w_fsw * a * a) * v * v; }
            }
            if ((parts & 4) && lane < 3) s += c.gq * ((double)acc[lane] * (double)acc[lane] + (double)acc[3 + lane] * (double)acc[3 + lane]);
            if ((parts & 2) && SDDP_CONE_ON(c) && lane < 4) {      // inequality barriers of foot `lane`
                double val, g[3], h[6];
                cone_terms_cold(c, u + 6 * lane + 3, val, g, h);
                s += val;
            }
        }
        return s;
    }

    // ---- single-thread rigid-body pack (see file header) ------------------------------------
    // pk[PK_WD..] wd(3) rdd(3) nu(3) Hww(9) Jac[3][34] Ho[4][34]
    // One thread per node, and nothing is read back from memory: every per-thread array is indexed statically
    // (registers), each Jacobian column is turned into its four curvature entries Ho[a][q] = -(J_a nu) . Jac[:,q] the
    // moment it exists, and the only values that cross a rolled loop -- the ten second-derivative terms of the (o, o)
    // block -- go through the shared-memory scratch sc[e * ss] (ss = threads sharing it; the solve kernel's shared
    // memory is idle between the forward and the backward pass).  (Round 1 kept Ra[4][9], Ja[4][9] in local memory,
    // 1.1 K local loads per pack thrashing L1 with 200 packs in flight per SM, and re-read the Jacobian from global
    // memory for the Ho rows: the pack phase was 6-9 % of the solve.)
    // With v x e_b = (0, v2, -v1), (-v2, 0, v0), (v1, -v0, 0): the three columns M (v x e_b) of a lever arm or force.
    template <bool HO>
    SDDP_DEV static void put3(real* Jac, const real* M, const real (*Jan)[3], int z0, real v0, real v1, real v2) {
        real col[3][3];      // col[k][b]
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const real m0 = M[3 * k], m1 = M[3 * k + 1], m2 = M[3 * k + 2];
            col[k][0] = m1 * v2 - m2 * v1;
            col[k][1] = m2 * v0 - m0 * v2;
            col[k][2] = m0 * v1 - m1 * v0;
#pragma unroll
            for (int b = 0; b < 3; b++) Jac[k * NZ + z0 + b] = col[k][b];
        }
        if (HO) {
            real* Ho = Jac + (PK_HO - PK_JAC);
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) Ho[a * NZ + z0 + b] = -(Jan[a][0] * col[0][b] + Jan[a][1] * col[1][b] + Jan[a][2] * col[2][b]);
        }
    }
    template <bool EXACT>
    SDDP_DEV static void pack_impl(const DevCfg& c, const real* x, const real* u, real* pk, real* sc, int ss) {
        const real o[4] = {x[XO], x[XO + 1], x[XO + 2], x[XO + 3]};
        const real r[3] = {x[0], x[1], x[2]}, w[3] = {x[XW], x[XW + 1], x[XW + 2]};
        real R[9], J[9], M[9];
        quat_R(o, R);
        inertia(c, R, J);
        m3::inv(J, M);
        real* Jac = pk + PK_JAC;   // [3][34]: Jac[:,p] = M (dh/dp - J_p wd), h = tau - w x J w
        real* Ho = pk + PK_HO;     // [4][34]
        real tau[3] = {0, 0, 0}, F[3] = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const real* ci = x + XC + 3 * i;
            const real* fi = u + 6 * i + 3;
            const real f[3] = {fi[0], fi[1], fi[2]}, d[3] = {ci[0] - r[0], ci[1] - r[1], ci[2] - r[2]};
            real t[3];
            m3::cross(d, f, t);
#pragma unroll
            for (int k = 0; k < 3; k++) { tau[k] += t[k]; F[k] += f[k]; }
        }
        real Jw[3], wJw[3], wd[3];
        m3::mv(J, w, Jw);
        m3::cross(w, Jw, wJw);
        const real h[3] = {tau[0] - wJw[0], tau[1] - wJw[1], tau[2] - wJw[2]};
        m3::mv(M, h, wd);
#pragma unroll
        for (int k = 0; k < 3; k++) pk[PK_WD + k] = wd[k];
        pk[PK_RDD + 0] = F[0] * c.inv_ms; pk[PK_RDD + 1] = F[1] * c.inv_ms; pk[PK_RDD + 2] = F[2] * c.inv_ms - c.g;
        // curvature of lambda^T wdot at lambda = wd:  nu = M wd,  phi_pq = nu^T ( h_pq - J_p wd_q - J_q wd_p - J_pq wd )
        real nu[3], wxn[3], nxw[3];
        m3::mv(M, wd, nu);
        m3::cross(w, nu, wxn);       // w^T skew(nu) A w = (w x nu) . (A w)
        m3::cross(nu, w, nxw);
        if (EXACT) {
#pragma unroll
            for (int k = 0; k < 3; k++) pk[PK_NU + k] = nu[k];
            // second derivatives of the inertia inside the (o, o) block: v_ab = (w x nu) . (J_ab w) - nu . (J_ab wd), a <= b
            // (rolled: ten pairs through one copy of inertia_dd; results cross to the unrolled code below through sc)
            int e = 0;
#pragma unroll 1
            for (int a = 0; a < 4; a++) {
                real Ra[9];
                quat_dR(o, a, Ra);
#pragma unroll 1
                for (int b = a; b < 4; b++, e++) {
                    real eb[4], Rb[9], Rab[9], Jab[9], t3[3], t4[3];
#pragma unroll
                    for (int i = 0; i < 4; i++) eb[i] = (i == b) ? real(1.0) : 0.0;
                    quat_dR(o, b, Rb);
                    quat_dR(eb, a, Rab);
                    inertia_dd(c, R, Ra, Rb, Rab, Jab);
                    m3::mv(Jab, w, t3); m3::mv(Jab, wd, t4);
                    sc[e * ss] = m3::dot(wxn, t3) - m3::dot(nu, t4);
                }
            }
        }
        // d(-w x Jw)/dw_b = Jw x e_b - w x J[:,b]
        real colW[3][3];      // colW[k][b] = Jac[k][ZW + b]
        {
            real cw[9];      // cw[3 k + b]
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const real Jcol[3] = {J[b], J[3 + b], J[6 + b]};
                real t[3];
                m3::cross(w, Jcol, t);
#pragma unroll
                for (int k = 0; k < 3; k++) cw[3 * k + b] = -t[k];
            }
            cw[3 * 1 + 0] += Jw[2]; cw[3 * 2 + 0] -= Jw[1];
            cw[3 * 0 + 1] -= Jw[2]; cw[3 * 2 + 1] += Jw[0];
            cw[3 * 0 + 2] += Jw[1]; cw[3 * 1 + 2] -= Jw[0];
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const real col[3] = {cw[b], cw[3 + b], cw[6 + b]};
                real out[3];
                m3::mv(M, col, out);
#pragma unroll
                for (int k = 0; k < 3; k++) { colW[k][b] = out[k]; Jac[k * NZ + ZW + b] = out[k]; }
            }
        }
        // d/do_a: column M (-w x (J_a w) - J_a wd); J_a nu; the (o_a, w) curvature nu x (J_a w) - J_a (nu x w)
        real Jan[4][3], colO[4][3];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            real Ra[9], Ja[9], Jaw[3], t[3], Jawd[3];
            quat_dR(o, a, Ra);
            inertia_d(c, R, Ra, Ja);
            m3::mv(Ja, w, Jaw); m3::cross(w, Jaw, t); m3::mv(Ja, wd, Jawd);
            const real co[3] = {-t[0] - Jawd[0], -t[1] - Jawd[1], -t[2] - Jawd[2]};
            m3::mv(M, co, colO[a]);
#pragma unroll
            for (int k = 0; k < 3; k++) Jac[k * NZ + ZO + a] = colO[a][k];
            if (EXACT) {
                real t1[3], t2[3];
                m3::mv(Ja, nu, Jan[a]);
                m3::cross(nu, Jaw, t1); m3::mv(Ja, nxw, t2);
#pragma unroll
                for (int b = 0; b < 3; b++)
                    Ho[a * NZ + ZW + b] = -(Jan[a][0] * colW[0][b] + Jan[a][1] * colW[1][b] + Jan[a][2] * colW[2][b]) + (t1[b] - t2[b]);
            }
        }
        if (EXACT) {      // (o_a, o_b) = -(J_a nu) . Jac[:,o_b] - (J_b nu) . Jac[:,o_a] + v_ab
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = a; b < 4; b++) {
                    const int e = 4 * a - a * (a - 1) / 2 + (b - a);      // index of (a, b) in the rolled pair loop above
                    const real v = sc[e * ss] - (Jan[a][0] * colO[b][0] + Jan[a][1] * colO[b][1] + Jan[a][2] * colO[b][2])
                                               - (Jan[b][0] * colO[a][0] + Jan[b][1] * colO[a][1] + Jan[b][2] * colO[a][2]);
                    Ho[a * NZ + ZO + b] = v;
                    if (b != a) Ho[b * NZ + ZO + a] = v;
                }
            // (w,w): skew(nu) J - J skew(nu)
            pk[PK_HWW + 0] = (-nu[2] * J[3] + nu[1] * J[6]) - (J[1] * nu[2] - J[2] * nu[1]);
            pk[PK_HWW + 1] = (-nu[2] * J[4] + nu[1] * J[7]) - (-J[0] * nu[2] + J[2] * nu[0]);
            pk[PK_HWW + 2] = (-nu[2] * J[5] + nu[1] * J[8]) - (J[0] * nu[1] - J[1] * nu[0]);
            pk[PK_HWW + 3] = (nu[2] * J[0] - nu[0] * J[6]) - (J[4] * nu[2] - J[5] * nu[1]);
            pk[PK_HWW + 4] = (nu[2] * J[1] - nu[0] * J[7]) - (-J[3] * nu[2] + J[5] * nu[0]);
            pk[PK_HWW + 5] = (nu[2] * J[2] - nu[0] * J[8]) - (J[3] * nu[1] - J[4] * nu[0]);
            pk[PK_HWW + 6] = (-nu[1] * J[0] + nu[0] * J[3]) - (J[7] * nu[2] - J[8] * nu[1]);
            pk[PK_HWW + 7] = (-nu[1] * J[1] + nu[0] * J[4]) - (-J[6] * nu[2] + J[8] * nu[0]);
            pk[PK_HWW + 8] = (-nu[1] * J[2] + nu[0] * J[5]) - (J[6] * nu[1] - J[7] * nu[0]);
        }
        // d tau / d r_b = F x e_b;  d tau / d c_ib = -f_i x e_b;  d tau / d f_ib = (c_i - r) x e_b
        put3<EXACT>(Jac, M, Jan, ZR, F[0], F[1], F[2]);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const real* ci = x + XC + 3 * i;
            const real* fi = u + 6 * i + 3;
            put3<EXACT>(Jac, M, Jan, ZC + 3 * i, -fi[0], -fi[1], -fi[2]);
            put3<EXACT>(Jac, M, Jan, ZF + 3 * i, ci[0] - r[0], ci[1] - r[1], ci[2] - r[2]);
        }
    }
    __device__ SDDP_NOINLINE static void pack(const DevCfg& c, int kind, const real* x, const real* u, real* pk, real* sc, int ss) {
        if (kind == NODE_TERM) return;
        if (kind == NODE_TAIL) {      // LIP-style tail: wdot = 0 identically, so no Jacobian and no curvature; rddot as usual
            for (int i = 0; i < PACK; i++) pk[i] = 0.0;
            real fsum[3] = {0, 0, 0};
            for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) fsum[k] += u[6 * i + 3 + k];
            pk[PK_RDD + 0] = fsum[0] * c.inv_ms; pk[PK_RDD + 1] = fsum[1] * c.inv_ms; pk[PK_RDD + 2] = fsum[2] * c.inv_ms - c.g;
            return;
        }
        if (c.hessian_mode == 0) pack_impl<true>(c, x, u, pk, sc, ss);
        else pack_impl<false>(c, x, u, pk, sc, ss);
    }

    // ---- structured backward pass, phase e: Qxx, Qux, Qx, Qu (+ the y-recursion copies Qx2, Qu2) += lxx, lux, lx, lu ------
    // One pass over disjoint destinations, no barrier and no shared scratch (the round-1 version needed two barrier-
    // separated passes among warps 1-3): the 2 gq Jac^T Jac part already rides on the caller's fx^T / fu^T products, so
    // what is left is (1) the 184 curvature entries of the wdot block through the compact descriptor list at ZT_COFF,
    // each carrying the index of the affine Hessian value that lands on the same entry (orientation tracking on (o, o),
    // w tracking on the (w, w) diagonal), (2) the 44 other affine Hessian entries through the table at ZT_AOFF, (3) the
    // gradients lx (37) and lu (24), one per thread.  tid in [0, ZT_LAZY_THREADS); Qux = Qxx + ZT_QUX_OFF.
    // Compact descriptors: valid | hs << 1 | hoff << 3 | dst1 << 12 | dst2 << 24 | (affine value index + 1) << 36.
    // (measured: as a real call, one copy instead of three, this function costs 10 % of the whole solve -- the ABI spills)
    SDDP_DEV static real aff_value(const DevCfg& c, int kind, const real* x, const real* p, int l) {
        const bool track = kind != NODE_FIRST, tail = kind == NODE_TAIL;
        if (l < 12) {      // (cd, cd) diagonal: relative_vel + cdotxy_tracking (prb.py:166-181) [+ velocity box]
            const int foot = l / 3, ax = l - 3 * foot;
            const real sw_ = p[8 + 2 * foot];
            real v = ax < 2 ? real(2.0) * c.cw * (real(1.0) + sw_ * sw_) : 0.0;
            if (SDDP_CONE_ON(c) && c.w_cdb != 0.0) { double val, g_, h_; cdot_box_cold(c, x[XCD + l], val, g_, h_); v += h_; }
            return v;
        }
        if (l >= AF_OO) {      // (a, b), a <= b in row-major order: orientation tracking 2 otg^2 (E^T E)_ab (prb.py:185-189)
            // E(q) is the matrix of a quaternion product, so E^T E = |q|^2 I for any q: the diagonal entries sit at
            // l - AF_OO = 0, 4, 7, 9 of the row-major upper triangle
            const real* qr = p + 15;
            const int t = l - AF_OO;
            const bool diag = t == 0 || t == 4 || t == 7 || t == 9;
            return (track && diag) ? real(2.0) * p[6] * p[6] * (qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3]) : 0.0;
        }
        const real tw = (l == AF_RELP) ? real(2.0) * c.w_rel : (l == AF_RELN ? -real(2.0) * c.w_rel : (l == AF_RDOT ? real(2.0) * c.w_rdot : (l == AF_RZ ? real(2.0) * c.w_r : (l == AF_WW ? real(2.0) * c.w_w : 0.0))));
        real v = track ? tw : 0.0;
        if (l == AF_NCW) v = -real(2.0) * c.cw;
        if (l == AF_CW) v = real(2.0) * c.cw;
        if (tail && (l == AF_RZ || l == AF_WW)) v += real(2.0) * c.cw;      // lip_com_height, lip_zero_angular_momentum
        return v;
    }
    SDDP_DEV static void apply_rec(const DevCfg& c, int kind, const real* x, const real* u, const real* p, const real* pk,
                                   real* Qxx, real* Qx, real* Qx2, real* Qu, real* Qu2, int tid) {
        const bool track = kind != NODE_FIRST, tail = kind == NODE_TAIL;
        const bool exact = c.hessian_mode == 0 && !tail;      // Gauss-Newton / LIP-style tail: no curvature, affine parts only
        const real g2 = real(2.0) * c.gq;
        constexpr int EU = SDDP_E_UNROLL;
#pragma unroll EU
        for (int r = 0; r < ZT_CROUNDS; r++) {
            const unsigned long long d = __ldg(c.ztab + ZT_COFF + r * ZT_LAZY_THREADS + tid);
            if (!(d & 1ull)) continue;
            const real v = exact ? g2 * pk[(int)(d >> 3) & 511] : 0.0;
            real hv = ((d >> 1) & 3) == 1 ? v : -v;
            const int xi = (int)(d >> 36) & 63;
            if (xi) hv += aff_value(c, kind, x, p, xi - 1);
            const int o1 = (int)(d >> 12) & 4095, o2 = (int)(d >> 24) & 4095;
            Qxx[o1] += hv;
            if (o2 != o1) Qxx[o2] += hv;
        }
        if (tid < ZT_NAFF) {
            const unsigned d = (unsigned)__ldg(c.ztab + ZT_AOFF + tid);
            Qxx[d & 0xfffu] += aff_value(c, kind, x, p, (int)(d >> 12));
        }
        const int e = tid - (ZT_LAZY_THREADS - NX - NU);      // the last 61 threads: lu, lx
        if (e < 0) return;
        auto jtw = [&](int z) {      // 2 gq Jac[:, z] . wdot
            return g2 * (pk[PK_JAC + z] * pk[PK_WD] + pk[PK_JAC + NZ + z] * pk[PK_WD + 1] + pk[PK_JAC + 2 * NZ + z] * pk[PK_WD + 2]);
        };
        if (e < NU) {                                  // lu
            const int i = e / 6, r_ = e - 6 * i;
            real g;
            if (r_ < 3) g = g2 * u[e];
            else {
                const real a = real(1.0) - p[8 + 2 * i];
                g = g2 * c.inv_ms * pk[PK_RDD + r_ - 3] + real(2.0) * (c.w_minf + c.w_fsw * a * a) * u[e] + jtw(ZF + 3 * i + r_ - 3);
                if (SDDP_CONE_ON(c)) { double val, cg[3], ch[6]; cone_terms_cold(c, u + 6 * i + 3, val, cg, ch); g += cg[r_ - 3]; }
            }
            const int j = e * ZT_LDUX;
            Qu[j] += g; Qu2[j] += g;
        } else {                                       // lx
            const int i = e - NU;
            real g = 0.0;
            if (i < XRD) g = jtw(i); else if (i >= XW && i < XCD) g = jtw(ZW + i - XW);
            if (tail && (i == 2 || (i >= XW && i < XCD))) g += real(2.0) * c.cw * ((i == 2) ? x[2] - c.com[2] : x[i]);
            if (i == 2) { if (track) g += real(2.0) * c.w_r * (x[2] - c.com[2]); }
            else if (i >= XO && i < XC) {
                if (track) {      // E^T (E o - e_4) = |q|^2 o - E[3][:],  E[3][:] = (-q0, -q1, -q2, q3)
                    const real* qr = p + 15;
                    const int a = i - XO;
                    const real qq = qr[0] * qr[0] + qr[1] * qr[1] + qr[2] * qr[2] + qr[3] * qr[3];
                    g += real(2.0) * p[6] * p[6] * (qq * x[i] - (a == 3 ? qr[3] : -qr[a]));
                }
            } else if (i >= XC && i < XRD) {
                const int foot = (i - XC) / 3, ax = (i - XC) % 3;
                if (ax == 2) g += real(2.0) * c.cw * (x[i] - p[7 + 2 * foot]);
                else if (track) {
                    const int j = foot & 1, ia = XC + 3 * j + ax, ib = ia + 6;
                    const real res = -x[ia] + x[ib] - c.drel[j][ax];
                    g += (foot < 2) ? -real(2.0) * c.w_rel * res : real(2.0) * c.w_rel * res;
                }
            } else if (i >= XRD && i < XW) { if (track) g += real(2.0) * c.w_rdot * (x[i] - p[i - XRD]); }
            else if (i >= XW && i < XCD) { if (track) g += real(2.0) * c.w_w * (x[i] - p[3 + i - XW]); }
            else if (i >= XCD) {
                const int foot = (i - XCD) / 3, ax = (i - XCD) % 3;
                if (ax < 2) {
                    const int ia = XCD + 3 * (foot & ~1) + ax, ib = ia + 3;
                    const real res = x[ia] - x[ib], sw_ = p[8 + 2 * foot];
                    g += real(2.0) * c.cw * (((foot & 1) ? -res : res) + sw_ * sw_ * x[i]);
                }
                if (SDDP_CONE_ON(c) && c.w_cdb != 0.0) { double val, g_, h_; cdot_box_cold(c, x[i], val, g_, h_); g += g_; }
            }
            Qx[i] += g; Qx2[i] += g;
        }
    }

    SDDP_DEV static int zmap_x(int pi) { return pi < 19 ? pi : (pi < 22 ? XW + (pi - 19) : -1); }
    SDDP_DEV static int zmap_u(int pi) { return 6 * ((pi - 22) / 3) + 3 + (pi - 22) % 3; }

    // curvature entry Hc[pi][qi] from the pack
    SDDP_DEV static real hc(const real* pk, int pi, int qi) {
        if (pi >= ZO && pi < ZC) return pk[PK_HO + (pi - ZO) * NZ + qi];
        if (qi >= ZO && qi < ZC) return pk[PK_HO + (qi - ZO) * NZ + pi];
        if (pi >= ZW && pi < ZF && qi >= ZW && qi < ZF) return pk[PK_HWW + 3 * (pi - ZW) + (qi - ZW)];
        const real* nu = pk + PK_NU;
        if (pi >= ZF && qi < ZF) { int t = pi; pi = qi; qi = t; }     // make pi the x-side, qi the f-side
        if (qi >= ZF) {
            int fi = (qi - ZF) / 3, b = (qi - ZF) % 3;
            if (pi < ZO) return m3::skew_ab(nu, pi, b);                                         // (r_a, f_ib): +skew(nu)[a][b]
            if (pi >= ZC && pi < ZW && (pi - ZC) / 3 == fi) return -m3::skew_ab(nu, (pi - ZC) % 3, b);   // (c_ia, f_ib)
        }
        return 0.0;
    }

    // Q buffers <- lx, lu, lxx, lux, luu of this node.  Every thread of the block must call.  (Stage-1 kernel, generic
    // dense solve kernel and the terminal node of the structured one; the structured backward pass itself adds the node
    // terms in its phase e: apply_rec.)
    // Fast path (128 threads + descriptor table): three barrier-separated passes, each with one or a few
    // entries per thread and short uniform branches:
    //   1. zero fill   2. wdot block 2 gq (Jac^T Jac + Hc) through the descriptor table   3. affine residuals
    SDDP_DEV static void prep_E(const real* p, real* scratch, int t) {   // t in [0, 16)
        scratch[t] = E_row(p + 15, t >> 2, t & 3);
    }
    template <int LDUX = NX, class Sync>
    __device__ static void expand(const DevCfg& c, int kind, const real* x, const real* u, const real* p, const real* pk,
                                  real* Qx, real* Qu, real* Qxx, real* Qux, real* Quu, int tid, int nthr, Sync sync,
                                  real* scratch = nullptr) {
        if (c.ztab == nullptr || nthr != ZT_THREADS || scratch == nullptr) {
            expand_generic<LDUX>(c, kind, x, u, p, pk, Qx, Qu, Qxx, Qux, Quu, tid, nthr, sync);
            return;
        }
        constexpr int NTH = ZT_THREADS, ROUNDS = ZT_ROUNDS, EB = NTH - 32;
        const bool track = kind != NODE_FIRST, input = kind != NODE_TERM, tail = kind == NODE_TAIL;
        unsigned long long zd[ROUNDS];
        if (input) {      // latency hidden by the zero fill
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) zd[r] = __ldg(c.ztab + r * NTH + tid);
        }
        real* Es = scratch;          // E(oref) 4x4, then the four orientation residuals
        for (int e = tid; e < NX * NX; e += NTH) Qxx[e] = 0.0;
        for (int e = tid; e < NU * LDUX; e += NTH) Qux[e] = 0.0;
        for (int e = tid; e < NU * NU; e += NTH) Quu[e] = 0.0;
        if (tid < NX) Qx[tid] = 0.0;
        if (tid < NU) Qu[tid] = 0.0;
        if (track && tid >= EB && tid < EB + 16) prep_E(p, Es, tid - EB);
        sync();
        PROF(20);
        if (track && tid >= EB && tid < EB + 4) {
            const int r = tid - EB;
            const real* o = x + XO;
            Es[16 + r] = Es[4 * r] * o[0] + Es[4 * r + 1] * o[1] + Es[4 * r + 2] * o[2] + Es[4 * r + 3] * o[3] - (r == 3 ? real(1.0) : 0.0);
        }
        if (input) {   // gq * ||wdot||^2 : 2 gq (Jac^T Jac + Hc), upper triangle mirrored (all zero on the LIP-style tail)
            const real* Jac = pk + PK_JAC;
            const real g2 = real(2.0) * c.gq;
            const bool exact = c.hessian_mode == 0 && !tail;
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) {
                const unsigned long long d = zd[r];
                if (!ZT_VALID(d)) continue;
                const int pi = ZT_PI(d), qi = ZT_QI(d), da = ZT_DA(d), db = ZT_DB(d), hs = ZT_HSIGN(d);
                real hh = Jac[pi] * Jac[qi] + Jac[NZ + pi] * Jac[NZ + qi] + Jac[2 * NZ + pi] * Jac[2 * NZ + qi];
                if (exact && hs) { const real v = pk[ZT_HOFF(d)]; hh += (hs == 1) ? v : -v; }
                hh *= g2;
                const int kd = ZT_KIND(d);
                if (kd == 0) { Qxx[da * NX + db] = hh; Qxx[db * NX + da] = hh; }
                else if (kd == 1) Qux[da * LDUX + db] = hh;
                else { Quu[da * NU + db] = hh; Quu[db * NU + da] = hh; }
            }
            if (tid < NZ) {
                const int pi = tid;
                const real g = g2 * (Jac[pi] * pk[PK_WD] + Jac[NZ + pi] * pk[PK_WD + 1] + Jac[2 * NZ + pi] * pk[PK_WD + 2]);
                const int xi = zmap_x(pi);
                if (xi >= 0) Qx[xi] = g; else Qu[zmap_u(pi)] = g;
            }
        }
        sync();
        PROF(21);
        // ---- affine residuals, Hessian: one entry per thread (119 entries)
        const int t = tid;
        // (each case only computes the destination and the value; one read-modify-write at the end keeps the cases
        //  short enough to be predicated instead of branched over)
        int hdst = -1;
        bool huu = false;
        real hval = 0.0;
        if (t < 48) {                                   // rddot rows of min_qddot + min_f + f_active   (prb.py:200-204)
            if (input) {
                const int k = t >> 4, i = (t >> 2) & 3, j = t & 3;
                hval = real(2.0) * c.gq * c.inv_ms * c.inv_ms;
                if (i == j) { const real a = real(1.0) - p[8 + 2 * i]; hval += real(2.0) * (c.w_minf + c.w_fsw * a * a); }
                hdst = (6 * i + 3 + k) * NU + 6 * j + 3 + k; huu = true;
            }
        } else if (t < 64) {                            // o_tracking_xyz / _w   (prb.py:185-189)
            if (track) {
                const int a = (t - 48) >> 2, b = (t - 48) & 3;
                const real hh = Es[a] * Es[b] + Es[4 + a] * Es[4 + b] + Es[8 + a] * Es[8 + b] + Es[12 + a] * Es[12 + b];
                hdst = (XO + a) * NX + XO + b; hval = real(2.0) * p[6] * p[6] * hh;
            }
        } else if (t < 80) {                            // rel_pos_*   (prb.py:192-199)
            if (track) {
                const int r = (t - 64) >> 2, e = (t - 64) & 3, j = r >> 1, ax = r & 1;
                const int ia = XC + 3 * j + ax, ib = ia + 6;
                const int row = (e & 2) ? ib : ia, col = (e & 1) ? ib : ia;
                hdst = row * NX + col; hval = (row == col) ? real(2.0) * c.w_rel : -real(2.0) * c.w_rel;
            }
        } else if (t < 96) {                            // relative_vel_* and cdotxy_tracking_*   (prb.py:166-181)
            if (input) {
                const int g = (t - 80) >> 2, e = (t - 80) & 3, leg = g >> 1, ax = g & 1;
                const int ia = XCD + 3 * (2 * leg) + ax, ib = ia + 3;
                const int row = (e & 2) ? ib : ia, col = (e & 1) ? ib : ia;
                const real sw = p[8 + 2 * (2 * leg + ((e & 2) ? 1 : 0))];
                hdst = row * NX + col; hval = (row == col) ? real(2.0) * c.cw * (real(1.0) + sw * sw) : -real(2.0) * c.cw;
            }
        } else if (t < 108) {                           // cddot rows of min_qddot
            if (input) { const int i = t - 96, ui = 6 * (i / 3) + i % 3; hdst = ui * NU + ui; hval = real(2.0) * c.gq; huu = true; }
        } else if (t < 112) {                           // cz_tracking_i   (prb.py:180)
            if (input) { const int id = XC + 3 * (t - 108) + 2; hdst = id * NX + id; hval = real(2.0) * c.cw; }
        } else if (t < 115) {                           // rdot_tracking   (prb.py:190)
            if (track) { const int id = XRD + t - 112; hdst = id * NX + id; hval = real(2.0) * c.w_rdot; }
        } else if (t < 118) {                           // w_tracking   (prb.py:191) + lip_zero_angular_momentum on the tail
            if (track) { const int id = XW + t - 115; hdst = id * NX + id; hval = real(2.0) * c.w_w + (tail ? real(2.0) * c.cw : 0.0); }
        } else if (t == 118) {                          // rz_tracking   (prb.py:184) + lip_com_height on the tail
            if (track) { hdst = 2 * NX + 2; hval = real(2.0) * c.w_r + (tail ? real(2.0) * c.cw : 0.0); }
        }
        if (hdst >= 0) {
            if (huu) Quu[hdst] += hval; else Qxx[hdst] += hval;
        }
        // ---- affine residuals, gradient: thread tg < 24 owns lu[tg], thread 32 + i owns lx[i]
        const int tg = tid;
        if (tg < NU) {
            if (input) {
                const int i = tg / 6, r = tg % 6;
                real g;
                if (r < 3) g = real(2.0) * c.gq * u[tg];
                else {
                    const real a = real(1.0) - p[8 + 2 * i];
                    g = real(2.0) * c.gq * c.inv_ms * pk[PK_RDD + r - 3] + real(2.0) * (c.w_minf + c.w_fsw * a * a) * u[tg];
                    if (SDDP_CONE_ON(c)) { double val, cg[3], ch[6]; cone_terms_cold(c, u + 6 * i + 3, val, cg, ch); g += cg[r - 3]; }
                }
                Qu[tg] += g;
            }
        } else if (tg >= 32 && tg < 32 + NX) {
            const int i = tg - 32;
            real g = 0.0;
            if (tail && (i == 2 || (i >= XW && i < XCD))) g = real(2.0) * c.cw * ((i == 2) ? x[2] - c.com[2] : x[i]);
            if (i == 2) { if (track) g += real(2.0) * c.w_r * (x[2] - c.com[2]); }
            else if (i >= XO && i < XC) {
                if (track) {
                    const int a = i - XO;
                    g = real(2.0) * p[6] * p[6] * (Es[a] * Es[16] + Es[4 + a] * Es[17] + Es[8 + a] * Es[18] + Es[12 + a] * Es[19]);
                }
            } else if (i >= XC && i < XRD) {
                const int foot = (i - XC) / 3, ax = (i - XC) % 3;
                if (ax == 2) { if (input) g = real(2.0) * c.cw * (x[i] - p[7 + 2 * foot]); }
                else if (track) {
                    const int j = foot & 1, ia = XC + 3 * j + ax, ib = ia + 6;
                    const real res = -x[ia] + x[ib] - c.drel[j][ax];
                    g = (foot < 2) ? -real(2.0) * c.w_rel * res : real(2.0) * c.w_rel * res;
                }
            } else if (i >= XRD && i < XW) { if (track) g = real(2.0) * c.w_rdot * (x[i] - p[i - XRD]); }
            else if (i >= XW && i < XCD) { if (track) g += real(2.0) * c.w_w * (x[i] - p[3 + i - XW]); }
            else if (i >= XCD) {
                const int foot = (i - XCD) / 3, ax = (i - XCD) % 3;
                if (ax < 2 && input) {
                    const int ia = XCD + 3 * (foot & ~1) + ax, ib = ia + 3;
                    const real res = x[ia] - x[ib], sw = p[8 + 2 * foot];
                    g = real(2.0) * c.cw * (((foot & 1) ? -res : res) + sw * sw * x[i]);
                }
            }
            Qx[i] += g;
        }
        if (input && SDDP_CONE_ON(c)) {      // inequality barriers: luu blocks per foot, contact-point velocity box on lx / lxx
            sync();
            if (tid < 36) {
                const int i = tid / 9, ka = (tid % 9) / 3, kb = tid % 3;
                double val, cg[3], ch[6];
                cone_terms_cold(c, u + 6 * i + 3, val, cg, ch);
                Quu[(6 * i + 3 + ka) * NU + 6 * i + 3 + kb] += ch[cone_hidx(ka, kb)];
            } else if (tid < 48 && c.w_cdb != 0.0) {
                const int id = XCD + tid - 36;
                double val, g_, h_;
                cdot_box_cold(c, x[id], val, g_, h_);
                Qx[id] += g_; Qxx[id * NX + id] += h_;
            }
        }
        sync();
        PROF(22);
    }

    // generic fallback (any thread count, no descriptor table)
    template <int LDUX = NX, class Sync>
    __device__ static void expand_generic(const DevCfg& c, int kind, const real* x, const real* u, const real* p, const real* pk,
                                  real* Qx, real* Qu, real* Qxx, real* Qux, real* Quu, int tid, int nthr, Sync sync) {
        for (int e = tid; e < NX * NX; e += nthr) Qxx[e] = 0.0;
        for (int e = tid; e < NU * LDUX; e += nthr) Qux[e] = 0.0;
        for (int e = tid; e < NU * NU; e += nthr) Quu[e] = 0.0;
        for (int e = tid; e < NX; e += nthr) Qx[e] = 0.0;
        for (int e = tid; e < NU; e += nthr) Qu[e] = 0.0;
        sync();
        const bool track = kind != NODE_FIRST, input = kind != NODE_TERM, tail = kind == NODE_TAIL;
        if (input) {   // gq * ||wdot||^2 : 2 gq (Jac^T Jac + Hc)  (all zero on the LIP-style tail)
            const real* Jac = pk + PK_JAC;
            const real g2 = real(2.0) * c.gq;
            const bool exact = c.hessian_mode == 0 && !tail;
            if (c.ztab != nullptr && nthr == ZT_THREADS) {
#pragma unroll
                for (int r = 0; r < ZT_ROUNDS; r++) {
                    const unsigned long long d = __ldg(c.ztab + r * ZT_THREADS + tid);
                    if (!ZT_VALID(d)) continue;
                    const int pi = ZT_PI(d), qi = ZT_QI(d), da = ZT_DA(d), db = ZT_DB(d), hs = ZT_HSIGN(d);
                    real hh = Jac[pi] * Jac[qi] + Jac[NZ + pi] * Jac[NZ + qi] + Jac[2 * NZ + pi] * Jac[2 * NZ + qi];
                    if (exact && hs) { const real v = pk[ZT_HOFF(d)]; hh += (hs == 1) ? v : -v; }
                    hh *= g2;
                    const int kd = ZT_KIND(d);
                    if (kd == 0) { Qxx[da * NX + db] = hh; Qxx[db * NX + da] = hh; }
                    else if (kd == 1) Qux[da * LDUX + db] = hh;
                    else { Quu[da * NU + db] = hh; Quu[db * NU + da] = hh; }
                }
            } else
            for (int e = tid; e < NZ * NZ; e += nthr) {
                int pi = e / NZ, qi = e % NZ;
                int xi = zmap_x(pi), xj = zmap_x(qi);
                if (xi >= 0 && xj < 0) continue;       // (x,u) pairs are stored once, as lux[u][x]
                real hh = Jac[pi] * Jac[qi] + Jac[NZ + pi] * Jac[NZ + qi] + Jac[2 * NZ + pi] * Jac[2 * NZ + qi];
                if (exact) hh += hc(pk, pi, qi);
                hh *= g2;
                if (xi >= 0) Qxx[xi * NX + xj] = hh;
                else if (xj >= 0) Qux[zmap_u(pi) * LDUX + xj] = hh;
                else Quu[zmap_u(pi) * NU + zmap_u(qi)] = hh;
            }
            for (int pi = tid; pi < NZ; pi += nthr) {
                real g = g2 * (Jac[pi] * pk[PK_WD] + Jac[NZ + pi] * pk[PK_WD + 1] + Jac[2 * NZ + pi] * pk[PK_WD + 2]);
                int xi = zmap_x(pi);
                if (xi >= 0) Qx[xi] = g; else Qu[zmap_u(pi)] = g;
            }
        }
        sync();
        // affine residuals: each task owns a disjoint set of entries
        const int t = tid;
        if (t == 0 && track) {                                      // rz_tracking, prb.py:184
            Qxx[2 * NX + 2] += real(2.0) * c.w_r; Qx[2] += real(2.0) * c.w_r * (x[2] - c.com[2]);
        } else if (t == 1 && track) {                               // o_tracking_xyz / _w, prb.py:185-189
            const real* q = p + 15;
            const real* o = x + XO;
            real w2 = real(2.0) * p[6] * p[6], res[4];
            for (int i = 0; i < 4; i++)
                res[i] = E_row(q, i, 0) * o[0] + E_row(q, i, 1) * o[1] + E_row(q, i, 2) * o[2] + E_row(q, i, 3) * o[3] - (i == 3 ? real(1.0) : 0.0);
            for (int a = 0; a < 4; a++) {
                real g = 0;
                for (int i = 0; i < 4; i++) g += E_row(q, i, a) * res[i];
                Qx[XO + a] += w2 * g;
                for (int b = 0; b < 4; b++) {
                    real hh = 0;
                    for (int i = 0; i < 4; i++) hh += E_row(q, i, a) * E_row(q, i, b);
                    Qxx[(XO + a) * NX + XO + b] += w2 * hh;
                }
            }
        } else if (t >= 2 && t < 5 && track) {                      // rdot_tracking, prb.py:190
            int i = XRD + t - 2;
            Qxx[i * NX + i] += real(2.0) * c.w_rdot; Qx[i] += real(2.0) * c.w_rdot * (x[i] - p[t - 2]);
        } else if (t >= 5 && t < 8 && track) {                      // w_tracking, prb.py:191
            int i = XW + t - 5;
            Qxx[i * NX + i] += real(2.0) * c.w_w; Qx[i] += real(2.0) * c.w_w * (x[i] - p[3 + t - 5]);
        } else if (t >= 8 && t < 12 && track) {                     // rel_pos_*, prb.py:192-199
            int j = (t - 8) / 2, ax = (t - 8) % 2;
            int ia = XC + 3 * j + ax, ib = ia + 6;
            real w2 = real(2.0) * c.w_rel, res = -x[ia] + x[ib] - c.drel[j][ax];
            Qxx[ia * NX + ia] += w2; Qxx[ib * NX + ib] += w2; Qxx[ia * NX + ib] -= w2; Qxx[ib * NX + ia] -= w2;
            Qx[ia] -= w2 * res; Qx[ib] += w2 * res;
        } else if (t >= 12 && t < 16 && input) {                    // cz_tracking_i, prb.py:180
            int i = t - 12, id = XC + 3 * i + 2;
            Qxx[id * NX + id] += real(2.0) * c.cw; Qx[id] += real(2.0) * c.cw * (x[id] - p[7 + 2 * i]);
        } else if (t >= 16 && t < 19 && input) {                    // rddot rows of min_qddot + min_f + f_active
            int k = t - 16;
            real h2 = real(2.0) * c.gq * c.inv_ms * c.inv_ms, g = real(2.0) * c.gq * c.inv_ms * pk[PK_RDD + k];
            for (int i = 0; i < 4; i++) {
                int ui = 6 * i + 3 + k;
                real a = real(1.0) - p[8 + 2 * i], wf = real(2.0) * (c.w_minf + c.w_fsw * a * a);
                Qu[ui] += g + wf * u[ui];
                Quu[ui * NU + ui] += wf;
                for (int j = 0; j < 4; j++) Quu[ui * NU + 6 * j + 3 + k] += h2;
            }
        } else if (t >= 19 && t < 31 && input) {                    // cddot rows of min_qddot
            int i = (t - 19) / 3, k = (t - 19) % 3, ui = 6 * i + k;
            Quu[ui * NU + ui] += real(2.0) * c.gq; Qu[ui] += real(2.0) * c.gq * u[ui];
        } else if (t >= 31 && t < 35 && input) {                    // relative_vel_* and cdotxy_tracking_*, prb.py:166-181
            int leg = (t - 31) / 2, ax = (t - 31) % 2;
            int ia = XCD + 3 * (2 * leg) + ax, ib = ia + 3;
            real w2 = real(2.0) * c.cw, res = x[ia] - x[ib];
            real sa = p[8 + 2 * (2 * leg)], sb = p[8 + 2 * (2 * leg + 1)];
            Qxx[ia * NX + ia] += w2 * (real(1.0) + sa * sa); Qxx[ib * NX + ib] += w2 * (real(1.0) + sb * sb);
            Qxx[ia * NX + ib] -= w2; Qxx[ib * NX + ia] -= w2;
            Qx[ia] += w2 * (res + sa * sa * x[ia]); Qx[ib] += w2 * (-res + sb * sb * x[ib]);
        }
        sync();
        if (tail && tid == 0) {                                       // lip_com_height, lip_zero_angular_momentum (isrbd_example.py:352-353)
            Qxx[2 * NX + 2] += real(2.0) * c.cw; Qx[2] += real(2.0) * c.cw * (x[2] - c.com[2]);
            for (int i = XW; i < XW + 3; i++) { Qxx[i * NX + i] += real(2.0) * c.cw; Qx[i] += real(2.0) * c.cw * x[i]; }
        }
        if (input && SDDP_CONE_ON(c) && c.w_cdb != 0.0 && tid >= 4 && tid < 16) {      // contact-point velocity box
            const int id = XCD + tid - 4;
            double val, g_, h_;
            cdot_box_cold(c, x[id], val, g_, h_);
            Qx[id] += g_; Qxx[id * NX + id] += h_;
        }
        if (tail || (input && SDDP_CONE_ON(c))) sync();
        if (input && SDDP_CONE_ON(c)) {       // inequality barriers of the feet (extension): gradient and 3 x 3 Hessian per foot
            if (tid < 4) {
                double val, cg[3], ch[6];
                cone_terms_cold(c, u + 6 * tid + 3, val, cg, ch);
                for (int a = 0; a < 3; a++) {
                    Qu[6 * tid + 3 + a] += cg[a];
                    for (int b = 0; b < 3; b++) Quu[(6 * tid + 3 + a) * NU + 6 * tid + 3 + b] += ch[cone_hidx(a, b)];
                }
            }
            sync();
        }
    }

    // dense fx = I + dt A, fu = dt B.  Every thread of the block must call.
    template <class Sync>
    __device__ static void expand_f(const DevCfg& c, const real* x, const real* u, const real* pk, real* fx, real* fu,
                                    int tid, int nthr, Sync sync) {
        for (int e = tid; e < NX * NX; e += nthr) fx[e] = (e / NX == e % NX) ? real(1.0) : 0.0;
        for (int e = tid; e < NX * NU; e += nthr) fu[e] = 0.0;
        sync();
        const real dt = c.dt;
        const real* o = x + XO;
        const real* w = x + XW;
        const real* Jac = pk + PK_JAC;
        for (int e = tid; e < 105 + 60; e += nthr) {
            if (e < 3) fx[(XR + e) * NX + XRD + e] += dt;
            else if (e < 15) fx[(XC + e - 3) * NX + XCD + e - 3] += dt;
            else if (e < 24) {            // d odot_v / d o_v = skew(w)/2
                int a = (e - 15) / 3, b = (e - 15) % 3;
                if (a != b) fx[(XO + a) * NX + XO + b] += real(0.5) * dt * m3::skew_ab(w, a, b);
            } else if (e < 27) {          // d odot_v / d o_w = w/2 ; d odot_w / d o_v = -w/2
                int a = e - 24;
                fx[(XO + a) * NX + XO + 3] += real(0.5) * dt * w[a];
                fx[(XO + 3) * NX + XO + a] += -real(0.5) * dt * w[a];
            } else if (e < 36) {          // d odot_v / d w = (o_w I - skew(o_v))/2
                int a = (e - 27) / 3, b = (e - 27) % 3;
                fx[(XO + a) * NX + XW + b] += real(0.5) * dt * ((a == b ? o[3] : 0.0) - m3::skew_ab(o, a, b));
            } else if (e < 39) {          // d odot_w / d w = -o_v/2
                int a = e - 36;
                fx[(XO + 3) * NX + XW + a] += -real(0.5) * dt * o[a];
            } else if (e < 105) {         // wdot rows: r(3) o(4) c(12) w(3) = first 22 z-columns
                int a = (e - 39) / 22, pz = (e - 39) % 22;
                fx[(XW + a) * NX + zmap_x(pz)] += dt * Jac[a * NZ + pz];
            } else if (e < 117) {         // rddot / f
                int i = (e - 105) / 3, k = (e - 105) % 3;
                fu[(XRD + k) * NU + 6 * i + 3 + k] += dt * c.inv_ms;
            } else if (e < 129) {         // cddot
                int i = (e - 117) / 3, k = (e - 117) % 3;
                fu[(XCD + 3 * i + k) * NU + 6 * i + k] += dt;
            } else {                      // wdot / f
                int a = (e - 129) / 12, pf = (e - 129) % 12;
                fu[(XW + a) * NU + zmap_u(ZF + pf)] += dt * Jac[a * NZ + ZF + pf];
            }
        }
        sync();
    }
};
using Srbd = SrbdT<false>;       // the reference's problem: inequalities dropped (prb.py:173-177, ddp.py:197-209)
using SrbdI = SrbdT<true>;       // with the inequality barriers compiled in


// =====================================================================================  LIP
struct Lip {
    static constexpr int NX = 30, NU = 15, NP = 11, NACC = 1, PACK = 1;
    // x: r[0:3] c_i[3+3i] rdot[15:18] cdot_i[18+3i];  u: z[0:3] cddot_i[3+3i];  p: rdot_ref[0:3] (c_ref_i, sw_i)[3+2i, 4+2i]
    enum { XR = 0, XC = 3, XRD = 15, XCD = 18 };

    static constexpr int NPRE = 1;
    SDDP_DEV static void accel(const DevCfg&, const real*, const real*, real*, bool = false) {}
    SDDP_DEV static void accel_pre(const DevCfg&, const real*, real*) {}
    SDDP_DEV static void accel_post(const DevCfg&, const real*, const real*, const real*, real*, bool = false) {}
    SDDP_DEV static real xdot_i(const DevCfg& c, int i, const real* x, const real* u, const real*) {
        if (i < 15) return x[i + 15];
        if (i < 18) { int k = i - 15; return c.eta2 * (x[k] - u[k]) - (k == 2 ? c.g : 0.0); }   // prb.py:317-318
        return u[3 + i - 18];
    }
    SDDP_DEV static double cost_lane(const DevCfg& c, int kind, int lane, const real* x, const real* u, const real* p, const real*,
                                     int parts = 15) {
        double s = 0.0;
        if ((parts & 1) && lane < NX) {
            int i = lane;
            if (kind != NODE_FIRST) {   // prb.py:390-392, 394-401
                if (i == 2) { double r = (double)x[2] - c.com[2]; s += c.w_r * r * r; }
                else if (i < 2) { double r = (double)x[i] - 0.25 * ((double)x[3 + i] + (double)x[6 + i] + (double)x[9 + i] + (double)x[12 + i]); s += c.w_r * r * r; }
                else if (i >= 3 && i < 9 && (i - 3) % 3 < 2) {
                    int j = (i - 3) / 3, ax = (i - 3) % 3;
                    double r = -(double)x[i] + (double)x[i + 6] - c.drel[j][ax];
                    s += c.w_rel * r * r;
                } else if (i >= 15 && i < 18) { double r = (double)x[i] - (double)p[i - 15]; s += c.w_rdot * r * r; }
            }
            if (kind != NODE_TERM) {    // prb.py:379-387
                if (i >= 3 && i < 15 && (i - 3) % 3 == 2) { double r = (double)x[i] - (double)p[3 + 2 * ((i - 3) / 3)]; s += c.cw * r * r; }
                else if (i >= 18 && (i - 18) % 3 < 2) {
                    int j = (i - 18) / 3;
                    double r = (double)p[4 + 2 * j] * (double)x[i];
                    s += c.cw * r * r;
                    if (j == 0 || j == 2) { double r2 = (double)x[i] - (double)x[i + 3]; s += c.cw * r2 * r2; }
                }
            }
        }
        if ((parts & 2) && kind != NODE_TERM && lane < NU) {   // prb.py:393, 402
            if (lane < 3) {
                int k = lane;
                double r = (double)u[k] - 0.25 * ((double)x[3 + k] + (double)x[6 + k] + (double)x[9 + k] + (double)x[12 + k]);
                double a = c.eta2 * ((double)x[k] - (double)u[k]) - (k == 2 ? c.g : 0.0);
                s += c.w_zmp * r * r + c.gq * a * a;
            } else s += c.gq * (double)u[lane] * (double)u[lane];
        }
        return s;
    }
    __device__ static void pack(const DevCfg&, int, const real*, const real*, real*, real*, int) {}

    template <int LDUX = NX, class Sync>
    __device__ static void expand(const DevCfg& c, int kind, const real* x, const real* u, const real* p, const real*,
                                  real* Qx, real* Qu, real* Qxx, real* Qux, real* Quu, int tid, int nthr, Sync sync) {
        for (int e = tid; e < NX * NX; e += nthr) Qxx[e] = 0.0;
        for (int e = tid; e < NU * LDUX; e += nthr) Qux[e] = 0.0;
        for (int e = tid; e < NU * NU; e += nthr) Quu[e] = 0.0;
        for (int e = tid; e < NX; e += nthr) Qx[e] = 0.0;
        for (int e = tid; e < NU; e += nthr) Qu[e] = 0.0;
        sync();
        const bool track = kind != NODE_FIRST, input = kind != NODE_TERM;
        const int t = tid;
        if (t < 3) {   // axis group {r_k, c_0k..c_3k, z_k}
            int k = t;
            int ci[4] = {XC + k, XC + 3 + k, XC + 6 + k, XC + 9 + k};
            real csum = real(0.25) * (x[ci[0]] + x[ci[1]] + x[ci[2]] + x[ci[3]]);
            if (track) {
                if (k == 2) { Qxx[2 * NX + 2] += real(2.0) * c.w_r; Qx[2] += real(2.0) * c.w_r * (x[2] - c.com[2]); }
                else {     // rxy_tracking, prb.py:391
                    real w2 = real(2.0) * c.w_r, res = x[k] - csum;
                    Qxx[k * NX + k] += w2; Qx[k] += w2 * res;
                    for (int a = 0; a < 4; a++) {
                        Qxx[k * NX + ci[a]] -= real(0.25) * w2; Qxx[ci[a] * NX + k] -= real(0.25) * w2; Qx[ci[a]] -= real(0.25) * w2 * res;
                        for (int b = 0; b < 4; b++) Qxx[ci[a] * NX + ci[b]] += w2 / real(16.0);
                    }
                    for (int j = 0; j < 2; j++) {   // rel_pos, prb.py:394-401
                        int ia = ci[j], ib = ci[j + 2];
                        real wr = real(2.0) * c.w_rel, rr = -x[ia] + x[ib] - c.drel[j][k];
                        Qxx[ia * NX + ia] += wr; Qxx[ib * NX + ib] += wr; Qxx[ia * NX + ib] -= wr; Qxx[ib * NX + ia] -= wr;
                        Qx[ia] -= wr * rr; Qx[ib] += wr * rr;
                    }
                }
            }
            if (input) {
                real w2 = real(2.0) * c.w_zmp, res = u[k] - csum;   // zmp_tracking, prb.py:393
                Quu[k * NU + k] += w2; Qu[k] += w2 * res;
                for (int a = 0; a < 4; a++) {
                    Qux[k * LDUX + ci[a]] -= real(0.25) * w2; Qx[ci[a]] -= real(0.25) * w2 * res;
                    for (int b = 0; b < 4; b++) Qxx[ci[a] * NX + ci[b]] += w2 / real(16.0);
                }
                real e2 = c.eta2, wq = real(2.0) * c.gq * e2 * e2;   // min_qddot rddot rows, prb.py:402
                real acc = e2 * (x[k] - u[k]) - (k == 2 ? c.g : 0.0);
                Qxx[k * NX + k] += wq; Quu[k * NU + k] += wq; Qux[k * LDUX + k] -= wq;
                Qx[k] += real(2.0) * c.gq * e2 * acc; Qu[k] -= real(2.0) * c.gq * e2 * acc;
                if (k == 2)
                    for (int i = 0; i < 4; i++) { Qxx[ci[i] * NX + ci[i]] += real(2.0) * c.cw; Qx[ci[i]] += real(2.0) * c.cw * (x[ci[i]] - p[3 + 2 * i]); }
            }
        } else if (t < 6 && track) {
            int i = XRD + t - 3;
            Qxx[i * NX + i] += real(2.0) * c.w_rdot; Qx[i] += real(2.0) * c.w_rdot * (x[i] - p[t - 3]);
        } else if (t >= 6 && t < 10 && input) {
            int leg = (t - 6) / 2, ax = (t - 6) % 2;
            int ia = XCD + 3 * (2 * leg) + ax, ib = ia + 3;
            real w2 = real(2.0) * c.cw, res = x[ia] - x[ib];
            real sa = p[4 + 2 * (2 * leg)], sb = p[4 + 2 * (2 * leg + 1)];
            Qxx[ia * NX + ia] += w2 * (real(1.0) + sa * sa); Qxx[ib * NX + ib] += w2 * (real(1.0) + sb * sb);
            Qxx[ia * NX + ib] -= w2; Qxx[ib * NX + ia] -= w2;
            Qx[ia] += w2 * (res + sa * sa * x[ia]); Qx[ib] += w2 * (-res + sb * sb * x[ib]);
        } else if (t >= 10 && t < 22 && input) {
            int ui = 3 + t - 10;
            Quu[ui * NU + ui] += real(2.0) * c.gq; Qu[ui] += real(2.0) * c.gq * u[ui];
        }
        sync();
    }

    template <class Sync>
    __device__ static void expand_f(const DevCfg& c, const real*, const real*, const real*, real* fx, real* fu,
                                    int tid, int nthr, Sync sync) {
        for (int e = tid; e < NX * NX; e += nthr) fx[e] = (e / NX == e % NX) ? real(1.0) : 0.0;
        for (int e = tid; e < NX * NU; e += nthr) fu[e] = 0.0;
        sync();
        const real dt = c.dt;
        for (int e = tid; e < 33; e += nthr) {
            if (e < 15) { fx[e * NX + e + 15] += dt; if (e >= 3) fu[(XCD + e - 3) * NU + 3 + e - 3] += dt; }
            else if (e < 18) { int k = e - 15; fx[(XRD + k) * NX + k] += dt * c.eta2; fu[(XRD + k) * NU + k] += -dt * c.eta2; }
        }
        sync();
    }
};
