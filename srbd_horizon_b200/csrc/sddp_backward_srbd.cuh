// sddp_backward_srbd.cuh -- stage 2 (backward Riccati pass) specialised to the SRBD block structure.
//
// State blocks  x = [ r(0:3) | o(3:7) | c(7:19) | rd(19:22) | w(22:25) | cd(25:37) ]   (prb.py:32-59)
// fx = I + dt A,  fu = dt B  with (prb.py:97-109)
//   A rows r : [rd: I]                    B rows rd : f_i -> (fs/m) I
//   A rows o : [o: Aoo(4x4), w: Aow(4x3)] B rows w  : f_i -> G_i = d wdot / d f_i
//   A rows c : [cd: I]                    B rows cd_i: cddot_i -> I
//   A rows w : [r,o,c,w: d wdot / d(r,o,c,w)]  (3x22, from the node pack)
// so every product with fx / fu is a handful of 3x3 / 4x4 block operations:
//   T   = V fx           in place, one thread per row of V           (4.0 K FMA instead of 50.7 K)
//   Qxx = lxx + fx^T T,  Qux = lux + fu^T T   one thread per (column, row group)
//   Quu = luu + fu^T V fu                      one thread per entry of the lower triangle
// Gains (W-form of sddp_solver.cuh, square-root free):
//   warp 0 factorises Quu_r = Lt D Lt^T (one lane per column held in registers, 24 pivot steps, only
//   __syncwarp between them, branch-free updates, Newton-refined reciprocal started one element early) WHILE warps 1-3 build
//   T, Qxx, Qux (they do not depend on the factor).  Then one thread per right-hand side applies
//   Lt^-1 to [Qux | Qu | I] with no barrier at all: frozen rows U = Lt^-1 [Qux Qu], E = Lt^-1.
//   With rs = D^-1/2:  Wn = rs.U (= L^-1 [Qux Qu] of the Cholesky form), Es = rs.E
//   K = -Es^T Wn  and  [Vxx Vx; . |w0|^2] = [sym(Qxx) Qx; . 0] - Wn^T Wn  on the FP64 tensor cores
//   (mma.sync m8n8k4 f64, 8x8 tiles, fragments straight from shared memory: two 8-byte loads feed 256 FMA,
//   where a 3x3 register tile needs six loads for nine -- these products were shared-memory-bandwidth bound)
// Node inputs (x, u, p, d, pack) of node k-1 are fetched with cp.async while node k is processed.
#pragma once
#include <cstddef>
#include "sddp_solver.cuh"

#ifndef SDDP_D1BLOCK
// 1: the factorisation takes one 2 x 2 pivot block per step.  Default: on in the fp32 build only.  Measured on B200, 8192
// problems: fp32 55.6 -> 48.6 ms (its factorisation is bound by the publish -> sync -> load chain that the block step
// halves); fp64 60.70 -> 60.55 ms and a single solve 0.957 -> 0.981 ms (there the warp is held up by the FP64 / shared-
// memory traffic of the three warps beside it, not by its own chain), so the fp64 library keeps one pivot per step.
#ifdef SDDP_F32
#define SDDP_D1BLOCK 1
#else
#define SDDP_D1BLOCK 0
#endif
#endif
// LAT: the latency variant (batches smaller than the grid): unrolled factorisation and interleaved f/g tiles; otherwise the
// throughput variant (rolled, a third of the code).  Same arithmetic up to the order of independent operations.
template <class MT, bool LAT = false>
struct alignas(16) SmemSrbdT {
    static constexpr int NX = 37, NU = 24, NP = 19;
    static constexpr int LDW = 44;   // row pitch of W: 88 words = 24 mod 32, so the 4 x 8 DMMA fragment loads are conflict free
    real VT[NX * NX + 1];    // Vxx', then T = Vxx' fx in place, then the new Vxx; forward: scratch
    alignas(16) real Qxx[NX * NX + 1];   // forward: K of the current / next node (with W: 2 x 888 doubles)
    real W[NU * LDW];        // B = [Qux | Qu | quy | .] -> Wn = Es B  (quy: lu + fu^T ys of the y recursion)
    real Quu[NU * NU];       // Quu -> Et = Lt^-1 (unit lower triangular, zeros above the diagonal); Es = diag(rs) Et
    real Vx[NX + 1], y[NX + 1], Qx[NX + 1], vp[NX + 1], ys[NX + 1], qxy[NX + 1], sv[NX + 1];
    real Qu[NU], kk[NU];
    real invp[NU], rs[NU];   // d1: 1 / pivot_j; 1 / sqrt(pivot_j) (row scaling of Et, applied in h and g)
    real prow[SDDP_D1BLOCK ? 4 : 2][NU];   // d1: the published column(s) of the current pivot step / 2 x 2 block (ping-pong)
    alignas(16) real nb[2][NodeBuf<MT>::SIZE];
    real escr[24];           // expand scratch: E(oref) and the orientation residuals
    real sacc[NWARP][8];
    double red[16];          // reductions and the expected-decrease accumulators: double in both builds
    double alpha[NCAND], rho[NCAND], Jc[NCAND];
    const real* gp[8];       // base pointers of the per-node prefetches (registers are scarce in the node loops)
    unsigned long long mbar[2];   // completion barriers of the bulk copies of K_k (forward_wave)
    int iflag[4];
    __device__ real* Kbuf(int b) { return Qxx + b * (NU * NX); }
    __device__ real* scr() { return VT; }
    __device__ static int backward(const DevCfg& c, SmemSrbdT& S, const real* X, const real* U, const real* P, const real* D,
                                   const real* packs, real mu, real* Kg, real* kg, double* dV3, bool has_gap, int tid);
    // rigid-body packs of nodes 0..N-1, one thread per node
    __device__ static void prep(const DevCfg& c, SmemSrbdT& S, const real* X, const real* U, const real*, real* packs, int tid) {
        // scratch: VT, Qxx, W, Quu and the first vectors (all dead between the forward and the backward pass)
        static_assert(offsetof(SmemSrbdT, Qu) >= PACK_SCRATCH * NT * sizeof(real), "pack scratch");
        compute_packs<MT>(c, X, U, packs, S.VT, tid);
    }
};
using SmemSrbd = SmemSrbdT<Srbd>;
using SmemSrbdI = SmemSrbdT<SrbdI>;
using SmemSrbdL = SmemSrbdT<Srbd, true>;
using SmemSrbdIL = SmemSrbdT<SrbdI, true>;
static_assert(2 * 24 * 37 <= (37 * 37 + 1) + 24 * 44, "forward K real buffer must fit in Qxx + W");
static_assert(offsetof(SmemSrbd, W) - offsetof(SmemSrbd, Qxx) == ZT_QUX_OFF * sizeof(real) && SmemSrbd::LDW == ZT_LDUX, "descriptor table destinations (sddp.cu:build_ztab)");
#define SDDP_INEQ_ON(c) (MT::HAS_INEQ && (c).ineq != 0)
SDDP_DEV void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
struct Bar96 { __device__ void operator()() const { bar_named(2, 96); } };      // warps 1-3

// D(8x8) += A(8x4) B(4x8) on the FP64 tensor cores.  Fragments (PTX ISA, m8n8k4 .f64): lane holds
// A[lane/4][lane%4], B[lane%4][lane/4], C[lane/4][2*(lane%4) + {0,1}].
// (fp32 build: operands converted on the way in, accumulators stay double -- the FP64 pipe is otherwise idle there)
SDDP_DEV void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 1/p from the hardware double-precision seed (MUFU.RCP64H, ~20 bits) and two Newton steps (about 2 ulp).
// The pivot chain of the factorisation is latency critical; a full IEEE division is ~3x longer.
SDDP_DEV double fast_rcp(double p) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(p));
    double e = fma(-p, x, 1.0);
    x = fma(x, e, x);
    e = fma(-p, x, 1.0);
    return fma(x, e, x);
}
// fp32 build: MUFU.RCP (about 1 ulp) and one Newton step
SDDP_DEV float fast_rcp(float p) {
    float x;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(p));
    const float e = fmaf(-p, x, 1.0f);
    return fmaf(x, e, x);
}

#ifndef SDDP_ROW128
#define SDDP_ROW128 1
#endif
// a[i] -= row[i] * s for lo <= i < n; `lo` is a compile-time constant after unrolling.  Every thread of a warp reads
// the same addresses (broadcast).  Loads first, then the FMAs, in batches of 8 (keeps the loads in flight together
// without holding a whole second column in registers).  SDDP_ROW128: 128-bit loads (`row` is 16-byte aligned).
template <int n, int I0 = 0, int I1 = n>
SDDP_DEV void axpy_row(real* a, const real* row, int lo, real s) {
    static_assert(n % 8 == 0 && I0 % 8 == 0 && I1 % 8 == 0, "row length");
#pragma unroll
    for (int i0 = I0; i0 < I1; i0 += 8) {
        real2 c[4];
#pragma unroll
        for (int q = 0; q < 4; q++) if (i0 + 2 * q + 1 >= lo) c[q] = *reinterpret_cast<const real2*>(row + i0 + 2 * q);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i0 + 2 * q;
            if (i >= lo) a[i] -= c[q].x * s;
            if (i + 1 >= lo) a[i + 1] -= c[q].y * s;
        }
    }
}
// Two pivot steps at once (2 x 2 block, see ldlt_warp): a[i] -= row0[i] * s1;  a[i] -= (row1[i] - row0[i] * l) * s2  for lo <= i.
// row1 is column j+1 WITHOUT the update of step j: every lane forms the updated entry itself, exactly as lane j+1 would have
// (one more FMA per entry; applying row1 raw with a combined multiplier s1 - l s2 instead saves it, but the two large terms
// then cancel in the accumulator and the factorisation loses the accuracy the 1e6-weighted blocks need).
template <int n, int I0 = 0, int I1 = n>
SDDP_DEV void axpy2_row(real* a, const real* row0, const real* row1, int lo, real l, real s1, real s2) {
    static_assert(n % 8 == 0 && I0 % 8 == 0 && I1 % 8 == 0, "row length");
#pragma unroll
    for (int i0 = I0; i0 < I1; i0 += 8) {
        real2 c0[4], c1[4];
#pragma unroll
        for (int q = 0; q < 4; q++) if (i0 + 2 * q + 1 >= lo) { c0[q] = *reinterpret_cast<const real2*>(row0 + i0 + 2 * q); c1[q] = *reinterpret_cast<const real2*>(row1 + i0 + 2 * q); }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = i0 + 2 * q;
            if (i >= lo) { a[i] -= c0[q].x * s1; a[i] -= fma(-c0[q].x, l, c1[q].x) * s2; }
            if (i + 1 >= lo) { a[i + 1] -= c0[q].y * s1; a[i + 1] -= fma(-c0[q].y, l, c1[q].y) * s2; }
        }
    }
}
// row[i] = a[i] for lo <= i < n
template <int n>
SDDP_DEV void store_row(real* row, const real* a, int lo) {
#if SDDP_ROW128
#pragma unroll
    for (int i = 0; i < n; i += 2) {
        if (i >= lo) *reinterpret_cast<real2*>(row + i) = make_real2(a[i], a[i + 1]);
        else if (i + 1 >= lo) row[i + 1] = a[i + 1];
    }
#else
#pragma unroll
    for (int i = 0; i < n; i++) if (i >= lo) row[i] = a[i];
#endif
}

// out[b] = sum_a v[a] * (dt Aoo)[a][b],  Aoo = d odot / d o = [[skew(w)/2, w/2], [-w^T/2, 0]];  hw = dt w / 2
SDDP_DEV void contract_Aoo(const real* v, const real* hw, real* out) {
    out[0] = v[1] * hw[2] - v[2] * hw[1] - v[3] * hw[0];
    out[1] = -v[0] * hw[2] + v[2] * hw[0] - v[3] * hw[1];
    out[2] = v[0] * hw[1] - v[1] * hw[0] - v[3] * hw[2];
    out[3] = v[0] * hw[0] + v[1] * hw[1] + v[2] * hw[2];
}
// out[b] = sum_a v[a] * (dt Aow)[a][b],  Aow = d odot / d w = [[(o_w I - skew(o_v))/2], [-o_v^T/2]];  ho = dt o / 2
SDDP_DEV void contract_Aow(const real* v, const real* ho, real* out) {
    out[0] = v[0] * ho[3] - v[1] * ho[2] + v[2] * ho[1] - v[3] * ho[0];
    out[1] = v[0] * ho[2] + v[1] * ho[3] - v[2] * ho[0] - v[3] * ho[1];
    out[2] = -v[0] * ho[1] + v[1] * ho[0] + v[2] * ho[3] - v[3] * ho[2];
}

// d1 as a function of its own (SDDP_D1CALL): a call gives ptxas a fresh register budget for the 24-entry column frame;
// inlined into the node loop, next to ~40 registers of long-lived solver state, the shared-memory loads of a pivot step
// were serialised through one register quad (9.3 K cycles per factorisation against 6.5 K in tools/microbench/ldlt.cu).
// Returns whether a pivot was not positive and finite.  LAT: the fully unrolled form (lowest latency, 21 KB of code).
#ifndef SDDP_D1CALL
#define SDDP_D1CALL 1
#endif
#ifndef SDDP_C3_UNROLL
#define SDDP_C3_UNROLL 3
#define SDDP_C3F_UNROLL 1
#endif
#ifndef SDDP_C2_UNROLL
#define SDDP_C2_UNROLL 19
#endif
#ifndef SDDP_D1R
#define SDDP_D1R 2         // pivot steps per trip of the rolled factorisation (the register frame is shifted once per trip)
#endif
#ifndef SDDP_BULK_PACK
#define SDDP_BULK_PACK 0
#endif
#ifndef SDDP_NO_DMMA
#define SDDP_NO_DMMA 0     // 1: A/B build without tensor cores (vector FP64 register tiles for Wn = Es B, Vxx -= Wn^T Wn, K = -Es^T Wn)
#endif
#ifndef SDDP_ROTATE
#define SDDP_ROTATE 0      // 148: rotate the warp roles of the d phase by blockIdx / 148 (A/B experiment, profiles/README.md)
#endif
#ifndef SDDP_D1SKIP
#define SDDP_D1SKIP 1
#endif
#if SDDP_D1CALL
#define SDDP_D1ATTR __device__ __noinline__
#else
#define SDDP_D1ATTR __device__ __forceinline__
#endif
template <bool LAT, class SMT>
SDDP_D1ATTR bool ldlt_warp(SMT& S, int lane) {
    constexpr int NU = SMT::NU;
    bool bad = false;
    const int t = lane < NU ? lane : NU - 1;
    real npinv = 0.0;                                          // -1 / pivot of this lane's column
    if constexpr (LAT) {
    // the fully unrolled form: lowest latency (small batches), 21 KB of straight-line code
    real a[NU];
#pragma unroll
    for (int i = 0; i < NU; i++) a[i] = S.Quu[i * NU + t];
    __syncwarp();
    real myinv = fast_rcp(a[0]);          // lane 0's pivot
#if SDDP_D1BLOCK
    // 2 x 2 pivot blocks (see the rolled form below: same operations in the same order, so both variants give the same bits)
#pragma unroll
    for (int jb = 0; jb < NU; jb += 2) {
        const real* pr0 = S.Quu + jb * NU;
        const real* pr1 = S.Quu + (jb + 1) * NU;
        const real m0 = a[jb];
        if (lane == jb || lane == jb + 1) {
            if (lane == jb) {
                bad = !(m0 > 0.0) || !isfinite(m0);
                npinv = -myinv;
                S.invp[jb] = myinv;
            }
            store_row<NU>(S.Quu + lane * NU, a, jb);
        }
        __syncwarp();
        const real p0inv = S.invp[jb], c01 = pr0[jb + 1], p1raw = pr1[jb + 1];
        const real l = pr1[jb] * p0inv;      // lane j+1's own multiplier of step j (its copy of the symmetric entry)
        const real p1 = fma(-c01, l, p1raw);
        const real p1inv = fast_rcp(p1);
        const real s1 = (lane == jb) ? real(0.0) : m0 * p0inv;
        const real m1 = fma(-c01, s1, a[jb + 1]);
        a[jb + 1] = m1;
        if (lane == jb + 1) {
            bad = !(p1 > 0.0) || !isfinite(p1);
            npinv = -p1inv;
        }
        const real s2 = (lane == jb + 1) ? real(0.0) : m1 * p1inv;
        if (jb + 2 < NU) {
            a[jb + 2] -= pr0[jb + 2] * s1;
            a[jb + 2] -= fma(-pr0[jb + 2], l, pr1[jb + 2]) * s2;
            myinv = fast_rcp(a[jb + 2]);     // meaningful on lane jb + 2
            axpy2_row<NU>(a, pr0, pr1, jb + 3, l, s1, s2);
        }
    }
#else
#pragma unroll
    for (int j = 0; j < NU; j++) {
        if (lane == j) {
            const real p = a[j];
            bad = !(p > 0.0) || !isfinite(p);
            npinv = -myinv;
            S.invp[j] = myinv;
            store_row<NU>(S.Quu + j * NU, a, j + 1);
        }
        __syncwarp();
        if (j + 1 < NU) {
            const real sj = (lane == j) ? 0.0 : S.invp[j] * a[j];
            a[j + 1] -= S.Quu[j * NU + j + 1] * sj;
            myinv = fast_rcp(a[j + 1]);     // meaningful on lane j+1
            if (j + 2 < NU) axpy_row<NU>(a, S.Quu + j * NU, j + 2, sj);
        }
    }
#endif
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NU; i += 2) {
        const real v0 = (i > t) ? npinv * a[i] : (i == t ? real(1.0) : 0.0);
        const real v1 = (i + 1 > t) ? npinv * a[i + 1] : (i + 1 == t ? real(1.0) : 0.0);
        if (lane < NU) { S.Quu[i * NU + lane] = v0; S.Quu[(i + 1) * NU + lane] = v1; }
    }
    } else {
    // Rolled over the pivot steps (R per loop trip): the kernel is instruction-fetch bound (profiles/README.md), and
    // the fully unrolled factorisation was 21 KB of straight-line code per node.  The column lives in a rotating
    // register frame b[q] = a_t[jb + q] (shifted down by R at the end of a trip, zero filled), so every index is
    // static; the column of the pivot goes to a ping-pong row in the same frame, and row j of Et = Lt^-1 is stored
    // at step j, when lane t < j holds its final E[j][t] = -a_t[j] / pivot_t (the multiplier of that step), which
    // replaces the separate write-out pass.  The row scaling D^-1/2 of Es = D^-1/2 Et is applied where Es is used
    // (phases h and g).  The frame is always updated whole: its zero tail costs a few FMAs in the shadow of the
    // pivot chain, and a step stays one basic block.
    // (Measured and rejected: nobody publishes a column, every lane stores its own a_i[j] = col_j[i] by symmetry.  25 % fewer
    //  cycles per step, tools/microbench/ldlt.cu V10 / V16 -- but the elimination then loses the scaling invariance of
    //  symmetric LDL^T: at the first SRBD node behind a LIP-style tail, cond(Quu) = 2e9, the gains come out at 2e-8
    //  instead of 1e-13.)
    real b[NU];
#pragma unroll
    for (int i = 0; i < NU; i++) b[i] = S.Quu[i * NU + t];
    __syncwarp();
    real myinv = fast_rcp(b[0]);                               // lane 0's pivot
#if SDDP_D1BLOCK
    // One 2 x 2 pivot BLOCK per trip (round 2, tools/microbench/ldlt.cu V21 against V15): lanes j and j+1 publish their columns together -- column j final, column j+1 without the update of
    // step j -- and after ONE warp sync every lane forms l = c' / p_j (c' = lane j+1's own copy of the symmetric entry:
    // with it every lane reproduces lane j+1's update of its column bit for bit, which keeps the scaling invariance of
    // symmetric LDL^T that the 1e6-weighted blocks need) and the Schur pivot p1 = p_{j+1} - c l itself
    // (c = col_j[j+1]), its two multipliers s1 = a[j] / p_j and s2 = (a[j+1] - c s1) / p1, and applies both steps at once:
    //   a[i] -= col_j[i] s1;  a[i] -= (col_{j+1}[i] - col_j[i] l) s2      (axpy2_row).
    // Same elimination with the same shared-memory loads and one more FMA per entry, but one publish -> sync -> load round
    // trip (the latency chain of a step) per two pivots instead of two.  Lane j skips its own step (s1 = 0), lane j+1 its own
    // (s2 = 0), and both keep taking part afterwards (their registers then carry the columns of E, see above).
    static_assert(NU % 2 == 0, "d1 frame");
    int par = 0;
#pragma unroll 1
    for (int jb = 0; jb < NU; jb += 2, par ^= 2) {
        const real* pr0 = S.prow[par];
        const real* pr1 = S.prow[par + 1];
        const real m0 = b[0];
        if (lane == jb || lane == jb + 1) {
            if (lane == jb) {
                bad = !(m0 > 0.0) || !isfinite(m0);
                npinv = -myinv;
                S.invp[jb] = myinv;
            }
            real* pr = S.prow[par + (lane - jb)];
            store_row<8>(pr, b, 0);
            if (jb < 16) store_row<8>(pr + 8, b + 8, -8);
            if (jb < 8) store_row<8>(pr + 16, b + 16, -16);
        }
        __syncwarp();
        const real p0inv = S.invp[jb], c01 = pr0[1], p1raw = pr1[1];
        const real l = pr1[0] * p0inv;       // lane j+1's own multiplier of step j (its copy of the symmetric entry)
        const real p1 = fma(-c01, l, p1raw);
        const real p1inv = fast_rcp(p1);
        const real s1 = (lane == jb) ? real(0.0) : m0 * p0inv;
        const real m1 = fma(-c01, s1, b[1]);
        if (lane == jb + 1) {
            bad = !(p1 > 0.0) || !isfinite(p1);
            npinv = -p1inv;
        }
        const real s2 = (lane == jb + 1) ? real(0.0) : m1 * p1inv;
        // the element that becomes the next pivot first, its reciprocal started at once (meaningful on lane jb + 2)
        b[2] -= pr0[2] * s1;
        b[2] -= fma(-pr0[2], l, pr1[2]) * s2;
        myinv = fast_rcp(b[2]);
        // chunks of 8 entries wholly in the zero tail of the frame (q >= NU - jb) are skipped (warp-uniform)
        axpy2_row<NU, 0, 8>(b, pr0, pr1, 3, l, s1, s2);
        if (jb < 16) axpy2_row<NU, 8, 16>(b, pr0, pr1, 3, l, s1, s2);
        if (jb < 8) axpy2_row<NU, 16, 24>(b, pr0, pr1, 3, l, s1, s2);
        // rows j and j+1 of Et: (t < row) -multiplier / pivot_t, (t == row) 1, (t > row) 0
        const real e0 = (lane < jb) ? npinv * m0 : (lane == jb ? real(1.0) : real(0.0));
        const real e1 = (lane < jb + 1) ? npinv * m1 : (lane == jb + 1 ? real(1.0) : real(0.0));
        if (lane < NU) { S.Quu[jb * NU + lane] = e0; S.Quu[(jb + 1) * NU + lane] = e1; }
#pragma unroll
        for (int q = 0; q < NU; q++) b[q] = (q + 2 < NU) ? b[q + 2] : real(0.0);
    }
#else
    constexpr int R = SDDP_D1R;
    static_assert(NU % R == 0 && R % 2 == 0, "d1 frame");
#pragma unroll 1
    for (int jb = 0; jb < NU; jb += R) {
#pragma unroll
        for (int s_ = 0; s_ < R; s_++) {
            const int j = jb + s_;
            real* pr = S.prow[s_ & 1];                         // (R is even: j & 1 == s_ & 1)
            const real m = b[s_];
            if (lane == j) {
                bad = !(m > 0.0) || !isfinite(m);
                npinv = -myinv;
                S.invp[j] = myinv;
#if SDDP_D1SKIP
                store_row<8>(pr, b, s_ + 1);
                if (jb < 16) store_row<8>(pr + 8, b + 8, s_ + 1 - 8);
                if (jb < 8) store_row<8>(pr + 16, b + 16, s_ + 1 - 16);
#else
                store_row<NU>(pr, b, s_ + 1);
#endif
            }
            __syncwarp();
            const real sj = (lane == j) ? 0.0 : S.invp[j] * m;
            b[s_ + 1] -= pr[s_ + 1] * sj;
            myinv = fast_rcp(b[s_ + 1]);                         // meaningful on lane j + 1
#if SDDP_D1SKIP
            // chunks of 8 entries wholly in the zero tail of the frame (q >= NU - jb) are skipped (warp-uniform)
            axpy_row<NU, 0, 8>(b, pr, s_ + 2, sj);
            if (jb < 16) axpy_row<NU, 8, 16>(b, pr, s_ + 2, sj);
            if (jb < 8) axpy_row<NU, 16, 24>(b, pr, s_ + 2, sj);
#else
            axpy_row<NU>(b, pr, s_ + 2, sj);
#endif
            // row j of Et: (t < j) -multiplier / pivot_t, (t == j) 1, (t > j) 0
            const real ev = (lane < j) ? npinv * m : (lane == j ? real(1.0) : 0.0);
            if (lane < NU) S.Quu[j * NU + lane] = ev;
        }
#pragma unroll
        for (int q = 0; q < NU; q++) b[q] = (q + R < NU) ? b[q + R] : 0.0;
    }
#endif
    }
    if (lane < NU) S.rs[lane] = sqrt(-npinv);
    return bad;
}

template <class MT, bool LAT>
__device__ int SmemSrbdT<MT, LAT>::backward(const DevCfg& c, SmemSrbdT<MT, LAT>& S, const real* X, const real* U, const real* P, const real* D,
                                       const real* packs, real mu, real* Kg, real* kg, double* dV3, bool has_gap, int tid) {
    using M = MT;
    using NBL = NodeBuf<MT>;
    constexpr int NZ = MT::NZ;
    const int N = c.N, lane = tid & 31, warp = tid >> 5;
    const bool fixed = c.rho_fixed > 0.0;
    const real rho_b = fixed ? (real)c.rho_fixed : real(1.0);
    const real dt = c.dt;
    SyncBlock sync;

    // (the 2 KB pack as one bulk copy: measured 1 % slower than 128 16-byte cp.async here -- every thread then polls a barrier
    //  at every node -- so it is off by default; the 7.1 KB gain tile of the forward pass keeps its bulk copy)
    const bool bulk = SDDP_BULK_PACK && M::PACK % RV == 0 && ((((size_t)packs) | ((size_t)(S.nb[0] + NBL::OK)) | ((size_t)(S.nb[1] + NBL::OK))) & 15) == 0;
    auto prefetch = [&](int k) {       // node k -> buffer k & 1 (base pointers from shared memory: see forward_wave)
        real* nb = S.nb[k & 1];
        const real* Xs = S.gp[2] + (size_t)k * NX;
        const real* Ds = S.gp[3] + (size_t)k * NX;
        for (int i = tid; i < NX; i += NT) {
            cp_async8(nb + NBL::OX + i, Xs + i);
            if (has_gap) cp_async8(nb + NBL::OD + i, Ds + i);
        }
        const real* Us = S.gp[4] + (size_t)k * NU;
        for (int i = tid; i < NU; i += NT) cp_async8(nb + NBL::OU + i, Us + i);
        const real* Ps = S.gp[5] + (size_t)k * NP;
        for (int i = tid; i < NP; i += NT) cp_async8(nb + NBL::OP + i, Ps + i);
        const real* ps = S.gp[6] + (size_t)k * M::PACK;
        if (bulk) { if (tid == 0) bulk_g2s(nb + NBL::OK, ps, M::PACK * (int)sizeof(real), &S.mbar[k & 1]); }      // the 2 KB pack: one bulk copy
        else if (M::PACK % RV == 0 && ((((size_t)ps) | ((size_t)(nb + NBL::OK))) & 15) == 0) { for (int i = RV * tid; i < M::PACK; i += RV * NT) cp_async16(nb + NBL::OK + i, ps + i); }
        else { for (int i = tid; i < M::PACK; i += NT) cp_async8(nb + NBL::OK + i, ps + i); }
        cp_commit();
    };
    // buffer k & 1 is used by nodes N-1, N-3, ... (or N-2, N-4, ...): phase parity of its barrier at node k
    auto landed = [&](int k) { if (bulk) mbar_wait(&S.mbar[k & 1], ((N - 1 - k) >> 1) & 1); };

    // terminal node: Vx = l_Nx, Vxx = l_Nxx (ddp.py:216-226: costs only)
    __syncthreads();
    if (tid == 0) {
        S.gp[0] = Kg; S.gp[1] = kg; S.gp[2] = X; S.gp[3] = D; S.gp[4] = U; S.gp[5] = P; S.gp[6] = packs;
        if (bulk) { mbar_init(&S.mbar[0], 1); mbar_init(&S.mbar[1], 1); fence_async_proxy(); }
    }
    __syncthreads();
    {
        real* nb = S.nb[N & 1];
        for (int i = tid; i < NX; i += NT) nb[NBL::OX + i] = X[(size_t)N * NX + i];
        for (int i = tid; i < NP; i += NT) nb[NBL::OP + i] = P[(size_t)N * NP + i];
    }
    prefetch(N - 1);
    if (tid == 0) { S.red[R_TOT] = 0.0; S.red[R_ACC1] = 0.0; S.red[R_ACC2] = 0.0; S.iflag[1] = 0; S.Qx[NX] = 0.0; S.qxy[NX] = 0.0; }
    __syncthreads();
    M::template expand<LDW>(c, NODE_TERM, S.nb[N & 1] + NBL::OX, nullptr, S.nb[N & 1] + NBL::OP, nullptr, S.Vx, S.Qu, S.VT, S.W, S.Quu, tid, NT, sync, S.escr);
    for (int i = tid; i < NX; i += NT) S.y[i] = S.Vx[i];
    cp_wait_all();
    landed(N - 1);
    __syncthreads();                           // node N-1 landed
    PROF(19);

    const unsigned long long c1d = __ldg(c.ztab + ZT_C1OFF + tid);     // this thread's Quu entries (see c1)
    for (int k = N - 1; k >= 0; k--) {
        const int kind = node_kind(c, k);
        real* nb = S.nb[k & 1];
        const real* xk = nb + NBL::OX;
        const real* uk = nb + NBL::OU;
        const real* pk = nb + NBL::OP;
        real* cg = nb + NBL::OD;
        const real* pack = nb + NBL::OK;
        STAMP(0);                              // node k has landed and everyone is done with node k+1: see the barrier that
        STAMP(1);                              // ends the f / g phase (and the one in front of the loop)
        PROF(8);
        const real* Jac = pack + M::PK_JAC;
        PROF(9);
        // ---- c1: everything that needs Vxx' itself: gap shift, Quu = luu + fu^T Vxx' fu + mu I (written whole: luu of
        //          the wdot block is 2 gq Jac_f^T Jac_f, of the affine residuals a few constants, prb.py:200-204)
        if (tid < NX) {
            real s = 0.0;
            if (has_gap) {
#pragma unroll 4
                for (int j = 0; j < NX; j++) s += S.VT[tid * NX + j] * cg[j];
                s *= rho_b;
            }
            S.sv[tid] = s;
            S.vp[tid] = S.Vx[tid] + s;
            S.ys[tid] = fixed ? S.y[tid] + s : S.y[tid];
        }
        {
            // u = (cddot_i, f_i) x 4: "c" index ci = 3 i + r -> u index 6 i + r, "f" index fj = 3 j + k -> 6 j + 3 + k.
            // B^T rows: cddot -> e(cd); f_jk -> (fs/m) e(rd_k) + sum_q G_j[q][k] e(w_q),  G_j[q][k] = Jac[q][ZF + fj].
            // Threads 0..77 take one (f, f) entry (16 products) and one (cddot, cddot) entry, threads 78..127 three of
            // the remaining (cddot, f) / (cddot, cddot) entries (4 or 1 products): the assignment is a host-built
            // table (sddp.cu:build_ztab), four 16-bit descriptors per thread.
            const real dt2 = dt * dt, g2 = real(2.0) * c.gq;
            if (SDDP_INEQ_ON(c)) {          // inequality barriers (extension, off by default): 3 x 3 Hessian per foot -> escr
                if (tid < 4) {
                    double val, cg_[3], ch_[6];
                    M::cone_terms_cold(c, uk + 6 * tid + 3, val, cg_, ch_);
#pragma unroll
                    for (int q = 0; q < 6; q++) S.escr[6 * tid + q] = ch_[q];
                }
                __syncthreads();
            }
            if (tid < 78) {
                const int fa = C1_I1(c1d), fb = C1_I2(c1d);
                const int ka = fa % 3, kb = fb % 3;
                const real* Ga = Jac + M::ZF + fa;
                const real* Gb = Jac + M::ZF + fb;
                const real a1 = Ga[0], a2 = Ga[NZ], a3 = Ga[2 * NZ], b1 = Gb[0], b2 = Gb[NZ], b3 = Gb[2 * NZ], im = c.inv_ms;
                const real* Vr = S.VT + (M::XRD + ka) * NX;
                const real* Vw = S.VT + M::XW * NX;
                const int cr = M::XRD + kb, cw = M::XW;
                const real t0 = Vr[cr] * im + Vr[cw] * b1 + Vr[cw + 1] * b2 + Vr[cw + 2] * b3;
                const real t1 = Vw[cr] * im + Vw[cw] * b1 + Vw[cw + 1] * b2 + Vw[cw + 2] * b3;
                const real t2 = Vw[NX + cr] * im + Vw[NX + cw] * b1 + Vw[NX + cw + 1] * b2 + Vw[NX + cw + 2] * b3;
                const real t3 = Vw[2 * NX + cr] * im + Vw[2 * NX + cw] * b1 + Vw[2 * NX + cw + 1] * b2 + Vw[2 * NX + cw + 2] * b3;
                real v = dt2 * (im * t0 + a1 * t1 + a2 * t2 + a3 * t3) + g2 * (a1 * b1 + a2 * b2 + a3 * b3);
                if (ka == kb) {                       // rddot rows of min_qddot; min_f and f_active on the diagonal
                    v += g2 * im * im;
                    if (fa == fb) { const real sw1 = real(1.0) - pk[8 + 2 * (fa / 3)]; v += real(2.0) * (c.w_minf + c.w_fsw * sw1 * sw1) + mu; }
                }
                if (SDDP_INEQ_ON(c) && fa / 3 == fb / 3) v += S.escr[6 * (fa / 3) + M::cone_hidx(ka, kb)];      // inequality barriers of the foot
                const int ua = 6 * (fa / 3) + 3 + ka, ub = 6 * (fb / 3) + 3 + kb;
                S.Quu[ua * NU + ub] = v;
                S.Quu[ub * NU + ua] = v;
            }
#pragma unroll 1
            for (int sl = 1; sl < 4; sl++) {
                const unsigned d = (unsigned)(c1d >> (16 * sl)) & 0xffffu;
                const int ty = C1_TYPE(d), i1 = C1_I1(d), i2 = C1_I2(d);
                if (ty == 0) continue;
                int ua, ub;
                real v;
                if (ty == 2) {                        // (cddot ci, f fj)
                    const int kb = i2 % 3;
                    const real* G = Jac + M::ZF + i2;
                    const real* Vc = S.VT + (M::XCD + i1) * NX;
                    v = dt2 * (Vc[M::XRD + kb] * c.inv_ms + Vc[M::XW] * G[0] + Vc[M::XW + 1] * G[NZ] + Vc[M::XW + 2] * G[2 * NZ]);
                    ua = 6 * (i1 / 3) + i1 % 3; ub = 6 * (i2 / 3) + 3 + kb;
                } else {                              // (cddot ca, cddot cb)
                    v = dt2 * S.VT[(M::XCD + i1) * NX + M::XCD + i2];
                    if (i1 == i2) v += g2 + mu;      // cddot rows of min_qddot
                    ua = 6 * (i1 / 3) + i1 % 3; ub = 6 * (i2 / 3) + i2 % 3;
                }
                S.Quu[ua * NU + ub] = v;
                S.Quu[ub * NU + ua] = v;
            }
        }
        if (k > 0) prefetch(k - 1);
        STAMP(2);
        __syncthreads();
        STAMP(3);
        PROF(10);

        const real* o = xk + M::XO;
        const real* w = xk + M::XW;
#if SDDP_ROTATE
        // Role rotation (experiment): the factorisation warp of co-resident CTAs sits on different SM sub-partitions
        const int vwarp = (warp + 4 - ((blockIdx.x / SDDP_ROTATE) & 3)) & 3, vt = vwarp * 32 + lane - 32;
#else
        const int vwarp = warp, vt = tid - 32;
#endif
        if (vwarp == 0) {
            // ---- d1: Quu + mu I = Lt D Lt^T; lane t owns column t.  Es = D^-1/2 Lt^-1 overwrites S.Quu row by row.
            if (has_gap) {   // gap terms of the model (the only use of cg after c1)
                real g1 = 0, g2 = 0, yg = 0;
                for (int i = lane; i < NX; i += 32) { g1 += S.Vx[i] * cg[i]; g2 += cg[i] * S.sv[i]; yg += S.y[i] * cg[i]; }
                g1 = rho_b * warp_sum(g1); g2 = rho_b * warp_sum(g2); yg = rho_b * warp_sum(yg);
                if (lane == 0) { S.red[R_G1] = g1; S.red[R_G2] = g2; S.red[R_YG] = yg; }
            } else if (lane == 0) { S.red[R_G1] = 0.0; S.red[R_G2] = 0.0; S.red[R_YG] = 0.0; }
            // Lane t holds column t.  Step j: lane j publishes its (final) column raw and 1/pivot, 1/sqrt(pivot); every other
            // lane then applies  a[i] -= col_j[i] * (a[j] / pivot_j),  i > j.  The element that becomes the next pivot
            // (i = j+1) is updated first and its reciprocal started at once, so the rest of the update overlaps the
            // reciprocal latency.  Lane j skips step j and keeps taking part afterwards: its registers then carry
            // -pivot_j times column j of E = Lt^-1 (E[:,j] starts as -col_j / pivot_j and obeys the same linear recurrence),
            // so the inverse factor costs no extra arithmetic in lanes the factorisation no longer needs.
            const bool bad = ldlt_warp<LAT>(S, lane);
            if (__any_sync(FULL, bad) && lane == 0) S.iflag[1] = 1;
            PROF_T(14, 0);
            STAMP(4); STAMP(5); STAMP(6); STAMP(7);
        } else {
            // ---- c2: T = Vxx' fx = V + dt V A, in place, one thread per row (warps 1-2)
            const int r_ = vt;
            if (r_ < NX) {
                real* row = S.VT + r_ * NX;
                real vr[3], vo[4], vw[3];
#pragma unroll
                for (int q = 0; q < 3; q++) { vr[q] = row[M::XR + q]; vw[q] = row[M::XW + q]; }
#pragma unroll
                for (int q = 0; q < 4; q++) vo[q] = row[M::XO + q];
                const real hw[3] = {real(0.5) * dt * w[0], real(0.5) * dt * w[1], real(0.5) * dt * w[2]};
                const real ho[4] = {real(0.5) * dt * o[0], real(0.5) * dt * o[1], real(0.5) * dt * o[2], real(0.5) * dt * o[3]};
                real to[4], tw[3];
                contract_Aoo(vo, hw, to);     // (V dt Aoo): o columns
                contract_Aow(vo, ho, tw);     // (V dt Aow): w columns
                const real dw0 = dt * vw[0], dw1 = dt * vw[1], dw2 = dt * vw[2];
                // cd and rd columns first (they read the c and r entries before those are overwritten)
                constexpr int C2U = LAT ? 19 : SDDP_C2_UNROLL;
#pragma unroll C2U
                for (int q = 0; q < 12; q++) row[M::XCD + q] += dt * row[M::XC + q];
#pragma unroll
                for (int q = 0; q < 3; q++) row[M::XRD + q] += dt * vr[q];
#pragma unroll C2U
                for (int z = 0; z < 19; z++) {        // r, o, c columns: + dt vw . dwdot/dz
                    real v = row[z] + dw0 * Jac[z] + dw1 * Jac[NZ + z] + dw2 * Jac[2 * NZ + z];
                    if (z >= 3 && z < 7) v += (z == 3 ? to[0] : (z == 4 ? to[1] : (z == 5 ? to[2] : to[3])));
                    row[z] = v;
                }
#pragma unroll
                for (int q = 0; q < 3; q++)
                    row[M::XW + q] = vw[q] + tw[q] + dw0 * Jac[M::ZW + q] + dw1 * Jac[NZ + M::ZW + q] + dw2 * Jac[2 * NZ + M::ZW + q];
            }
            PROF_T(15, 32);
            STAMP(4);
            bar_named(2, 96);
            // ---- c3: columns of fx^T (.) and fu^T (.) applied to T (j < 37), v+ (j = 37) and ys (j = 38); warps 1-3.
            //          (throughput variant: the 12-row loops are only partly unrolled, the kernel is bound by instruction fetch)
            constexpr int C3U = LAT ? 12 : SDDP_C3_UNROLL, C3F = LAT ? 4 : SDDP_C3F_UNROLL;
            //          Plain stores: lxx, lux, lx, lu are added afterwards (expand MODE 1)
            for (int task = vt; task < 39 * 3; task += 96) {
                const int j = task % 39, g = task / 39;
                const real* col = (j < NX) ? S.VT + j : (j == NX ? S.vp : S.ys);
                const int cs = (j < NX) ? NX : 1;      // stride between rows of this "column"
                real* oxx = (j < NX) ? S.Qxx + j : (j == NX ? S.Qx : S.qxy);
                const int os = (j < NX) ? NX : 1;
                // tw = dt T[w, j] (+ 2 gq Jac[:, z_j] when column j is one of the wdot arguments r, o, c, w): the second
                // term makes  Jac[:, z]^T tw  deliver the 2 gq Jac^T Jac part of lxx / lux on top of dt A^T T, so that
                // only the curvature entries of the wdot block are left for the descriptor table (phase e)
                const int zj = (j < M::XRD) ? j : ((j >= M::XW && j < M::XCD) ? j - 3 : -1);
                real tw0 = dt * col[(M::XW + 0) * cs], tw1 = dt * col[(M::XW + 1) * cs], tw2 = dt * col[(M::XW + 2) * cs];
                if (zj >= 0) { const real g2 = real(2.0) * c.gq; tw0 += g2 * Jac[zj]; tw1 += g2 * Jac[NZ + zj]; tw2 += g2 * Jac[2 * NZ + zj]; }
                if (g == 0) {          // rows r, o, rd, w
                    real tov[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) tov[q] = col[(M::XO + q) * cs];
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        oxx[(M::XR + q) * os] = col[(M::XR + q) * cs] + (Jac[M::ZR + q] * tw0 + Jac[NZ + M::ZR + q] * tw1 + Jac[2 * NZ + M::ZR + q] * tw2);
                        oxx[(M::XRD + q) * os] = col[(M::XRD + q) * cs] + dt * col[(M::XR + q) * cs];
                    }
                    const real hw[3] = {real(0.5) * dt * w[0], real(0.5) * dt * w[1], real(0.5) * dt * w[2]};
                    const real ho[4] = {real(0.5) * dt * o[0], real(0.5) * dt * o[1], real(0.5) * dt * o[2], real(0.5) * dt * o[3]};
                    real ao[4], aw[3];
                    contract_Aoo(tov, hw, ao);    // (dt Aoo)^T T[o,j]: row o_b = sum_a Aoo[a][b] T[o_a][j]
                    contract_Aow(tov, ho, aw);    // (dt Aow)^T T[o,j]: row w_b
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        oxx[(M::XO + q) * os] = col[(M::XO + q) * cs] + ao[q] + (Jac[M::ZO + q] * tw0 + Jac[NZ + M::ZO + q] * tw1 + Jac[2 * NZ + M::ZO + q] * tw2);
#pragma unroll
                    for (int q = 0; q < 3; q++)
                        oxx[(M::XW + q) * os] = col[(M::XW + q) * cs] + aw[q] + (Jac[M::ZW + q] * tw0 + Jac[NZ + M::ZW + q] * tw1 + Jac[2 * NZ + M::ZW + q] * tw2);
                } else if (g == 1) {   // rows c, cd
#pragma unroll C3U
                    for (int q = 0; q < 12; q++) {
                        real tc = col[(M::XC + q) * cs];
                        oxx[(M::XC + q) * os] = tc + (Jac[M::ZC + q] * tw0 + Jac[NZ + M::ZC + q] * tw1 + Jac[2 * NZ + M::ZC + q] * tw2);
                        oxx[(M::XCD + q) * os] = col[(M::XCD + q) * cs] + dt * tc;
                    }
                } else {               // fu^T (.): rows cddot_i, f_i
                    real* oux = S.W + j;           // columns NX, NX+1 of W: Qu = lu + fu^T v+ and quy = lu + fu^T ys
                    const int us = LDW;
                    const real trd[3] = {c.inv_ms * col[(M::XRD + 0) * cs], c.inv_ms * col[(M::XRD + 1) * cs], c.inv_ms * col[(M::XRD + 2) * cs]};
#pragma unroll C3F
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int q = 0; q < 3; q++) {
                            oux[(6 * i + q) * us] = dt * col[(M::XCD + 3 * i + q) * cs];
                            const int zf = M::ZF + 3 * i + q;
                            oux[(6 * i + 3 + q) * us] = dt * trd[q] + (Jac[zf] * tw0 + Jac[NZ + zf] * tw1 + Jac[2 * NZ + zf] * tw2);
                        }
                }
            }
            PROF_T(16, 32);
            STAMP(5);
            bar_named(2, 96);
            // ---- e: + lx, lu, lxx, lux of the node (one pass, disjoint destinations, no barrier inside)
            M::apply_rec(c, kind, xk, uk, pk, pack, S.Qxx, S.Qx, S.qxy, S.W + NX, S.W + NX + 1, vt);
            PROF_T(17, 32);
            STAMP(6);
        }
        if (vwarp != 0) STAMP(7);
        __syncthreads();
        STAMP(8);
        PROF(11);
        if (S.iflag[1]) {      // a pivot was not positive: leave (the prefetch of node k-1 is in flight: wait for it first)
            __syncthreads();
            cp_wait_all();
            if (k > 0) landed(k - 1);
            __syncthreads();                   // nobody polls a barrier any more
            if (tid == 0) { S.iflag[1] = 0; if (bulk) { mbar_inval(&S.mbar[0]); mbar_inval(&S.mbar[1]); } }
            __syncthreads();
            return k + 1;
        }

        // ---- h: Wn = Es B = diag(rs) Et B (B = [Qux | Qu | quy], 24 x 40 in S.W; Et lower triangular in S.Quu, entries
        //         above the diagonal are zero).  In place: a warp owns whole 8-column blocks
        //         (warp 0: blocks 0 and 4) and reads all of a block before it writes; the three row tiles of a block
        //         are independent DMMA chains of 2, 4 and 6 steps.
#if SDDP_NO_DMMA
        // (A/B build, profiles/README.md: the three dense products of a node on the vector FP64 pipe, register tiles fed
        //  from shared memory, no tensor cores.)  h: one thread per (column, group of 8 rows); the column of B goes to
        //  registers first because the product is formed in place.
        {
            const int c_ = tid % 40, g_ = tid / 40;
            real bcol[NU];
            if (tid < 120) {
#pragma unroll
                for (int l = 0; l < NU; l++) bcol[l] = S.W[l * LDW + c_];
            }
            __syncthreads();
            if (tid < 120) {
#pragma unroll 1
                for (int ii = 0; ii < 8; ii++) {
                    const int i = 8 * g_ + ii;
                    real s_ = 0.0;
#pragma unroll
                    for (int l = 0; l < NU; l++) if (l <= i) s_ += S.Quu[i * NU + l] * bcol[l];
                    S.W[i * LDW + c_] = S.rs[i] * s_;
                }
            }
        }
#else
        {
            const int fr = lane >> 2, fc = lane & 3;
            for (int J = warp; J < 5; J += NWARP) {
                double h0[3] = {0, 0, 0}, h1[3] = {0, 0, 0};
#pragma unroll
                for (int k0 = 0; k0 < NU; k0 += 4) {
                    const int l = k0 + fc;
                    const real bv = S.W[l * LDW + 8 * J + fr];
#pragma unroll
                    for (int I = 0; I < 3; I++) {
                        if (k0 > 8 * I + 7) continue;
                        const int i = 8 * I + fr;
                        const real av = S.Quu[i * NU + l];
                        dmma884(h0[I], h1[I], av, bv);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int I = 0; I < 3; I++) {
                    const real r = S.rs[8 * I + fr];
                    *reinterpret_cast<real2*>(S.W + (8 * I + fr) * LDW + 8 * J + 2 * fc) = make_real2(r * h0[I], r * h1[I]);
                }
            }
        }
#endif
        STAMP(11);
        __syncthreads();
        PROF(12);
        // ---- f: [Vxx Vx y] = [sym(Qxx) Qx qxy] - Wn^T Wn: upper-triangular 8x8 tiles of the 40x40 product, K = 24 in
        //         six DMMA steps.  Results go to VT (T is dead), Vx, y (column 38 = Wn^T Es quy = -K^T quy),
        //         Vx[37] = -|w0|^2 and y[37] = quy . k.
        // ---- g: [K | k] = -Es^T Wn: 3 x 5 tiles (rows i, columns c); Es is lower triangular, so row tile I starts
        //         at k0 = 8 I.
#if SDDP_NO_DMMA
        {
            // f: 2 x 4 register tiles of the upper triangle of the 40 x 40 product Wn^T Wn (110 tiles, one per thread)
            if (tid < 110) {
                int pr_ = 0, rem = tid;
                while (rem >= 2 * (10 - pr_)) { rem -= 2 * (10 - pr_); pr_++; }
                const int ri = 2 * pr_ + (rem >= 10 - pr_ ? 1 : 0), cj = pr_ + (rem >= 10 - pr_ ? rem - (10 - pr_) : rem);
                real acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll 4
                for (int l = 0; l < NU; l++) {
                    const real2 a = *reinterpret_cast<const real2*>(S.W + l * LDW + 2 * ri);
                    const real2 b0 = *reinterpret_cast<const real2*>(S.W + l * LDW + 4 * cj), b1 = *reinterpret_cast<const real2*>(S.W + l * LDW + 4 * cj + 2);
                    acc[0][0] += a.x * b0.x; acc[0][1] += a.x * b0.y; acc[0][2] += a.x * b1.x; acc[0][3] += a.x * b1.y;
                    acc[1][0] += a.y * b0.x; acc[1][1] += a.y * b0.y; acc[1][2] += a.y * b1.x; acc[1][3] += a.y * b1.y;
                }
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int gi = 2 * ri + r, gj = 4 * cj + q;
                        const bool m = gj < NX, ok = gj >= gi && gj <= NX + 1 && gi <= NX;
                        const real* s1 = m ? S.Qxx + gi * NX + gj : (gj == NX ? S.Qx : S.qxy) + gi;
                        const real* s2 = m ? S.Qxx + gj * NX + gi : s1;
                        real* d1 = m ? S.VT + gi * NX + gj : (gj == NX ? S.Vx : S.y) + gi;
                        real* d2 = m ? S.VT + gj * NX + gi : d1;
                        if (ok) {
                            const real v = real(0.5) * (*s1 + *s2) - acc[r][q];
                            *d1 = v;
                            *d2 = v;
                        }
                    }
            }
            // g: [K | k] = -Et^T (diag(rs) Wn), 2 x 4 register tiles (12 x 10, one per thread); Et is lower triangular
            if (tid < 120) {
                const int I2 = tid / 10, J4 = tid - 10 * I2;
                real acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
                for (int l = 2 * I2; l < NU; l++) {
                    const real2 e_ = *reinterpret_cast<const real2*>(S.Quu + l * NU + 2 * I2);
                    const real r_ = S.rs[l];
                    const real2 b0 = *reinterpret_cast<const real2*>(S.W + l * LDW + 4 * J4), b1 = *reinterpret_cast<const real2*>(S.W + l * LDW + 4 * J4 + 2);
                    const real w0 = r_ * b0.x, w1 = r_ * b0.y, w2 = r_ * b1.x, w3 = r_ * b1.y;
                    acc[0][0] += e_.x * w0; acc[0][1] += e_.x * w1; acc[0][2] += e_.x * w2; acc[0][3] += e_.x * w3;
                    acc[1][0] += e_.y * w0; acc[1][1] += e_.y * w1; acc[1][2] += e_.y * w2; acc[1][3] += e_.y * w3;
                }
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int i = 2 * I2 + r, cc = 4 * J4 + q;
                        const real kv = -acc[r][q];
                        if (cc < NX) Kg[((size_t)k * NU + i) * NX + cc] = kv;
                        if (cc == NX) { S.kk[i] = kv; kg[(size_t)k * NU + i] = kv; }
                    }
            }
        }
#else
        if constexpr (!LAT) {
        // Warp w takes tiles w, w+4, w+8, w+12 of each product, one after the other through the same code (a rolled loop:
        // the kernel is bound by instruction fetch, not by the latency of the DMMA chains, and four interleaved copies of
        // the chains and epilogues were 12 KB of straight-line code per warp); the f and the g chain of a slot run
        // interleaved.
        {
            const int fr = lane >> 2, fc = lane & 3;       // fragment row / column of this lane
#pragma unroll 1
            for (int t = warp; t < 15; t += NWARP) {
                // tile t of the upper triangle of the 5 x 5 grid, row-major: (fI, fJ) from two packed tables (4 bits each)
                const int fI = (int)((0x433222111100000ull >> (4 * t)) & 15), fJ = (int)((0x443432432143210ull >> (4 * t)) & 15);
                const int gI = t / 5, gJ = t - 5 * gI;
                double f0 = 0.0, f1 = 0.0, g0 = 0.0, g1 = 0.0;
                const real* rf = S.W + fc * LDW + fr;
                const real* qa = S.Quu + fc * NU + 8 * gI + fr;
#pragma unroll
                for (int k0 = 0; k0 < NU; k0 += 4) {
                    dmma884(f0, f1, rf[k0 * LDW + 8 * fI], rf[k0 * LDW + 8 * fJ]);
                    // Es^T Wn = Et^T (diag(rs) Wn); Et is lower triangular: row tile gI starts at k0 = 8 gI
                    if (k0 >= 8 * gI) dmma884(g0, g1, qa[k0 * NU], S.rs[k0 + fc] * rf[k0 * LDW + 8 * gJ]);
                }
                const int gi = 8 * fI + fr;
                if (fJ < 4) {
                    // interior tile (rows and columns < 32): every entry is a matrix entry (on a diagonal tile the upper half)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int gj = 8 * fJ + 2 * fc + e;
                        if (gj >= gi) {
                            const real v = real(0.5) * (S.Qxx[gi * NX + gj] + S.Qxx[gj * NX + gi]) - (e ? f1 : f0);
                            S.VT[gi * NX + gj] = v;
                            S.VT[gj * NX + gi] = v;
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        // Branch free (a divergent branch per case costs more than the work): columns 37 / 38 use the
                        // vectors Qx -> Vx and qxy -> y as one more row / column of the matrices, whose spare entry 37 is
                        // kept zero, so Vx[37] = -|w0|^2 and y[37] = quy . k; inactive lanes hit dummy addresses.
                        const int gj = 8 * fJ + 2 * fc + e;
                        const real acc = e ? f1 : f0;
                        const bool m = gj < NX, ok = gj >= gi && gj <= NX + 1 && gi <= NX;
                        const real* s1 = m ? S.Qxx + gi * NX + gj : (gj == NX ? S.Qx : S.qxy) + gi;
                        const real* s2 = m ? S.Qxx + gj * NX + gi : s1;
                        real* d1 = m ? S.VT + gi * NX + gj : (gj == NX ? S.Vx : S.y) + gi;
                        real* d2 = m ? S.VT + gj * NX + gi : d1;
                        if (!ok) { s1 = s2 = S.Qx + NX; d1 = d2 = S.escr + 23; }
                        const real v = real(0.5) * (*s1 + *s2) - acc;
                        *d1 = v;
                        *d2 = v;
                    }
                }
                {
                    const int i = 8 * gI + fr, cc = 8 * gJ + 2 * fc;
                    real* Krow = const_cast<real*>(S.gp[0]) + ((size_t)k * NU + i) * NX;      // (pointers from shared memory: see prefetch)
                    if (cc < NX) Krow[cc] = -g0;
                    if (cc + 1 < NX) Krow[cc + 1] = -g1;
                    if (cc + 1 == NX) { S.kk[i] = -g1; const_cast<real*>(S.gp[1])[(size_t)k * NU + i] = -g1; }      // column 37 is odd
                }
            }
        }
        } else {
        // Warp w takes tiles w, w+4, w+8, w+12 of each product and runs their accumulation chains interleaved (a
        // single chain of six dependent DMMAs is latency bound); slot 3 of warp 3 is a dummy.
        {
            const int fr = lane >> 2, fc = lane & 3;       // fragment row / column of this lane
            int fI[4], fJ[4], gI[4], gJ[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int t = min(warp + 4 * q, 14);
                int I = 0, rem = t;
                while (rem >= 5 - I) { rem -= 5 - I; I++; }
                fI[q] = I; fJ[q] = I + rem;
                gI[q] = t / 5; gJ[q] = t % 5;
            }
            double fc0[4] = {0, 0, 0, 0}, fc1[4] = {0, 0, 0, 0}, gc0[4] = {0, 0, 0, 0}, gc1[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k0 = 0; k0 < NU; k0 += 4) {
                const real* r = S.W + (k0 + fc) * LDW + fr;
                real av[4], bv[4];
#pragma unroll
                for (int q = 0; q < 4; q++) { av[q] = r[8 * fI[q]]; bv[q] = r[8 * fJ[q]]; }
#pragma unroll
                for (int q = 0; q < 4; q++) dmma884(fc0[q], fc1[q], av[q], bv[q]);
            }
#pragma unroll
            for (int k0 = 0; k0 < NU; k0 += 4) {
                const int l = k0 + fc;
                real av[4], bv[4];
                const real rl = S.rs[l];                  // Es^T Wn = Et^T (diag(rs) Wn)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = 8 * gI[q] + fr;
                    av[q] = S.Quu[l * NU + i];
                    bv[q] = rl * S.W[l * LDW + 8 * gJ[q] + fr];
                }
#pragma unroll
                for (int q = 0; q < 4; q++) if (k0 >= 8 * gI[q]) dmma884(gc0[q], gc1[q], av[q], bv[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (warp + 4 * q >= 15) continue;
                const int gi = 8 * fI[q] + fr;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    // Branch free (a divergent branch per case costs more than the work): columns 37 / 38 use the
                    // vectors Qx -> Vx and qxy -> y as one more row / column of the matrices, whose spare entry 37 is
                    // kept zero, so Vx[37] = -|w0|^2 and y[37] = quy . k; inactive lanes hit dummy addresses.
                    const int gj = 8 * fJ[q] + 2 * fc + e;
                    const real acc = e ? fc1[q] : fc0[q];
                    const bool m = gj < NX, ok = gj >= gi && gj <= NX + 1 && gi <= NX;
                    const real* s1 = m ? S.Qxx + gi * NX + gj : (gj == NX ? S.Qx : S.qxy) + gi;
                    const real* s2 = m ? S.Qxx + gj * NX + gi : s1;
                    real* d1 = m ? S.VT + gi * NX + gj : (gj == NX ? S.Vx : S.y) + gi;
                    real* d2 = m ? S.VT + gj * NX + gi : d1;
                    if (!ok) { s1 = s2 = S.Qx + NX; d1 = d2 = S.escr + 23; }
                    const real v = real(0.5) * (*s1 + *s2) - acc;
                    *d1 = v;
                    *d2 = v;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (warp + 4 * q >= 15) continue;
                const int J = gJ[q], i = 8 * gI[q] + fr;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int cc = 8 * J + 2 * fc + e;
                    const real kv = e ? -gc1[q] : -gc0[q];
                    if (cc < NX) Kg[((size_t)k * NU + i) * NX + cc] = kv;
                    if (cc == NX) S.kk[i] = kv;
                    if (cc == NX) kg[(size_t)k * NU + i] = kv;
                }
            }
        }
        }
#endif
        STAMP(9);
        cp_wait_all();                         // the prefetch of node k-1 (issued in c1) is long done: this barrier also
        if (k > 0) landed(k - 1);
        __syncthreads();                       // publishes it, so the next node starts without one of its own
        STAMP(10);
        PROF(13);
        if (mu != 0.0) {
            // Regularised step (every backward pass after a failed factorisation or line search: 17 % of the node-iterations
            // of BASELINE configs[4]): [Vxx Vx] -= mu [K k]^T [K k].  Qxx is dead after the barrier above, so [K | k] of this
            // node (the gains this CTA just wrote; L2 hits) is staged there at the pitch of W and the product runs as the
            // same upper-triangular tile syrk as phase f on the FP64 tensor cores; entry (37, 37) is |k|^2.
            real* Ks = S.Qxx;
            const real* Kn = S.gp[0] + (size_t)k * NU * NX;
            for (int e = tid; e < NU * LDW; e += NT) {
                const int l = e / LDW, cc = e - l * LDW;
                Ks[e] = cc < NX ? Kn[l * NX + cc] : (cc == NX ? S.kk[l] : 0.0);
            }
            __syncthreads();
            const int fr = lane >> 2, fc = lane & 3;
#pragma unroll 1
            for (int t = warp; t < 15; t += NWARP) {
                const int fI = (int)((0x433222111100000ull >> (4 * t)) & 15), fJ = (int)((0x443432432143210ull >> (4 * t)) & 15);
                double f0 = 0.0, f1 = 0.0;
                const real* rf = Ks + fc * LDW + fr;
#pragma unroll
                for (int k0 = 0; k0 < NU; k0 += 4) dmma884(f0, f1, rf[k0 * LDW + 8 * fI], rf[k0 * LDW + 8 * fJ]);
                const int gi = 8 * fI + fr;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int gj = 8 * fJ + 2 * fc + e;
                    const real acc = e ? f1 : f0;
                    if (gj >= gi && gj < NX) {              // (gi <= gj < NX)
                        const real v = S.VT[gi * NX + gj] - mu * acc;
                        S.VT[gi * NX + gj] = v;
                        S.VT[gj * NX + gi] = v;
                    } else if (gj == NX && gi < NX) S.Vx[gi] -= mu * acc;
                    else if (gj == NX && gi == NX) S.red[6] = acc;      // |k|^2
                }
            }
            __syncthreads();                   // the next node reads Vxx, Vx right away
        }
        if (tid == 0) {   // model accumulators
            const double sw = -S.Vx[NX];        // |w0|^2 (see f)
            const double sk = (mu != 0.0) ? S.red[6] : 0.0;
            const double sq = S.y[NX];          // quy . k
            const double kQk = sw - mu * sk;
            S.red[R_TOT] += S.red[R_G1] + real(0.5) * S.red[R_G2] + (-sw) + real(0.5) * kQk;
            S.red[R_ACC2] += real(0.5) * kQk;
            S.red[R_ACC1] += fixed ? (S.red[R_YG] + real(0.5) * S.red[R_G2]) : (S.red[R_YG] + sq);
        }
    }
    __syncthreads();
    if (tid == 0) {
        double tot = S.red[R_TOT], a1 = S.red[R_ACC1], a2 = S.red[R_ACC2];
        if (fixed) { dV3[2] = a1; dV3[1] = a2; dV3[0] = tot - a1 - a2; }
        else       { dV3[2] = 0.0; dV3[0] = a1; dV3[1] = tot - a1; }
        if (bulk) { mbar_inval(&S.mbar[0]); mbar_inval(&S.mbar[1]); }
    }
    __syncthreads();
    return 0;
}
