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
#include "sddp_solver.cuh"

struct alignas(16) SmemSrbd {
    static constexpr int NX = 37, NU = 24, NP = 19;
    static constexpr int LDW = 44;   // row pitch of W: 88 words = 24 mod 32, so the 4 x 8 DMMA fragment loads are conflict free
    double VT[NX * NX + 1];    // Vxx', then T = Vxx' fx in place, then the new Vxx; forward: scratch
    double Qxx[NX * NX + 1];   // forward: K of the current / next node (with W: 2 x 888 doubles)
    double W[NU * LDW];        // [Qux | w0 | 0 ...] -> Wn
    double Quu[NU * NU];       // Quu -> (strict upper) D Lt^T = frozen raw columns, (lower) Es
    double Vx[NX + 1], y[NX + 1], Qx[NX + 1], vp[NX + 1], ys[NX + 1], qxy[NX + 1], sv[NX + 1];
    double Qu[NU], quy[NU], kk[NU], invp[NU], rs[NU];
    double nb[2][NodeBuf<Srbd>::SIZE];
    double ypart[3][40];
    double sacc[NWARP][8];
    double red[16];
    double alpha[NCAND], rho[NCAND], Jc[NCAND];
    int iflag[4];
    __device__ double* Kbuf(int b) { return Qxx + b * (NU * NX); }
    __device__ double* scr() { return VT; }
    __device__ static int backward(const DevCfg& c, SmemSrbd& S, const double* X, const double* U, const double* P, const double* D,
                                   const double* packs, double mu, double* Kg, double* kg, double* dV3, bool has_gap, int tid);
};
static_assert(2 * 24 * 37 <= (37 * 37 + 1) + 24 * 44, "forward K double buffer must fit in Qxx + W");
enum { R_SW = 6 };
SDDP_DEV void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// D(8x8) += A(8x4) B(4x8) on the FP64 tensor cores.  Fragments (PTX ISA, m8n8k4 .f64): lane holds
// A[lane/4][lane%4], B[lane%4][lane/4], C[lane/4][2*(lane%4) + {0,1}].
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

// rows of V that make up row `a` of B^T (.) : cddot(i,k) -> cd_ik ; f(i,k) -> (fs/m) rd_k + G_i[:,k]^T w
SDDP_DEV int bt_rows(const DevCfg& c, const double* Jac, int a, int* idx, double* coef) {
    int i = a / 6, r = a % 6;
    if (r < 3) { idx[0] = Srbd::XCD + 3 * i + r; coef[0] = 1.0; return 1; }
    int k = r - 3;
    idx[0] = Srbd::XRD + k; coef[0] = c.inv_ms;
#pragma unroll
    for (int q = 0; q < 3; q++) { idx[1 + q] = Srbd::XW + q; coef[1 + q] = Jac[q * Srbd::NZ + Srbd::ZF + 3 * i + k]; }
    return 4;
}

// out[b] = sum_a v[a] * (dt Aoo)[a][b],  Aoo = d odot / d o = [[skew(w)/2, w/2], [-w^T/2, 0]];  hw = dt w / 2
SDDP_DEV void contract_Aoo(const double* v, const double* hw, double* out) {
    out[0] = v[1] * hw[2] - v[2] * hw[1] - v[3] * hw[0];
    out[1] = -v[0] * hw[2] + v[2] * hw[0] - v[3] * hw[1];
    out[2] = v[0] * hw[1] - v[1] * hw[0] - v[3] * hw[2];
    out[3] = v[0] * hw[0] + v[1] * hw[1] + v[2] * hw[2];
}
// out[b] = sum_a v[a] * (dt Aow)[a][b],  Aow = d odot / d w = [[(o_w I - skew(o_v))/2], [-o_v^T/2]];  ho = dt o / 2
SDDP_DEV void contract_Aow(const double* v, const double* ho, double* out) {
    out[0] = v[0] * ho[3] - v[1] * ho[2] + v[2] * ho[1] - v[3] * ho[0];
    out[1] = v[0] * ho[2] + v[1] * ho[3] - v[2] * ho[0] - v[3] * ho[1];
    out[2] = -v[0] * ho[1] + v[1] * ho[0] + v[2] * ho[3] - v[3] * ho[2];
}

__device__ int SmemSrbd::backward(const DevCfg& c, SmemSrbd& S, const double* X, const double* U, const double* P, const double* D,
                                  const double* packs, double mu, double* Kg, double* kg, double* dV3, bool has_gap, int tid) {
    using M = Srbd;
    using NBL = NodeBuf<Srbd>;
    constexpr int NZ = Srbd::NZ;
    const int N = c.N, lane = tid & 31, warp = tid >> 5;
    const bool fixed = c.rho_fixed > 0.0;
    const double rho_b = fixed ? c.rho_fixed : 1.0;
    const double dt = c.dt;
    SyncBlock sync;

    auto prefetch = [&](int k) {       // node k -> buffer k & 1
        double* nb = S.nb[k & 1];
        for (int i = tid; i < NX; i += NT) {
            cp_async8(nb + NBL::OX + i, X + (size_t)k * NX + i);
            if (has_gap) cp_async8(nb + NBL::OD + i, D + (size_t)k * NX + i);
        }
        for (int i = tid; i < NU; i += NT) cp_async8(nb + NBL::OU + i, U + (size_t)k * NU + i);
        for (int i = tid; i < NP; i += NT) cp_async8(nb + NBL::OP + i, P + (size_t)k * NP + i);
        const double* ps = packs + (size_t)k * M::PACK;
        if (((((size_t)ps) | ((size_t)(nb + NBL::OK))) & 15) == 0) { for (int i = 2 * tid; i < M::PACK; i += 2 * NT) cp_async16(nb + NBL::OK + i, ps + i); }
        else { for (int i = tid; i < M::PACK; i += NT) cp_async8(nb + NBL::OK + i, ps + i); }
        cp_commit();
    };

    // terminal node: Vx = l_Nx, Vxx = l_Nxx (ddp.py:216-226: costs only)
    __syncthreads();
    {
        double* nb = S.nb[N & 1];
        for (int i = tid; i < NX; i += NT) nb[NBL::OX + i] = X[(size_t)N * NX + i];
        for (int i = tid; i < NP; i += NT) nb[NBL::OP + i] = P[(size_t)N * NP + i];
    }
    prefetch(N - 1);
    if (tid == 0) { S.red[R_TOT] = 0.0; S.red[R_ACC1] = 0.0; S.red[R_ACC2] = 0.0; S.iflag[1] = 0; }
    __syncthreads();
    M::expand<LDW>(c, NODE_TERM, S.nb[N & 1] + NBL::OX, nullptr, S.nb[N & 1] + NBL::OP, nullptr, S.Vx, S.Qu, S.VT, S.W, S.Quu, tid, NT, sync, &S.ypart[0][0]);
    for (int i = tid; i < NX; i += NT) S.y[i] = S.Vx[i];

    for (int k = N - 1; k >= 0; k--) {
        const int kind = node_kind(k, N);
        double* nb = S.nb[k & 1];
        const double* xk = nb + NBL::OX;
        const double* uk = nb + NBL::OU;
        const double* pk = nb + NBL::OP;
        double* cg = nb + NBL::OD;
        const double* pack = nb + NBL::OK;
        cp_wait_all();
        __syncthreads();                       // node k landed; everyone is done with node k+1
        PROF(8);
        if (tid < NX) cg[tid] = has_gap ? rho_b * cg[tid] : 0.0;      // consumed after expand's barriers
        M::expand<LDW>(c, kind, xk, uk, pk, pack, S.Qx, S.Qu, S.Qxx, S.W, S.Quu, tid, NT, sync, &S.ypart[0][0]);
        const double* Jac = pack + M::PK_JAC;
        PROF(9);
        if (k > 0) prefetch(k - 1);      // after the zero fill: shared stores queue behind outstanding cp.async

        // ---- c1: everything that needs Vxx' itself: gap shift, Quu = luu + fu^T Vxx' fu, copies of lx, lu
        if (tid < NX) {
            double s = 0.0;
            if (has_gap) {
#pragma unroll 4
                for (int j = 0; j < NX; j++) s += S.VT[tid * NX + j] * cg[j];
            }
            S.sv[tid] = s;
            S.vp[tid] = S.Vx[tid] + s;
            S.ys[tid] = fixed ? S.y[tid] + s : S.y[tid];
            S.qxy[tid] = S.Qx[tid];
        } else if (tid >= 64 && tid < 64 + NU) {
            S.quy[tid - 64] = S.Qu[tid - 64];
        } else if (tid == 127) {
            S.iflag[2] = 0;            // number of factor columns published by warp 0 (see d1 / d2)
        }
        for (int e = tid; e < NU * (NU + 1) / 2; e += NT) {   // lower triangle (a >= b), mirrored
            int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
            while ((a + 1) * (a + 2) / 2 <= e) a++;
            while (a * (a + 1) / 2 > e) a--;
            int b = e - a * (a + 1) / 2;
            int ia[4], ib[4];
            double ca[4], cb[4];
            int na = bt_rows(c, Jac, a, ia, ca), nb_ = bt_rows(c, Jac, b, ib, cb);
            double s = 0.0;
            for (int p = 0; p < na; p++) {
                double t = 0.0;
                for (int q = 0; q < nb_; q++) t += S.VT[ia[p] * NX + ib[q]] * cb[q];
                s += ca[p] * t;
            }
            double v = S.Quu[a * NU + b] + dt * dt * s + (a == b ? mu : 0.0);
            S.Quu[a * NU + b] = v;
            S.Quu[b * NU + a] = v;
        }
        __syncthreads();
        PROF(10);

        const double* o = xk + M::XO;
        const double* w = xk + M::XW;
        if (warp == 0) {
            // ---- d1: Quu + mu I = Lt D Lt^T; lane t owns column t.  Row j of Lt^T (the multipliers of step j)
            //          overwrites the strict upper triangle of S.Quu; D^-1 goes to S.invp.
            if (has_gap) {   // gap terms of the model (the only use of cg after c1)
                double g1 = 0, g2 = 0, yg = 0;
                for (int i = lane; i < NX; i += 32) { g1 += S.Vx[i] * cg[i]; g2 += cg[i] * S.sv[i]; yg += S.y[i] * cg[i]; }
                g1 = warp_sum(g1); g2 = warp_sum(g2); yg = warp_sum(yg);
                if (lane == 0) { S.red[R_G1] = g1; S.red[R_G2] = g2; S.red[R_YG] = yg; }
            } else if (lane == 0) { S.red[R_G1] = 0.0; S.red[R_G2] = 0.0; S.red[R_YG] = 0.0; }
            // Lane t holds column t.  Step j: lane j publishes its (final) column raw, row j of the strict upper
            // triangle of S.Quu, and 1/pivot; every lane then applies  a[i] -= col_j[i] * (a[j] / pivot_j).
            // The element that becomes the next pivot (i = j+1) is updated first and its reciprocal started at
            // once, so the rest of the update overlaps the reciprocal latency.  Lanes <= j update dead values.
            double a[NU];
            const int t = lane < NU ? lane : NU - 1;
#pragma unroll
            for (int i = 0; i < NU; i++) a[i] = S.Quu[i * NU + t];
            __syncwarp();
            bool bad = false;
            double myinv = fast_rcp(a[0]);          // lane 0's pivot
#pragma unroll
            for (int j = 0; j < NU; j++) {
                if (lane == j) {
                    const double p = a[j];
                    bad = !(p > 0.0) || !isfinite(p);
                    S.invp[j] = myinv;
#pragma unroll
                    for (int i = j + 1; i < NU; i++) S.Quu[j * NU + i] = a[i];
                    __threadfence_block();
                    *(volatile int*)&S.iflag[2] = j + 1;      // column j is visible to the right-hand-side warps
                }
                __syncwarp();
                if (j + 1 < NU) {
                    const double sj = S.invp[j] * a[j];
                    a[j + 1] -= S.Quu[j * NU + j + 1] * sj;
                    myinv = fast_rcp(a[j + 1]);     // meaningful on lane j+1
                    // the rest of the column in batches of 8: loads first, then the FMAs (keeps the loads in
                    // flight together without holding a whole second column in registers)
#pragma unroll
                    for (int i0 = j + 2; i0 < NU; i0 += 8) {
                        double col[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) if (i0 + q < NU) col[q] = S.Quu[j * NU + i0 + q];
#pragma unroll
                        for (int q = 0; q < 8; q++) if (i0 + q < NU) a[i0 + q] -= col[q] * sj;
                    }
                }
            }
            if (__any_sync(FULL, bad) && lane == 0) S.iflag[1] = 1;
            if (lane < NU) S.rs[lane] = sqrt(S.invp[lane]);
            __syncwarp();
            __threadfence_block();
            if (lane == 0) *(volatile int*)&S.iflag[2] = NU + 1;           // rs is visible too
            PROF_T(14, 0);
        } else {
            // ---- c2: T = Vxx' fx = V + dt V A, in place, one thread per row (warps 1-2)
            const int r_ = tid - 32;
            if (r_ < NX) {
                double* row = S.VT + r_ * NX;
                double vr[3], vo[4], vw[3];
#pragma unroll
                for (int q = 0; q < 3; q++) { vr[q] = row[M::XR + q]; vw[q] = row[M::XW + q]; }
#pragma unroll
                for (int q = 0; q < 4; q++) vo[q] = row[M::XO + q];
                const double hw[3] = {0.5 * dt * w[0], 0.5 * dt * w[1], 0.5 * dt * w[2]};
                const double ho[4] = {0.5 * dt * o[0], 0.5 * dt * o[1], 0.5 * dt * o[2], 0.5 * dt * o[3]};
                double to[4], tw[3];
                contract_Aoo(vo, hw, to);     // (V dt Aoo): o columns
                contract_Aow(vo, ho, tw);     // (V dt Aow): w columns
                const double dw0 = dt * vw[0], dw1 = dt * vw[1], dw2 = dt * vw[2];
                // cd and rd columns first (they read the c and r entries before those are overwritten)
#pragma unroll
                for (int q = 0; q < 12; q++) row[M::XCD + q] += dt * row[M::XC + q];
#pragma unroll
                for (int q = 0; q < 3; q++) row[M::XRD + q] += dt * vr[q];
#pragma unroll
                for (int z = 0; z < 19; z++) {        // r, o, c columns: + dt vw . dwdot/dz
                    double v = row[z] + dw0 * Jac[z] + dw1 * Jac[NZ + z] + dw2 * Jac[2 * NZ + z];
                    if (z >= 3 && z < 7) v += to[z - 3];
                    row[z] = v;
                }
#pragma unroll
                for (int q = 0; q < 3; q++)
                    row[M::XW + q] = vw[q] + tw[q] + dw0 * Jac[M::ZW + q] + dw1 * Jac[NZ + M::ZW + q] + dw2 * Jac[2 * NZ + M::ZW + q];
            }
            PROF_T(15, 32);
            bar_named(2, 96);
            // ---- c3: columns of fx^T (.) and fu^T (.) applied to T (j < 37), v+ (j = 37) and ys (j = 38); warps 1-3
            for (int task = tid - 32; task < 39 * 3; task += 96) {
                const int j = task % 39, g = task / 39;
                const double* col = (j < NX) ? S.VT + j : (j == NX ? S.vp : S.ys);
                const int cs = (j < NX) ? NX : 1;      // stride between rows of this "column"
                double* oxx = (j < NX) ? S.Qxx + j : (j == NX ? S.Qx : S.qxy);
                const int os = (j < NX) ? NX : 1;
                const double tw0 = col[(M::XW + 0) * cs], tw1 = col[(M::XW + 1) * cs], tw2 = col[(M::XW + 2) * cs];
                if (g == 0) {          // rows r, o, rd, w
                    double tov[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) tov[q] = col[(M::XO + q) * cs];
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        oxx[(M::XR + q) * os] += col[(M::XR + q) * cs] + dt * (Jac[M::ZR + q] * tw0 + Jac[NZ + M::ZR + q] * tw1 + Jac[2 * NZ + M::ZR + q] * tw2);
                        oxx[(M::XRD + q) * os] += col[(M::XRD + q) * cs] + dt * col[(M::XR + q) * cs];
                    }
                    const double hw[3] = {0.5 * dt * w[0], 0.5 * dt * w[1], 0.5 * dt * w[2]};
                    const double ho[4] = {0.5 * dt * o[0], 0.5 * dt * o[1], 0.5 * dt * o[2], 0.5 * dt * o[3]};
                    double ao[4], aw[3];
                    contract_Aoo(tov, hw, ao);    // (dt Aoo)^T T[o,j]: row o_b = sum_a Aoo[a][b] T[o_a][j]
                    contract_Aow(tov, ho, aw);    // (dt Aow)^T T[o,j]: row w_b
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        oxx[(M::XO + q) * os] += col[(M::XO + q) * cs] + ao[q] + dt * (Jac[M::ZO + q] * tw0 + Jac[NZ + M::ZO + q] * tw1 + Jac[2 * NZ + M::ZO + q] * tw2);
#pragma unroll
                    for (int q = 0; q < 3; q++)
                        oxx[(M::XW + q) * os] += col[(M::XW + q) * cs] + aw[q] + dt * (Jac[M::ZW + q] * tw0 + Jac[NZ + M::ZW + q] * tw1 + Jac[2 * NZ + M::ZW + q] * tw2);
                } else if (g == 1) {   // rows c, cd
#pragma unroll
                    for (int q = 0; q < 12; q++) {
                        double tc = col[(M::XC + q) * cs];
                        oxx[(M::XC + q) * os] += tc + dt * (Jac[M::ZC + q] * tw0 + Jac[NZ + M::ZC + q] * tw1 + Jac[2 * NZ + M::ZC + q] * tw2);
                        oxx[(M::XCD + q) * os] += col[(M::XCD + q) * cs] + dt * tc;
                    }
                } else {               // fu^T (.): rows cddot_i, f_i
                    double* oux = (j < NX) ? S.W + j : (j == NX ? S.Qu : S.quy);
                    const int us = (j < NX) ? LDW : 1;
                    const double trd[3] = {c.inv_ms * col[(M::XRD + 0) * cs], c.inv_ms * col[(M::XRD + 1) * cs], c.inv_ms * col[(M::XRD + 2) * cs]};
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int q = 0; q < 3; q++) {
                            oux[(6 * i + q) * us] += dt * col[(M::XCD + 3 * i + q) * cs];
                            const int zf = M::ZF + 3 * i + q;
                            oux[(6 * i + 3 + q) * us] += dt * (trd[q] + Jac[zf] * tw0 + Jac[NZ + zf] * tw1 + Jac[2 * NZ + zf] * tw2);
                        }
                }
            }
            PROF_T(16, 32);
            // ---- d2: Lt^-1 applied to [Qux | Qu | I], one right-hand side per thread (warps 1-2), trailing the
            //          factorisation of warp 0 column by column (S.iflag[2] counts the published columns)
            bar_named(2, 96);                      // Qux, Qu complete
            const int t = tid - 32;
            if (t < NX + 1 + NU) {
                volatile int* published = (volatile int*)&S.iflag[2];
                double a[NU];
                if (t < NX) {
#pragma unroll
                    for (int i = 0; i < NU; i++) a[i] = S.W[i * LDW + t];
                } else if (t == NX) {
#pragma unroll
                    for (int i = 0; i < NU; i++) a[i] = S.Qu[i];
                } else {
#pragma unroll
                    for (int i = 0; i < NU; i++) a[i] = (i == t - (NX + 1)) ? 1.0 : 0.0;
                }
#pragma unroll
                for (int j = 0; j < NU - 1; j++) {
                    while (*published <= j) __nanosleep(40);      // polite wait: spinning warps steal issue slots
                    asm volatile("" ::: "memory");
                    const double sj = S.invp[j] * a[j];
#pragma unroll
                    for (int i0 = j + 1; i0 < NU; i0 += 8) {
                        double col[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) if (i0 + q < NU) col[q] = S.Quu[j * NU + i0 + q];
#pragma unroll
                        for (int q = 0; q < 8; q++) if (i0 + q < NU) a[i0 + q] -= col[q] * sj;
                    }
                }
                while (*published <= NU) __nanosleep(40);
                asm volatile("" ::: "memory");
                if (t <= NX) {                      // Wn = rs . frozen rows (column NX is w0)
#pragma unroll
                    for (int l = 0; l < NU; l++) S.W[l * LDW + t] = a[l] * S.rs[l];
                } else {                            // Es = rs . Lt^-1 -> lower triangle (incl. diagonal) of S.Quu
                    const int m = t - (NX + 1);
#pragma unroll
                    for (int l = 0; l < NU; l++) if (l >= m) S.Quu[l * NU + m] = a[l] * S.rs[l];
                }
            }
        }
        __syncthreads();
        PROF(11);
        if (S.iflag[1]) { __syncthreads(); if (tid == 0) S.iflag[1] = 0; cp_wait_all(); __syncthreads(); return k + 1; }

        // ---- f: [Vxx Vx] = [sym(Qxx) Qx] - Wn^T Wn: upper-triangular 8x8 tiles of the 40x40 product, K = 24 in six
        //         DMMA steps; warp w takes tiles w, w+4, ...  Results go to VT (T is dead), Vx and red[R_SW] = |w0|^2.
        {
            const int fr = lane >> 2, fc = lane & 3;       // fragment row / column of this lane
            for (int t = warp; t < 15; t += NWARP) {
                int I = 0, rem = t;
                while (rem >= 5 - I) { rem -= 5 - I; I++; }
                const int J = I + rem;
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int k0 = 0; k0 < NU; k0 += 4) {
                    const double* r = S.W + (k0 + fc) * LDW + fr;
                    dmma884(c0, c1, r[8 * I], r[8 * J]);
                }
                const int gi = 8 * I + fr;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int gj = 8 * J + 2 * fc + e;
                    const double acc = e ? c1 : c0;
                    if (gj < gi) continue;
                    if (gj < NX) {
                        const double v = 0.5 * (S.Qxx[gi * NX + gj] + S.Qxx[gj * NX + gi]) - acc;
                        S.VT[gi * NX + gj] = v;
                        S.VT[gj * NX + gi] = v;
                    } else if (gj == NX) {
                        if (gi < NX) S.Vx[gi] = S.Qx[gi] - acc;
                        else S.red[R_SW] = acc;            // |w0|^2
                    }
                }
            }
            // ---- g: [K | k] = -Es^T Wn: 3 x 5 tiles (rows i, columns c); Es is lower triangular, so row tile I starts
            //         at k0 = 8 I and entries above the diagonal are masked (that part of S.Quu holds the raw factor).
            for (int t = warp; t < 15; t += NWARP) {
                const int I = t / 5, J = t % 5;
                const int i = 8 * I + fr;
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int k0 = 0; k0 < NU; k0 += 4) {
                    if (k0 < 8 * I) continue;
                    const int l = k0 + fc;
                    const double a = (l >= i) ? S.Quu[l * NU + i] : 0.0;
                    dmma884(c0, c1, a, S.W[l * LDW + 8 * J + fr]);
                }
                const double q = S.quy[i];
                double y0 = -c0 * q, y1 = -c1 * q;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int cc = 8 * J + 2 * fc + e;
                    const double kv = e ? -c1 : -c0;
                    if (cc < NX) Kg[((size_t)k * NU + i) * NX + cc] = kv;
                    else if (cc == NX) { S.kk[i] = kv; kg[(size_t)k * NU + i] = kv; }
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { y0 += __shfl_xor_sync(FULL, y0, o); y1 += __shfl_xor_sync(FULL, y1, o); }
                if (fr == 0) { S.ypart[I][8 * J + 2 * fc] = y0; S.ypart[I][8 * J + 2 * fc + 1] = y1; }
            }
        }
        __syncthreads();
        PROF(13);
        if (tid < NX) S.y[tid] = S.qxy[tid] + (S.ypart[0][tid] + S.ypart[1][tid] + S.ypart[2][tid]);
        if (mu != 0.0) {   // regularised step (rare): Vxx -= mu K^T K, Vx -= mu K^T k, gains re-read from global
            __syncthreads();
            for (int e = tid; e < NX * NX + NX; e += NT) {
                const int i = e / NX, j = e % NX;
                if (i < NX && j < i) continue;
                double t = 0.0;
                if (i < NX) {
                    for (int l = 0; l < NU; l++) t += Kg[((size_t)k * NU + l) * NX + i] * Kg[((size_t)k * NU + l) * NX + j];
                    const double v = S.VT[i * NX + j] - mu * t;
                    S.VT[i * NX + j] = v;
                    S.VT[j * NX + i] = v;
                } else {
                    for (int l = 0; l < NU; l++) t += Kg[((size_t)k * NU + l) * NX + j] * S.kk[l];
                    S.Vx[j] -= mu * t;
                }
            }
        }
        if (tid == 0) {   // model accumulators
            const double sw = S.red[R_SW];
            double sk = 0.0;
            if (mu != 0.0) for (int i = 0; i < NU; i++) sk += S.kk[i] * S.kk[i];
            const double sq = S.ypart[0][NX] + S.ypart[1][NX] + S.ypart[2][NX];      // quy . k
            const double kQk = sw - mu * sk;
            S.red[R_TOT] += S.red[R_G1] + 0.5 * S.red[R_G2] + (-sw) + 0.5 * kQk;
            S.red[R_ACC2] += 0.5 * kQk;
            S.red[R_ACC1] += fixed ? (S.red[R_YG] + 0.5 * S.red[R_G2]) : (S.red[R_YG] + sq);
        }
    }
    __syncthreads();
    if (tid == 0) {
        double tot = S.red[R_TOT], a1 = S.red[R_ACC1], a2 = S.red[R_ACC2];
        if (fixed) { dV3[2] = a1; dV3[1] = a2; dV3[0] = tot - a1 - a2; }
        else       { dV3[2] = 0.0; dV3[0] = a1; dV3[1] = tot - a1; }
    }
    __syncthreads();
    return 0;
}
