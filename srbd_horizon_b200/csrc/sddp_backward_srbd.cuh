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
//   [Quu_r | Qux | Qu | I] is eliminated column-wise, every thread owning one of the 86 columns in
//   registers (24 steps, one named barrier each): Quu_r = Lt D Lt^T, frozen rows U = Lt^-1 [Qux Qu],
//   E = Lt^-1.  With rs = D^-1/2:  Wn = rs.U (= L^-1 [Qux Qu] of the Cholesky form), Es = rs.E
//   K = -Es^T Wn  (register-tiled matmul over all threads, no substitution chain)
//   [Vxx Vx; . |w0|^2] = [sym(Qxx) Qx; . 0] - Wn^T Wn   (3x3 register tiles over the 39x39 product)
#pragma once
#include "sddp_solver.cuh"

struct SmemSrbd {
    static constexpr int NX = 37, NU = 24, NP = 19, LDW = 39;
    double VT[NX * NX];        // Vxx', then T = Vxx' fx in place, then the new Vxx; forward: scratch
    double Qxx[NX * NX];       // forward: K of the current node
    double W[NU * LDW];        // [Qux | w0 | 0] -> Wn
    double Quu[NU * NU];       // -> Es
    double Vx[NX], y[NX], Qx[NX], vp[NX], ys[NX], qxy[NX], cg[NX], sv[NX];
    double Qu[NU], quy[NU], kk[NU];
    double xk[NX], uk[NU], pk[NP], pack[Srbd::PACK];
    double mult[2][NU], invp[NU], rs[NU];
    double ypart[3][40];
    double red[16];
    double alpha[NCAND], rho[NCAND], Jc[NCAND];
    int iflag[4];
    __device__ double* Kbuf() { return Qxx; }
    __device__ double* scr() { return VT; }
    __device__ static int backward(const DevCfg& c, SmemSrbd& S, const double* X, const double* U, const double* P, const double* D,
                                   const double* packs, double mu, double* Kg, double* kg, double* dV3, bool has_gap, int tid);
};
enum { R_SW = 6 };

// upper-triangular 3x3 tiles of the 39x39 product (13 x 13 tile grid): tile t -> (ti, tj), ti <= tj
__device__ const unsigned char kTileI[91] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 8, 8, 8, 8, 8, 9, 9, 9, 9, 10, 10, 10, 11, 11, 12};
__device__ const unsigned char kTileJ[91] = {
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12,
    4, 5, 6, 7, 8, 9, 10, 11, 12, 5, 6, 7, 8, 9, 10, 11, 12, 6, 7, 8, 9, 10, 11, 12, 7, 8, 9, 10, 11, 12, 8, 9, 10, 11, 12, 9, 10, 11, 12, 10, 11, 12, 11, 12, 12};

SDDP_DEV void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

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
    constexpr int NZ = Srbd::NZ;
    const int N = c.N, lane = tid & 31, warp = tid >> 5;
    const bool fixed = c.rho_fixed > 0.0;
    const double rho_b = fixed ? c.rho_fixed : 1.0;
    const double dt = c.dt;
    SyncBlock sync;

    // terminal node: Vx = l_Nx, Vxx = l_Nxx (ddp.py:216-226: costs only)
    for (int i = tid; i < NX; i += NT) S.xk[i] = X[(size_t)N * NX + i];
    for (int i = tid; i < NP; i += NT) S.pk[i] = P[(size_t)N * NP + i];
    if (tid == 0) { S.red[R_TOT] = 0.0; S.red[R_ACC1] = 0.0; S.red[R_ACC2] = 0.0; S.iflag[1] = 0; }
    __syncthreads();
    M::expand<LDW>(c, NODE_TERM, S.xk, nullptr, S.pk, nullptr, S.Vx, S.Qu, S.VT, S.W, S.Quu, tid, NT, sync);
    for (int i = tid; i < NX; i += NT) S.y[i] = S.Vx[i];
    __syncthreads();

    for (int k = N - 1; k >= 0; k--) {
        const int kind = node_kind(k, N);
        for (int i = tid; i < NX; i += NT) {
            S.xk[i] = X[(size_t)k * NX + i];
            S.cg[i] = has_gap ? rho_b * D[(size_t)k * NX + i] : 0.0;
        }
        for (int i = tid; i < NU; i += NT) S.uk[i] = U[(size_t)k * NU + i];
        for (int i = tid; i < NP; i += NT) S.pk[i] = P[(size_t)k * NP + i];
        for (int i = tid; i < M::PACK; i += NT) S.pack[i] = packs[(size_t)k * M::PACK + i];
        __syncthreads();
        M::expand<LDW>(c, kind, S.xk, S.uk, S.pk, S.pack, S.Qx, S.Qu, S.Qxx, S.W, S.Quu, tid, NT, sync);
        const double* Jac = S.pack + M::PK_JAC;

        // ---- c1: everything that needs Vxx' itself: gap shift, Quu = luu + fu^T Vxx' fu, copies of lx, lu
        if (tid < NX) {
            double s = 0.0;
            if (has_gap) {
#pragma unroll 4
                for (int j = 0; j < NX; j++) s += S.VT[tid * NX + j] * S.cg[j];
            }
            S.sv[tid] = s;
            S.vp[tid] = S.Vx[tid] + s;
            S.ys[tid] = fixed ? S.y[tid] + s : S.y[tid];
            S.qxy[tid] = S.Qx[tid];
        } else if (tid >= 64 && tid < 64 + NU) {
            S.quy[tid - 64] = S.Qu[tid - 64];
        }
        for (int e = tid; e < NU * (NU + 1) / 2; e += NT) {   // lower triangle (a >= b), mirrored
            int a = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
            while ((a + 1) * (a + 2) / 2 <= e) a++;
            while (a * (a + 1) / 2 > e) a--;
            int b = e - a * (a + 1) / 2;
            int ia[4], ib[4];
            double ca[4], cb[4];
            int na = bt_rows(c, Jac, a, ia, ca), nb = bt_rows(c, Jac, b, ib, cb);
            double s = 0.0;
            for (int p = 0; p < na; p++) {
                double t = 0.0;
                for (int q = 0; q < nb; q++) t += S.VT[ia[p] * NX + ib[q]] * cb[q];
                s += ca[p] * t;
            }
            double v = S.Quu[a * NU + b] + dt * dt * s;
            S.Quu[a * NU + b] = v;
            S.Quu[b * NU + a] = v;
        }
        __syncthreads();
        if (warp == 0 && has_gap) {   // gap terms of the model
            double g1 = 0, g2 = 0, yg = 0;
            for (int i = lane; i < NX; i += 32) { g1 += S.Vx[i] * S.cg[i]; g2 += S.cg[i] * S.sv[i]; yg += S.y[i] * S.cg[i]; }
            g1 = warp_sum(g1); g2 = warp_sum(g2); yg = warp_sum(yg);
            if (lane == 0) { S.red[R_G1] = g1; S.red[R_G2] = g2; S.red[R_YG] = yg; }
        } else if (tid == 0) { S.red[R_G1] = 0.0; S.red[R_G2] = 0.0; S.red[R_YG] = 0.0; }

        // ---- c2: T = Vxx' fx = V + dt V A, in place, one thread per row
        const double* o = S.xk + M::XO;
        const double* w = S.xk + M::XW;
        if (tid < NX) {
            double* row = S.VT + tid * NX;
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
        __syncthreads();

        // ---- c3: columns of fx^T (.) and fu^T (.) applied to T (j < 37), v+ (j = 37) and ys (j = 38)
        if (tid < 39 * 3) {
            const int j = tid % 39, g = tid / 39;
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
        __syncthreads();

        // ---- d: square-root-free elimination of [Quu + mu I | Qux | Qu | I], one column per thread
        double a[NU];
        if (tid < 96) {
            const int t = tid;
            if (t < NU) {
#pragma unroll
                for (int i = 0; i < NU; i++) a[i] = S.Quu[i * NU + t] + (i == t ? mu : 0.0);
            } else if (t < NU + NX) {
#pragma unroll
                for (int i = 0; i < NU; i++) a[i] = S.W[i * LDW + (t - NU)];
            } else if (t == NU + NX) {
#pragma unroll
                for (int i = 0; i < NU; i++) a[i] = S.Qu[i];
            } else {
#pragma unroll
                for (int i = 0; i < NU; i++) a[i] = (i == t - (NU + NX + 1)) ? 1.0 : 0.0;
            }
#pragma unroll
            for (int j = 0; j < NU; j++) {
                if (t == j) {
                    const double p = a[j];
                    if (!(p > 0.0) || !isfinite(p)) S.iflag[1] = 1;
                    const double inv = 1.0 / p;
                    S.invp[j] = inv;
#pragma unroll
                    for (int i = j + 1; i < NU; i++) S.mult[j & 1][i] = a[i] * inv;
                }
                bar_named(1, 96);
                if (t > j) {
                    const double aj = a[j];
#pragma unroll
                    for (int i = j + 1; i < NU; i++) a[i] -= S.mult[j & 1][i] * aj;
                }
            }
            if (t < NU) S.rs[t] = sqrt(S.invp[t]);
            bar_named(1, 96);
            if (t >= NU && t < NU + NX + 1) {          // Wn = rs . frozen rows (column NX is w0)
#pragma unroll
                for (int l = 0; l < NU; l++) S.W[l * LDW + (t - NU)] = a[l] * S.rs[l];
            } else if (t >= NU + NX + 1 && t < 2 * NU + NX + 1) {   // Es = rs . Lt^-1 (lower triangular)
                const int m = t - (NU + NX + 1);
#pragma unroll
                for (int l = 0; l < NU; l++) S.Quu[l * NU + m] = (l >= m) ? a[l] * S.rs[l] : 0.0;
            }
        }
        __syncthreads();
        if (S.iflag[1]) { __syncthreads(); if (tid == 0) S.iflag[1] = 0; __syncthreads(); return k + 1; }

        // ---- f: [Vxx Vx] = [sym(Qxx) Qx] - Wn^T Wn, 3x3 register tiles of the upper triangle (into VT)
        if (tid < 91) {
            const int ti = kTileI[tid], tj = kTileJ[tid];
            double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll 4
            for (int l = 0; l < NU; l++) {
                const double* r = S.W + l * LDW;
                const double u0 = r[3 * ti], u1 = r[3 * ti + 1], u2 = r[3 * ti + 2];
                const double v0 = r[3 * tj], v1 = r[3 * tj + 1], v2 = r[3 * tj + 2];
                acc[0][0] += u0 * v0; acc[0][1] += u0 * v1; acc[0][2] += u0 * v2;
                acc[1][0] += u1 * v0; acc[1][1] += u1 * v1; acc[1][2] += u1 * v2;
                acc[2][0] += u2 * v0; acc[2][1] += u2 * v1; acc[2][2] += u2 * v2;
            }
#pragma unroll
            for (int p = 0; p < 3; p++)
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const int gi = 3 * ti + p, gj = 3 * tj + q;
                    if (gj < gi) continue;
                    if (gj < NX) {
                        const double v = 0.5 * (S.Qxx[gi * NX + gj] + S.Qxx[gj * NX + gi]) - acc[p][q];
                        S.VT[gi * NX + gj] = v;
                        S.VT[gj * NX + gi] = v;
                    } else if (gj == NX) {
                        if (gi < NX) S.Vx[gi] = S.Qx[gi] - acc[p][q];
                        else S.red[R_SW] = acc[p][q];            // |w0|^2
                    }
                }
        }
        // ---- g: K = -Es^T Wn (column c of [K | k], interleaved row groups), y partial sums
        if (tid < 38 * 3) {
            const int cc = tid % 38, g = tid / 38;
            double wn[NU];
#pragma unroll
            for (int l = 0; l < NU; l++) wn[l] = S.W[l * LDW + cc];
            double yp = 0.0;
#pragma unroll
            for (int r8 = 0; r8 < 8; r8++) {
                const int i = g + 3 * r8;
                double s = 0.0;
#pragma unroll
                for (int l = 3 * r8; l < NU; l++) s += S.Quu[l * NU + i] * wn[l];   // Es[l][i] = 0 for l < i
                const double kv = -s;
                yp += kv * S.quy[i];
                if (cc < NX) Kg[((size_t)k * NU + i) * NX + cc] = kv;
                else { S.kk[i] = kv; kg[(size_t)k * NU + i] = kv; }
            }
            S.ypart[g][cc] = yp;
        }
        __syncthreads();
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
        __syncthreads();
    }
    if (tid == 0) {
        double tot = S.red[R_TOT], a1 = S.red[R_ACC1], a2 = S.red[R_ACC2];
        if (fixed) { dV3[2] = a1; dV3[1] = a2; dV3[0] = tot - a1 - a2; }
        else       { dV3[2] = 0.0; dV3[0] = a1; dV3[1] = tot - a1; }
    }
    __syncthreads();
    return 0;
}
