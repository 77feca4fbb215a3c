// sddp_solver.cuh -- the DDP iteration as sm_100a device code: one CTA per problem.
//
// Stages (BASELINE.json north_star):
//   1. derivatives   node packs, one thread per horizon node (Model::pack), expanded block-parallel
//   2. backward      sequential Riccati recursion over nodes; matrices in shared memory; Quu
//                    regularisation, in-warp Cholesky (one lane per row, warp shuffles for the pivot),
//                    one thread per right-hand side for the gain solves
//   3. forward       one warp per candidate step size (parallel line search), lanes own the rows of
//                    K dx and the state components; warp-shuffle cost reduction
//   4. defects       multiple-shooting gaps d_k = f(x_k,u_k) - x_{k+1}, contracted by (1 - rho)
//
// The algorithm is the one documented in oracle/sddp_oracle.c (it restates pyddp's role; the reference's
// ddp.py:96-106 only calls it).  Algebraic form used here (W-form, Quu_r = L L^T):
//   W = L^-1 Qux, w0 = L^-1 Qu, K = -L^-T W, k = -L^-T w0
//   Vx = Qx - W^T w0 - mu K^T k,  Vxx = sym(Qxx) - W^T W - mu K^T K
//   Qu.k = -|w0|^2,  k^T Quu k = |w0|^2 - mu |k|^2
#pragma once
#include "sddp_model.cuh"

#ifndef SDDP_FIRST_WAVE_SINGLE
#define SDDP_FIRST_WAVE_SINGLE 1
#endif
#ifndef SDDP_MAX_PEERS
#define SDDP_MAX_PEERS 8     // GPUs of one NVLink box
#endif
constexpr int NT = 128;      // threads per CTA
constexpr int NWARP = NT / 32;
constexpr int NCAND = NWARP; // line-search candidates evaluated per wave
constexpr int NSLOT = NCAND + 1;   // trial trajectories per CTA: one of them holds the current trajectory after the first accepted step
constexpr unsigned FULL = 0xffffffffu;

// per-node input buffer: x | u | p | d | (pack in the backward pass, k in the forward pass)
template <class M>
struct NodeBuf {
    static constexpr int OX = 0, OU = M::NX, OP = M::NX + M::NU, OD = M::NX + M::NU + M::NP, OK = (2 * M::NX + M::NU + M::NP + RV - 1) & ~(RV - 1);
    static constexpr int TAIL = M::PACK > M::NU ? M::PACK : M::NU;
    static constexpr int SIZE = (OK + TAIL + RV - 1) & ~(RV - 1);
};

// one element (8 bytes; 4 in the fp32 build)
SDDP_DEV void cp_async8(real* smem, const real* g) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
#ifdef SDDP_F32
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(g) : "memory");
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(g) : "memory");
#endif
}
// RV elements (16 bytes)
SDDP_DEV void cp_async16(real* smem, const real* g) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
SDDP_DEV void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
SDDP_DEV void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- 1-D bulk copies (the TMA unit without a tensor map: cp.async.bulk + mbarrier), for tiles whose size and addresses are
//      multiples of 16 bytes: one thread issues one instruction for the whole tile instead of every thread a few cp.async
#ifndef SDDP_BULK
#define SDDP_BULK 1
#endif
SDDP_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SDDP_DEV void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SDDP_DEV void mbar_inval(unsigned long long* bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
SDDP_DEV void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the issuing thread: expect `bytes` on the barrier, then the copy that delivers them
SDDP_DEV void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
SDDP_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

template <class M>
struct Smem {
    static constexpr int NX = M::NX, NU = M::NU, NP = M::NP;
    real Vxx[NX * NX], Qxx[NX * NX], Qux[NU * NX], Quu[NU * NU];
    real T[NX * (NX + NU)];            // Vxx [fx fu]; afterwards K of the node (NU*NX); forward: K_k
    real fx[NX * NX], fu[NX * NU];     // dense; forward: per-candidate x^, u^, x^+
    real Vx[NX], y[NX], Qx[NX], Qu[NU], vp[NX], sv[NX], ys[NX], quy[NU], qxy[NX], kk[NU], w0[NU];
    real nb[2][NodeBuf<M>::SIZE];      // double-buffered per-node inputs (x, u, p, d, pack | k)
    real sacc[NWARP][8];
    double red[16];                      // reductions and the expected-decrease accumulators: double in both builds
    double alpha[NCAND], rho[NCAND], Jc[NCAND];
    const real* gp[8];                 // base pointers of the per-node prefetches (see forward_wave)
    unsigned long long mbar[2];          // completion barriers of the bulk copies of K_k (forward_wave)
    int iflag[4];
    __device__ static int backward(const DevCfg& c, Smem<M>& S, const real* X, const real* U, const real* P, const real* D,
                                   const real* packs, real mu, real* Kg, real* kg, double* dV3, bool has_gap, int tid);
    __device__ static void prep(const DevCfg& c, Smem<M>&, const real* X, const real* U, const real*, real* packs, int tid);
    __device__ real* Kbuf(int b) { return T + b * (NU * NX); }   // forward pass: K of node k in buffer k & 1
    __device__ real* scr() { return fx; }    // per-warp scratch
};
enum { R_TOT = 0, R_ACC1 = 1, R_ACC2 = 2, R_G1 = 3, R_G2 = 4, R_YG = 5, R_W0 = 8 };

struct SyncBlock { __device__ void operator()() const { __syncthreads(); } };

// FIRST: node 0 (no trackers), TERM: node N (trackers only), TAIL: a node 1..N-1 of the LIP-style tail (SddpConfig.lip_tail_start), else MID
SDDP_DEV int node_kind(const DevCfg& c, int k) {
    if (k == 0) return NODE_FIRST;
    if (k == c.N) return NODE_TERM;
    return (c.lip_tail > 0 && k >= c.lip_tail) ? NODE_TAIL : NODE_MID;
}

template <class T>
SDDP_DEV T warp_sum(T s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    return s;
}
template <class T>
SDDP_DEV T warp_max(T s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(FULL, s, o));
    return s;
}

// cost of one node (all lanes get the sum) and, if xnext != null, the Euler step
//   xnext = xs + dt*ode(xs,us) - omr*dk
template <class M>
__device__ SDDP_NOINLINE double warp_node(const DevCfg& c, int kind, const real* xs, const real* us, const real* ps,
                            real* xnext, const real* dk, real omr, real* sacc, int lane) {
    if (kind != NODE_TERM && M::NACC > 1) {
        real acc[M::NACC];
        M::accel(c, xs, us, acc, kind == NODE_TAIL);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < M::NACC; q++) sacc[q] = acc[q];
        }
        __syncwarp();
    }
    const double s = warp_sum(M::cost_lane(c, kind, lane, xs, us, ps, sacc));
    if (xnext != nullptr && kind != NODE_TERM) {
        for (int i = lane; i < M::NX; i += 32) {
            real v = xs[i] + c.dt * M::xdot_i(c, i, xs, us, sacc);
            if (dk != nullptr) v -= omr * dk[i];
            xnext[i] = v;
        }
    }
    __syncwarp();
    return s;
}

// ------------------------------------------------------------------------------------------------
// defects d_k = f(X_k,U_k) - X_{k+1} (if dout) and the total cost; warps stride over nodes.
// Returns J to every thread.  Uses S.fx as per-warp scratch.
template <class M, class SM>
__device__ double defects_and_cost(const DevCfg& c, SM& S, const real* X, const real* U, const real* P,
                                   real* dout, int tid) {
    constexpr int NX = M::NX, NU = M::NU, NP = M::NP;
    const int N = c.N, lane = tid & 31, w = tid >> 5;
    real* xs = S.scr() + w * (2 * NX + NU + NP);
    real* us = xs + NX;
    real* ps = us + NU;
    real* xn = ps + NP;
    double part = 0.0;
    for (int k = w; k <= N; k += NWARP) {
        const int kind = node_kind(c, k);
        for (int i = lane; i < NX; i += 32) xs[i] = X[(size_t)k * NX + i];
        if (k < N) for (int i = lane; i < NU; i += 32) us[i] = U[(size_t)k * NU + i];
        for (int i = lane; i < NP; i += 32) ps[i] = P[(size_t)k * NP + i];
        __syncwarp();
        part += warp_node<M>(c, kind, xs, us, ps, (dout && k < N) ? xn : nullptr, nullptr, 0.0, S.sacc[w], lane);
        __syncwarp();
        if (dout && k < N)
            for (int i = lane; i < NX; i += 32) dout[(size_t)k * NX + i] = xn[i] - X[(size_t)(k + 1) * NX + i];
        __syncwarp();
    }
    if (lane == 0) S.red[R_W0 + w] = part;
    __syncthreads();
    double J = 0.0;
#pragma unroll
    for (int i = 0; i < NWARP; i++) J += S.red[R_W0 + i];
    __syncthreads();
    return J;
}

// open-loop rollout X_{k+1} = f(X_k, U_k) by warp 0 (single-shooting initialisation)
template <class M, class SM>
__device__ void open_loop_rollout(const DevCfg& c, SM& S, real* X, const real* U, int tid) {
    constexpr int NX = M::NX, NU = M::NU;
    const int lane = tid & 31;
    if (tid < 32) {
        real* xs = S.scr();
        real* us = xs + NX;
        real* xn = us + NU;
        for (int i = lane; i < NX; i += 32) xs[i] = X[i];
        for (int k = 0; k < c.N; k++) {
            for (int i = lane; i < NU; i += 32) us[i] = U[(size_t)k * NU + i];
            __syncwarp();
            if (M::NACC > 1) {
                real acc[M::NACC];
                M::accel(c, xs, us, acc, node_kind(c, k) == NODE_TAIL);
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < M::NACC; q++) S.sacc[0][q] = acc[q];
                }
                __syncwarp();
            }
            for (int i = lane; i < NX; i += 32) xn[i] = xs[i] + c.dt * M::xdot_i(c, i, xs, us, S.sacc[0]);
            __syncwarp();
            for (int i = lane; i < NX; i += 32) { xs[i] = xn[i]; X[(size_t)(k + 1) * NX + i] = xn[i]; }
            __syncwarp();
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// Stage 2.  Returns 0 or (failing node + 1) to every thread.  dV3[0..2] (shared) = {D1, D2, C0}.
template <class M>
__device__ int backward_pass(const DevCfg& c, Smem<M>& S, const real* X, const real* U, const real* P,
                             const real* D, const real* packs, real mu, real* Kg, real* kg, double* dV3, int tid) {
    constexpr int NX = M::NX, NU = M::NU, NP = M::NP, LD = NX + NU;
    const int N = c.N, lane = tid & 31, warp = tid >> 5;
    const bool fixed = c.rho_fixed > 0.0;
    const real rho_b = fixed ? (real)c.rho_fixed : real(1.0);
    SyncBlock sync;
    using NBL = NodeBuf<M>;
    real* xk = S.nb[0] + NBL::OX;
    real* uk = S.nb[0] + NBL::OU;
    real* pk = S.nb[0] + NBL::OP;
    real* cg = S.nb[0] + NBL::OD;
    real* pack = S.nb[0] + NBL::OK;

    // terminal node: Vx = l_Nx, Vxx = l_Nxx (ddp.py:216-226: costs only)
    for (int i = tid; i < NX; i += NT) xk[i] = X[(size_t)N * NX + i];
    for (int i = tid; i < NP; i += NT) pk[i] = P[(size_t)N * NP + i];
    if (tid == 0) { S.red[R_TOT] = 0.0; S.red[R_ACC1] = 0.0; S.red[R_ACC2] = 0.0; }
    __syncthreads();
    M::expand(c, NODE_TERM, xk, nullptr, pk, nullptr, S.Vx, S.Qu, S.Vxx, S.Qux, S.Quu, tid, NT, sync);
    for (int i = tid; i < NX; i += NT) S.y[i] = S.Vx[i];
    __syncthreads();

    for (int k = N - 1; k >= 0; k--) {
        const int kind = node_kind(c, k);
        for (int i = tid; i < NX; i += NT) {
            xk[i] = X[(size_t)k * NX + i];
            cg[i] = (D != nullptr) ? rho_b * D[(size_t)k * NX + i] : 0.0;
        }
        for (int i = tid; i < NU; i += NT) uk[i] = U[(size_t)k * NU + i];
        for (int i = tid; i < NP; i += NT) pk[i] = P[(size_t)k * NP + i];
        for (int i = tid; i < M::PACK; i += NT) pack[i] = packs[(size_t)k * M::PACK + i];
        __syncthreads();
        M::expand(c, kind, xk, uk, pk, pack, S.Qx, S.Qu, S.Qxx, S.Qux, S.Quu, tid, NT, sync);
        M::expand_f(c, xk, uk, pack, S.fx, S.fu, tid, NT, sync);

        // sv = Vxx' c, v+ = Vx' + sv, ys = y' (+ sv)
        if (tid < NX) {
            real s = 0.0;
            for (int j = 0; j < NX; j++) s += S.Vxx[tid * NX + j] * cg[j];
            S.sv[tid] = s;
            S.vp[tid] = S.Vx[tid] + s;
            S.ys[tid] = fixed ? S.y[tid] + s : S.y[tid];
        }
        // T = Vxx' [fx fu]
        for (int e = tid; e < NX * LD; e += NT) {
            int i = e / LD, j = e % LD;
            real s = 0.0;
            if (j < NX) for (int l = 0; l < NX; l++) s += S.Vxx[i * NX + l] * S.fx[l * NX + j];
            else        for (int l = 0; l < NX; l++) s += S.Vxx[i * NX + l] * S.fu[l * NU + (j - NX)];
            S.T[e] = s;
        }
        __syncthreads();
        if (warp == 0) {   // gap terms of the model
            real g1 = 0, g2 = 0, yg = 0;
            for (int i = lane; i < NX; i += 32) { g1 += S.Vx[i] * cg[i]; g2 += cg[i] * S.sv[i]; yg += S.y[i] * cg[i]; }
            g1 = warp_sum(g1); g2 = warp_sum(g2); yg = warp_sum(yg);
            if (lane == 0) { S.red[R_G1] = g1; S.red[R_G2] = g2; S.red[R_YG] = yg; }
        }
        // first-order quantities (Qx, Qu still hold lx, lu)
        if (tid < NX) {
            real a = 0.0, b = 0.0;
            for (int l = 0; l < NX; l++) { a += S.fx[l * NX + tid] * S.ys[l]; b += S.fx[l * NX + tid] * S.vp[l]; }
            S.qxy[tid] = S.Qx[tid] + a;
            S.Qx[tid] += b;
        } else if (tid >= 64 && tid < 64 + NU) {
            int j = tid - 64;
            real a = 0.0, b = 0.0;
            for (int l = 0; l < NX; l++) { a += S.fu[l * NU + j] * S.ys[l]; b += S.fu[l * NU + j] * S.vp[l]; }
            S.quy[j] = S.Qu[j] + a;
            S.Qu[j] += b;
        }
        // Qxx += fx^T Tx, Qux += fu^T Tx, Quu += fu^T Tu
        for (int e = tid; e < NX * NX; e += NT) {
            int i = e / NX, j = e % NX;
            real s = 0.0;
            for (int l = 0; l < NX; l++) s += S.fx[l * NX + i] * S.T[l * LD + j];
            S.Qxx[e] += s;
        }
        for (int e = tid; e < NU * NX; e += NT) {
            int i = e / NX, j = e % NX;
            real s = 0.0;
            for (int l = 0; l < NX; l++) s += S.fu[l * NU + i] * S.T[l * LD + j];
            S.Qux[e] += s;
        }
        for (int e = tid; e < NU * NU; e += NT) {
            int i = e / NU, j = e % NU;
            real s = 0.0;
            for (int l = 0; l < NX; l++) s += S.fu[l * NU + i] * S.T[l * LD + NX + j];
            S.Quu[e] += s;
        }
        __syncthreads();

        // Cholesky of sym(Quu) + mu I, left-looking, lane i owns row i; L overwrites the lower triangle
        if (warp == 0) {
            int ok = 1;
            for (int j = 0; j < NU; j++) {
                real s = 0.0;
                if (lane < NU && lane >= j) {
                    s = real(0.5) * (S.Quu[lane * NU + j] + S.Quu[j * NU + lane]) + (lane == j ? mu : 0.0);
                    for (int l = 0; l < j; l++) s -= S.Quu[lane * NU + l] * S.Quu[j * NU + l];
                }
                real piv = __shfl_sync(FULL, s, j);
                if (!(piv > 0.0) || !isfinite(piv)) { ok = 0; break; }
                real d = sqrt(piv);
                if (lane < NU && lane >= j) S.Quu[lane * NU + j] = (lane == j) ? d : s / d;
                __syncwarp();
            }
            if (lane == 0) S.iflag[0] = ok;
        }
        __syncthreads();
        if (!S.iflag[0]) { __syncthreads(); return k + 1; }

        // gains: one thread per right-hand side (NX columns of Qux, then Qu)
        if (tid <= NX) {
            const int t = tid;
            real b[NU], kc[NU];
#pragma unroll
            for (int i = 0; i < NU; i++) b[i] = (t < NX) ? S.Qux[i * NX + t] : S.Qu[i];
#pragma unroll
            for (int i = 0; i < NU; i++) {
                real s = b[i];
#pragma unroll
                for (int l = 0; l < i; l++) s -= S.Quu[i * NU + l] * b[l];
                b[i] = s / S.Quu[i * NU + i];
            }
#pragma unroll
            for (int i = NU - 1; i >= 0; i--) {
                real s = -b[i];
#pragma unroll
                for (int l = i + 1; l < NU; l++) s -= S.Quu[l * NU + i] * kc[l];
                kc[i] = s / S.Quu[i * NU + i];
            }
            if (t < NX) {
                real yv = S.qxy[t];
#pragma unroll
                for (int i = 0; i < NU; i++) {
                    S.Qux[i * NX + t] = b[i];                       // W
                    S.T[i * NX + t] = kc[i];                        // K (shared copy)
                    Kg[((size_t)k * NU + i) * NX + t] = kc[i];
                    yv += kc[i] * S.quy[i];
                }
                S.y[t] = yv;
            } else {
#pragma unroll
                for (int i = 0; i < NU; i++) { S.w0[i] = b[i]; S.kk[i] = kc[i]; kg[(size_t)k * NU + i] = kc[i]; }
            }
        }
        __syncthreads();
        if (warp == 0) {   // model accumulators
            real sw = 0, sk = 0, sq = 0;
            for (int i = lane; i < NU; i += 32) { sw += S.w0[i] * S.w0[i]; sk += S.kk[i] * S.kk[i]; sq += S.quy[i] * S.kk[i]; }
            sw = warp_sum(sw); sk = warp_sum(sk); sq = warp_sum(sq);
            if (lane == 0) {
                real kQk = sw - mu * sk;
                S.red[R_TOT] += S.red[R_G1] + real(0.5) * S.red[R_G2] + (-sw) + real(0.5) * kQk;
                S.red[R_ACC2] += real(0.5) * kQk;
                S.red[R_ACC1] += fixed ? (S.red[R_YG] + real(0.5) * S.red[R_G2]) : (S.red[R_YG] + sq);
            }
        }
        // Vx = Qx - W^T w0 - mu K^T k
        if (tid >= 64 && tid < 64 + NX) {
            int t = tid - 64;
            real s = S.Qx[t];
            for (int l = 0; l < NU; l++) s -= S.Qux[l * NX + t] * S.w0[l] + mu * S.T[l * NX + t] * S.kk[l];
            S.Vx[t] = s;
        }
        // Vxx = sym(Qxx) - W^T W - mu K^T K   (upper triangle, mirrored)
        for (int e = tid; e < NX * NX; e += NT) {
            int i = e / NX, j = e % NX;
            if (j < i) continue;
            real s = real(0.5) * (S.Qxx[i * NX + j] + S.Qxx[j * NX + i]);
            for (int l = 0; l < NU; l++) s -= S.Qux[l * NX + i] * S.Qux[l * NX + j];
            if (mu != 0.0) {
                real t = 0.0;
                for (int l = 0; l < NU; l++) t += S.T[l * NX + i] * S.T[l * NX + j];
                s -= mu * t;
            }
            S.Vxx[i * NX + j] = s;
            S.Vxx[j * NX + i] = s;
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

// ------------------------------------------------------------------------------------------------
// Stage 3.  ncand (<= NCAND) candidate step sizes S.alpha[], S.rho[] rolled out in parallel, one warp each.
// Trial trajectories go to Xn + cand*xn_stride / Un + cand*un_stride (global); costs to S.Jc[].
template <class M, class SM>
__device__ void forward_wave(const DevCfg& c, SM& S, const real* x0, const real* X, const real* U, const real* P,
                             const real* D, const real* Kg, const real* kg, int ncand, real* Xn, size_t xn_stride,
                             real* Un, size_t un_stride, int tid, int skip = 1 << 30) {
    constexpr int NX = M::NX, NU = M::NU, NP = M::NP;
    using NBL = NodeBuf<M>;
    const int N = c.N, lane = tid & 31, w = tid >> 5;
    real* xh = S.scr() + w * (2 * NX + NU);
    real* uh = xh + NX;
    real* xn = uh + NU;
    // candidate w writes trial slot w, stepping over slot `skip` (it holds the current trajectory, see solve_one)
    real* Xo = Xn + (size_t)(w + (w >= skip ? 1 : 0)) * xn_stride;
    real* Uo = Un + (size_t)(w + (w >= skip ? 1 : 0)) * un_stride;
    const bool active = w < ncand;
    const real alpha = active ? (real)S.alpha[w] : 0.0, omr = active ? real(1.0) - (real)S.rho[w] : 0.0;
    double J = 0.0;      // costs are summed in double in both builds (cost_lane)
    // K_k by one bulk copy when the tile and both addresses are multiples of 16 bytes (always for SRBD with caller-owned gains)
    const bool bulk = SDDP_BULK && (NU * NX) % RV == 0 && ((((size_t)S.Kbuf(0)) | ((size_t)S.Kbuf(1)) | ((size_t)Kg)) & 15) == 0;
    // node k's inputs (K_k, k_k, X_k, U_k, d_k, p_k) are fetched with cp.async while node k-1 is computed
    // The six base pointers live in shared memory during the rollout: held in registers across the node loop they were
    // spilled (the kernel sits at its 128-register cap) and re-read from local memory at every node, a quarter of them L1 misses.
    auto prefetch = [&](int k) {
        real* nb = S.nb[k & 1];
        real* Kb = S.Kbuf(k & 1);
        const real* Ks = S.gp[0] + (size_t)k * NU * NX;
        if (bulk) { if (tid == 0) bulk_g2s(Kb, Ks, NU * NX * (int)sizeof(real), &S.mbar[k & 1]); }      // one instruction for the 7.1 KB tile
        else if ((NU * NX) % RV == 0 && ((((size_t)Kb) | ((size_t)Ks)) & 15) == 0) { for (int e = RV * tid; e < NU * NX; e += RV * NT) cp_async16(Kb + e, Ks + e); }
        else { for (int e = tid; e < NU * NX; e += NT) cp_async8(Kb + e, Ks + e); }
        const real* Xs = S.gp[2] + (size_t)k * NX;
        const real* Ds = S.gp[3] + (size_t)k * NX;
        for (int i = tid; i < NX; i += NT) { cp_async8(nb + NBL::OX + i, Xs + i); cp_async8(nb + NBL::OD + i, Ds + i); }
        const real* Us = S.gp[4] + (size_t)k * NU;
        const real* ks = S.gp[1] + (size_t)k * NU;
        for (int i = tid; i < NU; i += NT) { cp_async8(nb + NBL::OU + i, Us + i); cp_async8(nb + NBL::OK + i, ks + i); }
        const real* Ps = S.gp[5] + (size_t)k * NP;
        for (int i = tid; i < NP; i += NT) cp_async8(nb + NBL::OP + i, Ps + i);
        cp_commit();
    };
    __syncthreads();      // the buffers may still be in use by the caller's previous phase
    if (tid == 0) {
        S.gp[0] = Kg; S.gp[1] = kg; S.gp[2] = X; S.gp[3] = D; S.gp[4] = U; S.gp[5] = P;
        if (bulk) { mbar_init(&S.mbar[0], 1); mbar_init(&S.mbar[1], 1); fence_async_proxy(); }      // (also orders the earlier generic writes of the K buffers before the async-proxy writes)
    }
    __syncthreads();
    prefetch(0);
    PROF(23);
    if (ncand == 1) {
        Xn += (size_t)(skip == 0 ? 1 : 0) * xn_stride;      // candidate 0's trial slot
        Un += (size_t)(skip == 0 ? 1 : 0) * un_stride;
        // One candidate (the usual first wave): its work is spread over the warps instead of leaving three idle.
        // Per node: warp 0 forms u = U + alpha k + K dx; then warp 0 integrates (accel, Euler step) while warp 1 sums
        // the state-indexed cost terms and warp 2 the input-indexed ones.  Two block barriers per node.
        real* const xb0 = S.scr();                   // x^_k ping-pong: node k in xb0 + (k & 1) * NX (no pointer array: it would live in local memory)
        real* ub = S.scr() + 2 * NX;
        for (int i = tid; i < NX; i += NT) xb0[i] = x0[i];
        if (tid == 0) { S.gp[6] = Xn; S.gp[7] = Un; }   // trial trajectory of this candidate (published by the barrier below)
        double Jw = 0.0;
        for (int k = 0; k < N; k++) {
            cp_wait_all();
            if (bulk) mbar_wait(&S.mbar[k & 1], (k >> 1) & 1);
            __syncthreads();                  // node k landed; everyone is done with node k-1
            PROF(24);
            if (k + 1 < N) prefetch(k + 1);
            else {
                real* nbt = S.nb[(k + 1) & 1];
                for (int i = tid; i < NP; i += NT) cp_async8(nbt + NBL::OP + i, S.gp[5] + (size_t)N * NP + i);
                cp_commit();
            }
            STAMP(12);
            const real* nb = S.nb[k & 1];
            const real* xc = xb0 + (k & 1) * NX;
            constexpr int MV = (4 * NU + 31) & ~31;       // whole warps take part in the shuffles
            static_assert(MV <= NT - 32, "the last warp copies x^");
            static_assert(M::NPRE <= 2 * 8, "accel_pre result lives in rows 2-3 of sacc");
            if (tid < MV) {                    // u^ = U + alpha k + K dx: four threads per row of K, partial sums by shuffle
                const real* Kb = S.Kbuf(k & 1);
                const real* xk = nb + NBL::OX;
                const int j = tid >> 2, part = tid & 3;
                constexpr int CH = (NX + 3) / 4;
                real t = 0.0;
                if (j < NU) {
#pragma unroll
                    for (int q = 0; q < CH; q++) {
                        const int i = part * CH + q;
                        if (i < NX) t += Kb[j * NX + i] * (xc[i] - xk[i]);
                    }
                }
                t += __shfl_xor_sync(FULL, t, 1);
                t += __shfl_xor_sync(FULL, t, 2);
                if (j < NU && part == 0) {
                    const real v = nb[NBL::OU + j] + (real)S.alpha[0] * nb[NBL::OK + j] + t;      // (`alpha` is per warp = per candidate)
                    ub[j] = v;
                    const_cast<real*>(S.gp[7])[(size_t)k * NU + j] = v;
                }
            } else if (w == NWARP - 1) {       // meanwhile: copy x^ out, and the state-only part of the accelerations
                for (int i = lane; i < NX; i += 32) const_cast<real*>(S.gp[6])[(size_t)k * NX + i] = xc[i];
                if (M::NACC > 1) {
                    real pre[M::NPRE];
                    M::accel_pre(c, xc, pre);
                    if (lane == 0) {
#pragma unroll
                        for (int q = 0; q < M::NPRE; q++) (&S.sacc[0][0])[16 + q] = pre[q];      // rows 2-3 of sacc (unused on this path)
                    }
                }
            }
            STAMP(13);
            __syncthreads();                  // u^_k visible
            PROF(25);
            STAMP(14);
            const int kind = node_kind(c, k);
            if (w == 0) {
                real* xn_ = xb0 + ((k + 1) & 1) * NX;
                if (M::NACC > 1) {
                    real acc[M::NACC];
                    M::accel_post(c, xc, ub, &S.sacc[0][0] + 16, acc, kind == NODE_TAIL);
                    if (lane == 0) {
#pragma unroll
                        for (int q = 0; q < M::NACC; q++) S.sacc[0][q] = acc[q];
                    }
                    __syncwarp();
                }
                Jw += M::cost_lane(c, kind, lane, xc, ub, nb + NBL::OP, S.sacc[0], 4);
                for (int i = lane; i < NX; i += 32)
                    xn_[i] = xc[i] + c.dt * M::xdot_i(c, i, xc, ub, S.sacc[0]) - omr * nb[NBL::OD + i];
            } else if (w == 1) {
                Jw += M::cost_lane(c, kind, lane, xc, ub, nb + NBL::OP, S.sacc[0], 1);
            } else if (w == 2) {
                Jw += M::cost_lane(c, kind, lane, xc, ub, nb + NBL::OP, S.sacc[0], 2);
            } else {
                Jw += M::cost_lane(c, kind, lane, xc, ub, nb + NBL::OP, S.sacc[0], 8);
            }
            STAMP(15);
            PROF(26);
        }
        cp_wait_all();
        __syncthreads();
        const real* xT = xb0 + (N & 1) * NX;
        if (w == 0) for (int i = lane; i < NX; i += 32) const_cast<real*>(S.gp[6])[(size_t)N * NX + i] = xT[i];
        if (w == 1) Jw += M::cost_lane(c, NODE_TERM, lane, xT, nullptr, S.nb[N & 1] + NBL::OP, S.sacc[0], 1);
        Jw = warp_sum(Jw);
        if (lane == 0) S.red[R_W0 + w] = Jw;
        __syncthreads();
        if (tid == 0) {
            S.Jc[0] = S.red[R_W0] + S.red[R_W0 + 1] + S.red[R_W0 + 2] + S.red[R_W0 + 3];
            if (bulk) { mbar_inval(&S.mbar[0]); mbar_inval(&S.mbar[1]); }
        }
        __syncthreads();
        PROF(27);
        return;
    }
    for (int i = lane; i < NX; i += 32) xh[i] = x0[i];
    for (int k = 0; k < N; k++) {
        cp_wait_all();
        if (bulk) mbar_wait(&S.mbar[k & 1], (k >> 1) & 1);
        __syncthreads();                  // node k landed for everyone; everyone is done with node k-1
        if (k + 1 < N) prefetch(k + 1);
        else {                            // terminal parameters go to the free buffer
            real* nb = S.nb[(k + 1) & 1];
            for (int i = tid; i < NP; i += NT) cp_async8(nb + NBL::OP + i, S.gp[5] + (size_t)N * NP + i);
            cp_commit();
        }
        if (active) {
            const real* nb = S.nb[k & 1];
            const real* Kb = S.Kbuf(k & 1);
            const real* xk = nb + NBL::OX;
            for (int j = lane; j < NU; j += 32) {
                real t = 0.0;
                for (int i = 0; i < NX; i++) t += Kb[j * NX + i] * (xh[i] - xk[i]);
                real v = nb[NBL::OU + j] + alpha * nb[NBL::OK + j] + t;
                uh[j] = v;
                Uo[(size_t)k * NU + j] = v;
            }
            for (int i = lane; i < NX; i += 32) Xo[(size_t)k * NX + i] = xh[i];
            __syncwarp();
            J += warp_node<M>(c, node_kind(c, k), xh, uh, nb + NBL::OP, xn, nb + NBL::OD, omr, S.sacc[w], lane);
            for (int i = lane; i < NX; i += 32) xh[i] = xn[i];
            __syncwarp();
        }
    }
    cp_wait_all();
    __syncthreads();
    if (active) {
        for (int i = lane; i < NX; i += 32) Xo[(size_t)N * NX + i] = xh[i];
        J += warp_node<M>(c, NODE_TERM, xh, nullptr, S.nb[N & 1] + NBL::OP, nullptr, nullptr, 0.0, S.sacc[w], lane);
        if (lane == 0) S.Jc[w] = J;
    }
    if (bulk && tid == 0) { mbar_inval(&S.mbar[0]); mbar_inval(&S.mbar[1]); }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
struct SolveArgs {
    int B;
    const real* x0; const real* params;
    real* X; real* U; real* K; real* kff; real* hist;
    int* iters; int* status; real* cost;
    // workspace, per resident CTA
    real* ws_d; real* ws_pack; real* ws_xn; real* ws_un; real* ws_K; real* ws_k;
    int* counter;
    const int* order;     // dispatch order (permutation of 0..B-1) or nullptr
    int sms;
    // result records (include/sddp.h, sddp_set_result_peers): X | U | cost | iters | status of problem b goes to
    // peers[p] + (first + b) * rec for every p < n_peers -- own slab and, over NVLink, the slabs of the other GPUs
    int n_peers, rec;
    long long first;
    real* peers[SDDP_MAX_PEERS];
    // host-direct mode (sddp_solve_batch_host on pinned buffers): mapped HOST pointers.  The CTA that takes problem b pulls its
    // inputs over PCIe into the device arrays above (x0, params, X, U are then the handle's staging arrays) and stores its
    // results straight into the caller's host arrays: one launch, every transfer rides under the solves of the other CTAs.
    const real* h_x0; const real* h_params; const real* h_X0; const real* h_U0;
    real* h_X; real* h_U; real* h_kff; real* h_hist; real* h_cost; int* h_iters; int* h_status;
};

// `scratch`: PACK_SCRATCH * NT doubles of shared memory nobody else uses during the call (one column per thread)
constexpr int PACK_SCRATCH = 10;
template <class M>
__device__ void compute_packs(const DevCfg& c, const real* X, const real* U, real* packs, real* scratch, int tid) {
    if (M::PACK > 1)
        for (int k = tid; k < c.N; k += NT)
            M::pack(c, node_kind(c, k), X + (size_t)k * M::NX, U + (size_t)k * M::NU, packs + (size_t)k * M::PACK, scratch + tid, NT);
    __syncthreads();
}

template <class M, class SM>
__device__ void solve_one(const DevCfg& c, const SolveArgs& a, SM& S, int b, int slot, int tid) {
    constexpr int NX = M::NX, NU = M::NU, NP = M::NP;
    const int N = c.N;
    const size_t xsz = (size_t)(N + 1) * NX, usz = (size_t)N * NU;
    const real* x0 = a.x0 + (size_t)b * NX;
    const real* P = a.params + (size_t)b * (N + 1) * NP;
    if (a.h_params) {      // host-direct: this problem's inputs, straight from the caller's pinned buffers
        const size_t psz = (size_t)(N + 1) * NP;
        real* Pd = const_cast<real*>(P);
        real* xd = const_cast<real*>(x0);
        const real* hp = a.h_params + (size_t)b * psz;
        const real* hX = a.h_X0 + (size_t)b * xsz;
        const real* hU = a.h_U0 + (size_t)b * usz;
        real* Xd = a.X + (size_t)b * xsz;
        real* Ud = a.U + (size_t)b * usz;
#pragma unroll 8
        for (size_t i = tid; i < psz; i += NT) Pd[i] = hp[i];
#pragma unroll 8
        for (size_t i = tid; i < xsz; i += NT) Xd[i] = hX[i];
#pragma unroll 8
        for (size_t i = tid; i < usz; i += NT) Ud[i] = hU[i];
        for (int i = tid; i < NX; i += NT) xd[i] = a.h_x0[(size_t)b * NX + i];
        __syncthreads();
    }
    real* X = a.X + (size_t)b * xsz;
    real* U = a.U + (size_t)b * usz;
    real* Kg = a.K ? a.K + (size_t)b * N * NU * NX : a.ws_K + (size_t)slot * N * NU * NX;
    real* kg = a.kff ? a.kff + (size_t)b * usz : a.ws_k + (size_t)slot * usz;
    real* hist = a.hist ? a.hist + (size_t)b * c.max_iters * 4 : nullptr;
    real* d = a.ws_d + (size_t)slot * N * NX;
    real* packs = a.ws_pack + (size_t)slot * N * M::PACK;
    // NSLOT trial slots per CTA: the accepted candidate's slot BECOMES the current trajectory (no copy per iteration),
    // the next waves use the other slots; the caller's X, U are written once, at the end.
    real* Xn = a.ws_xn + (size_t)slot * NSLOT * xsz;
    real* Un = a.ws_un + (size_t)slot * NSLOT * usz;
    real* const Xuser = X;
    real* const Uuser = U;
    int cur = 1 << 30;      // trial slot that holds the current trajectory (none: it is still in the caller's buffers)
    const bool fixed = c.rho_fixed > 0.0;

    for (int i = tid; i < NX; i += NT) X[i] = x0[i];
    if (hist) for (int i = tid; i < c.max_iters * 4; i += NT) hist[i] = 0.0;
    __syncthreads();
    // The scalar logic of the iteration (costs, expected decrease, step sizes, regularisation) is double in both builds.
    double J, dmax = 0.0;
    bool bad_start = false;      // a non-finite initial gap (NaN / inf in the warm start or x0)
    if (!c.ms) {
        open_loop_rollout<M, SM>(c, S, X, U, tid);
        for (int i = tid; i < N * NX; i += NT) d[i] = 0.0;
        __syncthreads();
        J = defects_and_cost<M, SM>(c, S, X, U, P, nullptr, tid);
    } else {
        J = defects_and_cost<M, SM>(c, S, X, U, P, d, tid);
        real m = 0.0;
        bool nf = false;
        for (int i = tid; i < N * NX; i += NT) { const real v = fabs(d[i]); nf |= !(v <= 1.79e308); if (v > m) m = v; }      // fmax would drop a NaN
        m = warp_max(m);
        if ((tid & 31) == 0) S.red[R_W0 + (tid >> 5)] = m;
        bad_start = __syncthreads_or(nf) != 0;
        for (int i = 0; i < NWARP; i++) dmax = fmax(dmax, S.red[R_W0 + i]);
        __syncthreads();
    }
    double mu = c.mu0;
    int status = bad_start ? 4 /*NAN*/ : 1 /*MAX_ITERS*/, it = 0;
    PROF(4);
    PROF_INC(31);
    for (it = 0; it < (bad_start ? 0 : c.max_iters); it++) {
        PROF_INC(30);
        SM::prep(c, S, X, U, P, packs, tid);
        PROF(0);
        bool reg_fail = false;
        while (SM::backward(c, S, X, U, P, d, packs, mu, Kg, kg, &S.red[12], dmax != 0.0, tid)) {
            mu = fmax(mu * c.mu_factor, c.mu_min);
            if (mu > c.mu_max) { reg_fail = true; break; }
        }
        PROF(1);
        const double D1 = S.red[12], D2 = S.red[13], C0 = S.red[14];
        real* h = hist ? hist + it * 4 : nullptr;
        if (h && tid == 0) { h[0] = J; h[1] = 0.0; h[2] = mu; h[3] = dmax; }
        if (reg_fail) { status = 3; it++; break; }
        if (!isfinite(D1) || !isfinite(D2) || !isfinite(J)) { status = 4; it++; break; }
        const double a0 = c.alpha0;
        const double dJ0 = C0 + a0 * D1 + a0 * a0 * D2;
        if (fabs(dJ0) <= 1e-3 * c.cost_ths * (1.0 + fabs(J)) && dmax <= c.defect_ths) { status = 0; it++; break; }
        // parallel line search: waves of NCAND candidates, first (largest) passing alpha wins
        bool accepted = false;
        double alpha = a0, Jn = J, alpha_acc = 0.0, rho_acc = 0.0;
        int sel = 0;
        bool first_wave = true;
        while (!accepted && alpha >= c.alpha_min) {
            // The first wave tries alpha_0 alone: it is accepted in the vast majority of iterations, and the three
            // spared rollouts leave issue slots and shared-memory bandwidth to the co-resident CTAs.  Later waves
            // evaluate NCAND candidates in parallel; the accepted step is the same as with sequential backtracking.
            const int wave = (first_wave && SDDP_FIRST_WAVE_SINGLE) ? 1 : NCAND;
            first_wave = false;
            int ncand = 0;
            double al[NCAND];
            for (int j = 0; j < NCAND; j++) {
                al[j] = alpha;
                if (j < wave) {
                    if (alpha >= c.alpha_min) ncand++;
                    alpha *= c.ls_factor;
                }
            }
            __syncthreads();
            if (tid < NCAND) { S.alpha[tid] = al[tid]; S.rho[tid] = fixed ? c.rho_fixed : al[tid]; }
            __syncthreads();
            PROF(5);
            PROF_INC(29);
            forward_wave<M, SM>(c, S, x0, X, U, P, d, Kg, kg, ncand, Xn, xsz, Un, usz, tid, cur);
            PROF(2);
            // fp32 build only: the expected decrease dJm comes out of a float Riccati recursion, so it carries a rounding
            // noise of about 1e-7 (1 + |J|) that does not shrink with the step size (C0, the gap-closing part, does not
            // depend on alpha at all).  Without a floor of that size the test below rejects every candidate once the
            // true decrease falls under the noise -- fixed-rate gap contraction (README.md:6) runs many such iterations
            // while the gaps halve -- and the solve ends in LS_FAILED instead of converging.  0 in the fp64 build.
#ifdef SDDP_F32
#define SDDP_LS_NOISE +1e-6 * (1.0 + fabs(J))
#else
#define SDDP_LS_NOISE
#endif
            for (int j = 0; j < ncand; j++) {
                double am = al[j], dJm = C0 + am * D1 + am * am * D2, Jj = S.Jc[j];
                if (isfinite(Jj) && Jj - J <= dJm + (1.0 - c.beta) * fabs(dJm) SDDP_LS_NOISE) {
                    accepted = true; sel = j; Jn = Jj; alpha_acc = am; rho_acc = fixed ? c.rho_fixed : am;
                    break;
                }
            }
        }
        if (accepted) {
            cur = sel + (sel >= cur ? 1 : 0);
            X = Xn + (size_t)cur * xsz;
            U = Un + (size_t)cur * usz;
            const double omr = 1.0 - rho_acc;
            const real omr_r = (real)omr;
            for (int i = tid; i < N * NX; i += NT) d[i] *= omr_r;
            dmax *= omr;
            const double dJ = J - Jn;
            J = Jn;
            if (h && tid == 0) { h[0] = J; h[1] = alpha_acc; h[3] = dmax; }
            mu = mu / c.mu_factor;
            if (mu < c.mu_min) mu = 0.0;
            if (mu < c.mu0) mu = c.mu0;
            __syncthreads();
            PROF(3);
            if (dJ <= c.cost_ths * (1.0 + fabs(J)) && dmax <= c.defect_ths) { status = 0; it++; break; }
        } else {
            mu = fmax(mu * c.mu_factor, c.mu_min);
            if (mu > c.mu_max) { status = 2; it++; break; }
        }
    }
    __syncthreads();
    if (tid == 0) { a.iters[b] = it; a.status[b] = status; a.cost[b] = J; }
    if (a.h_X) {           // host-direct: results straight into the caller's pinned buffers (posted writes over PCIe)
        real* hX = a.h_X + (size_t)b * xsz;
        real* hU = a.h_U + (size_t)b * usz;
        for (size_t i = tid; i < xsz; i += NT) hX[i] = X[i];
        for (size_t i = tid; i < usz; i += NT) hU[i] = U[i];
        if (a.h_kff) for (size_t i = tid; i < usz; i += NT) a.h_kff[(size_t)b * usz + i] = kg[i];
        if (a.h_hist) for (int i = tid; i < c.max_iters * 4; i += NT) a.h_hist[(size_t)b * c.max_iters * 4 + i] = hist[i];
        if (tid == 0) { a.h_iters[b] = it; a.h_status[b] = status; a.h_cost[b] = J; }
    } else if (X != Xuser) {
        for (size_t i = tid; i < xsz; i += NT) Xuser[i] = X[i];
        for (size_t i = tid; i < usz; i += NT) Uuser[i] = U[i];
    }
    if (a.n_peers > 0) {
        // Gather without a collective: this problem's record is stored straight into every GPU's whole-batch slab (peer
        // memory over NVLink / NVSwitch) while the other CTAs keep solving, so the transfer rides under the batch.
        const size_t off = (size_t)(a.first + b) * a.rec;
        for (size_t i = tid; i < xsz + usz + 3; i += NT) {
            real v;
            if (i < xsz) v = X[i];
            else if (i < xsz + usz) v = U[i - xsz];
            else v = (i == xsz + usz) ? J : (i == xsz + usz + 1 ? (real)it : (real)status);
#pragma unroll
            for (int p = 0; p < SDDP_MAX_PEERS; p++)
                if (p < a.n_peers) a.peers[p][off + i] = v;
        }
    }
    PROF(6);
}

template <class M>
__device__ void Smem<M>::prep(const DevCfg& c, Smem<M>& S, const real* X, const real* U, const real*, real* packs, int tid) {
    static_assert(M::PACK == 1 || sizeof(Smem<M>) >= PACK_SCRATCH * NT * sizeof(real), "pack scratch");
    compute_packs<M>(c, X, U, packs, S.Vxx, tid);      // the matrices are dead between the forward and the backward pass
}

template <class M>
__device__ int Smem<M>::backward(const DevCfg& c, Smem<M>& S, const real* X, const real* U, const real* P, const real* D,
                                 const real* packs, real mu, real* Kg, real* kg, double* dV3, bool, int tid) {
    return backward_pass<M>(c, S, X, U, P, D, packs, mu, Kg, kg, dV3, tid);
}
