// sddp.cu -- kernels and the C ABI (include/sddp.h) of the B200-native batched DDP solver.
// Built for sm_100a only:  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>
#include <new>
#include <vector>

#include "../../include/sddp.h"
#include "sddp_backward_srbd.cuh"

// ===================================================================================== kernels
template <class M, class SM, int MINB>
__global__ void __launch_bounds__(NT, MINB) solve_kernel(DevCfg c, SolveArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    __shared__ int s_prob;
    // (Warp w of every resident CTA shares one SM sub-partition.  Rotating the warp roles per co-resident CTA was
    //  measured slower: same-role warps share their code in the sub-partition's instruction cache.)
    const int tid = threadIdx.x;
    PROF_RESET;
    for (;;) {   // persistent CTA: pull problems from a queue (iteration counts differ per problem)
        if (tid == 0) s_prob = atomicAdd(a.counter, 1);
        __syncthreads();
        PROF(7);
        if (s_prob >= a.B) break;
        const int b = a.order ? a.order[s_prob] : s_prob;
        if (b < 0 || b >= a.B) continue;      // not a permutation (caller error, include/sddp.h): skip, status stays -1
        solve_one<M, SM>(c, a, S, b, blockIdx.x, tid);
        __syncthreads();
    }
}

// Stage 1 alone: one CTA per (x,u,p) point, dense outputs.
template <class M>
__global__ void __launch_bounds__(NT) eval_kernel(DevCfg c, int npts, const int* kind, const real* x, const real* u, const real* p,
                                                  real* f, real* fx, real* fu, real* l, real* lx, real* lu, real* lxx,
                                                  real* lux, real* luu) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<M>& S = *reinterpret_cast<Smem<M>*>(smem_raw);
    constexpr int NX = M::NX, NU = M::NU, NP = M::NP;
    const int tid = threadIdx.x;
    SyncBlock sync;
    using NBL = NodeBuf<M>;
    real* xk = S.nb[0] + NBL::OX;
    real* uk = S.nb[0] + NBL::OU;
    real* pk = S.nb[0] + NBL::OP;
    real* pack = S.nb[0] + NBL::OK;
    for (int m = blockIdx.x; m < npts; m += gridDim.x) {
        const int kd = kind[m];
        for (int i = tid; i < NX; i += NT) xk[i] = x[(size_t)m * NX + i];
        for (int i = tid; i < NU; i += NT) uk[i] = u[(size_t)m * NU + i];
        for (int i = tid; i < NP; i += NT) pk[i] = p[(size_t)m * NP + i];
        __syncthreads();
        if (tid == 0 && kd != NODE_TERM) M::pack(c, kd, xk, uk, pack, S.Vxx, 1);
        __syncthreads();
        M::expand(c, kd, xk, uk, pk, pack, S.Qx, S.Qu, S.Qxx, S.Qux, S.Quu, tid, NT, sync);
        if (kd != NODE_TERM) M::expand_f(c, xk, uk, pack, S.fx, S.fu, tid, NT, sync);
        else {
            for (int e = tid; e < NX * NX; e += NT) S.fx[e] = 0.0;
            for (int e = tid; e < NX * NU; e += NT) S.fu[e] = 0.0;
            for (int e = tid; e < NX; e += NT) S.vp[e] = 0.0;
            __syncthreads();
        }
        if (tid < 32) {
            real J = warp_node<M>(c, kd, xk, uk, pk, S.vp, nullptr, 0.0, S.sacc[0], tid);
            if (tid == 0 && l) l[m] = J;
        }
        __syncthreads();
        if (f)   for (int e = tid; e < NX; e += NT) f[(size_t)m * NX + e] = S.vp[e];
        if (fx)  for (int e = tid; e < NX * NX; e += NT) fx[(size_t)m * NX * NX + e] = S.fx[e];
        if (fu)  for (int e = tid; e < NX * NU; e += NT) fu[(size_t)m * NX * NU + e] = S.fu[e];
        if (lx)  for (int e = tid; e < NX; e += NT) lx[(size_t)m * NX + e] = S.Qx[e];
        if (lu)  for (int e = tid; e < NU; e += NT) lu[(size_t)m * NU + e] = S.Qu[e];
        if (lxx) for (int e = tid; e < NX * NX; e += NT) lxx[(size_t)m * NX * NX + e] = S.Qxx[e];
        if (lux) for (int e = tid; e < NU * NX; e += NT) lux[(size_t)m * NU * NX + e] = S.Qux[e];
        if (luu) for (int e = tid; e < NU * NU; e += NT) luu[(size_t)m * NU * NU + e] = S.Quu[e];
        __syncthreads();
    }
}

template <class M, class SM, int MINB>
__global__ void __launch_bounds__(NT, MINB) backward_kernel(DevCfg c, int B, const real* X, const real* U, const real* P, const real* D,
                                                      real mu, real* K, real* kff, real* dV, int* rc, real* ws_pack) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    constexpr int NX = M::NX, NU = M::NU, NP = M::NP;
    const int tid = threadIdx.x, N = c.N;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const real* Xb = X + (size_t)b * (N + 1) * NX;
        const real* Ub = U + (size_t)b * N * NU;
        real* packs = ws_pack + (size_t)blockIdx.x * N * M::PACK;
        SM::prep(c, S, Xb, Ub, P + (size_t)b * (N + 1) * NP, packs, tid);
        int r = SM::backward(c, S, Xb, Ub, P + (size_t)b * (N + 1) * NP, D + (size_t)b * N * NX, packs, mu,
                             K + (size_t)b * N * NU * NX, kff + (size_t)b * N * NU, &S.red[12], true, tid);
        if (tid == 0) { rc[b] = r; dV[3 * b] = (real)S.red[12]; dV[3 * b + 1] = (real)S.red[13]; dV[3 * b + 2] = (real)S.red[14]; }
        __syncthreads();
    }
}

template <class M, class SM, int MINB>
__global__ void __launch_bounds__(NT, MINB) forward_kernel(DevCfg c, int B, int n_alpha, const real* alpha, const real* rho, const real* x0,
                                                     const real* X, const real* U, const real* P, const real* D, const real* K,
                                                     const real* kff, real* Jn, real* Xn, real* Un, real* ws_xn, real* ws_un) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    constexpr int NX = M::NX, NU = M::NU, NP = M::NP;
    const int tid = threadIdx.x, N = c.N;
    const size_t xsz = (size_t)(N + 1) * NX, usz = (size_t)N * NU;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int base = 0; base < n_alpha; base += NCAND) {
            const int ncand = min(NCAND, n_alpha - base);
            __syncthreads();
            if (tid < ncand) { S.alpha[tid] = alpha[base + tid]; S.rho[tid] = rho[base + tid]; }
            __syncthreads();
            real* xo = Xn ? Xn + ((size_t)b * n_alpha + base) * xsz : ws_xn + (size_t)blockIdx.x * NSLOT * xsz;
            real* uo = Un ? Un + ((size_t)b * n_alpha + base) * usz : ws_un + (size_t)blockIdx.x * NSLOT * usz;
            forward_wave<M, SM>(c, S, x0 + (size_t)b * NX, X + (size_t)b * xsz, U + (size_t)b * usz, P + (size_t)b * (N + 1) * NP,
                            D + (size_t)b * N * NX, K + (size_t)b * N * NU * NX, kff + (size_t)b * usz, ncand, xo, xsz, uo, usz, tid);
            if (tid < ncand) Jn[(size_t)b * n_alpha + base + tid] = S.Jc[tid];
        }
    }
}

template <class M, class SM, int MINB>
__global__ void __launch_bounds__(NT, MINB) defects_kernel(DevCfg c, int B, const real* X, const real* U, const real* P, real* D, real* cost) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    constexpr int NX = M::NX, NU = M::NU, NP = M::NP;
    const int tid = threadIdx.x, N = c.N;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        real J = defects_and_cost<M, SM>(c, S, X + (size_t)b * (N + 1) * NX, U + (size_t)b * N * NU, P + (size_t)b * (N + 1) * NP,
                                       D ? D + (size_t)b * N * NX : nullptr, tid);
        if (tid == 0 && cost) cost[b] = J;
    }
}

// ---- receding-horizon glue: schedule shift + gait fill, plant step
template <class M>
__global__ void mpc_advance_kernel(DevCfg c, int B, real* params, const int* action, int* counter, const real* cmd,
                                   const real* tab /* l_cycle, l_switch, r_cycle, r_switch: 4 x 21 */) {
    constexpr int NP = M::NP;
    const int N = c.N;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= B * NP) return;
    const int b = g / NP, p = g % NP;
    real* P = params + (size_t)b * (N + 1) * NP;
    for (int j = 1; j <= N; j++) P[(size_t)(j - 1) * NP + p] = P[(size_t)j * NP + p];      // one node back
    const int act = action[b], ref_id = counter[b] % 20;                                     // wpg.py:71
    real* last = P + (size_t)N * NP;
    constexpr bool srbd = (M::NP == 19);
    constexpr int base = srbd ? 7 : 3;            // first (c_ref, cdot_switch) pair
    if (p < 3) last[p] = cmd[3 * b + p];                                                     // dsrbd_example.py:115-122
    else if (srbd && p < 6) last[p] = 0.0;                                                   // w_ref, wpg.py:81,90,95
    else if (srbd && p == 6) last[p] = (act == 2) ? 0.0 : 1e2;                               // wpg.py:82,91,96
    else if (p >= base && p < base + 8) {
        const int i = (p - base) >> 1, is_sw = (p - base) & 1;
        const bool left = i < 2;                                                             // contact_model = 2
        if (act == 0) last[p] = tab[(left ? 0 : 42) + (is_sw ? 21 : 0) + ref_id];            // wpg.py:83-88
        else if (act == 2) { if (is_sw) last[p] = 0.0; }                                     // wpg.py:92-93 (c_ref untouched)
        else last[p] = is_sw ? 1.0 : 0.0;                                                    // wpg.py:97-99
    }
}
__global__ void counter_inc_kernel(int B, int* counter) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) counter[b] += 1;                                                              // wpg.py:101
}

template <class M>
__global__ void plant_step_kernel(DevCfg c, int B, real* state, const real* u, long long u_stride) {
    constexpr int NX = M::NX, NU = M::NU;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    real x[NX], uu[NU], acc[M::NACC < 6 ? 6 : M::NACC], xn[NX];
    for (int i = 0; i < NX; i++) x[i] = state[(size_t)b * NX + i];
    for (int i = 0; i < NU; i++) uu[i] = u[(size_t)b * u_stride + i];
    M::accel(c, x, uu, acc);
    for (int i = 0; i < NX; i++) xn[i] = x[i] + c.dt * M::xdot_i(c, i, x, uu, acc);
    if (M::NP == 19) {      // SRBD: state[3:7] /= norm (dsrbd_example.py:160)
        const real n = sqrt(xn[3] * xn[3] + xn[4] * xn[4] + xn[5] * xn[5] + xn[6] * xn[6]);
        for (int i = 3; i < 7; i++) xn[i] /= n;
    }
    for (int i = 0; i < NX; i++) state[(size_t)b * NX + i] = xn[i];
}

// FP64 FMA-rate microbenchmark: 8 independent chains per thread
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = fma(v[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    if (s == 123.456) out[0] = s;
}

// ===================================================================================== host side
struct SddpHandle {
    SddpConfig cfg;
    DevCfg dc;
    int device, sms, slots, variant;
    size_t smem_bytes, eval_smem_bytes, ws_bytes;
    real *ws_d, *ws_pack, *ws_xn, *ws_un, *ws_K, *ws_k;
    int* counter;
    unsigned long long* ztab;
    real* gait;        // device copy of the four 21-entry wpg tables
    // staging for the *_host entry point
    void* stage; size_t stage_bytes;
    void* hstage; size_t hstage_bytes;       // pinned host staging of the small-batch path
    cudaStream_t st_in, st_cmp, st_out;
    cudaEvent_t ev_last; bool ev_last_valid;   // recorded after the last launch that uses the workspace
    int host_chunk;
    long long launches;
    real* slab; long long slab_records;       // handle-owned result slab (sddp_slab_alloc)
    int n_peers; real* peers[SDDP_MAX_PEERS]; long long first_record;      // sddp_set_result_peers
    const int32_t* order_dev; int order_n;      // caller-owned device permutation for sddp_solve_batch
    std::vector<int32_t> order_host;            // copy of the host permutation for sddp_solve_batch_host
    char err[512];
};
static char g_create_err[512] = "";

static int fail(SddpHandle* h, int code, const char* fmt, const char* a, const char* b) {
    char* dst = h ? h->err : g_create_err;
    snprintf(dst, 512, fmt, a, b);
    return code;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return fail(h, SDDP_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// The handle's workspace serves one solve at a time: remember the end of the last launch that used it.
static cudaError_t mark_last(SddpHandle* h, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    // Under CUDA-graph capture (mpc.BatchedMPC.capture: a whole MPC tick as one graph) nothing is recorded: an event recorded
    // into a capturing stream cannot be synchronised on later, and the graph's own edges order its launches.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) { h->ev_last_valid = false; return cudaSuccess; }
    if (!h->ev_last && (e = cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming)) != cudaSuccess) return e;
    e = cudaEventRecord(h->ev_last, st);
    h->ev_last_valid = (e == cudaSuccess);
    return e;
}

// A handle is bound to the device that was current at sddp_create (workspace, streams); calling with another current
// device would launch there against this device's memory.
static int check_device(SddpHandle* h) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != h->device)
        return fail(h, SDDP_EINVAL, "%s%s", "the current CUDA device is not the one this handle was created on (cudaSetDevice first)", "");
    return 0;
}

static int check_config(const SddpConfig* c, SddpHandle* h) {
    if (!c) return fail(h, SDDP_EINVAL, "%s%s", "config is NULL", "");
    if (c->model != SDDP_MODEL_SRBD && c->model != SDDP_MODEL_LIP) return fail(h, SDDP_EINVAL, "%s%s", "unknown model", "");
    if (c->N < 1 || c->N > 4096) return fail(h, SDDP_EINVAL, "%s%s", "N must be in 1..4096", "");
    if (c->max_iters < 1) return fail(h, SDDP_EINVAL, "%s%s", "max_iters must be >= 1", "");
    if (!(c->dt > 0.0)) return fail(h, SDDP_EINVAL, "%s%s", "dt must be > 0", "");
    if (!(c->line_search_decrease_factor > 0.0 && c->line_search_decrease_factor < 1.0))
        return fail(h, SDDP_EINVAL, "%s%s", "line_search_decrease_factor must be in (0,1)", "");
    if (!(c->alpha_converge_threshold > 0.0)) return fail(h, SDDP_EINVAL, "%s%s", "alpha_converge_threshold must be > 0", "");
    if (!(c->mu_factor > 1.0) || !std::isfinite(c->mu_factor)) return fail(h, SDDP_EINVAL, "%s%s", "mu_factor must be finite and > 1", "");
    // the in-kernel regularisation and line-search loops terminate only for these (sddp_solver.cuh solve_one)
    if (!std::isfinite(c->alpha_0) || !(c->alpha_0 > 0.0)) return fail(h, SDDP_EINVAL, "%s%s", "alpha_0 must be finite and > 0", "");
    if (!std::isfinite(c->alpha_converge_threshold)) return fail(h, SDDP_EINVAL, "%s%s", "alpha_converge_threshold must be finite", "");
    if (!std::isfinite(c->mu0) || !(c->mu0 >= 0.0)) return fail(h, SDDP_EINVAL, "%s%s", "mu0 must be finite and >= 0", "");
    if (!std::isfinite(c->mu_min) || !(c->mu_min > 0.0)) return fail(h, SDDP_EINVAL, "%s%s", "mu_min must be finite and > 0", "");
    if (!std::isfinite(c->mu_max) || !(c->mu_max >= c->mu_min)) return fail(h, SDDP_EINVAL, "%s%s", "mu_max must be finite and >= mu_min", "");
    if (!std::isfinite(c->beta) || !std::isfinite(c->cost_reduction_ths) || !std::isfinite(c->defect_ths))
        return fail(h, SDDP_EINVAL, "%s%s", "beta, cost_reduction_ths and defect_ths must be finite", "");
    if (!std::isfinite(c->dt)) return fail(h, SDDP_EINVAL, "%s%s", "dt must be finite", "");
    if (c->defect_contraction_rate > 1.0) return fail(h, SDDP_EINVAL, "%s%s", "defect_contraction_rate must be <= 1", "");
    if (!(c->friction_cone_weight >= 0.0) || !(c->friction_cone_mu >= 0.0) || !(c->friction_cone_sharpness >= 0.0))
        return fail(h, SDDP_EINVAL, "%s%s", "friction_cone_weight, _mu and _sharpness must be >= 0", "");
    if (!(c->force_bound_weight >= 0.0) || !(c->unilateral_weight >= 0.0) || !(c->cdot_bound_weight >= 0.0) || !(c->bound_sharpness >= 0.0) ||
        !(c->force_bound >= 0.0) || !(c->cdot_bound >= 0.0) || !std::isfinite(c->force_bound) || !std::isfinite(c->cdot_bound) ||
        !std::isfinite(c->bound_sharpness) || !std::isfinite(c->friction_cone_sharpness) || !std::isfinite(c->friction_cone_mu))
        return fail(h, SDDP_EINVAL, "%s%s", "bound weights, bounds and sharpness must be finite and >= 0", "");
    if (c->lip_tail_start < 0 || c->lip_tail_start > c->N) return fail(h, SDDP_EINVAL, "%s%s", "lip_tail_start must be in 0..N", "");
    if (c->model != SDDP_MODEL_SRBD && c->lip_tail_start != 0) return fail(h, SDDP_EINVAL, "%s%s", "lip_tail_start is an SRBD option", "");
    return 0;
}

static void make_devcfg(const SddpConfig& s, DevCfg& d) {
    d.model = s.model; d.N = s.N; d.inertia_mode = s.inertia_mode; d.hessian_mode = s.hessian_mode;
    d.ms = s.multiple_shooting; d.max_iters = s.max_iters;
    d.dt = s.dt; d.fs = s.force_scaling; d.g = s.gravity; d.eta2 = s.eta2;
    d.mscaled = s.mass / s.force_scaling; d.inv_ms = 1.0 / d.mscaled;
    for (int i = 0; i < 9; i++) d.Ib[i] = s.inertia[i] / s.force_scaling;
    for (int i = 0; i < 3; i++) d.com[i] = s.com[i];
    for (int i = 0; i < 12; i++) d.foot[i] = s.foot[i];
    d.w_r = s.r_tracking_gain; d.w_rdot = s.rdot_tracking_gain; d.w_w = s.w_tracking_gain; d.w_rel = s.rel_position_gain;
    d.w_fsw = s.force_scaling * s.force_scaling * s.force_switch_weight;
    d.w_minf = s.force_scaling * s.force_scaling * s.min_f_gain;
    d.gq = s.min_qddot_gain; d.w_zmp = s.zmp_tracking_gain; d.cw = s.constraint_weight;
    for (int pr = 0; pr < 2; pr++)
        for (int ax = 0; ax < 2; ax++) d.drel[pr][ax] = -(s.foot[3 * pr + ax] - s.foot[3 * (pr + 2) + ax]);
    d.alpha0 = s.alpha_0; d.alpha_min = s.alpha_converge_threshold; d.ls_factor = s.line_search_decrease_factor;
    d.beta = s.beta; d.cost_ths = s.cost_reduction_ths; d.mu0 = s.mu0; d.rho_fixed = s.defect_contraction_rate;
    d.mu_min = s.mu_min; d.mu_max = s.mu_max; d.mu_factor = s.mu_factor; d.defect_ths = s.defect_ths;
    const bool srbd = s.model == SDDP_MODEL_SRBD;
    d.w_cone = srbd ? s.friction_cone_weight : 0.0; d.cone_mu = s.friction_cone_mu; d.cone_k = s.friction_cone_sharpness;
    d.w_fb = srbd ? s.force_bound_weight : 0.0; d.fb = s.force_bound; d.w_uni = srbd ? s.unilateral_weight : 0.0;
    d.w_cdb = srbd ? s.cdot_bound_weight : 0.0; d.cdb = s.cdot_bound; d.kb = s.bound_sharpness;
    d.ineq = (d.w_cone != 0.0 || d.w_fb != 0.0 || d.w_uni != 0.0 || d.w_cdb != 0.0) ? 1 : 0;
    d.lip_tail = srbd ? s.lip_tail_start : 0;
    d.ztab = nullptr;
}

// Host mirror of Srbd::zmap_x / zmap_u / hc(): descriptors of the upper-triangular (pi <= qi) entries of the
// 34 x 34 Hessian block of gq*||wdot||^2, consumed by Srbd::expand (bit layout: sddp_model.cuh ZT_*).
// returns false if the tables do not fit their slots (a programming error caught at sddp_create, never at run time)
static bool build_ztab(std::vector<unsigned long long>& tab) {
    const int NZ = Srbd::NZ, ZO = Srbd::ZO, ZC = Srbd::ZC, ZW = Srbd::ZW, ZF = Srbd::ZF;
    auto zx = [&](int p) { return p < 19 ? p : (p < 22 ? (int)Srbd::XW + (p - 19) : -1); };
    auto zu = [&](int p) { return 6 * ((p - 22) / 3) + 3 + (p - 22) % 3; };
    tab.assign((size_t)ZT_TOTAL, 0ull);
    int e = 0, nc = 0;
    for (int pass = 0; pass < 3; pass++)      // xx entries first, then ux (ZT_NXX_NUX together), then uu: warps see one kind
    for (int pi = 0; pi < NZ; pi++)
        for (int qi = pi; qi < NZ; qi++) {
            const int xi = zx(pi), xj = zx(qi);
            if (pass != ((xi >= 0 && xj >= 0) ? 0 : (xi >= 0 ? 1 : 2))) continue;
            int kind, da, db;
            if (xi >= 0 && xj >= 0) { kind = 0; da = xi; db = xj; }
            else if (xi >= 0) { kind = 1; da = zu(qi); db = xi; }
            else { kind = 2; da = zu(pi); db = zu(qi); }
            int hs = 0, hoff = 0;   // curvature source: sign * pack[hoff]
            auto skew = [&](int a, int b, int flip) {      // +-nu[k] of skew(nu)[a][b]
                if (a == b) return;
                const int k = 3 - a - b;
                int neg = ((b - a + 3) % 3 == 1) ? 1 : 0;
                neg ^= flip;
                hs = neg ? 2 : 1; hoff = Srbd::PK_NU + k;
            };
            if (pi >= ZO && pi < ZC) { hs = 1; hoff = Srbd::PK_HO + (pi - ZO) * NZ + qi; }
            else if (qi >= ZO && qi < ZC) { hs = 1; hoff = Srbd::PK_HO + (qi - ZO) * NZ + pi; }
            else if (pi >= ZW && pi < ZF && qi >= ZW && qi < ZF) { hs = 1; hoff = Srbd::PK_HWW + 3 * (pi - ZW) + (qi - ZW); }
            else if (qi >= ZF) {
                const int fi = (qi - ZF) / 3, b = (qi - ZF) % 3;
                if (pi < ZO) skew(pi, b, 0);                                                   // (r_a, f_ib): +skew(nu)[a][b]
                else if (pi >= ZC && pi < ZW && (pi - ZC) / 3 == fi) skew((pi - ZC) % 3, b, 1);  // (c_ia, f_ib): -skew(nu)[a][b]
            }
            unsigned long long d = 1ull | ((unsigned long long)kind << 1) | ((unsigned long long)da << 3) | ((unsigned long long)db << 9) |
                                   ((unsigned long long)pi << 15) | ((unsigned long long)qi << 21) | ((unsigned long long)hs << 27) |
                                   ((unsigned long long)hoff << 29);
            if (kind == 0) d |= ((unsigned long long)(da * Srbd::NX + db) << 38) | ((unsigned long long)(db * Srbd::NX + da) << 50);
            if (kind == 1) d |= ((unsigned long long)(ZT_QUX_OFF + da * ZT_LDUX + db) << 38) | ((unsigned long long)(ZT_QUX_OFF + da * ZT_LDUX + db) << 50);
            tab[e++] = d;
            if (kind != 2 && hs != 0) {      // compact list of the structured backward pass (layout: Srbd::apply_rec)
                if (nc >= ZT_CROUNDS * ZT_LAZY_THREADS) return false;
                const unsigned long long d1 = kind == 0 ? da * Srbd::NX + db : ZT_QUX_OFF + da * ZT_LDUX + db;
                const unsigned long long d2 = kind == 0 ? db * Srbd::NX + da : d1;
                unsigned long long xi = 0;      // overlapping affine Hessian value (index + 1): (o, o) and the (w, w) diagonal
                if (pi >= ZO && pi < ZC && qi >= ZO && qi < ZC) { const int a = pi - ZO, b = qi - ZO; xi = 1 + Srbd::AF_OO + (a * 4 - a * (a - 1) / 2) + (b - a); }
                if (pi >= ZW && pi < ZF && qi == pi) xi = 1 + Srbd::AF_WW;
                tab[ZT_COFF + nc++] = 1ull | ((unsigned long long)hs << 1) | ((unsigned long long)hoff << 3) | (d1 << 12) | (d2 << 24) | (xi << 36);
            }
        }
    if (e != NZ * (NZ + 1) / 2) return false;
    // Quu work table (sddp_backward_srbd.cuh, phase c1).  Types: 1 (f_a, f_b) a >= b, 2 (cddot_a, f_b), 3 (cddot_a, cddot_b) a >= b.
    auto desc = [](int type, int i1, int i2) { return (unsigned long long)(type | (i1 << 2) | (i2 << 6)); };
    std::vector<unsigned long long> light;
    for (int a = 0; a < 12; a++) for (int b = 0; b < 12; b++) light.push_back(desc(2, a, b));
    std::vector<unsigned long long> cc;
    for (int a = 0; a < 12; a++) for (int b = 0; b <= a; b++) cc.push_back(desc(3, a, b));
    int t = 0;
    for (int a = 0; a < 12; a++) for (int b = 0; b <= a; b++, t++) {      // threads 0..77: one heavy entry + one (c, c) entry
        unsigned long long d = desc(1, a, b);
        if (!cc.empty() && t < 72) { d |= cc.back() << 16; cc.pop_back(); }
        tab[ZT_C1OFF + t] = d;
    }
    {   // affine Hessian entries scattered by Srbd::apply_rec: dst (offset in Qxx) | index into the record's AF_* values << 12
        const int NX = Srbd::NX, XC = Srbd::XC, XCD = Srbd::XCD, XRD = Srbd::XRD;
        std::vector<unsigned long long> aff;
        auto put = [&](int row, int col, int idx) { aff.push_back((unsigned long long)(row * NX + col) | ((unsigned long long)idx << 12)); };
        for (int e2 = 0; e2 < 12; e2++) put(XCD + e2, XCD + e2, Srbd::AF_CD + e2);            // relative_vel + cdotxy_tracking diagonals (prb.py:166-181)
        for (int leg = 0; leg < 2; leg++) for (int ax = 0; ax < 2; ax++) { const int ia = XCD + 6 * leg + ax, ib = ia + 3; put(ia, ib, Srbd::AF_NCW); put(ib, ia, Srbd::AF_NCW); }
        for (int j = 0; j < 2; j++) for (int ax = 0; ax < 2; ax++) {                            // rel_pos (prb.py:192-199)
            const int ia = XC + 3 * j + ax, ib = ia + 6;
            put(ia, ia, Srbd::AF_RELP); put(ib, ib, Srbd::AF_RELP); put(ia, ib, Srbd::AF_RELN); put(ib, ia, Srbd::AF_RELN);
        }
        for (int i = 0; i < 4; i++) put(XC + 3 * i + 2, XC + 3 * i + 2, Srbd::AF_CW);           // cz_tracking (prb.py:180)
        for (int i = 0; i < 3; i++) put(XRD + i, XRD + i, Srbd::AF_RDOT);                       // rdot_tracking (prb.py:190)
        put(2, 2, Srbd::AF_RZ);                                                                 // rz_tracking (prb.py:184)
        if ((int)aff.size() != ZT_NAFF) return false;
        for (size_t i = 0; i < aff.size(); i++) tab[ZT_AOFF + i] = aff[i];
    }
    for (auto v : cc) light.push_back(v);
    for (size_t i = 0; i < light.size(); i++) {                            // threads 78..127: three entries each
        const size_t th = 78 + i % 50, sl = 1 + i / 50;
        if (sl > 3) return false;
        tab[ZT_C1OFF + th] |= light[i] << (16 * sl);
    }
    return true;
}

// kernel variants: 0 = SRBD structured (default), 1 = SRBD dense (A/B check, SddpConfig.dense_backward = 1), 2 = LIP dense,
// 3 / 4 = 0 / 1 with the inequality barriers compiled in (any of their weights > 0)
#ifndef SDDP_MINB
#define SDDP_MINB 4
#endif
constexpr int MINB_FAST = SDDP_MINB, MINB_DENSE = 1;
static bool has_ineq(const SddpConfig& c) {
    return c.friction_cone_weight != 0.0 || c.force_bound_weight != 0.0 || c.unilateral_weight != 0.0 || c.cdot_bound_weight != 0.0;
}
static int variant_of(const SddpConfig& c) { return c.model == SDDP_MODEL_LIP ? 2 : (c.dense_backward == 1 ? 1 : 0) + (has_ineq(c) ? 3 : 0); }
static size_t smem_of_variant(int v) {
    return v == 0 ? sizeof(SmemSrbd) : (v == 1 ? sizeof(Smem<Srbd>) : (v == 2 ? sizeof(Smem<Lip>) : (v == 3 ? sizeof(SmemSrbdI) : sizeof(Smem<SrbdI>))));
}

template <class M, class SM, int MINB>
static cudaError_t set_smem_attr(int* occ) {
    cudaError_t e;
    int bytes = (int)sizeof(SM);
    if ((e = cudaFuncSetAttribute(solve_kernel<M, SM, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    if ((e = cudaFuncSetAttribute(backward_kernel<M, SM, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    if ((e = cudaFuncSetAttribute(forward_kernel<M, SM, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    if ((e = cudaFuncSetAttribute(defects_kernel<M, SM, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    if ((e = cudaFuncSetAttribute(eval_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<M>)))) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, solve_kernel<M, SM, MINB>, NT, bytes);
}

static void model_dims(int model, int& nx, int& nu, int& np, int& pack) {
    if (model == SDDP_MODEL_SRBD) { nx = Srbd::NX; nu = Srbd::NU; np = Srbd::NP; pack = Srbd::PACK; }
    else { nx = Lip::NX; nu = Lip::NU; np = Lip::NP; pack = Lip::PACK; }
}

struct WsLayout { size_t d, pack, xn, un, K, k, total; };
static WsLayout ws_layout(const SddpConfig& c, int slots) {
    int nx, nu, np, pack;
    model_dims(c.model, nx, nu, np, pack);
    WsLayout w;
    size_t N = (size_t)c.N;
    w.d = (size_t)slots * N * nx;
    w.pack = (size_t)slots * N * pack;
    w.xn = (size_t)slots * NSLOT * (N + 1) * nx;
    w.un = (size_t)slots * NSLOT * N * nu;
    w.K = (size_t)slots * N * nu * nx;
    w.k = (size_t)slots * N * nu;
    w.total = (w.d + w.pack + w.xn + w.un + w.K + w.k) * sizeof(real) + 256;
    return w;
}
static const int kMaxSlotsPerSM = 6;
static const size_t RB = sizeof(real);      // bytes per array element (include/sddp.h sddp_real)

extern "C" {

int sddp_abi_version(void) { return SDDP_ABI_VERSION; }
int sddp_real_bytes(void) { return (int)sizeof(real); }
size_t sddp_config_size(void) { return sizeof(SddpConfig); }

int sddp_dims(int model, int* nx, int* nu, int* np) {
    if (model != SDDP_MODEL_SRBD && model != SDDP_MODEL_LIP) return SDDP_EINVAL;
    int a, b, c, d;
    model_dims(model, a, b, c, d);
    if (nx) *nx = a;
    if (nu) *nu = b;
    if (np) *np = c;
    return 0;
}

size_t sddp_workspace_bytes(const SddpConfig* cfg) {
    if (!cfg || check_config(cfg, nullptr)) return 0;
    int dev = 0, sms = 148;      // a B200 has 148 SMs; the current device is asked when there is one
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    } else (void)cudaGetLastError();
    return ws_layout(*cfg, sms * kMaxSlotsPerSM).total;   // upper bound: kMaxSlotsPerSM resident CTAs per SM
}

const char* sddp_last_error(const SddpHandle* h) { return h ? h->err : g_create_err; }

int sddp_create(const SddpConfig* cfg, SddpHandle** out) {
    SddpHandle* h = nullptr;
    if (!out) return fail(h, SDDP_EINVAL, "%s%s", "out is NULL", "");
    *out = nullptr;
    int rc = check_config(cfg, nullptr);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SDDP_ECUDA, "no CUDA device (%s); this library has no CPU path%s", cudaGetErrorString(e), "");
    SddpHandle* hh = new (std::nothrow) SddpHandle();
    if (!hh) return fail(nullptr, SDDP_ENOMEM, "%s%s", "host allocation failed", "");
    h = hh;
    h->cfg = *cfg;
    make_devcfg(h->cfg, h->dc);
    cudaError_t ce;
#define CUC(call) do { ce = (call); if (ce != cudaSuccess) { fail(nullptr, SDDP_ECUDA, "%s: %s", #call, cudaGetErrorString(ce)); sddp_destroy(h); return SDDP_ECUDA; } } while (0)
    CUC(cudaGetDevice(&h->device));
    CUC(cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, h->device));
    int occ = 0;
    h->variant = variant_of(h->cfg);
    h->smem_bytes = smem_of_variant(h->variant);
    h->eval_smem_bytes = cfg->model == SDDP_MODEL_SRBD ? sizeof(Smem<Srbd>) : sizeof(Smem<Lip>);
    // the latency variants of the structured solve kernel (small batches)
    if (h->variant == 0) CUC((cudaFuncSetAttribute(solve_kernel<Srbd, SmemSrbdL, MINB_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSrbdL))));
    if (h->variant == 3) CUC((cudaFuncSetAttribute(solve_kernel<SrbdI, SmemSrbdIL, MINB_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSrbdIL))));
    if (h->variant == 0) CUC((set_smem_attr<Srbd, SmemSrbd, MINB_FAST>(&occ)));
    else if (h->variant == 1) CUC((set_smem_attr<Srbd, Smem<Srbd>, MINB_DENSE>(&occ)));
    else if (h->variant == 2) CUC((set_smem_attr<Lip, Smem<Lip>, MINB_DENSE>(&occ)));
    else if (h->variant == 3) CUC((set_smem_attr<SrbdI, SmemSrbdI, MINB_FAST>(&occ)));
    else CUC((set_smem_attr<SrbdI, Smem<SrbdI>, MINB_DENSE>(&occ)));
    if (occ < 1) { fail(nullptr, SDDP_ECUDA, "%s%s", "solve kernel does not fit on this device", ""); sddp_destroy(h); return SDDP_ECUDA; }
    if (occ > kMaxSlotsPerSM) occ = kMaxSlotsPerSM;
    h->slots = h->sms * occ;
    WsLayout w = ws_layout(h->cfg, h->slots);
    h->ws_bytes = w.total;
    real* base = nullptr;
    ce = cudaMalloc((void**)&base, w.total);
    if (ce != cudaSuccess) { fail(nullptr, SDDP_ENOMEM, "cudaMalloc(workspace): %s%s", cudaGetErrorString(ce), ""); sddp_destroy(h); return SDDP_ENOMEM; }
    h->ws_d = base;
    h->ws_pack = h->ws_d + w.d;
    h->ws_xn = h->ws_pack + w.pack;
    h->ws_un = h->ws_xn + w.xn;
    h->ws_K = h->ws_un + w.un;
    h->ws_k = h->ws_K + w.K;
    h->counter = (int*)(h->ws_k + w.k);
    CUC(cudaMemset(base, 0, w.total));
    if (cfg->model == SDDP_MODEL_SRBD) {
        std::vector<unsigned long long> tab;
        if (!build_ztab(tab)) { sddp_destroy(h); return fail(nullptr, SDDP_EINVAL, "%s%s", "internal: descriptor tables do not fit", ""); }
        CUC(cudaMalloc((void**)&h->ztab, tab.size() * sizeof(unsigned long long)));
        CUC(cudaMemcpy(h->ztab, tab.data(), tab.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
    h->dc.ztab = h->ztab;
#undef CUC
    *out = h;
    return 0;
}

int sddp_destroy(SddpHandle* h) {
    if (!h) return 0;
    if (h->ws_d) cudaFree(h->ws_d);
    if (h->slab) cudaFree(h->slab);
    if (h->ztab) cudaFree(h->ztab);
    if (h->gait) cudaFree(h->gait);
    if (h->stage) cudaFree(h->stage);
    if (h->hstage) cudaFreeHost(h->hstage);
    if (h->st_in) cudaStreamDestroy(h->st_in);
    if (h->st_cmp) cudaStreamDestroy(h->st_cmp);
    if (h->st_out) cudaStreamDestroy(h->st_out);
    if (h->ev_last) cudaEventDestroy(h->ev_last);
    delete h;
    return 0;
}

int sddp_set_config(SddpHandle* h, const SddpConfig* cfg) {
    if (!h) return SDDP_EINVAL;
    int rc = check_config(cfg, h);
    if (rc) return rc;
    if (cfg->model != h->cfg.model || cfg->N != h->cfg.N || variant_of(*cfg) != h->variant)
        return fail(h, SDDP_EINVAL, "%s%s", "model, N and the kernel variant are fixed at create", "");
    h->cfg = *cfg;
    make_devcfg(h->cfg, h->dc);
    h->dc.ztab = h->ztab;
    return 0;
}

int sddp_launch_count(const SddpHandle* h, long long* out) {
    if (!h || !out) return SDDP_EINVAL;
    *out = h->launches;
    return 0;
}

#define DISPATCH(h, KERNEL, grid, stream, ...)                                                                              \
    do {                                                                                                                    \
        if ((h)->variant == 0) KERNEL<Srbd, SmemSrbd, MINB_FAST><<<grid, NT, (h)->smem_bytes, stream>>>(__VA_ARGS__);        \
        else if ((h)->variant == 1) KERNEL<Srbd, Smem<Srbd>, MINB_DENSE><<<grid, NT, (h)->smem_bytes, stream>>>(__VA_ARGS__); \
        else if ((h)->variant == 2) KERNEL<Lip, Smem<Lip>, MINB_DENSE><<<grid, NT, (h)->smem_bytes, stream>>>(__VA_ARGS__);   \
        else if ((h)->variant == 3) KERNEL<SrbdI, SmemSrbdI, MINB_FAST><<<grid, NT, (h)->smem_bytes, stream>>>(__VA_ARGS__);  \
        else KERNEL<SrbdI, Smem<SrbdI>, MINB_DENSE><<<grid, NT, (h)->smem_bytes, stream>>>(__VA_ARGS__);                      \
        (h)->launches++;                                                                                                    \
        CU(cudaGetLastError());                                                                                             \
        CU(mark_last(h, stream));                                                                                           \
    } while (0)

int sddp_eval_derivatives(SddpHandle* h, int M, const int32_t* kind, const real* x, const real* u, const real* p,
                          real* f, real* fx, real* fu, real* l, real* lx, real* lu, real* lxx, real* lux,
                          real* luu, void* stream) {
    if (!h) return SDDP_EINVAL;
    if (int rcd = check_device(h)) return rcd;
    if (M < 0 || (M > 0 && (!kind || !x || !u || !p))) return fail(h, SDDP_EINVAL, "%s%s", "eval_derivatives: bad arguments", "");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = M < h->sms * 8 ? M : h->sms * 8;
    if (h->variant >= 3) eval_kernel<SrbdI><<<grid, NT, h->eval_smem_bytes, st>>>(h->dc, M, kind, x, u, p, f, fx, fu, l, lx, lu, lxx, lux, luu);
    else if (h->cfg.model == SDDP_MODEL_SRBD) eval_kernel<Srbd><<<grid, NT, h->eval_smem_bytes, st>>>(h->dc, M, kind, x, u, p, f, fx, fu, l, lx, lu, lxx, lux, luu);
    else eval_kernel<Lip><<<grid, NT, h->eval_smem_bytes, st>>>(h->dc, M, kind, x, u, p, f, fx, fu, l, lx, lu, lxx, lux, luu);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

// Launch of the solve kernel for the arrays named in `a` (device pointers; a.h_* optional mapped host pointers).
static int launch_solve(SddpHandle* h, int B, SolveArgs& a, cudaStream_t st) {
    CU(cudaMemsetAsync(h->counter, 0, sizeof(int), st));
    CU(cudaMemsetAsync(a.status, 0xff, sizeof(int32_t) * (size_t)B, st));      // -1 = not solved (only a bad dispatch order leaves it)
    a.B = B;
    a.ws_d = h->ws_d; a.ws_pack = h->ws_pack; a.ws_xn = h->ws_xn; a.ws_un = h->ws_un; a.ws_K = h->ws_K; a.ws_k = h->ws_k;
    a.counter = h->counter;
    a.order = (h->order_dev && h->order_n == B) ? h->order_dev : nullptr;
    a.sms = h->sms;
    a.n_peers = h->n_peers; a.first = h->first_record;
    {
        int nx, nu, np, pack;
        model_dims(h->cfg.model, nx, nu, np, pack);
        a.rec = (h->cfg.N + 1) * nx + h->cfg.N * nu + SDDP_RECORD_TAIL;
    }
    for (int p = 0; p < SDDP_MAX_PEERS; p++) a.peers[p] = p < h->n_peers ? h->peers[p] : nullptr;
    int grid = B < h->slots ? B : h->slots;
    // Batches that do not fill the GPU are latency bound: they take the latency variant of the structured kernel (unrolled
    // factorisation, interleaved tensor-core chains); full batches are bound by instruction fetch and take the rolled one.
    if (B < h->slots && (h->variant == 0 || h->variant == 3)) {
        if (h->variant == 0) solve_kernel<Srbd, SmemSrbdL, MINB_FAST><<<grid, NT, h->smem_bytes, st>>>(h->dc, a);
        else solve_kernel<SrbdI, SmemSrbdIL, MINB_FAST><<<grid, NT, h->smem_bytes, st>>>(h->dc, a);
        h->launches++;
        CU(cudaGetLastError());
        CU(mark_last(h, st));
        return 0;
    }
    DISPATCH(h, solve_kernel, grid, st, h->dc, a);
    return 0;
}

int sddp_solve_batch(SddpHandle* h, int B, const real* x0, const real* params, real* X, real* U, real* K,
                     real* kff, real* hist, int32_t* iters, int32_t* status, real* cost, void* stream) {
    if (!h) return SDDP_EINVAL;
    if (int rcd = check_device(h)) return rcd;
    if (B < 0 || (B > 0 && (!x0 || !params || !X || !U || !iters || !status || !cost)))
        return fail(h, SDDP_EINVAL, "%s%s", "solve_batch: x0, params, X, U, iters, status, cost are required", "");
    if (B == 0) return 0;
    SolveArgs a;
    memset(&a, 0, sizeof(a));
    a.x0 = x0; a.params = params; a.X = X; a.U = U; a.K = K; a.kff = kff; a.hist = hist;
    a.iters = iters; a.status = status; a.cost = cost;
    return launch_solve(h, B, a, (cudaStream_t)stream);
}

int sddp_backward_pass(SddpHandle* h, int B, const real* X, const real* U, const real* params, const real* defect,
                       double mu, real* K, real* kff, real* dV, int32_t* rc, void* stream) {
    if (!h) return SDDP_EINVAL;
    if (int rcd = check_device(h)) return rcd;
    if (B < 0 || (B > 0 && (!X || !U || !params || !defect || !K || !kff || !dV || !rc)))
        return fail(h, SDDP_EINVAL, "%s%s", "backward_pass: all arrays are required", "");
    if (B == 0) return 0;
    int grid = B < h->slots ? B : h->slots;
    DISPATCH(h, backward_kernel, grid, (cudaStream_t)stream, h->dc, B, X, U, params, defect, mu, K, kff, dV, rc, h->ws_pack);
    return 0;
}

int sddp_forward_pass(SddpHandle* h, int B, int n_alpha, const real* alpha, const real* rho, const real* x0,
                      const real* X, const real* U, const real* params, const real* defect, const real* K,
                      const real* kff, real* Jn, real* Xn, real* Un, void* stream) {
    if (!h) return SDDP_EINVAL;
    if (int rcd = check_device(h)) return rcd;
    if (B < 0 || n_alpha < 1 || (B > 0 && (!alpha || !rho || !x0 || !X || !U || !params || !defect || !K || !kff || !Jn)))
        return fail(h, SDDP_EINVAL, "%s%s", "forward_pass: bad arguments", "");
    if (B == 0) return 0;
    int grid = B < h->slots ? B : h->slots;
    DISPATCH(h, forward_kernel, grid, (cudaStream_t)stream, h->dc, B, n_alpha, alpha, rho, x0, X, U, params, defect, K, kff, Jn, Xn, Un,
             h->ws_xn, h->ws_un);
    return 0;
}

int sddp_defects(SddpHandle* h, int B, const real* X, const real* U, const real* params, real* defect, real* cost,
                 void* stream) {
    if (!h) return SDDP_EINVAL;
    if (int rcd = check_device(h)) return rcd;
    if (B < 0 || (B > 0 && (!X || !U || !params))) return fail(h, SDDP_EINVAL, "%s%s", "defects: bad arguments", "");
    if (B == 0) return 0;
    int grid = B < h->slots ? B : h->slots;
    DISPATCH(h, defects_kernel, grid, (cudaStream_t)stream, h->dc, B, X, U, params, defect, cost);
    return 0;
}

int sddp_solve_batch_host(SddpHandle* h, int B, const real* x0, const real* params, const real* X0, const real* U0,
                          real* X, real* U, real* K, real* kff, real* hist, int32_t* iters, int32_t* status, real* cost) {
    if (!h) return SDDP_EINVAL;
    if (int rcd = check_device(h)) return rcd;
    if (B < 0 || (B > 0 && (!x0 || !params || !X0 || !U0 || !X || !U || !iters || !status || !cost)))
        return fail(h, SDDP_EINVAL, "%s%s", "solve_batch_host: x0, params, X0, U0, X, U, iters, status, cost are required", "");
    if (B == 0) return 0;
    int nx, nu, np, pack;
    model_dims(h->cfg.model, nx, nu, np, pack);
    const size_t N = (size_t)h->cfg.N, Bz = (size_t)B;
    const size_t s_x0 = nx, s_p = (N + 1) * np, s_X = (N + 1) * nx, s_U = N * nu;              // doubles per problem
    const size_t s_K = K ? N * nu * nx : 0, s_k = kff ? s_U : 0, s_h = hist ? (size_t)h->cfg.max_iters * SDDP_HIST : 0;
    const size_t n_d = Bz * (s_x0 + s_p + s_X + s_U + s_K + s_k + s_h + 1);
    const size_t bytes = n_d * sizeof(real) + 3 * Bz * sizeof(int32_t);
    if (bytes > h->stage_bytes) {
        if (h->stage) cudaFree(h->stage);
        h->stage = nullptr; h->stage_bytes = 0;
        cudaError_t e = cudaMalloc(&h->stage, bytes);
        if (e != cudaSuccess) return fail(h, SDDP_ENOMEM, "cudaMalloc(staging): %s%s", cudaGetErrorString(e), "");
        h->stage_bytes = bytes;
    }
    // The handle's workspace serves one solve at a time; this entry point runs on its own streams, so work the
    // caller queued earlier on other streams with the same handle must finish first.
    // (An event recorded after the last launch of every device entry point; nothing else on the device is waited for.)
    if (h->ev_last_valid) CU(cudaEventSynchronize(h->ev_last));
    if (!h->st_in) {
        CU(cudaStreamCreateWithFlags(&h->st_in, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->st_cmp, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->st_out, cudaStreamNonBlocking));
        h->host_chunk = 0;
        const char* env = getenv("SDDP_HOST_CHUNK");
        if (env && atoi(env) > 0) h->host_chunk = atoi(env);
    }
    real* d_x0 = (real*)h->stage;
    real* d_p = d_x0 + Bz * s_x0;
    real* d_X = d_p + Bz * s_p;
    real* d_U = d_X + Bz * s_X;
    real* d_K = d_U + Bz * s_U;
    real* d_k = d_K + Bz * s_K;
    real* d_h = d_k + Bz * s_k;
    real* d_c = d_h + Bz * s_h;
    int32_t* d_it = (int32_t*)(d_c + Bz);
    int32_t* d_st = d_it + Bz;
    int32_t* d_ord = d_st + Bz;
    // Host-direct path: when every buffer of the caller is mapped pinned host memory (cudaHostAlloc / cudaHostRegister:
    // torch pin_memory() is) and K is not asked for, ONE launch solves the whole batch: the CTA that takes a problem pulls
    // its inputs over PCIe and stores its results straight into the caller's arrays (sddp_solver.cuh, solve_one), so every
    // transfer rides under the solves of the other CTAs and there are no chunk boundaries with their tails of slow
    // problems.  (K is 7.1 KB per node: 355 KB per problem would make the stores PCIe bound; it takes the staged path.)
    {
        static int direct_env = -1;
        if (direct_env < 0) { const char* e = getenv("SDDP_HOST_DIRECT"); direct_env = (e && e[0] == '0') ? 0 : 1; }
        auto mapped = [](const void* p, void** dp) -> bool {
            *dp = nullptr;
            if (!p) return true;
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
            if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
            *dp = at.devicePointer;
            return true;
        };
        void *m_x0, *m_p, *m_X0, *m_U0, *m_X, *m_U, *m_k, *m_h, *m_c, *m_it, *m_st;
        if (direct_env && !K && mapped(x0, &m_x0) && mapped(params, &m_p) && mapped(X0, &m_X0) && mapped(U0, &m_U0) && mapped(X, &m_X) &&
            mapped(U, &m_U) && mapped(kff, &m_k) && mapped(hist, &m_h) && mapped(cost, &m_c) && mapped(iters, &m_it) && mapped(status, &m_st)) {
            if (h->order_host.size() == Bz) CU(cudaMemcpyAsync(d_ord, h->order_host.data(), Bz * sizeof(int32_t), cudaMemcpyHostToDevice, h->st_cmp));
            const int32_t* saved_order = h->order_dev;
            const int saved_n = h->order_n, saved_peers = h->n_peers;
            h->order_dev = h->order_host.size() == Bz ? d_ord : nullptr; h->order_n = B; h->n_peers = 0;
            SolveArgs a;
            memset(&a, 0, sizeof(a));
            a.x0 = d_x0; a.params = d_p; a.X = d_X; a.U = d_U; a.K = nullptr; a.kff = kff ? d_k : nullptr; a.hist = hist ? d_h : nullptr;
            a.iters = d_it; a.status = d_st; a.cost = d_c;
            a.h_x0 = (const real*)m_x0; a.h_params = (const real*)m_p; a.h_X0 = (const real*)m_X0; a.h_U0 = (const real*)m_U0;
            a.h_X = (real*)m_X; a.h_U = (real*)m_U; a.h_kff = (real*)m_k; a.h_hist = (real*)m_h; a.h_cost = (real*)m_c;
            a.h_iters = (int*)m_it; a.h_status = (int*)m_st;
            int rcl = launch_solve(h, B, a, h->st_cmp);
            h->order_dev = saved_order; h->order_n = saved_n; h->n_peers = saved_peers;
            if (rcl) return rcl;
            CU(cudaStreamSynchronize(h->st_cmp));
            return 0;
        }
    }
    // Chunk size: the copies of one chunk overlap the solve of another, so a batch needs several chunks whatever its
    // size (a fixed 16K left the 8192-problem shards of an 8-GPU run with one chunk: copy, solve, copy in series).
    // Default: a quarter of the batch, at least two problems per CTA slot (so a chunk still fills the GPU and its tail
    // of slow problems stays small) and at most 16384; SDDP_HOST_CHUNK overrides.
    int chunk = h->host_chunk;
    if (chunk <= 0) {
        chunk = (B + 3) / 4;
        if (chunk < 2 * h->slots) chunk = 2 * h->slots;
        if (chunk > 16384) chunk = 16384;
    }
    const int nchunk = (B + chunk - 1) / chunk;
    // Small batches (the reference's own use: one problem per call): one pinned staging buffer, one copy in, one copy
    // out, one stream, no events -- the latency of a single solve is mostly launch and copy overhead otherwise.
    const size_t in_bytes = Bz * (s_x0 + s_p + s_X + s_U) * sizeof(real);
    const size_t out_bytes = bytes - Bz * (s_x0 + s_p) * sizeof(real) - Bz * sizeof(int32_t);      // X .. status
    if (nchunk == 1 && in_bytes + out_bytes <= ((size_t)4 << 20) && h->order_host.empty()) {
        if (in_bytes + out_bytes > h->hstage_bytes) {
            if (h->hstage) cudaFreeHost(h->hstage);
            h->hstage = nullptr; h->hstage_bytes = 0;
            cudaError_t e = cudaHostAlloc(&h->hstage, in_bytes + out_bytes, cudaHostAllocDefault);
            if (e != cudaSuccess) return fail(h, SDDP_ENOMEM, "cudaHostAlloc(staging): %s%s", cudaGetErrorString(e), "");
            h->hstage_bytes = in_bytes + out_bytes;
        }
        real* hi = (real*)h->hstage;
        memcpy(hi, x0, Bz * s_x0 * RB);
        memcpy(hi + Bz * s_x0, params, Bz * s_p * RB);
        memcpy(hi + Bz * (s_x0 + s_p), X0, Bz * s_X * RB);
        memcpy(hi + Bz * (s_x0 + s_p + s_X), U0, Bz * s_U * RB);
        CU(cudaMemcpyAsync(d_x0, hi, in_bytes, cudaMemcpyHostToDevice, h->st_cmp));
        const int keep_peers = h->n_peers;
        h->n_peers = 0;
        int rc1 = sddp_solve_batch(h, B, d_x0, d_p, d_X, d_U, K ? d_K : nullptr, kff ? d_k : nullptr, hist ? d_h : nullptr, d_it, d_st, d_c, h->st_cmp);
        h->n_peers = keep_peers;
        if (rc1) return rc1;
        char* ho = (char*)h->hstage + in_bytes;
        CU(cudaMemcpyAsync(ho, d_X, out_bytes, cudaMemcpyDeviceToHost, h->st_cmp));
        CU(cudaStreamSynchronize(h->st_cmp));
        const real* o = (const real*)ho;
        memcpy(X, o, Bz * s_X * RB); o += Bz * s_X;
        memcpy(U, o, Bz * s_U * RB); o += Bz * s_U;
        if (K) memcpy(K, o, Bz * s_K * RB);
        o += Bz * s_K;
        if (kff) memcpy(kff, o, Bz * s_k * RB);
        o += Bz * s_k;
        if (hist) memcpy(hist, o, Bz * s_h * RB);
        o += Bz * s_h;
        memcpy(cost, o, Bz * RB); o += Bz;
        memcpy(iters, o, Bz * 4);
        memcpy(status, (const int32_t*)o + Bz, Bz * 4);
        return 0;
    }
    // dispatch order: the global permutation restricted to each chunk, as chunk-local indices (same relative order)
    const bool ordered = h->order_host.size() == Bz;
    std::vector<int32_t> ordl;
    if (ordered) {
        std::vector<size_t> fill(nchunk);
        for (int c = 0; c < nchunk; c++) fill[c] = (size_t)c * chunk;
        ordl.resize(Bz);
        for (size_t i = 0; i < Bz; i++) { const int g = h->order_host[i], c = g / chunk; ordl[fill[c]++] = g - c * chunk; }
    }
    const int32_t* saved_order = h->order_dev;
    const int saved_n = h->order_n, saved_peers = h->n_peers;
    h->n_peers = 0;      // result records are a feature of the device entry point (chunk-local problem indices here)
    std::vector<cudaEvent_t> ev_in(nchunk), ev_cmp(nchunk);
    int rc = 0;
    for (int c = 0; c < nchunk; c++) {
        cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_cmp[c], cudaEventDisableTiming);
    }
    for (int c = 0; c < nchunk && rc == 0; c++) {
        const size_t o = (size_t)c * chunk, n = (size_t)((B - (int)o) < chunk ? (B - (int)o) : chunk);
        cudaError_t e = cudaSuccess;
        auto cp = [&](void* dst, const void* src, size_t nbytes, cudaMemcpyKind kind, cudaStream_t st) {
            if (e == cudaSuccess) e = cudaMemcpyAsync(dst, src, nbytes, kind, st);
        };
        cp(d_x0 + o * s_x0, x0 + o * s_x0, n * s_x0 * RB, cudaMemcpyHostToDevice, h->st_in);
        cp(d_p + o * s_p, params + o * s_p, n * s_p * RB, cudaMemcpyHostToDevice, h->st_in);
        cp(d_X + o * s_X, X0 + o * s_X, n * s_X * RB, cudaMemcpyHostToDevice, h->st_in);
        cp(d_U + o * s_U, U0 + o * s_U, n * s_U * RB, cudaMemcpyHostToDevice, h->st_in);
        if (ordered) cp(d_ord + o, ordl.data() + o, n * 4, cudaMemcpyHostToDevice, h->st_in);
        if (e == cudaSuccess) e = cudaEventRecord(ev_in[c], h->st_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->st_cmp, ev_in[c], 0);
        if (e != cudaSuccess) { rc = fail(h, SDDP_ECUDA, "solve_batch_host (copy in): %s%s", cudaGetErrorString(e), ""); break; }
        h->order_dev = ordered ? d_ord + o : nullptr; h->order_n = (int)n;
        rc = sddp_solve_batch(h, (int)n, d_x0 + o * s_x0, d_p + o * s_p, d_X + o * s_X, d_U + o * s_U, K ? d_K + o * s_K : nullptr,
                              kff ? d_k + o * s_k : nullptr, hist ? d_h + o * s_h : nullptr, d_it + o, d_st + o, d_c + o, h->st_cmp);
        if (rc) break;
        e = cudaEventRecord(ev_cmp[c], h->st_cmp);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(h->st_out, ev_cmp[c], 0);
        cp(X + o * s_X, d_X + o * s_X, n * s_X * RB, cudaMemcpyDeviceToHost, h->st_out);
        cp(U + o * s_U, d_U + o * s_U, n * s_U * RB, cudaMemcpyDeviceToHost, h->st_out);
        if (K) cp(K + o * s_K, d_K + o * s_K, n * s_K * RB, cudaMemcpyDeviceToHost, h->st_out);
        if (kff) cp(kff + o * s_k, d_k + o * s_k, n * s_k * RB, cudaMemcpyDeviceToHost, h->st_out);
        if (hist) cp(hist + o * s_h, d_h + o * s_h, n * s_h * RB, cudaMemcpyDeviceToHost, h->st_out);
        cp(cost + o, d_c + o, n * RB, cudaMemcpyDeviceToHost, h->st_out);
        cp(iters + o, d_it + o, n * 4, cudaMemcpyDeviceToHost, h->st_out);
        cp(status + o, d_st + o, n * 4, cudaMemcpyDeviceToHost, h->st_out);
        if (e != cudaSuccess) rc = fail(h, SDDP_ECUDA, "solve_batch_host (copy out): %s%s", cudaGetErrorString(e), "");
    }
    h->order_dev = saved_order; h->order_n = saved_n; h->n_peers = saved_peers;
    cudaError_t e1 = cudaStreamSynchronize(h->st_in), e2 = cudaStreamSynchronize(h->st_cmp), e3 = cudaStreamSynchronize(h->st_out);
    for (int c = 0; c < nchunk; c++) { cudaEventDestroy(ev_in[c]); cudaEventDestroy(ev_cmp[c]); }
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail(h, SDDP_ECUDA, "solve_batch_host (sync): %s%s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)), "");
    return 0;
}

// ---- result records and the multi-GPU gather
static_assert(SDDP_MAX_PEERS == SDDP_MAX_RESULT_PEERS, "include/sddp.h");
long long sddp_record_doubles(const SddpHandle* h) {
    if (!h) return 0;
    int nx, nu, np, pack;
    model_dims(h->cfg.model, nx, nu, np, pack);
    return (long long)(h->cfg.N + 1) * nx + (long long)h->cfg.N * nu + SDDP_RECORD_TAIL;
}

int sddp_slab_alloc(SddpHandle* h, long long n_records, real** out) {
    if (!h || !out || n_records < 0) return h ? fail(h, SDDP_EINVAL, "%s%s", "slab_alloc: bad arguments", "") : SDDP_EINVAL;
    if (int rcd = check_device(h)) return rcd;
    for (int p = 0; p < h->n_peers; p++)
        if (h->peers[p] == h->slab) return fail(h, SDDP_EINVAL, "%s%s", "slab_alloc: the current slab is still a result peer (sddp_set_result_peers(h, 0, NULL, 0) first)", "");
    if (h->slab) { cudaFree(h->slab); h->slab = nullptr; h->slab_records = 0; }
    *out = nullptr;
    if (n_records == 0) return 0;
    // plain cudaMalloc: the allocation base is what cudaIpcGetMemHandle exports
    cudaError_t e = cudaMalloc((void**)&h->slab, (size_t)n_records * (size_t)sddp_record_doubles(h) * sizeof(real));
    if (e != cudaSuccess) { h->slab = nullptr; return fail(h, SDDP_ENOMEM, "cudaMalloc(result slab): %s%s", cudaGetErrorString(e), ""); }
    h->slab_records = n_records;
    *out = h->slab;
    return 0;
}

int sddp_ipc_export(const void* dev_ptr, unsigned char handle[SDDP_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == SDDP_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    if (!dev_ptr || !handle) return SDDP_EINVAL;
    cudaIpcMemHandle_t m;
    cudaError_t e = cudaIpcGetMemHandle(&m, const_cast<void*>(dev_ptr));
    if (e != cudaSuccess) return fail(nullptr, SDDP_ECUDA, "cudaIpcGetMemHandle: %s%s", cudaGetErrorString(e), "");
    memcpy(handle, &m, sizeof(m));
    return 0;
}

int sddp_ipc_open(const unsigned char handle[SDDP_IPC_HANDLE_BYTES], void** dev_ptr) {
    if (!handle || !dev_ptr) return SDDP_EINVAL;
    cudaIpcMemHandle_t m;
    memcpy(&m, handle, sizeof(m));
    cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, m, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { *dev_ptr = nullptr; return fail(nullptr, SDDP_ECUDA, "cudaIpcOpenMemHandle: %s%s", cudaGetErrorString(e), ""); }
    return 0;
}

int sddp_ipc_close(void* dev_ptr) {
    if (!dev_ptr) return 0;
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    return e == cudaSuccess ? 0 : fail(nullptr, SDDP_ECUDA, "cudaIpcCloseMemHandle: %s%s", cudaGetErrorString(e), "");
}

int sddp_set_result_peers(SddpHandle* h, int n_peers, real* const* slabs, long long first_record) {
    if (!h) return SDDP_EINVAL;
    if (n_peers < 0 || n_peers > SDDP_MAX_PEERS || (n_peers > 0 && !slabs) || first_record < 0)
        return fail(h, SDDP_EINVAL, "%s%s", "set_result_peers: 0..8 slabs, first_record >= 0", "");
    for (int p = 0; p < n_peers; p++)
        if (!slabs[p]) return fail(h, SDDP_EINVAL, "%s%s", "set_result_peers: NULL slab", "");
    h->n_peers = n_peers;
    h->first_record = first_record;
    for (int p = 0; p < SDDP_MAX_PEERS; p++) h->peers[p] = p < n_peers ? slabs[p] : nullptr;
    return 0;
}

int sddp_set_dispatch_order(SddpHandle* h, const int32_t* order, int n, int on_host) {
    if (!h) return SDDP_EINVAL;
    if (n < 0 || (order && n == 0)) return fail(h, SDDP_EINVAL, "%s%s", "set_dispatch_order: bad length", "");
    if (!on_host) { h->order_dev = order; h->order_n = order ? n : 0; return 0; }
    h->order_host.clear();
    if (!order) return 0;
    std::vector<char> seen((size_t)n, 0);
    for (int i = 0; i < n; i++) {
        if (order[i] < 0 || order[i] >= n || seen[order[i]]) return fail(h, SDDP_EINVAL, "%s%s", "set_dispatch_order: not a permutation of 0..n-1", "");
        seen[order[i]] = 1;
    }
    h->order_host.assign(order, order + n);
    return 0;
}

int sddp_set_gait_tables(SddpHandle* h, const double* l_cycle, const double* l_switch, const double* r_cycle, const double* r_switch) {
    if (!h) return SDDP_EINVAL;
    if (!l_cycle || !l_switch || !r_cycle || !r_switch) return fail(h, SDDP_EINVAL, "%s%s", "set_gait_tables: four tables are required", "");
    real tab[84];
    for (int i = 0; i < 21; i++) { tab[i] = (real)l_cycle[i]; tab[21 + i] = (real)l_switch[i]; tab[42 + i] = (real)r_cycle[i]; tab[63 + i] = (real)r_switch[i]; }
    if (!h->gait) CU(cudaMalloc((void**)&h->gait, sizeof(tab)));
    CU(cudaMemcpy(h->gait, tab, sizeof(tab), cudaMemcpyHostToDevice));
    return 0;
}

int sddp_mpc_advance(SddpHandle* h, int B, real* params, const int32_t* action, int32_t* step_counter, const real* rdot_ref_cmd,
                     void* stream) {
    if (!h) return SDDP_EINVAL;
    if (int rcd = check_device(h)) return rcd;
    if (B < 0 || (B > 0 && (!params || !action || !step_counter || !rdot_ref_cmd))) return fail(h, SDDP_EINVAL, "%s%s", "mpc_advance: bad arguments", "");
    if (!h->gait) return fail(h, SDDP_EINVAL, "%s%s", "mpc_advance: call sddp_set_gait_tables first", "");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int nx, nu, np, pack;
    model_dims(h->cfg.model, nx, nu, np, pack);
    const int threads = 128, grid = (B * np + threads - 1) / threads;
    if (h->cfg.model == SDDP_MODEL_SRBD) mpc_advance_kernel<Srbd><<<grid, threads, 0, st>>>(h->dc, B, params, action, step_counter, rdot_ref_cmd, h->gait);
    else mpc_advance_kernel<Lip><<<grid, threads, 0, st>>>(h->dc, B, params, action, step_counter, rdot_ref_cmd, h->gait);
    counter_inc_kernel<<<(B + threads - 1) / threads, threads, 0, st>>>(B, step_counter);
    h->launches += 2;
    CU(cudaGetLastError());
    return 0;
}

int sddp_plant_step(SddpHandle* h, int B, real* state, const real* u, long long u_stride, void* stream) {
    if (!h) return SDDP_EINVAL;
    if (int rcd = check_device(h)) return rcd;
    if (B < 0 || (B > 0 && (!state || !u))) return fail(h, SDDP_EINVAL, "%s%s", "plant_step: bad arguments", "");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 64, grid = (B + threads - 1) / threads;
    if (h->cfg.model == SDDP_MODEL_SRBD) plant_step_kernel<Srbd><<<grid, threads, 0, st>>>(h->dc, B, state, u, u_stride);
    else plant_step_kernel<Lip><<<grid, threads, 0, st>>>(h->dc, B, state, u, u_stride);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

int sddp_fp64_peak_tflops(double* out, void* stream) {
    SddpHandle* h = nullptr;
    if (!out) return SDDP_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* d = nullptr;
    CU(cudaMalloc((void**)&d, 8));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int iters = 4096, grid = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        CU(cudaEventRecord(e0, st));
        fp64_peak_kernel<<<grid, 256, 0, st>>>(d, iters, 1.0000001, 1e-9);
        CU(cudaEventRecord(e1, st));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)grid;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *out = best;
    return 0;
}

#ifdef SDDP_PROFILE
/* developer builds only: cycles accumulated per phase by thread 0 of CTA 0 (see sddp_solver.cuh PROF) */
int sddp_debug_profile(long long* out32, int reset) {
    long long z[32] = {0};
    if (cudaMemcpyFromSymbol(out32, g_prof, sizeof(z)) != cudaSuccess) return SDDP_ECUDA;
    if (reset && cudaMemcpyToSymbol(g_prof, z, sizeof(z)) != cudaSuccess) return SDDP_ECUDA;
    return 0;
}
#endif

#ifdef SDDP_STAMP
int sddp_debug_stamps(long long* out64) {
    return cudaMemcpyFromSymbol(out64, g_stamp, sizeof(long long) * 64) == cudaSuccess ? 0 : SDDP_ECUDA;
}
#endif

}  // extern "C"
