/*
 * sddp.h -- C ABI of the B200-native batched DDP solver for srbd_horizon.
 *
 * This is the drop-in boundary for the one hot path of hucebot/srbd_horizon:
 * the DDP solve behind python/ddp.py's `DDPSolver`.  In the reference that
 * boundary is the pybind module `pyddp` (not in the tree), reached from
 *
 *   ddp.py:14-35    pyddp.DdpSolverOptions()         -> fields of SddpConfig
 *   ddp.py:93-94    pyddp.DdpSolver(nx,nu,f,L,L_N,o) -> sddp_create
 *   ddp.py:101      ddp_solver.solve(params)         -> sddp_solve_batch (B = 1 for the reference's use)
 *   ddp.py:106      ddp_solver.is_converged()        -> status[b] == SDDP_CONVERGED
 *   ddp.py:114,117  set_u_warmstart / set_x_warmstart-> the X / U buffers on entry
 *   ddp.py:123      set_initial_state(x0)            -> the x0 buffer
 *
 * pyddp receives CasADi Function objects for f_k, L_k, L_N (ddp.py:179-230); here
 * the two problems of prb.py are hand-written sm_100a CUDA, selected by
 * SddpConfig.model, and the numeric constants prb.py reads from ROS/URDF are
 * plain fields of SddpConfig.
 *
 * Conventions
 *   - plain C, no C++/torch types; all array arguments are DEVICE pointers to
 *     sddp_real (double; float in the optional fp32 build; int32 for iters/status)
 *     unless the function name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all
 *     work is enqueued on it, nothing synchronises except the *_host calls;
 *   - every function returns 0 on success, a negative SDDP_E* code otherwise,
 *     never throws; sddp_last_error() gives the message for the calling handle;
 *   - a handle is bound to the device current at sddp_create (every later call must be made with that device
 *     current, SDDP_EINVAL otherwise), owns only its workspace, is not thread-safe; distinct handles are independent;
 *   - per-problem numerical outcome is reported in status[b], not in the return code.
 *
 * Layouts (row-major, problem-major):
 *   x0[B][nx]  params[B][N+1][np]  X[B][N+1][nx]  U[B][N][nu]
 *   K[B][N][nu][nx]  kff[B][N][nu]  hist[B][max_iters][SDDP_HIST]
 *   iters[B] status[B] (int32)  cost[B]
 * State / input / parameter layouts: see srbd_horizon_b200/config.py (prb.py:32-68, 264-295).
 */
#ifndef SDDP_H
#define SDDP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDDP_ABI_VERSION 5

/* Element type of every array argument.  double in the product library (libsddp.so).  The optional fp32 build
 * (libsddp_f32.so, compiled from the same sources with -DSDDP_F32; north_star "optional fp32 build ... within a stated
 * tolerance") exports the same symbols with float arrays; sddp_real_bytes() tells a caller which one it has loaded.
 * SddpConfig and scalar arguments are double in both. */
#ifdef SDDP_F32
typedef float sddp_real;
#else
typedef double sddp_real;
#endif

enum { SDDP_MODEL_SRBD = 0, SDDP_MODEL_LIP = 1 };
enum { SDDP_INERTIA_LITERAL = 0, SDDP_INERTIA_ROTATED = 1 };   /* prb.py:99 as written / README.md:2 intent */
enum { SDDP_HESSIAN_EXACT = 0, SDDP_HESSIAN_GN = 1 };
enum { SDDP_NODE_FIRST = 0, SDDP_NODE_MID = 1, SDDP_NODE_TERM = 2, SDDP_NODE_TAIL = 3 };   /* TAIL: a MID node of the LIP-style tail (lip_tail_start) */
enum { SDDP_NOT_SOLVED = -1, SDDP_CONVERGED = 0, SDDP_MAX_ITERS = 1, SDDP_LS_FAILED = 2, SDDP_REG_FAILED = 3, SDDP_NAN = 4 };
enum { SDDP_HIST = 4 };   /* per iteration: cost, alpha (0 = no step), mu, max|defect| */
enum { SDDP_EINVAL = -1, SDDP_ECUDA = -2, SDDP_ENOMEM = -3, SDDP_ECAPACITY = -4 };

typedef struct SddpConfig {
    int32_t model;             /* SDDP_MODEL_*  (prb.py:16-246 / 248-441) */
    int32_t N;                 /* shooting intervals `ns`; nodes 0..N (prb.py:21) */
    int32_t inertia_mode;      /* SDDP_INERTIA_* */
    int32_t hessian_mode;      /* SDDP_HESSIAN_* */
    int32_t multiple_shooting; /* 1: keep the x warm start, carry defects (README.md:5-6) */
    int32_t max_iters;         /* ddp.py:17-19 */
    int32_t dense_backward;    /* SRBD only. 1: generic dense Riccati kernel instead of the structured one (A/B check; same results) */
    int32_t lip_tail_start;    /* SRBD only. 0: off. k >= 1: nodes k..N-1 use the LIP-style model of the reference's model scheduler
                                * (README.md:7, isrbd_example.py:344-353): no rotational dynamics (wdot = 0, hence no wdot term in
                                * min_qddot) and the constraints lip_com_height (r_z - com_z) and lip_zero_angular_momentum (w),
                                * penalised with constraint_weight like every equality constraint (ddp.py:191-196) */
    double dt;                 /* prb.py:110 */
    double mass;               /* prb.py:92 */
    double inertia[9];         /* prb.py:94-95, row-major */
    double com[3];             /* prb.py:138-139 */
    double foot[12];           /* prb.py:127-135 */
    double force_scaling;      /* prb.py:98 */
    double gravity;
    double eta2;               /* prb.py:317 */
    double r_tracking_gain;    /* prb.py:142 */
    double rdot_tracking_gain; /* prb.py:145 */
    double w_tracking_gain;    /* prb.py:146 */
    double rel_position_gain;  /* prb.py:147 */
    double force_switch_weight;/* prb.py:148 */
    double min_qddot_gain;     /* prb.py:149 */
    double min_f_gain;         /* prb.py:150 */
    double zmp_tracking_gain;  /* prb.py:361 */
    double constraint_weight;  /* ddp.py:181 */
    double alpha_0;                      /* ddp.py:20-22 */
    double alpha_converge_threshold;     /* ddp.py:23-25 */
    double line_search_decrease_factor;  /* ddp.py:26-28 */
    double beta;                         /* ddp.py:29-31 */
    double cost_reduction_ths;           /* ddp.py:32-33 */
    double mu0;                          /* ddp.py:34-35 */
    double defect_contraction_rate;      /* README.md:6; <= 0: rho = alpha */
    double mu_min, mu_max, mu_factor;
    double defect_ths;
    /* Inequality handling: the friction cone the reference computes and drops (prb.py:173-177), as the
     * exponential barrier its adapter sketches (ddp.py:197-203): on nodes 0..N-1
     *   L += friction_cone_weight * sum_feet sum_rows exp(friction_cone_sharpness * g_row(f_i)),
     * g = (f_x - mu f_z, -f_x - mu f_z, f_y - mu f_z, -f_y - mu f_z, -f_z) <= 0 (linearised cone, world frame).
     * weight = 0 (default) reproduces the reference: no inequality terms.  SRBD only. */
    double friction_cone_weight;
    double friction_cone_mu;          /* prb.py:174 friction_cone_coefficient, default 0.8 */
    double friction_cone_sharpness;   /* ddp.py:182 exp_parameter, default 6.0 */
    /* Bounds as the barriers the adapter sketches for variable bounds (ddp.py:204-209): for a bounded component v,
     *   L += weight * ( exp(bound_sharpness * (v - ub)) + exp(bound_sharpness * (lb - v)) )   on nodes 0..N-1.
     * force_bound: box |f_ik| <= force_bound on every contact-force component (isrbd_example.py:200 max_contact_force,
     * in units of force_scaling); unilateral_weight: f_iz >= 0 alone (the "unilaterality" of isrbd_example.py:198,
     * lb = 0, no ub); cdot_bound: box |cdot_ik| <= cdot_bound on the contact-point velocities (isrbd_example.py:195).
     * Every weight = 0 (default) reproduces the reference (bounds ignored by its DDP adapter).  SRBD only. */
    double force_bound_weight;
    double force_bound;
    double unilateral_weight;
    double cdot_bound_weight;
    double cdot_bound;
    double bound_sharpness;
} SddpConfig;

typedef struct SddpHandle SddpHandle;

int sddp_abi_version(void);
int sddp_real_bytes(void);   /* sizeof(sddp_real) of the loaded library: 8, or 4 for the fp32 build */
size_t sddp_config_size(void);
/* nx, nu, np of a model */
int sddp_dims(int model, int *nx, int *nu, int *np);

/* Upper bound of the bytes of device workspace a handle allocates (independent of the batch size: scratch is per
 * resident CTA, not per problem; sized for the current device, or for a B200 when there is none).  0: cfg is refused. */
size_t sddp_workspace_bytes(const SddpConfig *cfg);

int sddp_create(const SddpConfig *cfg, SddpHandle **out);
int sddp_destroy(SddpHandle *h);
const char *sddp_last_error(const SddpHandle *h);   /* h may be NULL: last create error */
/* replace the solver options / mode switches of a live handle (model and N must not change) */
int sddp_set_config(SddpHandle *h, const SddpConfig *cfg);

/* Stage 1 alone (north_star stage one; the CasADi evaluation of f_k, L_k, L_N built at
 * ddp.py:179-230): for M independent (x,u,p) points of node kind `kind[m]`, writes
 * f[M][nx], fx[M][nx][nx], fu[M][nx][nu], l[M], lx[M][nx], lu[M][nu], lxx[M][nx][nx],
 * lux[M][nu][nx], luu[M][nu][nu].  Any output pointer may be NULL.  For SDDP_NODE_TERM
 * the dynamics and input outputs are written as zeros. */
int sddp_eval_derivatives(SddpHandle *h, int M, const int32_t *kind, const sddp_real *x, const sddp_real *u,
                          const sddp_real *p, sddp_real *f, sddp_real *fx, sddp_real *fu, sddp_real *l, sddp_real *lx,
                          sddp_real *lu, sddp_real *lxx, sddp_real *lux, sddp_real *luu, void *stream);

/* The solve (ddp.py:101).  X and U carry the warm start in and the solution out.
 * K, kff, hist may be NULL (gains then stay in the workspace). */
int sddp_solve_batch(SddpHandle *h, int B, const sddp_real *x0, const sddp_real *params, sddp_real *X, sddp_real *U,
                     sddp_real *K, sddp_real *kff, sddp_real *hist, int32_t *iters, int32_t *status, sddp_real *cost,
                     void *stream);

/* Stage entry points used by the stage parity tests (north_star stages two to four).
 * defect[B][N][nx]; dV[B][3] = {D1, D2, C0}; rc[B] = 0 or failing node + 1. */
int sddp_backward_pass(SddpHandle *h, int B, const sddp_real *X, const sddp_real *U, const sddp_real *params,
                       const sddp_real *defect, double mu, sddp_real *K, sddp_real *kff, sddp_real *dV, int32_t *rc,
                       void *stream);
/* n_alpha candidate step sizes per problem, evaluated in parallel: Jn[B][n_alpha],
 * Xn[B][n_alpha][N+1][nx], Un[B][n_alpha][N][nu] (Xn/Un may be NULL). rho[n_alpha] as alpha. */
int sddp_forward_pass(SddpHandle *h, int B, int n_alpha, const sddp_real *alpha, const sddp_real *rho,
                      const sddp_real *x0, const sddp_real *X, const sddp_real *U, const sddp_real *params,
                      const sddp_real *defect, const sddp_real *K, const sddp_real *kff, sddp_real *Jn, sddp_real *Xn,
                      sddp_real *Un, void *stream);
/* defect[B][N][nx] = f(X_k,U_k) - X_{k+1},  cost[B] = total cost (either may be NULL) */
int sddp_defects(SddpHandle *h, int B, const sddp_real *X, const sddp_real *U, const sddp_real *params,
                 sddp_real *defect, sddp_real *cost, void *stream);

/* Same solve with HOST buffers: copies in, solves, copies out, synchronises.
 * This is what a non-CUDA caller (the reference's Python loop) binds.  X0 / U0 are the warm start
 * (ddp.py:114,117), X / U receive the solution and may alias X0 / U0.  Three paths, same results bit for bit:
 *  - host-direct: every buffer is mapped pinned host memory (cudaHostAlloc / cudaHostRegister) and K is NULL: ONE launch
 *    for the whole batch; the CTA that takes a problem pulls its inputs over PCIe and stores its results straight into
 *    the caller's arrays, so the transfers ride under the solves of the other CTAs (SDDP_HOST_DIRECT=0 disables it);
 *  - small batches (< 4 MB, the reference's one-problem use): one copy in and one copy out through a pinned staging
 *    buffer of the handle, one stream;
 *  - otherwise chunks on three streams, so that the copies of one chunk overlap the solve of another (pageable
 *    buffers work but serialise). */
int sddp_solve_batch_host(SddpHandle *h, int B, const sddp_real *x0, const sddp_real *params, const sddp_real *X0,
                          const sddp_real *U0, sddp_real *X, sddp_real *U, sddp_real *K, sddp_real *kff, sddp_real *hist,
                          int32_t *iters, int32_t *status, sddp_real *cost);

/* Dispatch order of the next solves (scheduling only: results do not depend on it).  The persistent CTAs take
 * problems order[0], order[1], ... instead of 0, 1, ...; `order` must be a permutation of 0..n-1 and applies to
 * solves of exactly n problems.  Problems that behave alike (same contact schedule, similar iteration counts)
 * should be neighbours: co-resident CTAs then run the same phases at the same time and share their instructions
 * in the SM's instruction cache (about 10 % on BASELINE configs[4]).  on_host = 0: `order` is a DEVICE array the
 * caller keeps alive, used by sddp_solve_batch; it is NOT checked: entries outside 0..n-1 are skipped and a problem
 * that no entry names stays unsolved with status[b] = -1 (duplicates solve a problem twice).  on_host = 1: HOST array,
 * copied, used by sddp_solve_batch_host (checked to be a permutation).  order = NULL clears.  The reference has no
 * counterpart (one problem per process). */
int sddp_set_dispatch_order(SddpHandle *h, const int32_t *order, int n, int on_host);

/* ---- result records and the multi-GPU gather (SURVEY.md 2.1 K5 "result pack", 8e) ----
 * A result record is one contiguous fp64 block per problem:  X[(N+1) nx] | U[N nu] | cost | iters | status
 * (sddp_record_doubles(h) doubles; iters and status as doubles).  A slab is an array of records, one per problem of the
 * WHOLE batch (all GPUs).  After sddp_set_result_peers(h, n, slabs, first) every sddp_solve_batch of this handle stores
 * the record of its problem b, the moment that problem is finished, into slabs[p] + (first + b) * record for every
 * p < n -- from inside the solve kernel, by the CTA that solved it.  With the slabs of all GPUs of the box in the list
 * (peer memory: cudaDeviceEnablePeerAccess in one process, or the IPC helpers below across processes) this IS the
 * all-gather of the results: it rides over NVLink / NVSwitch while the rest of the batch is still being solved, and no
 * collective follows the kernel.  The caller only has to order "all kernels finished" before anyone reads a slab
 * (any barrier: a 4-byte all-reduce, MPI_Barrier, cudaStreamWaitEvent on IPC events).  With n = 1 and the own slab it
 * is the packed send buffer for ONE ncclAllGather / MPI_Allgather instead of five.  n = 0 (default) switches it off.
 * The reference has no counterpart (one problem per process, results returned by pyddp in host memory). */
enum { SDDP_RECORD_TAIL = 3, SDDP_IPC_HANDLE_BYTES = 64, SDDP_MAX_RESULT_PEERS = 8 };
long long sddp_record_doubles(const SddpHandle *h);
/* (Re)allocates the handle-owned slab of n_records records with plain cudaMalloc (exportable with sddp_ipc_export);
 * n_records = 0 frees it.  Freed by sddp_destroy. */
int sddp_slab_alloc(SddpHandle *h, long long n_records, sddp_real **out);
/* Thin wrappers of cudaIpcGetMemHandle / cudaIpcOpenMemHandle (lazy peer access) / cudaIpcCloseMemHandle so that a
 * host language without CUDA bindings can exchange the 64 handle bytes over any channel it has. */
int sddp_ipc_export(const void *dev_ptr, unsigned char handle[SDDP_IPC_HANDLE_BYTES]);
int sddp_ipc_open(const unsigned char handle[SDDP_IPC_HANDLE_BYTES], void **dev_ptr);
int sddp_ipc_close(void *dev_ptr);
int sddp_set_result_peers(SddpHandle *h, int n_peers, sddp_real *const *slabs, long long first_record);

/* ---- receding-horizon glue on the device (the caller side of the path: dsrbd_example.py:102-131,158-160, wpg.py:68-101) ----
 * gait tables of wpg.steps_phase (wpg.py:19-64): four arrays of 21 entries, l_cycle, l_switch, r_cycle, r_switch (host pointers) */
int sddp_set_gait_tables(SddpHandle *h, const double *l_cycle, const double *l_switch, const double *r_cycle,
                         const double *r_switch);
/* One MPC tick of the parameter schedule for B problems: every per-node parameter moves one node back
 * (dsrbd_example.py:102-106, wpg.py:74-77), node N receives rdot_ref_cmd[b] and, per action[b]
 * (0 "step", 1 stance, 2 "jump"), the gait entries of wpg.py:80-99 at ref_id = step_counter[b] % 20;
 * step_counter[b] is incremented.  params[B][N+1][np], action/step_counter int32[B], rdot_ref_cmd[B][3]. */
int sddp_mpc_advance(SddpHandle *h, int B, sddp_real *params, const int32_t *action, int32_t *step_counter,
                     const sddp_real *rdot_ref_cmd, void *stream);
/* Plant step of the examples: state[b] <- EULER(state[b], u[b*u_stride .. +nu], dt), SRBD quaternion renormalised
 * (dsrbd_example.py:158-160, dlip_example.py:161-162).  state[B][nx] in place; u_stride in doubles (N*nu to use U[b][0]). */
int sddp_plant_step(SddpHandle *h, int B, sddp_real *state, const sddp_real *u, long long u_stride, void *stream);

/* Measured FP64 FMA rate of the device (microbenchmark, TFLOP/s); used as the roofline peak. */
int sddp_fp64_peak_tflops(double *out, void *stream);

/* counters since create: kernels launched by this handle */
int sddp_launch_count(const SddpHandle *h, long long *out);

#ifdef __cplusplus
}
#endif
#endif
