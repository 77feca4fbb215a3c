/*
 * sddp_oracle.h -- CPU oracle for the srbd_horizon DDP hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under srbd_horizon_b200/ may include,
 * link or call this; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * reported CPU baseline.
 *
 * PARITY UNPINNED: the reference's DDP iterations live in the external module
 * `pyddp` (python/ddp.py:1,93-94,101), which is neither vendored nor pinned
 * and cannot run here.  This oracle restates
 *   - the problem definitions of python/prb.py (dynamics, residuals,
 *     constraints, parameter layout) and the way python/ddp.py:179-230
 *     assembles them into f_k, L_k, L_N, and
 *   - a textbook multiple-shooting iLQR/DDP iteration (documented in
 *     sddp_oracle.c) for the part pyddp hides.
 * The model functions are pinned against sympy golden vectors transcribed
 * from prb.py (tests/golden/make_golden.py); the LIP solve is pinned against
 * a dense KKT solve (it is an exact LQR).
 */
#ifndef SDDP_ORACLE_H
#define SDDP_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order as include/sddp.h:SddpConfig (kept in sync by a test). */
typedef struct OrcConfig {
    int32_t model;             /* 0 = SRBD (prb.py:16-246), 1 = LIP (prb.py:248-441) */
    int32_t N;                 /* shooting intervals `ns`; nodes 0..N */
    int32_t inertia_mode;      /* 0 = literal element-wise R*I*R^T (prb.py:99), 1 = rotated R I R^T */
    int32_t hessian_mode;      /* 0 = exact Hessian of the scalar L (ddp.py:210-214), 1 = Gauss-Newton */
    int32_t multiple_shooting; /* 0 = single shooting, 1 = keep x warm start / defects */
    int32_t max_iters;         /* ddp.py:17-19 */
    int32_t dense_backward;    /* kernel selection of the CUDA library; ignored here */
    int32_t lip_tail_start;    /* 0 = off; nodes k..N-1 use the LIP-style model (include/sddp.h) */
    double dt;                 /* prb.py:110  T/ns */
    double mass;               /* kindyn.mass(), prb.py:92 (synthetic here) */
    double inertia[9];         /* CRBA block, prb.py:94-95 (synthetic), row-major */
    double com[3];             /* prb.py:138-139 */
    double foot[12];           /* initial_foot_position[i], prb.py:127-135 */
    double force_scaling;      /* prb.py:98 */
    double gravity;            /* 9.81 */
    double eta2;               /* prb.py:317 (LIP) */
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
    double defect_contraction_rate;      /* README.md:6; <=0 means rho = alpha */
    double mu_min, mu_max, mu_factor;    /* regularisation schedule */
    double defect_ths;                   /* MS convergence: max |defect| */
    double friction_cone_weight;         /* 0 = reference behaviour (cone dropped, prb.py:173-177); see include/sddp.h */
    double friction_cone_mu;             /* prb.py:174 */
    double friction_cone_sharpness;      /* ddp.py:182 exp_parameter */
    double force_bound_weight;           /* bounds as exponential barriers, ddp.py:204-209; see include/sddp.h */
    double force_bound;
    double unilateral_weight;
    double cdot_bound_weight;
    double cdot_bound;
    double bound_sharpness;
} OrcConfig;

enum { ORC_HIST = 4 };   /* per-iteration record: cost, alpha, mu, max|defect| */

/* node kinds: which cost groups are active (prb.py node ranges) */
enum { ORC_NODE_FIRST = 0, ORC_NODE_MID = 1, ORC_NODE_TERM = 2, ORC_NODE_TAIL = 3 };   /* TAIL: MID node of the LIP-style tail */

enum { ORC_OK = 0, ORC_MAX_ITERS = 1, ORC_LS_FAILED = 2, ORC_REG_FAILED = 3, ORC_NAN = 4 };

void orc_dims(int model, int *nx, int *nu, int *np);

/* x+ = x + dt * ode(x,u)   (ddp.py:228-230, explicit Euler) */
void orc_dynamics(const OrcConfig *c, const double *x, const double *u, double *xn);
/* the same for a node of the given kind (ORC_NODE_TAIL: LIP-style tail, no rotational dynamics) */
void orc_dynamics_kind(const OrcConfig *c, int kind, const double *x, const double *u, double *xn);

/* L_k or L_N  (ddp.py:179-226) */
double orc_cost(const OrcConfig *c, int kind, const double *x, const double *u, const double *p);

/* dense, row-major: fx[nx*nx], fu[nx*nu], lx[nx], lu[nu], lxx[nx*nx], lux[nu*nx], luu[nu*nu] */
void orc_derivs(const OrcConfig *c, int kind, const double *x, const double *u, const double *p,
                double *fx, double *fu, double *lx, double *lu, double *lxx, double *lux, double *luu);

/* One solve.  X[(N+1)*nx], U[N*nu] hold the warm start on entry and the
 * solution on exit (node-major rows).  K[N*nu*nx], kff[N*nu] gains of the last
 * backward pass, hist[max_iters*ORC_HIST].  Returns status. */
int orc_solve(const OrcConfig *c, const double *x0, const double *params,
              double *X, double *U, double *K, double *kff,
              double *hist, int *iters, double *cost);

/* Batch of independent problems spread over `nthreads` pthreads. */
void orc_solve_batch(const OrcConfig *c, int B, const double *x0, const double *params,
                     double *X, double *U, double *K, double *kff,
                     double *hist, int *iters, int *status, double *cost, int nthreads);

/* stage-level entry points used by the stage parity tests */
double orc_total_cost(const OrcConfig *c, const double *X, const double *U, const double *params);
/* backward pass at regularisation mu; returns 0 or the failing node+1.
 * dV[3] = {D1, D2, C0}: model change  C0 + alpha*D1 + alpha^2*D2  */
int orc_backward(const OrcConfig *c, const double *X, const double *U, const double *params,
                 const double *defect, double mu, double *K, double *kff, double *dV);
/* forward rollout for one step size; returns the new cost */
double orc_forward(const OrcConfig *c, const double *x0, const double *X, const double *U, const double *params,
                   const double *defect, const double *K, const double *kff,
                   double alpha, double rho, double *Xn, double *Un);

#ifdef __cplusplus
}
#endif
#endif
