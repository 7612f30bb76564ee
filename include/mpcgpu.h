/*
 * mpcgpu.h -- C ABI of libmpcgpu.so: batched closed-loop MPC evaluation for controller tuning on B200.
 *
 * Drop-in boundary.  Each entry point replaces what a MATLAB caller does today through
 *   [y,u,t,ys,uopt] = closedloop_toolbox(mpc_toolbox,r,v,N,Nu,delta,lambda,nit)
 *        /root/reference/MPC-Tuning/MPC_Tuning/closedloop_toolbox.m:1
 *   [g,h] = GAM_fun(X,Par)            /root/reference/MPC-Tuning/MPC_Tuning/GAM_fun.m:1   (cost mode GAM)
 *   VNS trial cost                    /root/reference/MPC-Tuning/MPC_Tuning/VNS2.m:147-195 (cost mode VNS)
 * and is what a MEX gateway binds (mex/mpcgpu_mex.c).
 *
 * Conventions: plain C, no exceptions, int return codes (0 = ok), caller-owned buffers, doubles are IEEE
 * fp64.  A handle owns one CUDA device context + stream and is thread-compatible (one thread at a
 * time).  There is NO CPU fallback: every call fails with MPCGPU_ERR_CUDA when no device is usable.
 */
#ifndef MPCGPU_H
#define MPCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCGPU_MAX_NY 8
#define MPCGPU_MAX_NU 4
#define MPCGPU_MAX_NW 8   /* nu + nd */
#define MPCGPU_MAX_P 255  /* nbp <= 8 bits (MPCTuning.m:283) */
#define MPCGPU_MAX_M 15   /* nbc  = 4 bits (MPCTuning.m:284) */

enum {
    MPCGPU_OK = 0,
    MPCGPU_ERR_ARG = 1,      /* bad argument / unsupported size                */
    MPCGPU_ERR_CUDA = 2,     /* CUDA runtime error, see mpcgpu_last_error()    */
    MPCGPU_ERR_STATE = 3,    /* call order (e.g. run before upload)            */
    MPCGPU_ERR_UNSUPPORTED = 4
};

/* per-candidate status codes (status[] output) */
enum {
    MPCGPU_CAND_OK = 0,
    MPCGPU_CAND_INFEASIBLE = 1,   /* QP infeasible (cannot happen with MV box + rate limits only) */
    MPCGPU_CAND_ITER_CAP = 2,     /* active-set iteration cap hit                                 */
    MPCGPU_CAND_NOT_PD = 3,       /* Hessian not positive definite (lambda == 0 and rank-deficient G) */
    MPCGPU_CAND_INVALID = 4,      /* horizons illegal: p < 2, m < 1, m >= p, beyond pmax / mmax; with MPCGPU_OPT_VNS_LEGALITY also
                                     VNS2.m:135 (N <= dmin_i for some output, Nu <= 1) -- PreCon.m:23 is m >= p            */
    MPCGPU_CAND_BOUND_CROSSED = 5 /* NMPC only: the closed loop crossed an OV bound (soft in nlmpc, Weights.ECR) or a state bound (hard in
                                     nlmpc) of VanDeVusse_NMPC.m:140-145.  The restated NLP does NOT enforce them (they are inactive in the
                                     reference's tuning scenario); the cost of the unconstrained trajectory is returned and flagged, and
                                     the host wrappers reject such a candidate by default (cost = inf) */
};

/* cost modes */
enum {
    MPCGPU_COST_RAW = 0,  /* trajectories only: closedloop_toolbox.m outputs                      */
    MPCGPU_COST_GAM = 1,  /* g_i = sum_k (y_i - Yref_i)^2, n x ny           (GAM_fun.m:110-115)   */
    MPCGPU_COST_VNS = 2   /* F = sum(j21+j22) + N + sum(Jnu), n             (VNS2.m:147-195)      */
};

/* The state `Par` carries into the objective functions (MPCTuning.m:307-340) for a LINEAR plant whose
 * channels are first-order-plus-dead-time, already L/R-scaled (MPCTuning.m:154-200):
 *    y_ij(k) = a*y_ij(k-1) + b0*w_j(k-d) + b1*w_j(k-d-1),   w = [MVs ; MDs]
 * All matrices are row-major ny x (nu+nd).  Bounds may be +-INFINITY. */
typedef struct {
    int32_t ny, nu, nd, nit;
    int32_t pmax, mmax;            /* 2^nbp - 1, 2^nbc - 1 (MPCTuning.m:283-284)                  */
    int32_t inK;                   /* 1-based first cost sample of the VNS objective (VNS2.m:43)  */
    int32_t reserved;
    const double *a, *b0, *b1;
    const int32_t *d;
    const double *umin, *umax, *dumin, *dumax;       /* nu : MV(i).Min/Max/RateMin/RateMax        */
    const double *ymin, *ymax, *ecr_min, *ecr_max;   /* ny : OV(i).Min/Max/MinECR/MaxECR          */
    const double *su, *sy;                           /* MV / OV ScaleFactor                       */
    double rho_ecr;                                  /* Weights.ECR                               */
    const double *r;       /* nit x ny, time-major: set-point (Par.Xsp')                          */
    const double *v;       /* nit x nd, time-major: measured disturbance (Par.mdv'), may be NULL  */
    const double *yref;    /* ny x nit: reference trajectory (Par.Yref)                           */
    const int32_t *dmin;   /* ny: per-output minimum dead time (MPCTuning.m:257-262)              */
} mpcgpu_problem;

typedef struct mpcgpu_handle mpcgpu_handle;

typedef struct {
    uint64_t candidates;        /* candidates evaluated since create                       */
    uint64_t closed_loops;      /* closed-loop simulations (VNS on a square plant: ny each) */
    uint64_t qp_solves;         /* closed_loops * nit + open-loop QPs                      */
    uint64_t qp_constrained;    /* QPs that left the unconstrained fast path               */
    uint64_t as_iterations;     /* active-set iterations (adds + drops)                    */
    uint64_t kernel_launches;   /* CUDA kernels launched by this handle                    */
    double last_build_ms;       /* device time of the prediction/Hessian builder kernel(s) */
    double last_sim_ms;         /* device time of the closed-loop kernel(s) (dominant)     */
    double last_total_ms;       /* device time of the whole last run                       */
} mpcgpu_counters;

/* Create an evaluator on CUDA device `device` (-1: current).  Copies the problem, builds the
 * candidate-independent prediction tables (step responses, prefix Grams, free-response maps). */
int mpcgpu_create(const mpcgpu_problem *problem, int device, mpcgpu_handle **out);
void mpcgpu_destroy(mpcgpu_handle *h);

/* Replace the signals (closedloop_toolbox takes r, v, nit per call: closedloop_toolbox.m:1).
 * r: nit x ny, v: nit x nd (may be NULL when nd == 0), yref: ny x nit; yref == NULL keeps the current reference trajectory
 * when nit is unchanged and zero-fills it otherwise. */
int mpcgpu_set_signals(mpcgpu_handle *h, int nit, const double *r, const double *v, const double *yref);

/* One call = one population.  HOST pointers; copies in, runs, copies out, synchronises.
 *   N, Nu   : n            prediction / control horizon per candidate (max(N), max(Nu) of the caller)
 *   delta   : n x ny       row-major (candidate-major)   Weights.OV
 *   lambda  : n x nu       row-major                     Weights.MVRate
 *   cost    : GAM: n x ny ; VNS: n ; RAW: ignored (may be NULL)
 *   y,u,ys,uopt : NULL or n x (ny|nu) x nit, signals x time per candidate (closedloop_toolbox.m:103-107)
 *   status  : NULL or n
 * Failed candidates get NaN cost and a non-zero status; the call itself still returns MPCGPU_OK, which
 * is what lets the reference's try/catch callers (GAM_fun.m:80-91, VNS2.m:151-163) keep working. */
int mpcgpu_eval_batch(mpcgpu_handle *h, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                      const double *lambda, int cost_mode, double *cost, double *y, double *u, double *ys,
                      double *uopt, int32_t *status);

/* Split form of the same call, for callers that keep a population resident on the device:
 * upload (H2D + host-side size bucketing) -> run (kernels only, asynchronous on `stream`) -> download. */
int mpcgpu_upload(mpcgpu_handle *h, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                  const double *lambda);
int mpcgpu_run(mpcgpu_handle *h, int cost_mode, int want_traj, void *cuda_stream /* cudaStream_t or NULL */);
int mpcgpu_download(mpcgpu_handle *h, int cost_mode, double *cost, double *y, double *u, double *ys,
                    double *uopt, int32_t *status);
/* Device pointer of the cost buffer of the last run (n x ny or n doubles), for on-device consumers
 * such as an NCCL all-gather of fitness.  The consumer must order itself after the run on the run's stream; the next
 * mpcgpu_upload / mpcgpu_run on this handle waits for the previous run before it touches the buffers. */
int mpcgpu_cost_device_ptr(mpcgpu_handle *h, int cost_mode, void **ptr, int *count);

/* Options.  MPCGPU_OPT_VNS_LEGALITY = 1: candidates the VNS search would reject before evaluating them
 * (VNS2.m:135  any(N<=dmin) | any(Nu<=1), dmin = problem.dmin, MPCTuning.m:257-262) get MPCGPU_CAND_INVALID and a NaN
 * cost instead of being simulated.  Off by default: closedloop_toolbox.m and GAM_fun.m themselves accept them. */
enum { MPCGPU_OPT_VNS_LEGALITY = 1 };
int mpcgpu_set_option(mpcgpu_handle *h, int option, int value);

/* Plant-model mismatch validation run (/root/reference/MPC-Tuning/Shell3x3.m:271-286, WoodBerry.m:263-278, Shell7x5.m:293-306:
 *   options = mpcsimopt(mpc_toolbox); options.Model = plant; [y,t,u] = sim(mpc_toolbox,nit,r,[],options)).
 * After this call every evaluation of the handle (GAM cost or RAW trajectories y, u; ys / uopt are zero) simulates the
 * controller against the REAL plant given here -- channels y(k) = a y(k-1) + b0 w(k-d) + b1 w(k-d-1), ny x nw row-major
 * like mpcgpu_problem's, scaled like the model (L*Psr*R) -- while the controller keeps predicting with the handle's model
 * and corrects its state at every sample with the estimator gain:
 *   x_c(k|k) = x_c(k|k-1) + gain * (y(k) - C x_c(k|k-1)),   x_c = [channel states (ny*nw); MV delay-line states w_j(k-1-q),
 *   q = 0..hl-1 (nu*hl); output-disturbance states (ny)],   gain: (ny*nw + nu*hl + ny) x ny row-major.
 * The gain is an INPUT: the Toolbox's own `getEstimator(mpcobj)` mapped to this state order, or the restated Toolbox default
 * (integrated white noise on every output, unit white noise on every MV and measurement, steady-state Kalman filter:
 * mpcgpu/estimator.py default_estimator_gain).  plant_a == NULL: back to the nominal evaluation.  Runs on the general
 * block-per-run kernel (hard MV limits and soft output limits alike); the VNS objective is not defined for it. */
int mpcgpu_set_mismatch(mpcgpu_handle *h, const double *plant_a, const double *plant_b0, const double *plant_b1,
                        const int32_t *plant_d, const double *gain, int hl);

/* [y,u,t,ys,uopt] = closedloop_toolbox(mpc_toolbox,r,v,N,Nu,delta,lambda,nit) as ONE call
 * (/root/reference/MPC-Tuning/MPC_Tuning/closedloop_toolbox.m:1): r nit x ny, v nit x nd (NULL when nd == 0), N / Nu the
 * max of the caller's vectors (:38-40), outputs signals x time (ny|nu x nit, :103-107; any may be NULL), t = (0..nit-1)*Ts
 * is the caller's.  The call's signals are used for this evaluation only: the handle's own set-point, disturbance,
 * reference trajectory and nit (Par.Xsp, Par.mdv, Par.Yref) are unchanged afterwards, so GAM / VNS costs evaluated
 * before and after are identical -- the reference's closedloop_toolbox never touches Par. */
int mpcgpu_closedloop(mpcgpu_handle *h, int nit, const double *r, const double *v, int32_t N, int32_t Nu,
                      const double *delta, const double *lambda, double *y, double *u, double *ys, double *uopt,
                      int32_t *status);

int mpcgpu_get_counters(mpcgpu_handle *h, mpcgpu_counters *out);
const char *mpcgpu_last_error(mpcgpu_handle *h); /* h may be NULL: last create() error */

/* Multi-GPU evaluator: one evaluator per listed device inside ONE process (what a MEX gateway needs
 * to use the 8 B200s of a box).  mpcgpu_multi_eval_batch deals the candidates by estimated work (sorted round-robin),
 * launches every device asynchronously and gathers the fitness (and status) into the caller's host arrays in population
 * order; costs only (GAM: n x ny, VNS: n).  Results are bit-identical to the single-device call: a candidate's cost does
 * not depend on its position in a population or on the device.  devices == NULL: 0 .. ndev-1.
 * (One process per GPU -- torchrun, bench.py -- shards with the same key, mpcgpu_work_estimate, and all-gathers with NCCL.) */
typedef struct mpcgpu_multi mpcgpu_multi;
int mpcgpu_create_multi(const mpcgpu_problem *problem, const int *devices, int ndev, mpcgpu_multi **out);
void mpcgpu_destroy_multi(mpcgpu_multi *m);
int mpcgpu_multi_device_count(mpcgpu_multi *m);
int mpcgpu_multi_set_option(mpcgpu_multi *m, int option, int value);
int mpcgpu_multi_set_signals(mpcgpu_multi *m, int nit, const double *r, const double *v, const double *yref);
int mpcgpu_multi_eval_batch(mpcgpu_multi *m, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                            const double *lambda, int cost_mode, double *cost, int32_t *status);
int mpcgpu_multi_get_counters(mpcgpu_multi *m, int device_index, mpcgpu_counters *out);
const char *mpcgpu_multi_last_error(mpcgpu_multi *m); /* m may be NULL: last create error */
/* work[c]: relative a-priori cost of candidate c (moves, aggressiveness of the weights, horizons barely past the dead time) */
int mpcgpu_work_estimate(const mpcgpu_problem *problem, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                         const double *lambda, double *work);
int mpcgpu_device_count(void);

/* fp64 FMA throughput microbenchmark (TFLOP/s, 2 flops per FMA) used as the roofline denominator of
 * this path (MEASURED_PEAKS.json has no fp64 entry). */
int mpcgpu_measure_fp64_peak(int device, double *tflops);


/* ------------------------------------------------------------------------------------------------
 * DTC-GPC batched sweep (BASELINE.json configs[3]): the unconstrained dead-time-compensated GPC of
 *   /root/reference/DTC-GPC/DTC_GPC_WW.m:56-164  (gain K = (H'QH+W)\H'Q, optimal predictor, control loop)
 * built from MatG.m:38-74, diophantine.m:15-79, diophantineMIMO.m:14-22, deltaUFree.m:12-63,
 * BA_MIMO.m:16-72, descompMPC.m:19-43, cell2mat2.m:25-58, OptimalPredictor2.m:24-40.
 * One candidate = (p[ny], m[nu], delta[ny], lambda[nu], robustness filter Fr per output).  The filter
 * coefficients are inputs (filter design is outside the hot path).  All channel matrices row-major.
 * ---------------------------------------------------------------------------------------------- */
#define MPCGPU_DTC_MAXF 8 /* max filter polynomial length */
typedef struct {
    int32_t ny, nu, nq, nit;
    int32_t pmax, mmax;          /* sweep bounds on the prediction window p_i and control horizon m_j */
    int32_t k_start;             /* 1-based first sample at which the controller acts (DTC_GPC_WW.m:128: 4) */
    int32_t reserved;
    const double *ma, *mb0, *mb1; const int32_t *md;   /* conditioned discrete model Pnz = c2d(L*Pn*R), ny x nu */
    const double *pa, *pb0, *pb1; const int32_t *pd;   /* discrete process P (unscaled), ny x nu              */
    const double *qa, *qb0, *qb1; const int32_t *qd;   /* discrete disturbance model Pq (unscaled), ny x nq   */
    const double *L, *R;                               /* conditioning diagonals, ny and nu                   */
    const double *r;                                   /* ny x nit reference (unscaled)                       */
    const double *q;                                   /* nq x nit disturbance                                */
} mpcgpu_dtc_problem;

typedef struct mpcgpu_dtc_handle mpcgpu_dtc_handle;

int mpcgpu_dtc_create(const mpcgpu_dtc_problem *problem, int device, mpcgpu_dtc_handle **out);
void mpcgpu_dtc_destroy(mpcgpu_dtc_handle *h);
/* p: n x ny, m: n x nu, delta: n x ny, lambda: n x nu (weights are NOT squared here: Q = delta*I, W = lambda*I,
 * DTC_GPC_WW.m:67-76); fr_num / fr_den: n x ny x MPCGPU_DTC_MAXF (descending powers of z, zero padded),
 * fr_len: n x ny x 2 = {len(num), len(den)}.  ise: n x ny = sum_k (y_i - r_i)^2;  y: n x ny x nit, u: n x nu x nit
 * (either may be NULL);  status: n. */
int mpcgpu_dtc_eval_batch(mpcgpu_dtc_handle *h, int n, const int32_t *p, const int32_t *m, const double *delta,
                          const double *lambda, const double *fr_num, const double *fr_den, const int32_t *fr_len,
                          double *ise, double *y, double *u, int32_t *status);
/* The same sweep with the robustness filter DESIGNED ON THE DEVICE per candidate from (alfa, raio)
 * (/root/reference/DTC-GPC/mimofilter.m:33-50, filtro_siso.m:26-96: Dr = (z - alfa)^ns over the ns poles of the output's
 * row model with |pole| >= raio, Nr = remainder of Dr z^d by (z - 1) prod(z - slow poles), d = the row's minimum dead
 * time; no slow pole: Fr = 1).  alfa, raio: n.  A row model with slow poles and no dead time (the reference's system is
 * under-determined there) flags the candidate MPCGPU_CAND_INVALID. */
int mpcgpu_dtc_eval_batch_design(mpcgpu_dtc_handle *h, int n, const int32_t *p, const int32_t *m, const double *delta,
                                 const double *lambda, const double *alfa, const double *raio, double *ise, double *y,
                                 double *u, int32_t *status);
int mpcgpu_dtc_get_counters(mpcgpu_dtc_handle *h, mpcgpu_counters *out); /* candidates, kernel_launches, last_sim_ms (device time of k_dtc) */
const char *mpcgpu_dtc_last_error(mpcgpu_dtc_handle *h);
/* Host-only (no CUDA call): the candidate-independent polynomial tables the sweep is built from -- step responses
 * (MatG.m:51), Diophantine F rows (diophantine.m:55-65), past-control rows (deltaUFree.m:36-57).  info: 64 ints
 * {step_len, pmax, MAXNA, MAXCP, sum duM, sum(na+1)}, [8+i] na_i, [16+i] dmin_i, [24+i*nu+j] cp_ij, [48+j] duM_j;
 * step: ny*nu*step_len, ftab: ny*(pmax+1)*MAXNA, ug: ny*nu*(pmax+1)*MAXCP.  Any output may be NULL. */
int mpcgpu_dtc_host_tables(const mpcgpu_dtc_problem *problem, int32_t *info, double *step, double *ftab, double *ug);


/* ------------------------------------------------------------------------------------------------
 * Nonlinear path (BASELINE.json configs[4]): batched
 *   [y,u,yopt,uopt] = closedloop_toolbox_nmpc(nmpcobj,model,init,r,N,Nu,delta,lambda,nit)
 *        /root/reference/MPC-Tuning/MPC_Tuning/closedloop_toolbox_nmpc.m:1
 * with the GAM / VNS objectives fused (GAM_fun.m:87,110-115; VNS2.m:147-195 nonlinear branch) for the plant of
 *   /root/reference/MPC-Tuning/vandevusse_model.m:39-77 (3 states, 2 MVs, outputs = states 2..3, :73).
 * The controller call nlmpcmove is restated as: MV levels over the control horizon as decision variables, cost
 * sum (delta/sy (r - y))^2 + sum (lambda/su du)^2 with r held over the horizon, hard MV bounds, prediction = plant =
 * RK4 with `nsub` sub-steps per sample; solved by Gauss-Newton SQP with an exact box-QP step.
 * ---------------------------------------------------------------------------------------------- */
enum { MPCGPU_MODEL_VANDEVUSSE = 0 };
typedef struct {
    int32_t nit, pmax, mmax;     /* samples; 2^nbp - 1; 2^nbc - 1 (<= 15)                                  */
    int32_t inK;                 /* 1-based first cost sample of the VNS objective (VNS2.m:43: 10)          */
    int32_t nsub, max_sqp;       /* RK4 sub-steps per sample (>= 4 for this plant); SQP iteration cap       */
    int32_t model, reserved;     /* MPCGPU_MODEL_VANDEVUSSE                                                 */
    double Ts;
    const double *x0, *u0;       /* init.x0 (3), init.u0 (2)                                                */
    const double *umin, *umax;   /* MV(i).Min / Max (2)                                                     */
    const double *xmin, *xmax;   /* States(i).Min / Max (3): checked, not enforced (may be NULL)            */
    const double *su, *sy;       /* MV / OV ScaleFactor (2, 2)                                              */
    const double *r, *yref;      /* 2 x nit set-point and reference trajectory, signals x time              */
} mpcgpu_nmpc_problem;
typedef struct mpcgpu_nmpc_handle mpcgpu_nmpc_handle;
int mpcgpu_nmpc_create(const mpcgpu_nmpc_problem *problem, int device, mpcgpu_nmpc_handle **out);
void mpcgpu_nmpc_destroy(mpcgpu_nmpc_handle *h);
/* N, Nu: n; delta, lambda: n x 2 row-major; cost_mode RAW | GAM (cost n x 2) | VNS (cost n);
 * r_override: NULL or 2 x nit set-point for this call (closedloop_toolbox_nmpc takes r per call);
 * y, u, yopt, uopt: NULL or n x 2 x nit; status: NULL or n. */
int mpcgpu_nmpc_eval_batch(mpcgpu_nmpc_handle *h, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                           const double *lambda, int cost_mode, const double *r_override, double *cost, double *y, double *u,
                           double *yopt, double *uopt, int32_t *status);
int mpcgpu_nmpc_get_counters(mpcgpu_nmpc_handle *h, mpcgpu_counters *out); /* qp_solves = controller calls, as_iterations = SQP iterations */
const char *mpcgpu_nmpc_last_error(mpcgpu_nmpc_handle *h);

/* ------------------------------------------------------------------------------------------------
 * Single-shooting NMPC: batched
 *   [y, u] = ClosedLoopNMPC(x0_model, x_control, u0, r, N, Nu, Q, W, nit, ub1, lb1, inK, Ts)
 *        /root/reference/Explicit NMPC/ClosedLoopNMPC.m:1
 * whose controller call is  duOt = NMPC_Controller(Par)  (/root/reference/Explicit NMPC/NMPC_Controller.m:1): per input a
 * block of Nu_j offsets from u_j(k-1), cost sum_j Q_j sum_i (r_j(k) - (y_j(k+i) + n_j))^2 + sum_j W_j sum_c X_jc^2 (weights
 * not squared, :128-138), n_j the model-deviation term of :106-123, bounds lb - u(k-1) <= X <= ub - u(k-1), start X = 0.
 * Plant and model: plant_model.m:1-56 (= the Van de Vusse right-hand side above), RK4 with `nsub` sub-steps where the
 * reference calls ode23t / ode45; minimiser by Gauss-Newton with an exact box-QP step where the reference calls fmincon-SQP
 * (csrc/mpc_ssnmpc_core.h: S1-S5).  The reference's loop adds 0.01*randn to the plant state every sample
 * (ClosedLoopNMPC.m:88-90): here the draws are an argument (`noise`), NULL = none.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t nit, pmax;           /* samples; largest prediction horizon a candidate may ask for               */
    int32_t inK;                 /* 1-based first simulated sample (main.m:52: 4)                             */
    int32_t nsub, max_sqp;       /* RK4 sub-steps per sample; Gauss-Newton iteration cap per controller call  */
    int32_t model;               /* MPCGPU_MODEL_VANDEVUSSE                                                   */
    int32_t x_control[2];        /* 0-based indices of the controlled states (main.m:75: [2 3] -> {1, 2})     */
    double Ts;
    const double *x0, *u0;       /* x0_model (3), u0 (2)                                                      */
    const double *lb, *ub;       /* lb1, ub1 (2)                                                              */
    const double *r;             /* 2 x nit set-point, signals x time                                         */
} mpcgpu_ssnmpc_problem;
typedef struct mpcgpu_ssnmpc_handle mpcgpu_ssnmpc_handle;
int mpcgpu_ssnmpc_create(const mpcgpu_ssnmpc_problem *problem, int device, mpcgpu_ssnmpc_handle **out);
void mpcgpu_ssnmpc_destroy(mpcgpu_ssnmpc_handle *h);
/* N: n (1 <= N <= pmax); Nu: n x 2 row-major, per-input control horizons (1 <= Nu_j <= N, sum <= 30); Q, W: n x 2;
 * r_override: NULL or 2 x nit; noise: NULL or 3 x nit (added to the plant state after the step of sample k);
 * cost: NULL or n x 2 = sum over k = inK..nit of (y_j(k) - r_j(k))^2 (NaN when status != 0); y, u: NULL or n x 2 x nit;
 * status: NULL or n (MPCGPU_CAND_*: 0 ok, 2 / 3 box-QP failure, 4 invalid horizons). */
int mpcgpu_ssnmpc_eval_batch(mpcgpu_ssnmpc_handle *h, int n, const int32_t *N, const int32_t *Nu, const double *Q, const double *W,
                             const double *r_override, const double *noise, double *cost, double *y, double *u, int32_t *status);
int mpcgpu_ssnmpc_get_counters(mpcgpu_ssnmpc_handle *h, mpcgpu_counters *out); /* qp_solves = controller calls, as_iterations = Gauss-Newton iterations */
const char *mpcgpu_ssnmpc_last_error(mpcgpu_ssnmpc_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* MPCGPU_H */
