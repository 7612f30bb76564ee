// nmpc_port.cpp -- CPU baseline of the nonlinear path (bench.py --config vdv): the SAME restated algorithm the GPU kernel
// runs (csrc/mpc_nmpc_core.h: one nlmpcmove = Gauss-Newton SQP on the NLP of closedloop_toolbox_nmpc.m:69 restated as N1-N4,
// Van de Vusse right-hand side of vandevusse_model.m:39-77, closed loop of closedloop_toolbox_nmpc.m:36-97, GAM / VNS sums of
// GAM_fun.m:110-115 / VNS2.m:147-195), compiled for the host and parallelised with OpenMP over candidates.
// TEST / BENCH INFRASTRUCTURE: a timing baseline ("kind": "port"), never loaded by the product.  The independent checker of
// the nonlinear path stays oracle/nmpc_oracle.py (scipy, a different solver on the same problem); the reference's own
// nlmpcmove (MATLAB, closed source) cannot run here: parity unpinned vs the Toolbox.
#include <cmath>
#include <cstddef>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../model-predictive-control-tuning_b200/csrc/mpc_nmpc_core.h"
#include "../model-predictive-control-tuning_b200/csrc/mpc_ssnmpc_core.h"

extern "C" int nmpc_port_eval_batch(int nit, int pmax, int mmax, int inK, int nsub, int max_sqp, double Ts, const double *x0,
                                    const double *u0, const double *umin, const double *umax, const double *xmin,
                                    const double *xmax, const double *su, const double *sy, const double *r, const double *yref,
                                    int n, const int *N, const int *Nu, const double *delta, const double *lambda, int mode,
                                    double *cost, int *status, int nthreads) {
    NmpcDev D;
    D.nit = nit; D.pmax = pmax; D.mmax = mmax; D.inK = inK; D.nsub = nsub; D.max_sqp = max_sqp; D.Ts = Ts;
    for (int i = 0; i < NX; ++i) { D.x0[i] = x0[i]; D.xmin[i] = xmin ? xmin[i] : -INFINITY; D.xmax[i] = xmax ? xmax[i] : INFINITY; }
    for (int j = 0; j < NU; ++j) { D.u0[j] = u0[j]; D.umin[j] = umin[j]; D.umax[j] = umax[j]; D.su[j] = su[j]; }
    for (int j = 0; j < NY; ++j) D.sy[j] = sy[j];
    const int runs = mode == 2 ? NY : 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        std::vector<double> H(NM_LD * NM_LD), Lc(NM_LD * NM_LD);
#pragma omp for schedule(dynamic, 1)
        for (int c = 0; c < n; ++c) {
            const int p = N[c], m = Nu[c];
            if (p < 2 || p > pmax || m < 1 || m > mmax || m >= p) {
                status[c] = 4;
                if (mode == 1) for (int j = 0; j < NY; ++j) cost[(size_t)c * NY + j] = NAN;
                if (mode == 2) cost[c] = NAN;
                continue;
            }
            int st = 0;
            double F = 0.0;
            for (int run = 0; run < runs; ++run) {
                double out[NY];
                unsigned nc = 0, ns = 0;
                const int s1 = nmpc_run(D, p, m, mode, mode == 2 ? run : -1, delta + (size_t)c * NY, lambda + (size_t)c * NU, r, yref,
                                        nullptr, nullptr, nullptr, nullptr, out, H.data(), Lc.data(), &nc, &ns);
                if (s1 > st) st = s1;
                if (mode == 1) for (int j = 0; j < NY; ++j) cost[(size_t)c * NY + j] = out[j];
                if (mode == 2) F += out[0];
            }
            if (mode == 2) cost[c] = (st == 0 || st == 5) ? F + (double)p : NAN;   // VNS2.m:195: ... + N(1)
            status[c] = st;
        }
    }
    return 0;
}

// Single-shooting formulation (`Explicit NMPC/`, csrc/mpc_ssnmpc_core.h) compiled for the host: the code of k_ssnmpc run per
// candidate with OpenMP.  Nu: n x 2 (per-input control horizons), Q, W: n x 2; noise: nx x nit or NULL; cost: n x 2 or NULL;
// y, u: n x 2 x nit or NULL.
extern "C" int ssnmpc_port_eval_batch(int nit, int pmax, int inK, int nsub, int max_sqp, double Ts, const double *x0, const double *u0,
                                      const double *lb, const double *ub, const int *xc, const double *r, const double *noise, int n,
                                      const int *N, const int *Nu, const double *Q, const double *W, double *cost, double *y, double *u,
                                      int *status, int nthreads, unsigned long long *counters_out) {
    SsnmpcDev S;
    NmpcDev &D = S.D;
    D.nit = nit; D.pmax = pmax; D.mmax = NM_MAXM; D.inK = inK; D.nsub = nsub; D.max_sqp = max_sqp; D.Ts = Ts;
    for (int i = 0; i < NX; ++i) { D.x0[i] = x0[i]; D.xmin[i] = -INFINITY; D.xmax[i] = INFINITY; }
    for (int j = 0; j < NU; ++j) { D.u0[j] = u0[j]; D.umin[j] = lb[j]; D.umax[j] = ub[j]; D.su[j] = ub[j] - lb[j]; }
    for (int j = 0; j < NY; ++j) { D.sy[j] = 1.0; S.xc[j] = xc[j]; }
    S.pmax = pmax;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    unsigned long long counters[2] = {0, 0};
    SsArgs A{N, Nu, Q, W, r, noise, cost, y, u, status, counters};
#pragma omp parallel
    {
        std::vector<double> H(NM_LD * NM_LD), Lc(NM_LD * NM_LD);
#pragma omp for schedule(dynamic, 1)
        for (int c = 0; c < n; ++c) ss_item(S, nullptr, A, c, H.data(), Lc.data());   // what a thread of k_ssnmpc runs
    }
    if (counters_out) { counters_out[0] = counters[0]; counters_out[1] = counters[1]; }
    return 0;
}

// one NMPC_Controller call of the host build (tests compare it with the scipy minimiser on the same objective); X: sum(Nu)
extern "C" int ssnmpc_port_controller(int nsub, int max_sqp, double Ts, const double *lb, const double *ub, const int *xc,
                                      const double *x, const double *uprev, const double *r, int N, const int *Nu, const double *Q,
                                      const double *W, double *X, int *n_sqp) {
    SsnmpcDev S;
    NmpcDev &D = S.D;
    D.nit = 0; D.pmax = N; D.mmax = NM_MAXM; D.inK = 1; D.nsub = nsub; D.max_sqp = max_sqp; D.Ts = Ts;
    for (int i = 0; i < NX; ++i) { D.x0[i] = x[i]; D.xmin[i] = -INFINITY; D.xmax[i] = INFINITY; }
    for (int j = 0; j < NU; ++j) { D.u0[j] = uprev[j]; D.umin[j] = lb[j]; D.umax[j] = ub[j]; D.su[j] = ub[j] - lb[j]; }
    for (int j = 0; j < NY; ++j) { D.sy[j] = 1.0; S.xc[j] = xc[j]; }
    S.pmax = N;
    std::vector<double> H(NM_LD * NM_LD), Lc(NM_LD * NM_LD);
    unsigned ns = 0;
    const int rc = ss_controller(S, x, uprev, r, N, Nu, Q, W, X, H.data(), Lc.data(), &ns);
    *n_sqp = (int)ns;
    return rc;
}
