// bound_activity.cpp -- DIAGNOSTIC (not product, not test): how often would the OV / state bounds of VanDeVusse_NMPC.m:140-145
// have been ACTIVE inside a controller call?  The restated NLP (N1-N4) does not enforce them; nlmpc does, on the PREDICTED
// trajectory of every call.  This replays the closed loop of csrc/mpc_nmpc_core.h for a population and counts the controller
// calls whose optimal plan predicts a state outside [xmin, xmax] (by more than tol): only there can nlmpc's plan differ.
// worst_excess: largest signed excess over a bound as a fraction of the bound's range (negative = margin left).
// build: g++ -O2 -fopenmp -shared -fPIC -o /tmp/libbound_activity.so tools/nmpc_bounds/bound_activity.cpp
#include <cmath>
#include <cstddef>
#include <vector>
#include "../../model-predictive-control-tuning_b200/csrc/mpc_nmpc_core.h"

extern "C" int bound_activity(int nit, int nsub, int max_sqp, double Ts, const double *x0, const double *u0, const double *umin,
                              const double *umax, const double *xmin, const double *xmax, const double *su, const double *sy,
                              const double *r, int n, const int *N, const int *Nu, const double *delta, const double *lambda, double tol,
                              int *calls_violating, double *worst_excess) {
    NmpcDev D;
    D.nit = nit; D.pmax = 31; D.mmax = NM_MAXM; D.inK = 10; D.nsub = nsub; D.max_sqp = max_sqp; D.Ts = Ts;
    for (int i = 0; i < NX; ++i) { D.x0[i] = x0[i]; D.xmin[i] = xmin[i]; D.xmax[i] = xmax[i]; }
    for (int j = 0; j < NU; ++j) { D.u0[j] = u0[j]; D.umin[j] = umin[j]; D.umax[j] = umax[j]; D.su[j] = su[j]; }
    for (int j = 0; j < NY; ++j) D.sy[j] = sy[j];
#pragma omp parallel
    {
        std::vector<double> H(NM_LD * NM_LD), Lc(NM_LD * NM_LD);
#pragma omp for schedule(dynamic, 1)
        for (int c = 0; c < n; ++c) {
            const int p = N[c], m = Nu[c];
            double wy2[NY], wu2[NU], v[NM_MAXZ], rr[NY];
            for (int j = 0; j < NY; ++j) { const double w = delta[c * NY + j] / D.sy[j]; wy2[j] = w * w; }
            for (int j = 0; j < NU; ++j) { const double w = lambda[c * NU + j] / D.su[j]; wu2[j] = w * w; }
            double x[NX] = {D.x0[0], D.x0[1], D.x0[2]}, uprev[NU] = {D.u0[0], D.u0[1]};
            for (int i = 0; i < NU * m; ++i) v[i] = D.u0[i % NU];
            int bad = 0; double worst = -INFINITY; unsigned ns = 0;
            for (int k = 1; k < nit; ++k) {
                for (int j = 0; j < NY; ++j) rr[j] = r[(size_t)j * nit + k];
                nlmpcmove(D, x, uprev, rr, p, m, wy2, wu2, v, H.data(), Lc.data(), &ns);
                double xp[NX] = {x[0], x[1], x[2]}, ex = -INFINITY;                       // the plan's own prediction
                for (int i = 0; i < p; ++i) {
                    rk4_sample(D, xp, v + NU * (i < m ? i : m - 1), nullptr);
                    for (int s = 0; s < NX; ++s) ex = fmax(ex, fmax(D.xmin[s] - xp[s], xp[s] - D.xmax[s]) / (D.xmax[s] - D.xmin[s]));
                }
                if (ex > tol) bad++;
                worst = fmax(worst, ex);
                uprev[0] = v[0]; uprev[1] = v[1];
                rk4_sample(D, x, uprev, nullptr);
            }
            calls_violating[c] = bad; worst_excess[c] = worst;
        }
    }
    return 0;
}
