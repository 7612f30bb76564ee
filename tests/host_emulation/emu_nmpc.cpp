// emu_nmpc.cpp -- TEST-ONLY host build of the NMPC group kernel (csrc/mpc_nmpc_group.cuh): one closed-loop run executed by a
// warp of 32 host threads (G = 32: one group, so every group-wide barrier is a warp barrier of the emulation), for the
// not-gpu test-suite to compare with the CPU port of the same algorithm (oracle/nmpc_port.cpp) and the scipy oracle.
// A translation unit of its own: mpc_nmpc_core.h defines NU / NX / NY as macros, the linear kernels use NU as a template parameter.
#include "simt.h"

#include <cmath>
#include <vector>

#include "../../include/mpcgpu.h"
#include "../../model-predictive-control-tuning_b200/csrc/mpc_nmpc_group.cuh"

// n candidates, GAM costs, G lanes per run (8, 16 or 32): 32 / G runs share a warp of host threads and may diverge
extern "C" int emu_nmpc_eval_group(int G, int nit, int pmax, int mmax, int inK, int nsub, int max_sqp, double Ts, const double *x0,
                                   const double *u0, const double *umin, const double *umax, const double *xmin, const double *xmax,
                                   const double *su, const double *sy, const double *r, const double *yref, int n, const int *N,
                                   const int *Nu, const double *delta, const double *lambda, double *cost, int *status) {
    NmpcDev D;
    D.nit = nit; D.pmax = pmax; D.mmax = mmax; D.inK = inK; D.nsub = nsub; D.max_sqp = max_sqp; D.Ts = Ts;
    for (int i = 0; i < NX; ++i) { D.x0[i] = x0[i]; D.xmin[i] = xmin ? xmin[i] : -INFINITY; D.xmax[i] = xmax ? xmax[i] : INFINITY; }
    for (int j = 0; j < NU; ++j) { D.u0[j] = u0[j]; D.umin[j] = umin[j]; D.umax[j] = umax[j]; D.su[j] = su[j]; }
    for (int j = 0; j < NY; ++j) D.sy[j] = sy[j];
    int maxnu = 1;
    std::vector<int> order(n);
    for (int c = 0; c < n; ++c) { order[c] = c; status[c] = 0; if (Nu[c] <= NM_MAXM && Nu[c] > maxnu) maxnu = Nu[c]; }
    const int maxz = NU * maxnu, rpw = 32 / G;
    unsigned long long counters[2] = {0, 0};
    NmpcArgs A{N, Nu, delta, lambda, r, yref, cost, nullptr, nullptr, nullptr, nullptr, nullptr, status, counters, nullptr};
    for (int block = 0; block * rpw < n; ++block) {
        std::vector<double> smem((size_t)nmg_doubles(maxz) * rpw + 8, std::nan(""));
        simt_run_warp([&]() {
            if (G == 8) nmg_run<8>(D, 0, n, 1, 1, order.data(), A, maxz, smem.data(), block);
            else if (G == 16) nmg_run<16>(D, 0, n, 1, 1, order.data(), A, maxz, smem.data(), block);
            else nmg_run<32>(D, 0, n, 1, 1, order.data(), A, maxz, smem.data(), block);
        });
    }
    return 0;
}

extern "C" int emu_nmpc_eval(int nit, int pmax, int mmax, int inK, int nsub, int max_sqp, double Ts, const double *x0, const double *u0,
                             const double *umin, const double *umax, const double *xmin, const double *xmax, const double *su,
                             const double *sy, const double *r, const double *yref, int N, int Nu, const double *delta,
                             const double *lambda, int mode, double *cost, double *y, double *u, unsigned long long *counters) {
    NmpcDev D;
    D.nit = nit; D.pmax = pmax; D.mmax = mmax; D.inK = inK; D.nsub = nsub; D.max_sqp = max_sqp; D.Ts = Ts;
    for (int i = 0; i < NX; ++i) { D.x0[i] = x0[i]; D.xmin[i] = xmin ? xmin[i] : -INFINITY; D.xmax[i] = xmax ? xmax[i] : INFINITY; }
    for (int j = 0; j < NU; ++j) { D.u0[j] = u0[j]; D.umin[j] = umin[j]; D.umax[j] = umax[j]; D.su[j] = su[j]; }
    for (int j = 0; j < NY; ++j) D.sy[j] = sy[j];
    const int runs = mode == 2 ? NY : 1, maxz = NU * (Nu >= 1 && Nu <= NM_MAXM ? Nu : 1);
    int order = 0, status = 0;
    double part[NY] = {0.0, 0.0};
    NmpcArgs A{&N, &Nu, delta, lambda, r, yref, cost, part, y, u, nullptr, nullptr, &status, counters, nullptr};
    for (int run = 0; run < runs; ++run) {
        std::vector<double> smem(nmg_doubles(maxz) + 8, std::nan(""));
        simt_run_warp([&]() { nmg_run<32>(D, run, run + 1, runs, mode, &order, A, maxz, smem.data(), 0); });
    }
    if (mode == 2) cost[0] = (status == 0 || status == 5) ? part[0] + part[1] + (double)N : NAN;   // k_nmpc_finish
    return status;
}
