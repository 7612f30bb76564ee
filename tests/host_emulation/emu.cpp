// emu.cpp -- TEST-ONLY host build of the device algorithm (csrc/mpc_core.cuh with MPC_HOST_EMULATION).
// Lets the not-gpu test-suite exercise the exact kernel source (lane loops serialised) against the
// oracle.  It is never part of libmpcgpu.so and the product never loads it.
#define MPC_HOST_EMULATION 1
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../model-predictive-control-tuning_b200/csrc/mpc_core.cuh"
#include "../../model-predictive-control-tuning_b200/csrc/mpc_tables.h"

template <typename... A>
static int sim_dispatch(int nu, A &&...a) {
    switch (nu) {
        case 1: return mpc_sim_run<1>(a...);
        case 2: return mpc_sim_run<2>(a...);
        case 3: return mpc_sim_run<3>(a...);
        default: return mpc_sim_run<4>(a...);
    }
}

extern "C" int emu_eval_batch(const mpcgpu_problem *pb, int n, const int *N, const int *Nu, const double *delta,
                              const double *lambda, int mode, double *cost, double *y, double *u, double *ys,
                              double *uopt, int *status, unsigned long long *counters, char *err, int errlen) {
    MpcHostTables ht;
    std::string e = mpc_build_tables(*pb, ht);
    if (!e.empty()) { std::strncpy(err, e.c_str(), errlen - 1); return 1; }
    const MpcLayout &L = ht.L;
    MpcTables T{ht.TG.data(), ht.TK.data(), ht.S1.data(), ht.r.data(), ht.v.data(), ht.yref.data()};
    const int ny = L.ny, nu = L.nu, nit = L.nit;
    const bool square = ny == nu;
    for (int c = 0; c < n; ++c) {
        const int p = N[c], m = Nu[c], nz = nu * m;
        if (p < 2 || p > L.pmax || m < 1 || m > L.mmax || m >= p) { if (status) status[c] = 4; continue; }
        std::vector<double> bsm(mpc_builder_smem_doubles(nz, L.nst));
        std::vector<double> Mg((size_t)L.nst * nz), Wg((size_t)nz * nz);
        int flag = 0;
        int st = mpc_build_candidate(L, T, p, m, delta + (size_t)c * ny, lambda + (size_t)c * nu, bsm.data(), Mg.data(),
                                     Wg.data(), &flag);
        std::vector<double> ssm(mpc_sim_smem_doubles(L, m) + 8);
        double part[MPC_MAXY + 1];
        if (st == 0) {
            if (mode == 2) {
                const int runs = square ? ny : 1;
                double F = 0.0;
                for (int rn = 0; rn < runs; ++rn) {
                    MpcRunOut out{part, y ? y + (size_t)c * ny * nit : nullptr, u ? u + (size_t)c * nu * nit : nullptr,
                                  ys ? ys + (size_t)c * ny * nit : nullptr, uopt ? uopt + (size_t)c * nu * nit : nullptr,
                                  counters};
                    int s2 = sim_dispatch(nu, L, T, p, m, Mg.data(), Wg.data(), 2, square ? rn : -1, ssm.data(), out);
                    if (s2) st = s2;
                    F += part[0];
                }
                cost[c] = st ? NAN : F + (double)p;
            } else {
                MpcRunOut out{mode == 1 ? cost + (size_t)c * ny : nullptr, y ? y + (size_t)c * ny * nit : nullptr,
                              u ? u + (size_t)c * nu * nit : nullptr, ys ? ys + (size_t)c * ny * nit : nullptr,
                              uopt ? uopt + (size_t)c * nu * nit : nullptr, counters};
                st = sim_dispatch(nu, L, T, p, m, Mg.data(), Wg.data(), mode, -2, ssm.data(), out);
            }
        } else if (cost) {
            if (mode == 1) for (int i = 0; i < ny; ++i) cost[(size_t)c * ny + i] = NAN;
            if (mode == 2) cost[c] = NAN;
        }
        if (status) status[c] = st;
    }
    return 0;
}
