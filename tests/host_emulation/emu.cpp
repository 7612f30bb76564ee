// emu.cpp -- TEST-ONLY host build of the device algorithm sources:
//   csrc/mpc_core.cuh (builder, thread-serialised via MPC_HOST_EMULATION) and
//   csrc/mpc_sim.cuh  (closed-loop warp kernel, executed thread-per-lane through simt.h).
// Lets the not-gpu test-suite exercise the exact kernel source against the oracle.  It is never part of
// libmpcgpu.so and the product never loads it.
#define MPC_HOST_EMULATION 1
#include "simt.h"
int g_sim_knob = 0;
double *g_soft_dump = nullptr;
int g_soft_dump_k = -1, g_soft_k = -2;
extern "C" void emu_soft_dump(double *p, int k) { g_soft_dump = p; g_soft_dump_k = k; }
int g_sim_verbose = 0;
long long g_spec_overflow = 0;
extern "C" long long emu_spec_overflows() { long long r = g_spec_overflow; g_spec_overflow = 0; return r; }
long long g_spec_blocks = 0, g_spec_samples = 0, g_spec_fail = 0, g_spec_replay = 0;
extern "C" void emu_spec_stats(long long *o) { o[0] = g_spec_blocks; o[1] = g_spec_samples; o[2] = g_spec_fail; o[3] = g_spec_replay; g_spec_blocks = g_spec_samples = g_spec_fail = g_spec_replay = 0; }
extern "C" void emu_set_verbose(int v) { g_sim_verbose = v; }
long long g_sim_reappends = 0, g_sim_rotations = 0;
extern "C" long long emu_rotations() { long long r = g_sim_rotations; g_sim_rotations = 0; return r; }
extern "C" void emu_set_knob(int k) { g_sim_knob = k; }
extern "C" long long emu_reappends() { long long r = g_sim_reappends; g_sim_reappends = 0; return r; }

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../model-predictive-control-tuning_b200/csrc/mpc_core.cuh"
#include "../../model-predictive-control-tuning_b200/csrc/mpc_sim.cuh"
#include "../../model-predictive-control-tuning_b200/csrc/mpc_sim_spec.cuh"
#include "../../model-predictive-control-tuning_b200/csrc/mpc_soft.cuh"
#include "../../model-predictive-control-tuning_b200/csrc/mpc_tables.h"

template <int NU>
static int sim_p(int P, const MpcLayout &L, const MpcTables &T, int m, const double *Mg, const double *Wg, int mode, int sel,
                 double *smem, double *gscr, const MpcRunOut &out) {
    double *slot = gscr + sim_scratch_doubles(NU * P);
    if (sim_spec_ok(L) && P == 16 && !(g_sim_knob & 256)) {   // the kernel the product runs for plants whose deviation state fits one warp
        int st;
        switch (P) {
            case 4: st = sim_run_spec<NU, 4>(L, T, m, Mg, Wg, mode, sel, smem, gscr, out, slot); break;
            case 8: st = sim_run_spec<NU, 8>(L, T, m, Mg, Wg, mode, sel, smem, gscr, out, slot); break;
            default: st = sim_run_spec<NU, 16>(L, T, m, Mg, Wg, mode, sel, smem, gscr, out, slot); break;
        }
        if (st != SIM_ST_OVERFLOW) return st;
        if (threadIdx.x == 0) g_spec_overflow += 1;
        __syncwarp();   // second pass: the kernel with the spill area, as mpcgpu_run does
    }
    switch (P) {
        case 4: return sim_run<NU, 4>(L, T, m, Mg, Wg, mode, sel, smem, gscr, out, slot);
        case 8: return sim_run<NU, 8>(L, T, m, Mg, Wg, mode, sel, smem, gscr, out, slot);
        default: return sim_run<NU, 16>(L, T, m, Mg, Wg, mode, sel, smem, gscr, out, slot);
    }
}
static int sim_dispatch(int nu, int P, const MpcLayout &L, const MpcTables &T, int m, const double *Mg, const double *Wg,
                        int mode, int sel, double *smem, double *gscr, const MpcRunOut &out) {
    switch (nu) {
        case 1: return sim_p<1>(P, L, T, m, Mg, Wg, mode, sel, smem, gscr, out);
        case 2: return sim_p<2>(P, L, T, m, Mg, Wg, mode, sel, smem, gscr, out);
        case 3: return sim_p<3>(P, L, T, m, Mg, Wg, mode, sel, smem, gscr, out);
        default: return sim_p<4>(P, L, T, m, Mg, Wg, mode, sel, smem, gscr, out);
    }
}

template <int NU>
static int soft_p(int P, const MpcLayout &L, const MpcTables &T, int p, int m, const double *Mg, const double *Wg, int mode, int sel,
                  double *smem, const MpcRunOut &out, double *j0g, const MpcEst *E = nullptr, const double *delta = nullptr) {
    if (E) return soft_run<NU, 16, true>(L, T, p, m, Mg, Wg, mode, sel, smem, out, E, delta, j0g);   // validation run: P = 16 image only
    switch (P) {
        case 4: return soft_run<NU, 4>(L, T, p, m, Mg, Wg, mode, sel, smem, out, nullptr, nullptr, j0g);
        case 8: return soft_run<NU, 8>(L, T, p, m, Mg, Wg, mode, sel, smem, out, nullptr, nullptr, j0g);
        default: return soft_run<NU, 16>(L, T, p, m, Mg, Wg, mode, sel, smem, out, nullptr, nullptr, j0g);
    }
}
// one CTA = SOFT_THREADS host threads (plants with soft output constraints)
static int run_block(int nu, int P, const MpcLayout &L, const MpcTables &T, int p, int m, const double *Mg, const double *Wg,
                     int mode, int sel, const MpcRunOut &out, const MpcEst *E = nullptr, const double *delta = nullptr) {
    std::vector<double> smem(soft_smem_doubles(L, nu, P) + (E ? soft_est_doubles(L, nu, P, E->hlp) : 0) + 8, std::nan(""));
    std::vector<double> j0((size_t)(nu * P + 1) * (nu * P + 1), std::nan(""));   // the run's J0 scratch (global memory on the device)
    int status[SOFT_THREADS];
    simt_run_block([&]() {
        int st;
        switch (nu) {
            case 1: st = soft_p<1>(P, L, T, p, m, Mg, Wg, mode, sel, smem.data(), out, j0.data(), E, delta); break;
            case 2: st = soft_p<2>(P, L, T, p, m, Mg, Wg, mode, sel, smem.data(), out, j0.data(), E, delta); break;
            case 3: st = soft_p<3>(P, L, T, p, m, Mg, Wg, mode, sel, smem.data(), out, j0.data(), E, delta); break;
            default: st = soft_p<4>(P, L, T, p, m, Mg, Wg, mode, sel, smem.data(), out, j0.data(), E, delta); break;
        }
        status[threadIdx.x] = st;
    }, SOFT_THREADS);
    for (int l = 1; l < SOFT_THREADS; ++l)
        if (status[l] != status[0]) return 99;
    return status[0];
}

// one warp = 32 host threads
static int run_warp(int nu, int P, const MpcLayout &L, const MpcTables &T, int m, const double *Mg, const double *Wg, int mode,
                    int sel, const MpcRunOut &out) {
    std::vector<double> smem(std::max(sim_spec_smem_doubles(L, nu, P), sim_smem_doubles(L, nu, P)) + 8, std::nan(""));   // NaN-poison: catches reads of unwritten shared memory
    std::vector<double> gscr(sim_scratch_doubles(nu * P) + std::max(sim_slot_doubles(nu * P), sim_spec_slot_doubles(nu * P)) + 8, std::nan(""));
    int status[32];
    simt_run_warp([&]() {
        status[threadIdx.x] = sim_dispatch(nu, P, L, T, m, Mg, Wg, mode, sel, smem.data(), gscr.data(), out);
    });
    for (int l = 1; l < 32; ++l)
        if (status[l] != status[0]) return 99;  // non-uniform control flow
    return status[0];
}

static int *g_trace = nullptr;
extern "C" void emu_set_trace(int *p) { g_trace = p; }

extern "C" int emu_eval_batch(const mpcgpu_problem *pb, int n, const int *N, const int *Nu, const double *delta,
                              const double *lambda, int mode, double *cost, double *y, double *u, double *ys,
                              double *uopt, int *status, unsigned long long *counters, char *err, int errlen) {
    MpcHostTables ht;
    std::string e = mpc_build_tables(*pb, ht);
    if (!e.empty()) { std::strncpy(err, e.c_str(), errlen - 1); return 1; }
    const MpcLayout &L = ht.L;
    MpcTables T{ht.TG.data(), ht.TK.data(), ht.S1.data(), ht.r.data(), ht.v.data(), ht.yref.data(),
                ht.step.data(), ht.pa.data(), L.pmax + L.mmax + 2, ht.sig.data()};
    const int ny = L.ny, nu = L.nu, nit = L.nit;
    const bool square = ny == nu;
    for (int c = 0; c < n; ++c) {
        const int p = N[c], m = Nu[c], nz = nu * m;
        if (p < 2 || p > L.pmax || m < 1 || m > L.mmax || m >= p) {
            if (status) status[c] = 4;
            if (cost && mode == 1) for (int i = 0; i < ny; ++i) cost[(size_t)c * ny + i] = NAN;
            if (cost && mode == 2) cost[c] = NAN;
            continue;
        }
        const int P = (sim_spec_ok(L) && !(g_sim_knob & 512)) ? 16 : sim_pad(m), R = nu * P;   // the product runs one P = 16 image (knob 512: size buckets, plain kernel)
        std::vector<double> bsm(mpc_builder_smem_doubles(nz, L.nst));
        std::vector<double> Mg((size_t)L.nst * R), Wg((size_t)2 * R * R);
        int flag = 0;
        int st = mpc_build_candidate(L, T, p, m, P, delta + (size_t)c * ny, lambda + (size_t)c * nu, bsm.data(), Mg.data(),
                                     Wg.data(), &flag);
        double part[MPC_MAXY + 1];
        if (st == 0) {
            if (mode == 2) {
                const int runs = square ? ny : 1;
                double F = 0.0;
                for (int rn = 0; rn < runs; ++rn) {
                    MpcRunOut out{part, y ? y + (size_t)c * ny * nit : nullptr, u ? u + (size_t)c * nu * nit : nullptr,
                                  ys ? ys + (size_t)c * ny * nit : nullptr, uopt ? uopt + (size_t)c * nu * nit : nullptr,
                                  counters, nullptr, nullptr};
                    int s2 = L.has_ov_bounds ? run_block(nu, P, L, T, p, m, Mg.data(), Wg.data(), 2, square ? rn : -1, out)
                                             : run_warp(nu, P, L, T, m, Mg.data(), Wg.data(), 2, square ? rn : -1, out);
                    if (s2) st = s2;
                    F += part[0];
                }
                cost[c] = st ? NAN : F + (double)p;
            } else {
                MpcRunOut out{mode == 1 ? cost + (size_t)c * ny : nullptr, y ? y + (size_t)c * ny * nit : nullptr,
                              u ? u + (size_t)c * nu * nit : nullptr, ys ? ys + (size_t)c * ny * nit : nullptr,
                              uopt ? uopt + (size_t)c * nu * nit : nullptr, counters, nullptr, g_trace};
                st = L.has_ov_bounds ? run_block(nu, P, L, T, p, m, Mg.data(), Wg.data(), mode, -2, out)
                                     : run_warp(nu, P, L, T, m, Mg.data(), Wg.data(), mode, -2, out);
            }
        } else if (cost) {
            if (mode == 1) for (int i = 0; i < ny; ++i) cost[(size_t)c * ny + i] = NAN;
            if (mode == 2) cost[c] = NAN;
        }
        if (status) status[c] = st;
    }
    return 0;
}

// Validation run against a mismatched plant (mpcgpu_set_mismatch / k_soft<NU,16,true>) for ONE candidate: the real plant's
// channels (ny x nw row-major), the estimator gain and hl as in include/mpcgpu.h.  y: ny x nit, u: nu x nit, cost: ny (GAM).
extern "C" int emu_eval_est(const mpcgpu_problem *pb, const double *pa, const double *pb0, const double *pb1, const int *pd,
                            const double *gain, int hl, int N, int Nu, const double *delta, const double *lambda, double *cost,
                            double *y, double *u, char *err, int errlen) {
    MpcHostTables ht;
    std::string e = mpc_build_tables(*pb, ht);
    if (!e.empty()) { std::strncpy(err, e.c_str(), errlen - 1); return -1; }
    const MpcLayout &L = ht.L;
    MpcTables T{ht.TG.data(), ht.TK.data(), ht.S1.data(), ht.r.data(), ht.v.data(), ht.yref.data(),
                ht.step.data(), ht.pa.data(), L.pmax + L.mmax + 2, ht.sig.data()};
    const int nu = L.nu, P = 16, R = nu * P, nz = nu * Nu, nch = L.ny * L.nw;
    MpcEst E;
    std::memset(&E, 0, sizeof(E));
    int dmax = 0;
    for (int c = 0; c < nch; ++c) { E.a[c] = pa[c]; E.b0[c] = pb0[c]; E.b1[c] = pb1[c]; E.d[c] = pd[c]; dmax = std::max(dmax, pd[c]); }
    E.hl = hl; E.hlp = dmax + 2; E.gain = gain;
    std::vector<double> bsm(mpc_builder_smem_doubles(nz, L.nst));
    std::vector<double> Mg((size_t)L.nst * R), Wg((size_t)2 * R * R);
    int flag = 0;
    int st = mpc_build_candidate(L, T, N, Nu, P, delta, lambda, bsm.data(), Mg.data(), Wg.data(), &flag);
    if (st) return st;
    unsigned long long counters[2] = {0, 0};
    MpcRunOut out{cost, y, u, nullptr, nullptr, counters, nullptr, nullptr};
    return run_block(nu, P, L, T, N, Nu, Mg.data(), Wg.data(), 1, -2, out, &E, delta);
}

// the builder's output for one candidate from the SERIAL host build of mpc_core.cuh (emu_build.cpp runs the same source with threads)
extern "C" int emu_build_serial(const mpcgpu_problem *pb, int p, int m, int P, const double *delta, const double *lambda, double *Mg,
                                double *Wg, char *err, int errlen) {
    MpcHostTables ht;
    std::string e = mpc_build_tables(*pb, ht);
    if (!e.empty()) { std::strncpy(err, e.c_str(), errlen - 1); return -1; }
    const MpcLayout &L = ht.L;
    MpcTables T{ht.TG.data(), ht.TK.data(), ht.S1.data(), ht.r.data(), ht.v.data(), ht.yref.data(),
                ht.step.data(), ht.pa.data(), L.pmax + L.mmax + 2, ht.sig.data()};
    std::vector<double> bsm(mpc_builder_smem_doubles(L.nu * m, L.nst));
    int flag = 0;
    return mpc_build_candidate(L, T, p, m, P, delta, lambda, bsm.data(), Mg, Wg, &flag);
}

// shared-memory footprint of one closed-loop run (bytes) for padded control horizon P: occupancy bookkeeping
extern "C" long long emu_sim_smem_bytes(const mpcgpu_problem *pb, int P) {
    MpcHostTables ht;
    if (!mpc_build_tables(*pb, ht).empty()) return -1;
    return (long long)((sim_spec_ok(ht.L) ? sim_spec_smem_doubles(ht.L, ht.L.nu, P) : sim_smem_doubles(ht.L, ht.L.nu, P)) * sizeof(double));
}
