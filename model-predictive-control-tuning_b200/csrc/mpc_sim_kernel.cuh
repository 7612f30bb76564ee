// mpc_sim_kernel.cuh -- the closed-loop kernel entry (one warp per run) and the argument blocks it shares with
// the host side (mpcgpu.cu).  Instantiated per NU in mpc_sim_inst.cu (one translation unit per NU so that the
// build parallelises); mpcgpu.cu only takes function pointers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/mpcgpu.h"
#include "mpc_sim.cuh"
#include "mpc_sim_spec.cuh"
#include "mpc_soft.cuh"

struct DevCand {
    const int *N, *Nu;
    const double *delta, *lambda;
    const long long *offM, *offW;
    double *M, *W;
    int *bstatus;  // builder status per candidate
    double *scratch;  // per-run spill area of the closed-loop kernel (V / Li beyond QC)
    long long scratch_stride;
    double *slot;     // per-run parking slots of the two-phase mode (sim_slot_doubles), or nullptr
    long long slot_stride;
    const MpcEst *est;   // validation run against a mismatched plant (mpcgpu_set_mismatch), or nullptr
};

struct DevOut {
    double *cost;   // GAM: n*ny ; VNS: n
    double *part;   // VNS partial sums n*runs
    int *status;    // n
    int *status2;   // n: status of the re-run of candidates whose QP outgrew the speculative kernel's factor (SIM_ST_OVERFLOW)
    unsigned long long *counters;  // [0] constrained QPs [1] active-set iterations
    double *y, *u, *ys, *uopt;     // optional trajectories
    unsigned long long *diag;      // optional per-run diagnostics (4 per run)
};


// One warp per (candidate, run).  mode: 0 RAW, 1 GAM, 2 VNS.
// SPEC: lane-resident state + speculative unconstrained stretches (mpc_sim_spec.cuh; plants with nst <= 32).
template <int NU, int P, bool LEAN = false, bool VLEAN = false, bool SPEC = false, bool MSM = false>
#ifndef SIM_SPEC_MINB
#define SIM_SPEC_MINB 12   /* resident runs per SM the speculative kernel's register budget is sized for */
#endif
#ifndef SIM_MSM_MINB
#define SIM_MSM_MINB 5     /* ... and of its small-population image (M in shared memory: at most 5 runs per SM are launched) */
#endif
__global__ void __launch_bounds__(32, SPEC ? (MSM ? SIM_MSM_MINB : SIM_SPEC_MINB) : 1) k_sim(const MpcLayout L, const MpcTables T, const int *order, int count, int runs,
                                            int mode, int square, long long item0, DevCand C, DevOut O) {
    extern __shared__ __align__(16) double smem_s[];
    const int item = blockIdx.x;
    if (item >= count * runs) return;
    const int c = order[item / runs];
    const int run = item - (item / runs) * runs;
    const int m = C.Nu[c];
    const int ny = L.ny, nit = L.nit;
    // mode bit 8: second pass over the candidates the speculative kernel gave up on (more than QC active constraints)
    const bool redo = (mode & 256) != 0;
    mode &= 255;
    if (redo && O.status[c] != SIM_ST_OVERFLOW) return;
    int *stat = redo ? O.status2 : O.status;
    if (C.bstatus[c] != 0) {
        if ((threadIdx.x & 31) == 0) {
            atomicMax(stat + c, C.bstatus[c]);
            if (mode == 1) for (int i = 0; i < ny; ++i) O.cost[(size_t)c * ny + i] = NAN;
            if (mode == 2) O.part[(size_t)c * runs + run] = NAN;
        }
        return;
    }
    MpcRunOut out;
    out.cost = mode == 1 ? O.cost + (size_t)c * ny : (mode == 2 ? O.part + (size_t)c * runs + run : nullptr);
    out.y = O.y ? O.y + (size_t)c * ny * nit : nullptr;
    out.u = O.u ? O.u + (size_t)c * NU * nit : nullptr;
    out.ys = O.ys ? O.ys + (size_t)c * ny * nit : nullptr;
    out.uopt = O.uopt ? O.uopt + (size_t)c * NU * nit : nullptr;
    out.counters = O.counters;
    out.diag = O.diag ? O.diag + 4 * ((size_t)c * runs + run) : nullptr;
    out.trace = nullptr;
    const long long t_start = clock64();
    const int sel = mode == 2 ? (square ? run : -1) : -2;
    double *gscr = C.scratch ? C.scratch + (size_t)(item0 + item) * C.scratch_stride : nullptr;
    double *pslot = C.slot ? C.slot + (size_t)(item0 + item) * C.slot_stride : nullptr;
    const int st = SPEC ? sim_run_spec<NU, P, LEAN, VLEAN, MSM>(L, T, m, C.M + C.offM[c], C.W + C.offW[c], mode, sel, smem_s, gscr, out, pslot)
                        : sim_run<NU, P, LEAN, VLEAN>(L, T, m, C.M + C.offM[c], C.W + C.offW[c], mode, sel, smem_s, gscr, out, pslot);
    if (st != 0 && (threadIdx.x & 31) == 0) atomicMax(stat + c, st);
    if (out.diag && (threadIdx.x & 31) == 0) out.diag[3] = (unsigned long long)(clock64() - t_start);
}


// One CTA per (candidate, run): plants with soft output constraints (mpc_soft.cuh).
#ifndef SOFT_MINB
#define SOFT_MINB 3   /* resident CTAs per SM the register budget is sized for (no spills; 575 -> 550 ms on 2048 Shell7x5 candidates) */
#endif
template <int NU, int P, bool EST = false>
__global__ void __launch_bounds__(SOFT_THREADS, SOFT_MINB) k_soft(const MpcLayout L, const MpcTables T, const int *order, int count, int runs,
                                                       int mode, int square, long long item0, DevCand C, DevOut O) {
    extern __shared__ double smem_f[];
    const int item = blockIdx.x;
    if (item >= count * runs) return;
    const int c = order[item / runs];
    const int run = item - (item / runs) * runs;
    const int ny = L.ny, nit = L.nit;
    if (C.bstatus[c] != 0) {
        if (threadIdx.x == 0) {
            atomicMax(O.status + c, C.bstatus[c]);
            if (mode == 1) for (int i = 0; i < ny; ++i) O.cost[(size_t)c * ny + i] = NAN;
            if (mode == 2) O.part[(size_t)c * runs + run] = NAN;
        }
        return;
    }
    MpcRunOut out;
    out.cost = mode == 1 ? O.cost + (size_t)c * ny : (mode == 2 ? O.part + (size_t)c * runs + run : nullptr);
    out.y = O.y ? O.y + (size_t)c * ny * nit : nullptr;
    out.u = O.u ? O.u + (size_t)c * NU * nit : nullptr;
    out.ys = O.ys ? O.ys + (size_t)c * ny * nit : nullptr;
    out.uopt = O.uopt ? O.uopt + (size_t)c * NU * nit : nullptr;
    out.counters = O.counters;
    out.diag = O.diag ? O.diag + 4 * ((size_t)c * runs + run) : nullptr;
    out.trace = nullptr;
    const long long t_start = clock64();
    const int sel = mode == 2 ? (square ? run : -1) : -2;
    const int st = soft_run<NU, P, EST>(L, T, C.N[c], C.Nu[c], C.M + C.offM[c], C.W + C.offW[c], mode, sel, smem_f, out,
                                        EST ? C.est : nullptr, EST ? C.delta + (size_t)c * ny : nullptr,
                                        C.scratch ? C.scratch + (size_t)(item0 + item) * C.scratch_stride : nullptr);
    if (st != 0 && threadIdx.x == 0) atomicMax(O.status + c, st);
    if (out.diag && threadIdx.x == 0) out.diag[3] = (unsigned long long)(clock64() - t_start);
}

typedef void (*sim_kernel_t)(const MpcLayout, const MpcTables, const int *, int, int, int, int, long long, DevCand, DevOut);
sim_kernel_t sim_kernel_nu1(int P);
sim_kernel_t sim_kernel_nu2(int P);
sim_kernel_t sim_kernel_nu3(int P);
sim_kernel_t sim_kernel_nu4(int P);
sim_kernel_t sim_lean_nu1(int P);
sim_kernel_t sim_lean_nu2(int P);
sim_kernel_t sim_lean_nu3(int P);
sim_kernel_t sim_lean_nu4(int P);
sim_kernel_t sim_vlean_nu1(int P);
sim_kernel_t sim_vlean_nu2(int P);
sim_kernel_t sim_vlean_nu3(int P);
sim_kernel_t sim_vlean_nu4(int P);
sim_kernel_t sim_spec_nu1(int variant);   // variant 0 full, 1 GAM cost-only, 2 VNS cost-only, 3 GAM cost-only with M in shared memory; P = 16 image only
sim_kernel_t sim_spec_nu2(int variant);
sim_kernel_t sim_spec_nu3(int variant);
sim_kernel_t sim_spec_nu4(int variant);
sim_kernel_t soft_kernel_nu1(int P);
sim_kernel_t soft_kernel_nu2(int P);
sim_kernel_t soft_kernel_nu3(int P);
sim_kernel_t soft_kernel_nu4(int P);
sim_kernel_t soft_est_kernel_nu1();   // validation run against a mismatched plant: P = 16 image only
sim_kernel_t soft_est_kernel_nu2();
sim_kernel_t soft_est_kernel_nu3();
sim_kernel_t soft_est_kernel_nu4();
