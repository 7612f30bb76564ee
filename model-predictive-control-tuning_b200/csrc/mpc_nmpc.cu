// mpc_nmpc.cu -- nonlinear path (BASELINE.json configs[4]): batched closedloop_toolbox_nmpc for the Van de Vusse
// CSTR (closedloop_toolbox_nmpc.m:36-97, vandevusse_model.m:39-77), GAM / VNS costs fused (GAM_fun.m:110-115,
// VNS2.m:147-195 nonlinear branch).  C ABI: include/mpcgpu.h, NMPC section.
//
// Default kernel: k_nmpc_g<16> (mpc_nmpc_group.cuh), sixteen lanes per closed-loop run, two runs per warp.  k_nmpc below is the
// first mapping, one THREAD per run (MPCGPU_NMPC_THREAD_PER_RUN=1), kept for A/B and as the plain statement of the algorithm
// (it compiles mpc_nmpc_core.h, the source the CPU port shares).  Per controller call (nlmpcmove restated as N1-N4,
// DESIGN.md section 2 / include/mpcgpu.h):
//   rollout with forward sensitivities (RK4 stage Jacobians chained: [A|B] per sample, X = dx/dv carried),
//   Gauss-Newton model  H = sum S'Wy^2 S + D'Wdu^2 D,  g,  accumulated on the fly (S never stored),
//   exact box-constrained QP step (primal active set on the MV bounds, Cholesky of the free block),
//   backtracking on the true cost;  stop when the scaled step is < 1e-10.
// HBM traffic per run: 48 B in, 8-16 B out (+ trajectories on request): compute/latency bound, fp64.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"

#include "mpc_nmpc_group.cuh"   /* NmpcArgs, k_nmpc_g */
#ifdef NM_GLOBAL_WORK
#define NM_WORK_ALWAYS 1
#else
#define NM_WORK_ALWAYS 0
#endif


// mode 0 RAW, 1 GAM, 2 VNS.  One thread per (candidate, run).  `order`: candidates sorted by (Nu, N) on the host so that
// the 32 runs of a warp share their loop bounds (the population mixes horizons 2..31 x 1..15: unsorted, 5 of 32 lanes were
// active per instruction).  H and its factor are THREAD-LOCAL arrays with a fixed leading dimension: local memory is
// interleaved across the warp by the hardware, equal (row, col) -> one coalesced access, and it stays in L1/L2 (the first
// version kept them in a per-run global slab: 15 GB of DRAM writes per 16384 candidates).
__global__ void __launch_bounds__(NM_THREADS) k_nmpc(const NmpcDev D, int n, int runs, int mode, const int *order, NmpcArgs A) {
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n * runs) return;
    const int c = order[item / runs], run = item - (item / runs) * runs;
    const int p = A.N[c], m = A.Nu[c], nit = D.nit;
    if (p < 2 || p > D.pmax || m < 1 || m > D.mmax || m >= p) {
        A.status[c] = MPCGPU_CAND_INVALID;
        if (mode == 1) for (int j = 0; j < NY; ++j) A.cost[(size_t)c * NY + j] = NAN;
        if (mode == 2) A.part[(size_t)c * runs + run] = NAN;
        return;
    }
    const int sel = mode == 2 ? run : -1;   // VNS: only set-point `sel` is kept, the others are ZEROED (VNS2.m:150-155)
#ifdef NM_GLOBAL_WORK   /* A/B: the first version's per-run slab in global memory */
    double *H = A.work + (size_t)item * 2 * NM_LD * NM_LD, *Lc = H + NM_LD * NM_LD;
#else
    double H[NM_LD * NM_LD], Lc[NM_LD * NM_LD];
#endif
    unsigned n_calls = 0, n_sqp = 0;
    const int status = nmpc_run(D, p, m, mode, sel, A.delta + (size_t)c * NY, A.lambda + (size_t)c * NU, A.r, A.yref,
                                A.y ? A.y + (size_t)c * NY * nit : nullptr, A.u ? A.u + (size_t)c * NU * nit : nullptr,
                                A.yopt ? A.yopt + (size_t)c * NY * nit : nullptr, A.uopt ? A.uopt + (size_t)c * NU * nit : nullptr,
                                mode == 1 ? A.cost + (size_t)c * NY : (mode == 2 ? A.part + (size_t)c * runs + run : nullptr), H, Lc,
                                &n_calls, &n_sqp);
    if (status) atomicMax(A.status + c, status);
    atomicAdd(A.counters + 0, (unsigned long long)n_calls);
    atomicAdd(A.counters + 1, (unsigned long long)n_sqp);
}


__global__ void k_nmpc_finish(int n, int runs, const int *N, const double *part, const int *status, double *cost) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double acc = 0.0;
    for (int r = 0; r < runs; ++r) acc += part[(size_t)c * runs + r];
    cost[c] = (status[c] == 0 || status[c] == 5) ? acc + (double)N[c] : NAN;
}

// ------------------------------------------------------------------------------------------------
struct mpcgpu_nmpc_handle {
    int device = 0;
    std::string err;
    NmpcDev D;
    std::vector<double> r, yref;
    double *dR = nullptr, *dYref = nullptr;
    cudaStream_t stream = nullptr;
    mpcgpu_counters cnt = {};
};
static std::string g_nmpc_create_error;

extern "C" const char *mpcgpu_nmpc_last_error(mpcgpu_nmpc_handle *h) { return h ? h->err.c_str() : g_nmpc_create_error.c_str(); }

extern "C" void mpcgpu_nmpc_destroy(mpcgpu_nmpc_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    cudaFree(h->dR); cudaFree(h->dYref);
    delete h;
}

extern "C" int mpcgpu_nmpc_create(const mpcgpu_nmpc_problem *pb, int device, mpcgpu_nmpc_handle **out) {
    if (!pb || !out) { g_nmpc_create_error = "NULL argument"; return MPCGPU_ERR_ARG; }
    *out = nullptr;
    if (pb->model != MPCGPU_MODEL_VANDEVUSSE) { g_nmpc_create_error = "unknown model id"; return MPCGPU_ERR_UNSUPPORTED; }
    if (pb->nit < 2 || pb->pmax < 2 || pb->mmax < 1 || pb->mmax > NM_MAXM || pb->nsub < 1 || !(pb->Ts > 0) || !pb->x0 || !pb->u0 ||
        !pb->umin || !pb->umax || !pb->su || !pb->sy || !pb->r) {
        g_nmpc_create_error = "bad NMPC problem (nit >= 2, 1 <= mmax <= 15, nsub >= 1, Ts > 0, pointers set)";
        return MPCGPU_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_nmpc_create_error = std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (mpcgpu has no CPU fallback)";
        return MPCGPU_ERR_CUDA;
    }
    mpcgpu_nmpc_handle *h = new mpcgpu_nmpc_handle();
    if (device < 0) cudaGetDevice(&device);
    h->device = device;
    NmpcDev &D = h->D;
    D.nit = pb->nit; D.pmax = pb->pmax; D.mmax = pb->mmax; D.inK = pb->inK > 0 ? pb->inK : 10; D.nsub = pb->nsub;
    D.max_sqp = pb->max_sqp > 0 ? pb->max_sqp : 30; D.Ts = pb->Ts;
    for (int i = 0; i < NX; ++i) { D.x0[i] = pb->x0[i]; D.xmin[i] = pb->xmin ? pb->xmin[i] : -INFINITY; D.xmax[i] = pb->xmax ? pb->xmax[i] : INFINITY; }
    for (int j = 0; j < NU; ++j) { D.u0[j] = pb->u0[j]; D.umin[j] = pb->umin[j]; D.umax[j] = pb->umax[j]; D.su[j] = pb->su[j]; }
    for (int j = 0; j < NY; ++j) D.sy[j] = pb->sy[j];
    h->r.assign(pb->r, pb->r + (size_t)NY * pb->nit);
    if (pb->yref) h->yref.assign(pb->yref, pb->yref + (size_t)NY * pb->nit); else h->yref.assign((size_t)NY * pb->nit, 0.0);
    auto fail = [&](const char *what, cudaError_t c2) {
        g_nmpc_create_error = std::string(what) + ": " + cudaGetErrorString(c2);
        mpcgpu_nmpc_destroy(h);
        return MPCGPU_ERR_CUDA;
    };
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", ce);
    if ((ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", ce);
    const size_t nb = sizeof(double) * NY * pb->nit;
    if ((ce = cudaMalloc((void **)&h->dR, nb)) != cudaSuccess) return fail("alloc", ce);
    if ((ce = cudaMalloc((void **)&h->dYref, nb)) != cudaSuccess) return fail("alloc", ce);
    cudaMemcpy(h->dR, h->r.data(), nb, cudaMemcpyHostToDevice);
    if ((ce = cudaMemcpy(h->dYref, h->yref.data(), nb, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("copy", ce);
    cudaFuncSetAttribute(k_nmpc_g<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * nmg_doubles(NM_MAXZ) * 4));
    cudaFuncSetAttribute(k_nmpc_g<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * nmg_doubles(NM_MAXZ) * 2));
    cudaFuncSetAttribute(k_nmpc_g<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * nmg_doubles(NM_MAXZ)));
    *out = h;
    return MPCGPU_OK;
}

extern "C" int mpcgpu_nmpc_eval_batch(mpcgpu_nmpc_handle *h, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                                      const double *lambda, int cost_mode, const double *r_override, double *cost, double *y,
                                      double *u, double *yopt, double *uopt, int32_t *status) {
    if (!h) return MPCGPU_ERR_ARG;
    if (n < 0 || cost_mode < 0 || cost_mode > 2 || (n > 0 && (!N || !Nu || !delta || !lambda))) { h->err = "bad arguments"; return MPCGPU_ERR_ARG; }
    if (n == 0) return MPCGPU_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return MPCGPU_ERR_CUDA; }
    const int nit = h->D.nit;
    const int runs = cost_mode == MPCGPU_COST_VNS ? NY : 1;     // square plant: one run per output (VNS2.m:148-165)
    const bool traj = y || u || yopt || uopt || cost_mode == MPCGPU_COST_RAW;
    cudaStream_t s = h->stream;
    const size_t nI = (size_t)n * 4, nT = (size_t)n * 2 * nit;
    // default: sixteen lanes per run, two runs per warp (mpc_nmpc_group.cuh); MPCGPU_NMPC_GROUP=8|32 for A/B (measured, 2048 / 16384
    // candidates: G = 8 286 / 545 ms, G = 16 231 / 510 ms, G = 32 237 / -- ms); MPCGPU_NMPC_THREAD_PER_RUN=1: one thread per run (1.6 s)
    const bool grouped = getenv("MPCGPU_NMPC_THREAD_PER_RUN") == nullptr;
    const int G = getenv("MPCGPU_NMPC_GROUP") ? atoi(getenv("MPCGPU_NMPC_GROUP")) : 16;
    const size_t nD = (size_t)n * (2 * NY + 2 * NU) + (size_t)n * runs + (traj ? 4 * nT : 0) + (size_t)NY * nit +
                      (NM_WORK_ALWAYS ? (size_t)n * runs * 2 * NM_MAXZ * NM_MAXZ : 0) + 8;
    int *dI = nullptr; double *dD = nullptr;
    if (cudaMalloc((void **)&dI, sizeof(int) * nI) != cudaSuccess || cudaMalloc((void **)&dD, sizeof(double) * nD) != cudaSuccess) {
        cudaFree(dI); h->err = "cudaMalloc failed"; return MPCGPU_ERR_CUDA;
    }
    int *dN = dI, *dNu = dN + n, *dSt = dNu + n, *dOrd = dSt + n;
    // candidates binned by horizons, longest first: the runs of a warp then share their loop trip counts (thread per run: by
    // (Nu, N), the box-QP dominates its divergence; grouped: by (N, Nu), the rollout length is what the groups of a warp share)
    std::vector<int> order(n);
    for (int c = 0; c < n; ++c) order[c] = c;
    if (grouped) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return N[a] != N[b] ? N[a] > N[b] : Nu[a] > Nu[b]; });
    int maxnu = 1;
    for (int c = 0; c < n; ++c) if (Nu[c] >= 1 && Nu[c] <= NM_MAXM && Nu[c] > maxnu) maxnu = Nu[c];
    double *q_ = dD;
    double *dDl = q_; q_ += (size_t)n * NY; double *dLm = q_; q_ += (size_t)n * NU;
    double *dCost = q_; q_ += (size_t)n * NY; double *dPart = q_; q_ += (size_t)n * runs;
    double *dY = nullptr, *dU = nullptr, *dYo = nullptr, *dUo = nullptr;
    if (traj) { dY = q_; q_ += nT; dU = q_; q_ += nT; dYo = q_; q_ += nT; dUo = q_; q_ += nT; }
    double *dRo = q_; q_ += (size_t)NY * nit;
    unsigned long long *dCnt = (unsigned long long *)q_; q_ += 2;
    double *dWork = q_;
    int rc = MPCGPU_OK;
    auto ck = [&](cudaError_t e2) { if (e2 != cudaSuccess && rc == MPCGPU_OK) { h->err = cudaGetErrorString(e2); rc = MPCGPU_ERR_CUDA; } };
    ck(cudaMemcpyAsync(dN, N, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dNu, Nu, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dOrd, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dDl, delta, sizeof(double) * n * NY, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dLm, lambda, sizeof(double) * n * NU, cudaMemcpyHostToDevice, s));
    if (r_override) ck(cudaMemcpyAsync(dRo, r_override, sizeof(double) * NY * nit, cudaMemcpyHostToDevice, s));
    ck(cudaMemsetAsync(dSt, 0, sizeof(int) * n, s));
    ck(cudaMemsetAsync(dCnt, 0, sizeof(unsigned long long) * 2, s));
    if (traj) ck(cudaMemsetAsync(dY, 0, sizeof(double) * 4 * nT, s));
    if (rc == MPCGPU_OK) {
        NmpcArgs A{dN, dNu, dDl, dLm, r_override ? dRo : h->dR, h->dYref, dCost, dPart, dY, dU, dYo, dUo, dSt, dCnt, dWork};
        const int items = n * runs;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        if (grouped) {
            // shared memory for the largest plan of THIS population (9.6 KB per run at Nu = 15).  Binning the population by
            // control horizon into launches with less shared memory each (more resident warps) was measured and dropped:
            // 810 vs 800 ms for 16384 candidates -- the time is the latency of the runs, not their residency (gpurun_out/nmpc3.log)
            const int rpw = G == 32 ? 1 : (G == 8 ? 4 : 2);   // runs per warp
            const int maxz = NU * maxnu;
            const size_t smem = sizeof(double) * nmg_doubles(maxz) * rpw;
            const int grid = (items + rpw - 1) / rpw;
            if (rpw == 1) k_nmpc_g<32><<<grid, 32, smem, s>>>(h->D, 0, items, runs, cost_mode, dOrd, A, maxz);
            else if (rpw == 2) k_nmpc_g<16><<<grid, 32, smem, s>>>(h->D, 0, items, runs, cost_mode, dOrd, A, maxz);
            else k_nmpc_g<8><<<grid, 32, smem, s>>>(h->D, 0, items, runs, cost_mode, dOrd, A, maxz);
        } else {
            k_nmpc<<<(items + NM_THREADS - 1) / NM_THREADS, NM_THREADS, 0, s>>>(h->D, n, runs, cost_mode, dOrd, A);
        }
        ck(cudaGetLastError());
        if (cost_mode == MPCGPU_COST_VNS) {
            k_nmpc_finish<<<(n + 127) / 128, 128, 0, s>>>(n, runs, dN, dPart, dSt, dCost);
            ck(cudaGetLastError());
        }
        cudaEventRecord(e1, s);
        unsigned long long cnt[2] = {0, 0};
        if (cost && cost_mode != MPCGPU_COST_RAW)
            ck(cudaMemcpyAsync(cost, dCost, sizeof(double) * (cost_mode == MPCGPU_COST_GAM ? (size_t)n * NY : (size_t)n), cudaMemcpyDeviceToHost, s));
        if (y) ck(cudaMemcpyAsync(y, dY, sizeof(double) * nT, cudaMemcpyDeviceToHost, s));
        if (u) ck(cudaMemcpyAsync(u, dU, sizeof(double) * nT, cudaMemcpyDeviceToHost, s));
        if (yopt) ck(cudaMemcpyAsync(yopt, dYo, sizeof(double) * nT, cudaMemcpyDeviceToHost, s));
        if (uopt) ck(cudaMemcpyAsync(uopt, dUo, sizeof(double) * nT, cudaMemcpyDeviceToHost, s));
        if (status) ck(cudaMemcpyAsync(status, dSt, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
        ck(cudaMemcpyAsync(cnt, dCnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
        ck(cudaStreamSynchronize(s));
        h->cnt.candidates += n; h->cnt.closed_loops += (uint64_t)n * runs; h->cnt.qp_solves += cnt[0]; h->cnt.as_iterations += cnt[1];
        h->cnt.kernel_launches += cost_mode == MPCGPU_COST_VNS ? 2 : 1;
        float ms = 0.f;
        if (rc == MPCGPU_OK && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) { h->cnt.last_sim_ms = ms; h->cnt.last_total_ms = ms; }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(dI); cudaFree(dD);
    return rc;
}

extern "C" int mpcgpu_nmpc_get_counters(mpcgpu_nmpc_handle *h, mpcgpu_counters *out) {
    if (!h || !out) return MPCGPU_ERR_ARG;
    *out = h->cnt;
    return MPCGPU_OK;
}
