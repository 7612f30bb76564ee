// mpcgpu.cu -- kernels + C ABI of libmpcgpu.so (include/mpcgpu.h).  sm_100a only, no CPU fallback.
//
// Launch structure of one population evaluation (mpcgpu_run):
//   candidates are bucketed on the host by padded control horizon P = 4 / 8 / 16 >= m (it fixes the row
//   count R = nu*P and with it the shared-memory footprint), largest first, each bucket on its own stream:
//     k_build      : one CTA (128 threads) per candidate -> M (nst x R), W = H^-1 (R x R) in HBM
//     k_sim<NU,P>  : one warp per closed-loop run (mpc_sim.cuh)
//   k_finish  : VNS only, F = sum_runs partial + N
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"
#include "mpc_core.cuh"
#include "mpc_sim_kernel.cuh"
#include "mpc_tables.h"

#define NSTREAM 4  /* >= number of size buckets (P = 4, 8, 16): every bucket runs concurrently */
#define BUILD_THREADS 128

static std::string g_create_error;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof(b_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            h->err = b_;                                                                           \
            return MPCGPU_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BUILD_THREADS) k_build(const MpcLayout L, const MpcTables T, const int *order,
                                                         int count, int P, DevCand C) {
    extern __shared__ double smem_b[];
    __shared__ int flag;
    const int item = blockIdx.x;
    if (item >= count) return;
    const int c = order[item];
    const int p = C.N[c], m = C.Nu[c];
    const int st = mpc_build_candidate(L, T, p, m, P, C.delta + (size_t)c * L.ny, C.lambda + (size_t)c * L.nu, smem_b,
                                       C.M + C.offM[c], C.W + C.offW[c], &flag);
    if (threadIdx.x == 0) C.bstatus[c] = st;
}

static sim_kernel_t soft_kernel(int nu, int P) {
    switch (nu) {
        case 1: return soft_kernel_nu1(P);
        case 2: return soft_kernel_nu2(P);
        case 3: return soft_kernel_nu3(P);
        case 4: return soft_kernel_nu4(P);
    }
    return nullptr;
}
static sim_kernel_t soft_est_kernel(int nu) {
    switch (nu) {
        case 1: return soft_est_kernel_nu1();
        case 2: return soft_est_kernel_nu2();
        case 3: return soft_est_kernel_nu3();
        case 4: return soft_est_kernel_nu4();
    }
    return nullptr;
}
static sim_kernel_t sim_lean(int nu, int P) {
    switch (nu) {
        case 1: return sim_lean_nu1(P);
        case 2: return sim_lean_nu2(P);
        case 3: return sim_lean_nu3(P);
        case 4: return sim_lean_nu4(P);
    }
    return nullptr;
}
static sim_kernel_t sim_vlean(int nu, int P) {
    switch (nu) {
        case 1: return sim_vlean_nu1(P);
        case 2: return sim_vlean_nu2(P);
        case 3: return sim_vlean_nu3(P);
        case 4: return sim_vlean_nu4(P);
    }
    return nullptr;
}
static sim_kernel_t sim_spec(int nu, int variant) {
    switch (nu) {
        case 1: return sim_spec_nu1(variant);
        case 2: return sim_spec_nu2(variant);
        case 3: return sim_spec_nu3(variant);
        case 4: return sim_spec_nu4(variant);
    }
    return nullptr;
}
static sim_kernel_t sim_kernel(int nu, int P) {
    switch (nu) {
        case 1: return sim_kernel_nu1(P);
        case 2: return sim_kernel_nu2(P);
        case 3: return sim_kernel_nu3(P);
        case 4: return sim_kernel_nu4(P);
    }
    return nullptr;
}

__global__ void k_finish_vns(int n, int runs, const int *N, const double *part, const int *status, double *cost) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double acc = 0.0;
    for (int r = 0; r < runs; ++r) acc += part[(size_t)c * runs + r];  // fixed order: deterministic
    cost[c] = status[c] ? NAN : acc + (double)N[c];
}

// candidates re-run on the kernel with the spill area take the status of that run
__global__ void k_merge_status(int n, int *status, const int *status2) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && status[c] == SIM_ST_OVERFLOW) status[c] = status2[c];
}

__global__ void k_mark_invalid(int n, const int *invalid, int ny, int mode, int *status, double *cost) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n || !invalid[c]) return;
    status[c] = MPCGPU_CAND_INVALID;
    if (mode == 1) for (int i = 0; i < ny; ++i) cost[(size_t)c * ny + i] = NAN;
    if (mode == 2) cost[c] = NAN;
}

// fp64 FMA peak: 8 independent chains per thread
__global__ void k_fp64_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------
template <typename Tp>
struct DBuf {
    Tp *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 16;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(Tp));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename Tp>
struct HBuf {  // pinned host staging
    Tp *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 16;
        cudaError_t e = cudaMallocHost((void **)&p, want * sizeof(Tp));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct mpcgpu_handle {
    int device = 0;
    std::string err;
    MpcHostTables ht;
    cudaStream_t stream = nullptr;          // own stream for eval_batch / upload / download
    cudaStream_t pool[NSTREAM] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[NSTREAM] = {}, ev_t0 = nullptr, ev_t1 = nullptr, ev_t2 = nullptr;
    DBuf<double> dTG, dTK, dS1, dR, dV, dYref, dST, dPA, dSig;
    // population
    int n = 0;
    bool uploaded = false, ran = false;
    int last_mode = -1, last_traj = 0;
    std::vector<int> hN, hNu, hOrder, hInvalid;
    std::vector<long long> hOffM, hOffW;
    struct Bucket { int P, mmax, off, count; };
    std::vector<Bucket> buckets;
    int n_valid = 0;
    DBuf<int> dN, dNu, dOrder, dInvalid, dBStatus, dStatus, dStatus2;
    DBuf<long long> dOffM, dOffW;
    DBuf<double> dDelta, dLambda, dM, dW, dCost, dPart, dY, dU, dYs, dUopt, dScratch;
    DBuf<unsigned long long> dCounters, dDiag;
    bool want_diag = false;
    int last_runs = 1;
    HBuf<double> pinD, pinOut;
    HBuf<int> pinI;
    mpcgpu_counters cnt = {};
    size_t smem_optin = 0;
    int sm_count = 148;
    int opt_vns_legality = 0;   // MPCGPU_OPT_VNS_LEGALITY
    // validation run against a mismatched plant (mpcgpu_set_mismatch): device copy of MpcEst + the gain, or nullptr
    MpcEst *dEst = nullptr;
    double *dEstGain = nullptr;
    int est_hlp = 0;
};

static MpcTables dev_tables(mpcgpu_handle *h) {
    MpcTables T;
    T.TG = h->dTG.p; T.TK = h->dTK.p; T.S1 = h->dS1.p; T.r = h->dR.p; T.v = h->dV.p; T.yref = h->dYref.p;
    T.ST = h->dST.p; T.PA = h->dPA.p; T.st_stride = h->ht.L.pmax + h->ht.L.mmax + 2;
    T.sig = h->dSig.p;
    return T;
}

static int upload_signals(mpcgpu_handle *h) {
    const MpcHostTables &t = h->ht;
    CK(h->dR.ensure(t.r.size()));
    CK(h->dV.ensure(t.v.size()));
    CK(h->dYref.ensure(t.yref.size()));
    CK(h->dSig.ensure(t.sig.size()));
    CK(cudaMemcpyAsync(h->dSig.p, t.sig.data(), t.sig.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->dR.p, t.r.data(), t.r.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->dV.p, t.v.data(), t.v.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->dYref.p, t.yref.data(), t.yref.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return MPCGPU_OK;
}

extern "C" int mpcgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" const char *mpcgpu_last_error(mpcgpu_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int mpcgpu_create(const mpcgpu_problem *problem, int device, mpcgpu_handle **out) {
    if (!problem || !out) { g_create_error = "NULL argument"; return MPCGPU_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (mpcgpu has no CPU fallback)";
        return MPCGPU_ERR_CUDA;
    }
    mpcgpu_handle *h = new mpcgpu_handle();
    if (device < 0) cudaGetDevice(&device);
    h->device = device;
    std::string e = mpc_build_tables(*problem, h->ht);
    if (!e.empty()) { g_create_error = e; delete h; return MPCGPU_ERR_ARG; }
    auto fail = [&](const char *what, cudaError_t ce2) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce2);
        mpcgpu_destroy(h);
        return MPCGPU_ERR_CUDA;
    };
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", ce);
    cudaDeviceProp prop;
    if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", ce);
    if (prop.major < 10) {
        g_create_error = "mpcgpu is built for sm_100a (B200) only; found compute capability " + std::to_string(prop.major) +
                         "." + std::to_string(prop.minor);
        mpcgpu_destroy(h);
        return MPCGPU_ERR_CUDA;
    }
    h->smem_optin = prop.sharedMemPerBlockOptin;
    h->sm_count = prop.multiProcessorCount;
    if ((ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", ce);
    for (int i = 0; i < NSTREAM; ++i) {
        if ((ce = cudaStreamCreateWithFlags(&h->pool[i], cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", ce);
        if ((ce = cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming)) != cudaSuccess) return fail("event", ce);
    }
    if ((ce = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail("event", ce);
    if ((ce = cudaEventCreate(&h->ev_t0)) != cudaSuccess) return fail("event", ce);
    if ((ce = cudaEventCreate(&h->ev_t1)) != cudaSuccess) return fail("event", ce);
    if ((ce = cudaEventCreate(&h->ev_t2)) != cudaSuccess) return fail("event", ce);
    const MpcHostTables &t = h->ht;
    if ((ce = h->dTG.ensure(t.TG.size())) != cudaSuccess) return fail("alloc TG", ce);
    if ((ce = h->dTK.ensure(t.TK.size())) != cudaSuccess) return fail("alloc TK", ce);
    if ((ce = h->dS1.ensure(t.S1.size())) != cudaSuccess) return fail("alloc S1", ce);
    if ((ce = h->dST.ensure(t.step.size())) != cudaSuccess) return fail("alloc ST", ce);
    if ((ce = h->dPA.ensure(t.pa.size())) != cudaSuccess) return fail("alloc PA", ce);
    if ((ce = cudaMemcpy(h->dST.p, t.step.data(), t.step.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("copy step table", ce);
    if ((ce = cudaMemcpy(h->dPA.p, t.pa.data(), t.pa.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("copy pole table", ce);
    if ((ce = cudaMemcpy(h->dTG.p, t.TG.data(), t.TG.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("copy TG", ce);
    if ((ce = cudaMemcpy(h->dTK.p, t.TK.data(), t.TK.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("copy TK", ce);
    if ((ce = cudaMemcpy(h->dS1.p, t.S1.data(), t.S1.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("copy S1", ce);
    if (upload_signals(h) != MPCGPU_OK) { g_create_error = h->err; mpcgpu_destroy(h); return MPCGPU_ERR_CUDA; }
    if ((ce = h->dCounters.ensure(4)) != cudaSuccess) return fail("alloc counters", ce);
    // allow large dynamic shared memory on both kernels
    // (k_build also has a few bytes of static shared memory: dynamic + static must stay within the opt-in limit)
    if ((ce = cudaFuncSetAttribute(k_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin - 1024)) != cudaSuccess)
        return fail("cudaFuncSetAttribute(k_build)", ce);
    for (int P = 4; P <= 16; P *= 2) {
        for (sim_kernel_t kf : {sim_kernel(t.L.nu, P), sim_lean(t.L.nu, P), sim_vlean(t.L.nu, P), soft_kernel(t.L.nu, P)})
            if ((ce = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin)) != cudaSuccess)
                return fail("cudaFuncSetAttribute(closed-loop kernel)", ce);
    }
    for (int v = 0; v < 4; ++v)
        if ((ce = cudaFuncSetAttribute(sim_spec(t.L.nu, v), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin)) != cudaSuccess)
            return fail("cudaFuncSetAttribute(k_sim spec)", ce);
    if ((ce = cudaFuncSetAttribute(soft_est_kernel(t.L.nu), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin)) != cudaSuccess)
        return fail("cudaFuncSetAttribute(k_soft est)", ce);
    // check the largest footprints fit
    const size_t sb = mpc_builder_smem_doubles(t.L.nu * t.L.mmax, t.L.nst) * sizeof(double);
    const size_t ss = (sim_spec_ok(t.L) ? sim_spec_smem_doubles(t.L, t.L.nu, 16) : sim_smem_doubles(t.L, t.L.nu, sim_pad(t.L.mmax))) * sizeof(double);
    if (sb > h->smem_optin - 1024 || ss > h->smem_optin) {
        g_create_error = "problem too large for shared memory (builder " + std::to_string(sb) + " B, sim " +
                         std::to_string(ss) + " B)";
        mpcgpu_destroy(h);
        return MPCGPU_ERR_UNSUPPORTED;
    }
    if (t.L.has_ov_bounds) {   // soft output constraints run on the block-per-run kernel (mpc_soft.cuh)
        const size_t sf = soft_smem_doubles(t.L, t.L.nu, sim_pad(t.L.mmax)) * sizeof(double);
        if (sf > h->smem_optin || t.L.nst > SOFT_THREADS || !(t.L.rho_ecr > 0.0) || (SOFT_THREADS / t.L.ny) * SOFT_TCH < t.L.pmax) {
            g_create_error = "soft output constraints: problem too large for the block kernel (shared memory " +
                             std::to_string(sf) + " B, nst " + std::to_string(t.L.nst) + ") or Weights.ECR <= 0";
            mpcgpu_destroy(h);
            return MPCGPU_ERR_UNSUPPORTED;
        }
    }
    *out = h;
    return MPCGPU_OK;
}

extern "C" void mpcgpu_destroy(mpcgpu_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->dEst) cudaFree(h->dEst);
    if (h->dEstGain) cudaFree(h->dEstGain);
    h->dTG.release(); h->dTK.release(); h->dS1.release(); h->dST.release(); h->dPA.release(); h->dR.release(); h->dV.release(); h->dYref.release(); h->dSig.release();
    h->dN.release(); h->dNu.release(); h->dOrder.release(); h->dInvalid.release(); h->dBStatus.release();
    h->dStatus.release(); h->dStatus2.release(); h->dOffM.release(); h->dOffW.release(); h->dDelta.release(); h->dLambda.release();
    h->dM.release(); h->dW.release(); h->dCost.release(); h->dPart.release(); h->dY.release(); h->dU.release();
    h->dYs.release(); h->dUopt.release(); h->dCounters.release(); h->dScratch.release(); h->dDiag.release();
    h->pinD.release(); h->pinOut.release(); h->pinI.release();
    for (int i = 0; i < NSTREAM; ++i) {
        if (h->pool[i]) cudaStreamDestroy(h->pool[i]);
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    if (h->ev_t2) cudaEventDestroy(h->ev_t2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int mpcgpu_set_signals(mpcgpu_handle *h, int nit, const double *r, const double *v, const double *yref) {
    if (!h) return MPCGPU_ERR_ARG;
    CK(cudaSetDevice(h->device));
    std::vector<double> keep;
    if (!yref && nit == h->ht.L.nit && (int)h->ht.yref.size() == nit * h->ht.L.ny) { keep = h->ht.yref; yref = keep.data(); }   // NULL: keep Par.Yref
    std::string e = mpc_set_signals(h->ht, nit, r, v, yref);
    if (!e.empty()) { h->err = e; return MPCGPU_ERR_ARG; }
    return upload_signals(h);
}

// A-priori work key of a candidate: what launches are ordered by (heaviest first) and what both multi-GPU front ends deal
// shards by.  Hard MV limits: the number of moves and how hard the controller pushes against the limits (large delta / small
// lambda), plus the short-horizon family that limit-cycles (N barely beyond the longest dead time: a QP at every sample).
// Soft output bands (block-per-run kernel): the iteration count grows with the moves per prediction row -- a control horizon
// close to the prediction horizon makes the band QPs degenerate -- and with small move weights; fitted on per-run cycle counters
// of 2048 Shell7x5 candidates (tools/diag_runs.py DIAG_CASE=shell7x5: rank correlation 0.94, list-schedule makespan 556 -> 524 ms).
static double mpc_work_key(int ny, int nu, bool soft, int dead_max, int N, int Nu, const double *delta, const double *lambda) {
    double dmax = 0.0, lmin = 1e300;
    for (int i = 0; i < ny; ++i) dmax = std::max(dmax, std::fabs(delta[i]));
    for (int j = 0; j < nu; ++j) lmin = std::min(lmin, std::fabs(lambda[j]));
    if (soft) return 0.15 * Nu + 1.15 * (double)Nu / (double)N - 0.09 * std::log10(lmin + 1e-300);
    return std::log10(dmax / (lmin + 1e-300) + 1e-300) + 0.15 * Nu + (N <= dead_max + 2 ? 2.0 : 0.0);
}

extern "C" int mpcgpu_upload(mpcgpu_handle *h, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                             const double *lambda) {
    if (!h) return MPCGPU_ERR_ARG;
    if (n < 0 || (n > 0 && (!N || !Nu || !delta || !lambda))) { h->err = "bad population arguments"; return MPCGPU_ERR_ARG; }
    CK(cudaSetDevice(h->device));
    // a previous run (possibly on a caller stream) may still be reading the population buffers this call overwrites
    if (h->ran) CK(cudaEventSynchronize(h->ev_t2));
    const MpcLayout &L = h->ht.L;
    const int ny = L.ny, nu = L.nu;
    h->n = n; h->uploaded = false; h->ran = false;
    h->hN.assign(N, N + n); h->hNu.assign(Nu, Nu + n);
    h->hInvalid.assign(n, 0);
    h->hOffM.assign(n, 0); h->hOffW.assign(n, 0);
    // bucket by padded horizon P (4/8/16), largest first; illegal horizons never reach a kernel
    std::vector<std::vector<int>> by_p(3);
    int mmax_p[3] = {0, 0, 0};
    long long offM = 0, offW = 0;
    h->n_valid = 0;
    for (int c = 0; c < n; ++c) {
        const int p = N[c], m = Nu[c];
        bool illegal = p < 2 || p > L.pmax || m < 1 || m > L.mmax || m >= p;
        if (!illegal && h->opt_vns_legality) {   // VNS2.m:135  any(N<=dmin) | any(Nu<=1)   (PreCon.m:23: min(N) > max(Nu))
            if (m <= 1) illegal = true;
            for (int i = 0; i < ny && !illegal; ++i) illegal = p <= h->ht.dmin[i];
        }
        if (illegal) { h->hInvalid[c] = 1; continue; }
        // One kernel image for the whole population: three concurrently running instantiations
        // (P = 4, 8, 16, ~240 KB of SASS each) thrashed the instruction cache -- measured 16.3 ms against 10.8 ms for
        // 4096 Shell3x3 candidates with everything on the P = 16 image (Shell7x5 / k_soft: 455 -> 418 ms).
        // MPCGPU_SIZE_BUCKETS=1 restores the buckets.
        static const bool size_buckets = getenv("MPCGPU_SIZE_BUCKETS") != nullptr;
        const int P = size_buckets ? sim_pad(m) : sim_pad(L.mmax), b = P == 4 ? 0 : (P == 8 ? 1 : 2);
        by_p[b].push_back(c);
        if (m > mmax_p[b]) mmax_p[b] = m;
        const long long R = (long long)nu * P;
        h->hOffM[c] = offM; offM += (long long)L.nst * R;
        h->hOffW[c] = offW; offW += 2 * R * R;
        h->n_valid++;
    }
    h->hOrder.clear(); h->buckets.clear();
    for (int b = 2; b >= 0; --b) {
        if (by_p[b].empty()) continue;
        // heaviest first, so that the longest closed loops start at t = 0: work grows with the number of
        // moves and with how hard the controller pushes against the MV limits (large delta / small lambda)
        std::vector<double> score(n, 0.0);
        int dead_max = 0;
        for (int ch = 0; ch < L.ny * L.nw; ++ch) dead_max = std::max(dead_max, (int)L.d[ch]);
        for (int c : by_p[b]) score[c] = mpc_work_key(ny, nu, L.has_ov_bounds != 0, dead_max, N[c], Nu[c], delta + (size_t)c * ny, lambda + (size_t)c * nu);
        std::stable_sort(by_p[b].begin(), by_p[b].end(), [&](int x, int y) { return score[x] > score[y]; });
        mpcgpu_handle::Bucket bk{4 << b, mmax_p[b], (int)h->hOrder.size(), (int)by_p[b].size()};
        h->hOrder.insert(h->hOrder.end(), by_p[b].begin(), by_p[b].end());
        h->buckets.push_back(bk);
    }
    const size_t nn = (size_t)(n > 0 ? n : 1);
    CK(h->dN.ensure(nn)); CK(h->dNu.ensure(nn)); CK(h->dOrder.ensure(nn)); CK(h->dInvalid.ensure(nn));
    CK(h->dBStatus.ensure(nn)); CK(h->dStatus.ensure(nn)); CK(h->dStatus2.ensure(nn)); CK(h->dOffM.ensure(nn)); CK(h->dOffW.ensure(nn));
    CK(h->dDelta.ensure(nn * ny)); CK(h->dLambda.ensure(nn * nu));
    CK(h->dM.ensure((size_t)offM + 1)); CK(h->dW.ensure((size_t)offW + 1));
    // stage through pinned memory: [delta | lambda] doubles, [N | Nu | order | invalid] ints, offsets
    CK(h->pinD.ensure(nn * (ny + nu) + 2 * nn));
    CK(h->pinI.ensure(4 * nn));
    std::memcpy(h->pinD.p, delta, sizeof(double) * (size_t)n * ny);
    std::memcpy(h->pinD.p + (size_t)n * ny, lambda, sizeof(double) * (size_t)n * nu);
    long long *poff = reinterpret_cast<long long *>(h->pinD.p + (size_t)n * (ny + nu));
    std::memcpy(poff, h->hOffM.data(), sizeof(long long) * n);
    std::memcpy(poff + n, h->hOffW.data(), sizeof(long long) * n);
    std::memcpy(h->pinI.p, N, sizeof(int) * n);
    std::memcpy(h->pinI.p + n, Nu, sizeof(int) * n);
    std::memcpy(h->pinI.p + 2 * (size_t)n, h->hOrder.data(), sizeof(int) * h->hOrder.size());
    std::memcpy(h->pinI.p + 3 * (size_t)n, h->hInvalid.data(), sizeof(int) * n);
    cudaStream_t s = h->stream;
    if (n > 0) {
        CK(cudaMemcpyAsync(h->dDelta.p, h->pinD.p, sizeof(double) * (size_t)n * ny, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(h->dLambda.p, h->pinD.p + (size_t)n * ny, sizeof(double) * (size_t)n * nu, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(h->dOffM.p, poff, sizeof(long long) * n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(h->dOffW.p, poff + n, sizeof(long long) * n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(h->dN.p, h->pinI.p, sizeof(int) * n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(h->dNu.p, h->pinI.p + n, sizeof(int) * n, cudaMemcpyHostToDevice, s));
        if (!h->hOrder.empty())
            CK(cudaMemcpyAsync(h->dOrder.p, h->pinI.p + 2 * (size_t)n, sizeof(int) * h->hOrder.size(), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(h->dInvalid.p, h->pinI.p + 3 * (size_t)n, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    }
    CK(cudaStreamSynchronize(s));  // pinned staging is reused by the next upload
    h->uploaded = true;
    return MPCGPU_OK;
}

extern "C" int mpcgpu_run(mpcgpu_handle *h, int cost_mode, int want_traj, void *cuda_stream) {
    if (!h) return MPCGPU_ERR_ARG;
    if (!h->uploaded) { h->err = "mpcgpu_run before mpcgpu_upload"; return MPCGPU_ERR_STATE; }
    if (cost_mode < 0 || cost_mode > 2) { h->err = "bad cost_mode"; return MPCGPU_ERR_ARG; }
    CK(cudaSetDevice(h->device));
    const MpcLayout &L = h->ht.L;
    const int n = h->n, ny = L.ny, nu = L.nu, nit = L.nit;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    const int square = (ny == nu) ? 1 : 0;
    const int runs = (cost_mode == MPCGPU_COST_VNS && square) ? ny : 1;
    if (cost_mode == MPCGPU_COST_RAW) want_traj = 1;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    CK(h->dCost.ensure(nn * ny));
    CK(h->dPart.ensure(nn * runs));
    if (want_traj) {
        CK(h->dY.ensure(nn * ny * nit)); CK(h->dYs.ensure(nn * ny * nit));
        CK(h->dU.ensure(nn * nu * nit)); CK(h->dUopt.ensure(nn * nu * nit));
    }
    // spill area: only rows beyond QC (= 16 active constraints) of the P = 16 bucket ever touch it
    long long scr_stride = 0, scr_items = 0;
    for (const auto &bk : h->buckets) {
        // hard-limit kernels: spill area of the active-set factor; block-per-run kernel (soft bands / validation run): J0 of the run
        const long long sd = (L.has_ov_bounds || h->dEst) ? (long long)(nu * bk.P + 1) * (nu * bk.P + 1) : (long long)sim_scratch_doubles(nu * bk.P);
        if (sd > scr_stride) scr_stride = sd;
        scr_items += (long long)bk.count * runs;
    }
    long long slot_stride = 0;
    for (const auto &bk : h->buckets) slot_stride = std::max(slot_stride, (long long)std::max(sim_slot_doubles(nu * bk.P), sim_spec_slot_doubles(nu * bk.P)));
    // the block kernel of the soft-constraint path has no such mode; MPCGPU_TWO_PHASE=0 switches it off (A/B runs)
    const char *tp_env = getenv("MPCGPU_TWO_PHASE");
    const bool two_phase = !L.has_ov_bounds && !(tp_env && atoi(tp_env) == 0);
    // one allocation: [spill areas | parking slots]
    if (scr_stride > 0 || two_phase) CK(h->dScratch.ensure((size_t)((scr_stride + (two_phase ? slot_stride : 0)) * scr_items + 1)));
    DevCand C{h->dN.p, h->dNu.p, h->dDelta.p, h->dLambda.p, h->dOffM.p, h->dOffW.p, h->dM.p, h->dW.p, h->dBStatus.p,
              scr_stride > 0 ? h->dScratch.p : nullptr, scr_stride,
              two_phase ? h->dScratch.p + scr_stride * scr_items : nullptr, slot_stride, h->dEst};
    DevOut O{h->dCost.p, h->dPart.p, h->dStatus.p, h->dStatus2.p, h->dCounters.p,
             want_traj ? h->dY.p : nullptr, want_traj ? h->dU.p : nullptr, want_traj ? h->dYs.p : nullptr,
             want_traj ? h->dUopt.p : nullptr, nullptr};
    if (h->want_diag) {
        CK(h->dDiag.ensure(4 * nn * runs));
        CK(cudaMemsetAsync(h->dDiag.p, 0, sizeof(unsigned long long) * 4 * nn * runs, s));
        O.diag = h->dDiag.p;
    }
    h->last_runs = runs;
    const MpcTables T = dev_tables(h);
    if (h->ran) CK(cudaStreamWaitEvent(s, h->ev_t2, 0));   // two runs in flight would race on status / counters / cost
    CK(cudaEventRecord(h->ev_t0, s));
    CK(cudaMemsetAsync(h->dStatus.p, 0, sizeof(int) * nn, s));
    CK(cudaMemsetAsync(h->dStatus2.p, 0, sizeof(int) * nn, s));
    CK(cudaMemsetAsync(h->dCounters.p, 0, sizeof(unsigned long long) * 4, s));
    if (want_traj) {  // rows a run does not own (VNS on a square plant) and invalid candidates read as zero
        CK(cudaMemsetAsync(h->dY.p, 0, sizeof(double) * nn * ny * nit, s));
        CK(cudaMemsetAsync(h->dYs.p, 0, sizeof(double) * nn * ny * nit, s));
        CK(cudaMemsetAsync(h->dU.p, 0, sizeof(double) * nn * nu * nit, s));
        CK(cudaMemsetAsync(h->dUopt.p, 0, sizeof(double) * nn * nu * nit, s));
    }
    uint64_t launches = 0;
    // ---- phase 1: builders, forked over the stream pool ----
    const int nb = (int)h->buckets.size();
    const int nfork = nb < NSTREAM ? nb : NSTREAM;
    CK(cudaEventRecord(h->ev_fork, s));
    for (int i = 0; i < nfork; ++i) CK(cudaStreamWaitEvent(h->pool[i], h->ev_fork, 0));
    for (int b = 0; b < nb; ++b) {
        const auto &bk = h->buckets[b];
        const size_t smem = mpc_builder_smem_doubles(nu * bk.mmax, L.nst) * sizeof(double);
        k_build<<<bk.count, BUILD_THREADS, smem, h->pool[b % NSTREAM]>>>(L, T, h->dOrder.p + bk.off, bk.count, bk.P, C);
        CK(cudaGetLastError());
        launches++;
    }
    for (int i = 0; i < nfork; ++i) {
        CK(cudaEventRecord(h->ev_join[i], h->pool[i]));
        CK(cudaStreamWaitEvent(s, h->ev_join[i], 0));
    }
    CK(cudaEventRecord(h->ev_t1, s));
    // ---- phase 2: closed loops ----
    CK(cudaEventRecord(h->ev_fork, s));
    for (int i = 0; i < nfork; ++i) CK(cudaStreamWaitEvent(h->pool[i], h->ev_fork, 0));
    long long item0 = 0;
    for (int b = 0; b < nb; ++b) {
        const auto &bk = h->buckets[b];
        const int grid = bk.count * runs;
        if (h->dEst) {
            // validation run: the block-per-run kernel with the estimator, whatever the plant's constraint set
            if (bk.P != 16 || cost_mode == MPCGPU_COST_VNS) { h->err = "mismatch validation run: GAM / RAW only, single P = 16 bucket (unset MPCGPU_SIZE_BUCKETS)"; return MPCGPU_ERR_UNSUPPORTED; }
            const size_t smem = (soft_smem_doubles(L, nu, 16) + soft_est_doubles(L, nu, 16, h->est_hlp)) * sizeof(double);
            if (smem > h->smem_optin) { h->err = "mismatch validation run: shared memory"; return MPCGPU_ERR_UNSUPPORTED; }
            soft_est_kernel(nu)<<<grid, SOFT_THREADS, smem, h->pool[b % NSTREAM]>>>(L, T, h->dOrder.p + bk.off, bk.count, runs,
                                                                                cost_mode, square, item0, C, O);
        } else if (L.has_ov_bounds) {
            const size_t smem = soft_smem_doubles(L, nu, bk.P) * sizeof(double);
            soft_kernel(nu, bk.P)<<<grid, SOFT_THREADS, smem, h->pool[b % NSTREAM]>>>(L, T, h->dOrder.p + bk.off, bk.count, runs,
                                                                                   cost_mode, square, item0, C, O);
        } else {
            // plants whose deviation state fits one warp run the speculative kernel (MPCGPU_SPEC=0: the plain one, A/B runs)
            static const bool spec_off = getenv("MPCGPU_SPEC") && atoi(getenv("MPCGPU_SPEC")) == 0;
            const bool spec = sim_spec_ok(L) && bk.P == 16 && !spec_off;
            const bool lean = cost_mode == MPCGPU_COST_GAM && !want_traj && !h->want_diag;   // the tuning loop's call
            const bool vlean = cost_mode == MPCGPU_COST_VNS && !want_traj && !h->want_diag;   // ... and its VNS phase
            size_t smem = (spec ? sim_spec_smem_doubles(L, nu, bk.P, lean) : sim_smem_doubles(L, nu, bk.P)) * sizeof(double);
            bool msm = false;
            {   // Resident runs per SM.  The kernel is bound by instruction fetch: co-resident runs slow
                // each other down, so a SMALL population, whose step lasts as long as its heaviest run, finishes sooner with
                // fewer runs per SM (measured on 4096 Shell3x3 candidates: 10.4 ms at 10 per SM, 9.1 ms at 5), while a large
                // one wants them all (16384: 21.0 ms at 10, 30.7 ms at 5).  MPCGPU_RUNS_PER_SM overrides the choice.
                static const int rps_env = getenv("MPCGPU_RUNS_PER_SM") ? atoi(getenv("MPCGPU_RUNS_PER_SM")) : 0;
                const int rps = rps_env > 0 ? rps_env : ((long long)grid <= 40LL * h->sm_count ? 5 : 0);
                if (rps > 0) { const size_t want = (size_t)(228 * 1024) / rps - 1024; if (want > smem && want <= h->smem_optin) smem = want; }
                // ... and such a launch has the shared memory to spare for M itself (GAM cost-only image; MPCGPU_MSM=0: off)
                static const bool msm_off = getenv("MPCGPU_MSM") && atoi(getenv("MPCGPU_MSM")) == 0;
                const size_t need = (((sim_spec_smem_doubles(L, nu, bk.P, true) + 1) & ~(size_t)1) + sim_spec_msm_doubles(nu, bk.P)) * sizeof(double);
                msm = spec && lean && rps > 0 && !msm_off && need <= smem;
            }
            (spec ? sim_spec(nu, msm ? 3 : (lean ? 1 : (vlean ? 2 : 0)))
                  : (lean ? sim_lean(nu, bk.P) : (vlean ? sim_vlean(nu, bk.P) : sim_kernel(nu, bk.P))))<<<grid, 32, smem, h->pool[b % NSTREAM]>>>(L, T, h->dOrder.p + bk.off, bk.count, runs, cost_mode,
                                                                           square, item0, C, O);
            if (spec) {
                // second pass, same stream: the (rare, ~1 %) candidates whose QP needs more active constraints than the
                // speculative kernel keeps in shared memory run again on the kernel with the spill area; every other
                // CTA of this launch exits on its first instruction
                CK(cudaGetLastError());
                const size_t smem2 = sim_smem_doubles(L, nu, bk.P) * sizeof(double);
                (lean ? sim_lean(nu, bk.P) : (vlean ? sim_vlean(nu, bk.P) : sim_kernel(nu, bk.P)))<<<grid, 32, smem2, h->pool[b % NSTREAM]>>>(
                    L, T, h->dOrder.p + bk.off, bk.count, runs, cost_mode | 256, square, item0, C, O);
                launches++;
            }
        }
        CK(cudaGetLastError());
        item0 += grid;
        launches++;
    }
    for (int i = 0; i < nfork; ++i) {
        CK(cudaEventRecord(h->ev_join[i], h->pool[i]));
        CK(cudaStreamWaitEvent(s, h->ev_join[i], 0));
    }
    if (!L.has_ov_bounds && n > 0) {
        k_merge_status<<<(n + 127) / 128, 128, 0, s>>>(n, h->dStatus.p, h->dStatus2.p);
        launches++;
    }
    if (cost_mode == MPCGPU_COST_VNS && n > 0) {
        k_finish_vns<<<(n + 127) / 128, 128, 0, s>>>(n, runs, h->dN.p, h->dPart.p, h->dStatus.p, h->dCost.p);
        launches++;
    }
    if (h->n_valid < n && n > 0) {
        k_mark_invalid<<<(n + 127) / 128, 128, 0, s>>>(n, h->dInvalid.p, ny, cost_mode, h->dStatus.p, h->dCost.p);
        launches++;
    }
    CK(cudaEventRecord(h->ev_t2, s));
    CK(cudaGetLastError());
    h->ran = true; h->last_mode = cost_mode; h->last_traj = want_traj;
    h->cnt.kernel_launches += launches;
    h->cnt.candidates += (uint64_t)h->n_valid;
    h->cnt.closed_loops += (uint64_t)h->n_valid * runs;
    const uint64_t ol = (cost_mode != MPCGPU_COST_GAM || want_traj) ? 1 : 0;
    h->cnt.qp_solves += (uint64_t)h->n_valid * runs * ((uint64_t)nit + ol);
    return MPCGPU_OK;
}

extern "C" int mpcgpu_download(mpcgpu_handle *h, int cost_mode, double *cost, double *y, double *u, double *ys,
                               double *uopt, int32_t *status) {
    if (!h) return MPCGPU_ERR_ARG;
    if (!h->ran || cost_mode != h->last_mode) { h->err = "mpcgpu_download without a matching mpcgpu_run"; return MPCGPU_ERR_STATE; }
    if ((y || u || ys || uopt) && !h->last_traj) { h->err = "trajectories were not requested in mpcgpu_run"; return MPCGPU_ERR_STATE; }
    CK(cudaSetDevice(h->device));
    const MpcLayout &L = h->ht.L;
    const int n = h->n, ny = L.ny, nu = L.nu, nit = L.nit;
    cudaStream_t s = h->stream;
    // order the copies after the run even when it was issued on a caller stream
    CK(cudaStreamWaitEvent(s, h->ev_t2, 0));
    const size_t ncost = cost_mode == MPCGPU_COST_GAM ? (size_t)n * ny : (cost_mode == MPCGPU_COST_VNS ? (size_t)n : 0);
    CK(h->pinOut.ensure(ncost + 8));
    CK(h->pinI.ensure(4 * (size_t)(n > 0 ? n : 1)));
    unsigned long long *pc = reinterpret_cast<unsigned long long *>(h->pinOut.p + ncost);
    if (cost && ncost) CK(cudaMemcpyAsync(h->pinOut.p, h->dCost.p, sizeof(double) * ncost, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(pc, h->dCounters.p, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, s));
    if (status && n) CK(cudaMemcpyAsync(h->pinI.p, h->dStatus.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    if (y) CK(cudaMemcpyAsync(y, h->dY.p, sizeof(double) * (size_t)n * ny * nit, cudaMemcpyDeviceToHost, s));
    if (ys) CK(cudaMemcpyAsync(ys, h->dYs.p, sizeof(double) * (size_t)n * ny * nit, cudaMemcpyDeviceToHost, s));
    if (u) CK(cudaMemcpyAsync(u, h->dU.p, sizeof(double) * (size_t)n * nu * nit, cudaMemcpyDeviceToHost, s));
    if (uopt) CK(cudaMemcpyAsync(uopt, h->dUopt.p, sizeof(double) * (size_t)n * nu * nit, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (cost && ncost) std::memcpy(cost, h->pinOut.p, sizeof(double) * ncost);
    if (status && n) std::memcpy(status, h->pinI.p, sizeof(int) * n);
    h->cnt.qp_constrained += pc[0];
    h->cnt.as_iterations += pc[1];
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1) == cudaSuccess) h->cnt.last_build_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev_t1, h->ev_t2) == cudaSuccess) h->cnt.last_sim_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t2) == cudaSuccess) h->cnt.last_total_ms = ms;
    return MPCGPU_OK;
}

extern "C" int mpcgpu_eval_batch(mpcgpu_handle *h, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                                 const double *lambda, int cost_mode, double *cost, double *y, double *u, double *ys,
                                 double *uopt, int32_t *status) {
    if (!h) return MPCGPU_ERR_ARG;
    int rc = mpcgpu_upload(h, n, N, Nu, delta, lambda);
    if (rc) return rc;
    const int want_traj = (y || u || ys || uopt) ? 1 : 0;
    rc = mpcgpu_run(h, cost_mode, want_traj, nullptr);
    if (rc) return rc;
    return mpcgpu_download(h, cost_mode, cost, y, u, ys, uopt, status);
}

extern "C" int mpcgpu_cost_device_ptr(mpcgpu_handle *h, int cost_mode, void **ptr, int *count) {
    if (!h || !ptr || !count) return MPCGPU_ERR_ARG;
    if (!h->ran || cost_mode != h->last_mode) { h->err = "no matching run"; return MPCGPU_ERR_STATE; }
    *ptr = h->dCost.p;
    *count = cost_mode == MPCGPU_COST_GAM ? h->n * h->ht.L.ny : h->n;
    return MPCGPU_OK;
}

extern "C" int mpcgpu_get_counters(mpcgpu_handle *h, mpcgpu_counters *out) {
    if (!h || !out) return MPCGPU_ERR_ARG;
    *out = h->cnt;
    return MPCGPU_OK;
}

extern "C" int mpcgpu_set_option(mpcgpu_handle *h, int option, int value) {
    if (!h) return MPCGPU_ERR_ARG;
    switch (option) {
        case MPCGPU_OPT_VNS_LEGALITY: h->opt_vns_legality = value != 0; return MPCGPU_OK;
    }
    h->err = "unknown option";
    return MPCGPU_ERR_ARG;
}

// Plant-model mismatch validation (Shell3x3.m:271-286  options.Model = plant; sim(mpc, nit, r, [], options)): from this call on
// every evaluation simulates the controller -- with its state estimator, gain given in the state order of mpcgpu/estimator.py
// E1 -- against the real plant (a, b0, b1, d: ny x nw row-major like the model's).  plant_a == NULL switches it off again.
extern "C" int mpcgpu_set_mismatch(mpcgpu_handle *h, const double *plant_a, const double *plant_b0, const double *plant_b1,
                                   const int32_t *plant_d, const double *gain, int hl) {
    if (!h) return MPCGPU_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (h->ran) CK(cudaEventSynchronize(h->ev_t2));
    if (h->dEst) { cudaFree(h->dEst); h->dEst = nullptr; }
    if (h->dEstGain) { cudaFree(h->dEstGain); h->dEstGain = nullptr; }
    if (!plant_a) return MPCGPU_OK;
    const MpcLayout &L = h->ht.L;
    if (!plant_b0 || !plant_b1 || !plant_d || !gain || hl < 1) { h->err = "mpcgpu_set_mismatch: NULL argument / hl < 1"; return MPCGPU_ERR_ARG; }
    const int nch = L.ny * L.nw;
    if (L.nst > SOFT_THREADS || (SOFT_THREADS / L.ny) * SOFT_TCH < L.pmax) { h->err = "mpcgpu_set_mismatch: problem too large for the block-per-run kernel"; return MPCGPU_ERR_UNSUPPORTED; }
    MpcEst e;
    memset(&e, 0, sizeof(e));
    int dmax = 0;
    for (int c = 0; c < nch; ++c) {
        e.a[c] = plant_a[c]; e.b0[c] = plant_b0[c]; e.b1[c] = plant_b1[c]; e.d[c] = plant_d[c];
        if (plant_d[c] < 0 || (plant_d[c] == 0 && plant_b0[c] != 0.0)) { h->err = "mpcgpu_set_mismatch: bad plant channel (d >= 0, b0 == 0 when d == 0)"; return MPCGPU_ERR_ARG; }
        dmax = std::max(dmax, (int)plant_d[c]);
    }
    e.hl = hl; e.hlp = dmax + 2;
    const size_t ng = (size_t)(nch + L.nu * hl + L.ny) * L.ny;
    CK(cudaMalloc((void **)&h->dEstGain, sizeof(double) * ng));
    CK(cudaMemcpy(h->dEstGain, gain, sizeof(double) * ng, cudaMemcpyHostToDevice));
    e.gain = h->dEstGain;
    CK(cudaMalloc((void **)&h->dEst, sizeof(MpcEst)));
    CK(cudaMemcpy(h->dEst, &e, sizeof(MpcEst), cudaMemcpyHostToDevice));
    h->est_hlp = e.hlp;
    return MPCGPU_OK;
}

// closedloop_toolbox.m:1 as ONE call: the signals of the call are installed for this evaluation only and the handle's
// own (Par.Xsp, Par.mdv, Par.Yref, nit) are put back afterwards -- the reference never touches Par in closedloop_toolbox,
// and VNS2 interleaves these calls with GAM_fun on the same Par.
extern "C" int mpcgpu_closedloop(mpcgpu_handle *h, int nit, const double *r, const double *v, int32_t N, int32_t Nu,
                                 const double *delta, const double *lambda, double *y, double *u, double *ys, double *uopt,
                                 int32_t *status) {
    if (!h || !r || !delta || !lambda) return MPCGPU_ERR_ARG;
    const std::vector<double> r0 = h->ht.r, v0 = h->ht.v, y0 = h->ht.yref;
    const int nit0 = h->ht.L.nit;
    int rc = mpcgpu_set_signals(h, nit, r, v, nullptr);
    if (rc == MPCGPU_OK) {
        if (nit != nit0) { /* yref was zero-filled: irrelevant for RAW */ }
        rc = mpcgpu_eval_batch(h, 1, &N, &Nu, delta, lambda, MPCGPU_COST_RAW, nullptr, y, u, ys, uopt, status);
    }
    const std::string err = h->err;
    const int rc2 = mpcgpu_set_signals(h, nit0, r0.data(), h->ht.L.nd > 0 ? v0.data() : nullptr, y0.data());
    if (rc != MPCGPU_OK) { h->err = err; return rc; }
    return rc2;
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU evaluator: one handle per device of ONE process, candidates dealt by estimated work (sorted round-robin,
// SURVEY.md 8e), every device runs asynchronously, fitness gathered into the caller's host arrays in population order.
// ------------------------------------------------------------------------------------------------
struct mpcgpu_multi {
    std::vector<mpcgpu_handle *> hs;
    std::string err;
    std::vector<std::vector<int>> shard;
    std::vector<std::vector<int32_t>> sN, sNu, sStatus;
    std::vector<std::vector<double>> sDelta, sLambda, sCost;
};

extern "C" int mpcgpu_create_multi(const mpcgpu_problem *problem, const int *devices, int ndev, mpcgpu_multi **out) {
    if (!problem || !out || ndev < 1) { g_create_error = "NULL argument / ndev < 1"; return MPCGPU_ERR_ARG; }
    *out = nullptr;
    mpcgpu_multi *m = new mpcgpu_multi();
    for (int d = 0; d < ndev; ++d) {
        mpcgpu_handle *h = nullptr;
        const int rc = mpcgpu_create(problem, devices ? devices[d] : d, &h);
        if (rc != MPCGPU_OK) { for (auto *x : m->hs) mpcgpu_destroy(x); delete m; return rc; }
        m->hs.push_back(h);
    }
    m->shard.resize(ndev); m->sN.resize(ndev); m->sNu.resize(ndev); m->sStatus.resize(ndev);
    m->sDelta.resize(ndev); m->sLambda.resize(ndev); m->sCost.resize(ndev);
    *out = m;
    return MPCGPU_OK;
}
extern "C" void mpcgpu_destroy_multi(mpcgpu_multi *m) {
    if (!m) return;
    for (auto *h : m->hs) mpcgpu_destroy(h);
    delete m;
}
extern "C" int mpcgpu_multi_device_count(mpcgpu_multi *m) { return m ? (int)m->hs.size() : 0; }
extern "C" const char *mpcgpu_multi_last_error(mpcgpu_multi *m) { return m ? m->err.c_str() : g_create_error.c_str(); }
extern "C" int mpcgpu_multi_set_option(mpcgpu_multi *m, int option, int value) {
    if (!m) return MPCGPU_ERR_ARG;
    for (auto *h : m->hs) { const int rc = mpcgpu_set_option(h, option, value); if (rc) { m->err = h->err; return rc; } }
    return MPCGPU_OK;
}
extern "C" int mpcgpu_multi_set_signals(mpcgpu_multi *m, int nit, const double *r, const double *v, const double *yref) {
    if (!m) return MPCGPU_ERR_ARG;
    for (auto *h : m->hs) { const int rc = mpcgpu_set_signals(h, nit, r, v, yref); if (rc) { m->err = h->err; return rc; } }
    return MPCGPU_OK;
}
// The a-priori work estimate the shards are balanced on (also exported: the multi-process path -- one rank per GPU,
// bench.py / mpcgpu.distributed -- deals with the same key).
extern "C" int mpcgpu_work_estimate(const mpcgpu_problem *pb, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                                    const double *lambda, double *work) {
    if (!pb || !N || !Nu || !delta || !lambda || !work) return MPCGPU_ERR_ARG;
    int dead_max = 0;
    for (int ch = 0; ch < pb->ny * (pb->nu + pb->nd); ++ch) dead_max = std::max(dead_max, (int)pb->d[ch]);
    bool soft = false;
    for (int i = 0; i < pb->ny; ++i) soft = soft || (pb->ymin && std::isfinite(pb->ymin[i])) || (pb->ymax && std::isfinite(pb->ymax[i]));
    for (int c = 0; c < n; ++c)
        work[c] = mpc_work_key(pb->ny, pb->nu, soft, dead_max, N[c], Nu[c], delta + (size_t)c * pb->ny, lambda + (size_t)c * pb->nu);
    return MPCGPU_OK;
}
extern "C" int mpcgpu_multi_eval_batch(mpcgpu_multi *m, int n, const int32_t *N, const int32_t *Nu, const double *delta,
                                       const double *lambda, int cost_mode, double *cost, int32_t *status) {
    if (!m) return MPCGPU_ERR_ARG;
    if (cost_mode != MPCGPU_COST_GAM && cost_mode != MPCGPU_COST_VNS) { m->err = "multi-GPU evaluation returns costs only (GAM or VNS)"; return MPCGPU_ERR_ARG; }
    if (n < 0 || (n > 0 && (!N || !Nu || !delta || !lambda || !cost))) { m->err = "bad population arguments"; return MPCGPU_ERR_ARG; }
    const int G = (int)m->hs.size();
    const MpcLayout &L = m->hs[0]->ht.L;
    const int ny = L.ny, nu = L.nu, width = cost_mode == MPCGPU_COST_GAM ? ny : 1;
    // deal by work: sorted, round-robin
    std::vector<double> work(n);
    {
        int dead_max = 0;
        for (int ch = 0; ch < ny * L.nw; ++ch) dead_max = std::max(dead_max, (int)L.d[ch]);
        for (int c = 0; c < n; ++c)
            work[c] = mpc_work_key(ny, nu, L.has_ov_bounds != 0, dead_max, N[c], Nu[c], delta + (size_t)c * ny, lambda + (size_t)c * nu);
    }
    std::vector<int> order(n);
    for (int c = 0; c < n; ++c) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return work[a] > work[b]; });
    for (int g = 0; g < G; ++g) { m->shard[g].clear(); }
    for (int k = 0; k < n; ++k) m->shard[k % G].push_back(order[k]);
    // phase 1: upload + launch on every device (mpcgpu_run is asynchronous)
    for (int g = 0; g < G; ++g) {
        const auto &sh = m->shard[g];
        const int ns = (int)sh.size();
        m->sN[g].resize(ns); m->sNu[g].resize(ns); m->sDelta[g].resize((size_t)ns * ny); m->sLambda[g].resize((size_t)ns * nu);
        for (int k = 0; k < ns; ++k) {
            const int c = sh[k];
            m->sN[g][k] = N[c]; m->sNu[g][k] = Nu[c];
            std::memcpy(&m->sDelta[g][(size_t)k * ny], delta + (size_t)c * ny, sizeof(double) * ny);
            std::memcpy(&m->sLambda[g][(size_t)k * nu], lambda + (size_t)c * nu, sizeof(double) * nu);
        }
        int rc = mpcgpu_upload(m->hs[g], ns, m->sN[g].data(), m->sNu[g].data(), m->sDelta[g].data(), m->sLambda[g].data());
        if (rc == MPCGPU_OK) rc = mpcgpu_run(m->hs[g], cost_mode, 0, nullptr);
        if (rc != MPCGPU_OK) { m->err = m->hs[g]->err; return rc; }
    }
    // phase 2: gather (the per-generation "all-gather" of a single-process caller: host arrays in population order)
    for (int g = 0; g < G; ++g) {
        const auto &sh = m->shard[g];
        const int ns = (int)sh.size();
        m->sCost[g].resize((size_t)ns * width + 1); m->sStatus[g].resize(ns + 1);
        const int rc = mpcgpu_download(m->hs[g], cost_mode, m->sCost[g].data(), nullptr, nullptr, nullptr, nullptr, m->sStatus[g].data());
        if (rc != MPCGPU_OK) { m->err = m->hs[g]->err; return rc; }
        for (int k = 0; k < ns; ++k) {
            const int c = sh[k];
            std::memcpy(cost + (size_t)c * width, &m->sCost[g][(size_t)k * width], sizeof(double) * width);
            if (status) status[c] = m->sStatus[g][k];
        }
    }
    return MPCGPU_OK;
}
extern "C" int mpcgpu_multi_get_counters(mpcgpu_multi *m, int device_index, mpcgpu_counters *out) {
    if (!m || !out || device_index < 0 || device_index >= (int)m->hs.size()) return MPCGPU_ERR_ARG;
    return mpcgpu_get_counters(m->hs[device_index], out);
}

extern "C" int mpcgpu_measure_fp64_peak(int device, double *tflops) {
    if (!tflops) return MPCGPU_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MPCGPU_ERR_CUDA;
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return MPCGPU_ERR_CUDA;
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return MPCGPU_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double *d = nullptr;
    if (cudaMalloc((void **)&d, sizeof(double) * blocks * threads) != cudaSuccess) return MPCGPU_ERR_CUDA;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k_fp64_peak<<<blocks, threads>>>(d, iters);
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { cudaFree(d); return MPCGPU_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return MPCGPU_OK;
}

// Internal diagnostics (not part of the public ABI): per-run {constrained QPs, active-set iterations,
// largest active set, SM clock cycles} of the last run; enable before mpcgpu_run.
extern "C" int mpcgpu_debug_enable_diag(mpcgpu_handle *h, int on) { if (!h) return 1; h->want_diag = on != 0; return 0; }
extern "C" int mpcgpu_debug_get_diag(mpcgpu_handle *h, unsigned long long *out) {
    if (!h || !out || !h->ran || !h->dDiag.p) return 1;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, h->dDiag.p, sizeof(unsigned long long) * 4 * (size_t)h->n * h->last_runs, cudaMemcpyDeviceToHost));
    return 0;
}
