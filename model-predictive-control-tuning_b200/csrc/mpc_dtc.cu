// mpc_dtc.cu -- DTC-GPC batched sweep (BASELINE.json configs[3]; include/mpcgpu.h, DTC section).
//
// One warp per candidate (p[ny], m[nu], delta[ny], lambda[nu], Fr per output), 4 candidates per CTA.
//   phase 1 (gain, DTC_GPC_WW.m:98-105): S1 = H'QH + W from the step-response table (H is never
//     materialised: H[(i,r),(j,k)] = s_ij(dmin_i + 1 + r - k), MatG.m:64-67), warp Cholesky in shared
//     memory, only the nu rows of K = S1^-1 H'Q the loop uses, and their products with the candidate-
//     independent free-response tables:  Kr = Km*[1 blocks], KHp = Km*Hp, KS = Km*S
//     so that the control law of :143-149,  dU = Km*(Ref - Hp*up - S*Yd),  is three short dot products.
//   phase 2 (loop, DTC_GPC_WW.m:128-164 + OptimalPredictor2.m:24-40): process, disturbance, model and
//     fast-model channels advance one sample per step (the reference re-simulates each from t = 0 with
//     lsim every step; a causal LTI system gives the same samples), lanes own channels / outputs / inputs.
// All fp64.  HBM traffic per candidate: 16*(ny+nu) + 8*ny*(2*MAXF+1) B in, 8*ny B out (+ trajectories on
// request); the tables (a few KB) stay in L1/L2.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mpc_dtc.h"

#include "mpc_dtc_kernel.cuh"


// ------------------------------------------------------------------------------------------------
struct mpcgpu_dtc_handle {
    int device = 0;
    std::string err;
    DtcHostTables ht;
    cudaStream_t stream = nullptr;
    double *dStep = nullptr, *dF = nullptr, *dUg = nullptr, *dR = nullptr, *dQ = nullptr;
    size_t smem_optin = 0;
    mpcgpu_counters cnt = {};
};
static std::string g_dtc_create_error;

#define DCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof(b_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            h->err = b_;                                                                           \
            return MPCGPU_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

extern "C" const char *mpcgpu_dtc_last_error(mpcgpu_dtc_handle *h) { return h ? h->err.c_str() : g_dtc_create_error.c_str(); }

extern "C" void mpcgpu_dtc_destroy(mpcgpu_dtc_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    cudaFree(h->dStep); cudaFree(h->dF); cudaFree(h->dUg); cudaFree(h->dR); cudaFree(h->dQ);
    delete h;
}

extern "C" int mpcgpu_dtc_create(const mpcgpu_dtc_problem *problem, int device, mpcgpu_dtc_handle **out) {
    if (!problem || !out) { g_dtc_create_error = "NULL argument"; return MPCGPU_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_dtc_create_error = std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (mpcgpu has no CPU fallback)";
        return MPCGPU_ERR_CUDA;
    }
    mpcgpu_dtc_handle *h = new mpcgpu_dtc_handle();
    if (device < 0) cudaGetDevice(&device);
    h->device = device;
    std::string e = dtc_build_tables(*problem, h->ht);
    if (!e.empty()) { g_dtc_create_error = e; delete h; return MPCGPU_ERR_ARG; }
    auto fail = [&](const char *what, cudaError_t c2) {
        g_dtc_create_error = std::string(what) + ": " + cudaGetErrorString(c2);
        mpcgpu_dtc_destroy(h);
        return MPCGPU_ERR_CUDA;
    };
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", ce);
    cudaDeviceProp prop;
    if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", ce);
    h->smem_optin = prop.sharedMemPerBlockOptin;
    if ((ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", ce);
    const DtcHostTables &t = h->ht;
    auto up = [&](double **d, const std::vector<double> &v) {
        cudaError_t c2 = cudaMalloc((void **)d, sizeof(double) * (v.size() + 1));
        if (c2 != cudaSuccess) return c2;
        return cudaMemcpy(*d, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice);
    };
    if ((ce = up(&h->dStep, t.step)) != cudaSuccess) return fail("tables", ce);
    if ((ce = up(&h->dF, t.ftab)) != cudaSuccess) return fail("tables", ce);
    if ((ce = up(&h->dUg, t.ug)) != cudaSuccess) return fail("tables", ce);
    if ((ce = up(&h->dR, t.r)) != cudaSuccess) return fail("tables", ce);
    if ((ce = up(&h->dQ, t.q)) != cudaSuccess) return fail("tables", ce);
    const size_t smem = dtc_plan(t.L).doubles * sizeof(double) * DTC_WARPS;
    if (smem > h->smem_optin) {
        g_dtc_create_error = "DTC-GPC: horizons too large for shared memory (" + std::to_string(smem) + " B)";
        mpcgpu_dtc_destroy(h);
        return MPCGPU_ERR_UNSUPPORTED;
    }
    cudaFuncSetAttribute(k_dtc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin);
    *out = h;
    return MPCGPU_OK;
}



static int dtc_eval_impl(mpcgpu_dtc_handle *h, int n, const int32_t *p, const int32_t *m, const double *delta,
                         const double *lambda, const double *fr_num, const double *fr_den, const int32_t *fr_len,
                         const double *alfa, const double *raio, double *ise, double *y, double *u, int32_t *status) {
    if (!h) return MPCGPU_ERR_ARG;
    const bool design = alfa != nullptr;
    if (n < 0 || (n > 0 && (!p || !m || !delta || !lambda || !ise || (design ? !raio : (!fr_num || !fr_den || !fr_len))))) {
        h->err = "bad population arguments";
        return MPCGPU_ERR_ARG;
    }
    if (n == 0) return MPCGPU_OK;
    DCK(cudaSetDevice(h->device));
    const DtcLayout &L = h->ht.L;
    const int ny = L.ny, nu = L.nu, nit = L.nit;
    cudaStream_t s = h->stream;
    // one device arena for the whole call
    const size_t nI = (size_t)n * (ny + nu + 2 * ny + 1);
    const size_t nD = (size_t)n * (ny + nu + 2 * ny * DTC_MAXF + ny + 2) + (y ? (size_t)n * ny * nit : 0) + (u ? (size_t)n * nu * nit : 0);
    int *dI = nullptr;
    double *dD = nullptr;
    DCK(cudaMalloc((void **)&dI, sizeof(int) * nI));
    if (cudaMalloc((void **)&dD, sizeof(double) * nD) != cudaSuccess) { cudaFree(dI); h->err = "cudaMalloc failed"; return MPCGPU_ERR_CUDA; }
    int *dp = dI, *dm = dp + (size_t)n * ny, *dfl = dm + (size_t)n * nu, *dst = dfl + (size_t)n * 2 * ny;
    double *ddl = dD, *dlm = ddl + (size_t)n * ny, *dfn = dlm + (size_t)n * nu, *dfd = dfn + (size_t)n * ny * DTC_MAXF;
    double *dise = dfd + (size_t)n * ny * DTC_MAXF, *dal = dise + (size_t)n * ny, *dra = dal + n, *dy = dra + n, *du = dy + (y ? (size_t)n * ny * nit : 0);
    int rc = MPCGPU_OK;
    auto ck = [&](cudaError_t e2) { if (e2 != cudaSuccess && rc == MPCGPU_OK) { h->err = cudaGetErrorString(e2); rc = MPCGPU_ERR_CUDA; } };
    ck(cudaMemcpyAsync(dp, p, sizeof(int) * (size_t)n * ny, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dm, m, sizeof(int) * (size_t)n * nu, cudaMemcpyHostToDevice, s));
    ck(cudaMemsetAsync(dst, 0, sizeof(int) * (size_t)n, s));
    if (!design) ck(cudaMemcpyAsync(dfl, fr_len, sizeof(int) * (size_t)n * 2 * ny, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(ddl, delta, sizeof(double) * (size_t)n * ny, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dlm, lambda, sizeof(double) * (size_t)n * nu, cudaMemcpyHostToDevice, s));
    if (design) {   // the filters are designed on the device from (alfa, raio): 16 B per candidate instead of 2 ny MAXF doubles
        ck(cudaMemcpyAsync(dal, alfa, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
        ck(cudaMemcpyAsync(dra, raio, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
        k_dtc_filter<<<(n * ny + 127) / 128, 128, 0, s>>>(L, n, dal, dra, dfn, dfd, dfl, dst);
        ck(cudaGetLastError());
        h->cnt.kernel_launches += 1;
    } else {
        ck(cudaMemcpyAsync(dfn, fr_num, sizeof(double) * (size_t)n * ny * DTC_MAXF, cudaMemcpyHostToDevice, s));
        ck(cudaMemcpyAsync(dfd, fr_den, sizeof(double) * (size_t)n * ny * DTC_MAXF, cudaMemcpyHostToDevice, s));
    }
    if (rc == MPCGPU_OK) {
        DtcCand Cd{dp, dm, ddl, dlm, dfn, dfd, dfl, dise, y ? dy : nullptr, u ? du : nullptr, dst};
        DtcTables T{h->dStep, h->dF, h->dUg, h->dR, h->dQ, h->ht.step_len};
        const size_t smem = dtc_plan(L).doubles * sizeof(double) * DTC_WARPS;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        k_dtc<<<(n + DTC_WARPS - 1) / DTC_WARPS, 32 * DTC_WARPS, smem, s>>>(L, T, n, Cd);
        ck(cudaGetLastError());
        cudaEventRecord(e1, s);
        ck(cudaMemcpyAsync(ise, dise, sizeof(double) * (size_t)n * ny, cudaMemcpyDeviceToHost, s));
        if (y) ck(cudaMemcpyAsync(y, dy, sizeof(double) * (size_t)n * ny * nit, cudaMemcpyDeviceToHost, s));
        if (u) ck(cudaMemcpyAsync(u, du, sizeof(double) * (size_t)n * nu * nit, cudaMemcpyDeviceToHost, s));
        if (status) ck(cudaMemcpyAsync(status, dst, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
        ck(cudaStreamSynchronize(s));
        float ms = 0.f;
        if (rc == MPCGPU_OK && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) { h->cnt.last_sim_ms = ms; h->cnt.last_total_ms = ms; }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        h->cnt.candidates += n; h->cnt.closed_loops += n; h->cnt.kernel_launches += 1;
    }
    cudaFree(dI); cudaFree(dD);
    return rc;
}

extern "C" int mpcgpu_dtc_eval_batch(mpcgpu_dtc_handle *h, int n, const int32_t *p, const int32_t *m, const double *delta,
                                     const double *lambda, const double *fr_num, const double *fr_den,
                                     const int32_t *fr_len, double *ise, double *y, double *u, int32_t *status) {
    return dtc_eval_impl(h, n, p, m, delta, lambda, fr_num, fr_den, fr_len, nullptr, nullptr, ise, y, u, status);
}

// The sweep with the robustness filter designed on the device from (alfa, raio) per candidate (mimofilter.m / filtro_siso.m)
extern "C" int mpcgpu_dtc_eval_batch_design(mpcgpu_dtc_handle *h, int n, const int32_t *p, const int32_t *m, const double *delta,
                                            const double *lambda, const double *alfa, const double *raio, double *ise, double *y,
                                            double *u, int32_t *status) {
    if (!alfa || !raio) { if (h) h->err = "alfa / raio missing"; return MPCGPU_ERR_ARG; }
    return dtc_eval_impl(h, n, p, m, delta, lambda, nullptr, nullptr, nullptr, alfa, raio, ise, y, u, status);
}

extern "C" int mpcgpu_dtc_get_counters(mpcgpu_dtc_handle *h, mpcgpu_counters *out) {
    if (!h || !out) return MPCGPU_ERR_ARG;
    *out = h->cnt;
    return MPCGPU_OK;
}
