// mpc_ssnmpc.cu -- single-shooting NMPC sweep: batched ClosedLoopNMPC
// (/root/reference/Explicit NMPC/ClosedLoopNMPC.m:1-110) with the controller of NMPC_Controller.m:1-141 on the Van de Vusse
// model (plant_model.m:1-56).  C ABI: include/mpcgpu.h, "Single-shooting NMPC".
//
// Kernel: k_ssnmpc, one THREAD per closed-loop run, running mpc_ssnmpc_core.h (S1-S5) -- the source the host build
// (oracle/nmpc_port) shares, so what the host tests pin is what the device executes.  The demo's problems are small (the
// reference runs N = 5, Nu = [2 2]: nz = 4; a sweep stays below nz ~ 10), a run is a chain of nit controller calls x up to
// max_sqp Gauss-Newton iterations x N RK4 samples with sensitivities: latency of one thread, parallel over candidates.
// Candidates are sorted by (N, sum Nu) so that the runs of a warp share their trip counts; H and its factor are thread-local
// arrays with the fixed leading dimension NM_LD (equal (row, col) -> equal offset across the warp, as in k_nmpc).
// HBM traffic per run: 44 B in, 16 B out (+ 2 x 2 x nit doubles of trajectories on request): compute / latency bound, fp64.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"

#define NM_FN static __device__   /* mpc_nmpc.cu holds the external-linkage copies of the shared functions */
#include "mpc_ssnmpc_core.h"

#define SS_THREADS 32

__global__ void __launch_bounds__(SS_THREADS) k_ssnmpc(const SsnmpcDev S, int n, const int *order, SsArgs A) {
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n) return;
    double H[NM_LD * NM_LD], Lc[NM_LD * NM_LD];
    ss_item(S, order, A, item, H, Lc);
}

// ------------------------------------------------------------------------------------------------
struct mpcgpu_ssnmpc_handle {
    int device = 0;
    std::string err;
    SsnmpcDev S;
    double *dR = nullptr;
    cudaStream_t stream = nullptr;
    mpcgpu_counters cnt = {};
};
static std::string g_ss_create_error;

extern "C" const char *mpcgpu_ssnmpc_last_error(mpcgpu_ssnmpc_handle *h) { return h ? h->err.c_str() : g_ss_create_error.c_str(); }

extern "C" void mpcgpu_ssnmpc_destroy(mpcgpu_ssnmpc_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    cudaFree(h->dR);
    delete h;
}

extern "C" int mpcgpu_ssnmpc_create(const mpcgpu_ssnmpc_problem *pb, int device, mpcgpu_ssnmpc_handle **out) {
    if (!pb || !out) { g_ss_create_error = "NULL argument"; return MPCGPU_ERR_ARG; }
    *out = nullptr;
    if (pb->model != MPCGPU_MODEL_VANDEVUSSE) { g_ss_create_error = "unknown model id"; return MPCGPU_ERR_UNSUPPORTED; }
    bool ok = pb->nit >= 2 && pb->pmax >= 1 && pb->inK >= 2 && pb->inK <= pb->nit && pb->nsub >= 1 && pb->Ts > 0 && pb->x0 && pb->u0 &&
              pb->lb && pb->ub && pb->r;
    for (int j = 0; ok && j < NY; ++j) ok = pb->x_control[j] >= 0 && pb->x_control[j] < NX;
    for (int j = 0; ok && j < NU; ++j) ok = pb->lb[j] < pb->ub[j] && pb->u0[j] >= pb->lb[j] && pb->u0[j] <= pb->ub[j];
    if (!ok) {
        g_ss_create_error = "bad single-shooting NMPC problem (nit >= 2, 2 <= inK <= nit, nsub >= 1, Ts > 0, x_control in 0..2, lb <= u0 <= ub, lb < ub, pointers set)";
        return MPCGPU_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_ss_create_error = std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (mpcgpu has no CPU fallback)";
        return MPCGPU_ERR_CUDA;
    }
    mpcgpu_ssnmpc_handle *h = new mpcgpu_ssnmpc_handle();
    if (device < 0) cudaGetDevice(&device);
    h->device = device;
    NmpcDev &D = h->S.D;
    D.nit = pb->nit; D.pmax = pb->pmax; D.mmax = NM_MAXM; D.inK = pb->inK; D.nsub = pb->nsub;
    D.max_sqp = pb->max_sqp > 0 ? pb->max_sqp : 400; D.Ts = pb->Ts;
    for (int i = 0; i < NX; ++i) { D.x0[i] = pb->x0[i]; D.xmin[i] = -INFINITY; D.xmax[i] = INFINITY; }
    for (int j = 0; j < NU; ++j) { D.u0[j] = pb->u0[j]; D.umin[j] = pb->lb[j]; D.umax[j] = pb->ub[j]; D.su[j] = pb->ub[j] - pb->lb[j]; }
    for (int j = 0; j < NY; ++j) { D.sy[j] = 1.0; h->S.xc[j] = pb->x_control[j]; }
    h->S.pmax = pb->pmax;
    auto fail = [&](const char *what, cudaError_t c2) {
        g_ss_create_error = std::string(what) + ": " + cudaGetErrorString(c2);
        mpcgpu_ssnmpc_destroy(h);
        return MPCGPU_ERR_CUDA;
    };
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", ce);
    if ((ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", ce);
    const size_t nb = sizeof(double) * NY * pb->nit;
    if ((ce = cudaMalloc((void **)&h->dR, nb)) != cudaSuccess) return fail("alloc", ce);
    if ((ce = cudaMemcpy(h->dR, pb->r, nb, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("copy", ce);
    *out = h;
    return MPCGPU_OK;
}

extern "C" int mpcgpu_ssnmpc_eval_batch(mpcgpu_ssnmpc_handle *h, int n, const int32_t *N, const int32_t *Nu, const double *Q,
                                        const double *W, const double *r_override, const double *noise, double *cost, double *y,
                                        double *u, int32_t *status) {
    if (!h) return MPCGPU_ERR_ARG;
    if (n < 0 || (n > 0 && (!N || !Nu || !Q || !W))) { h->err = "bad arguments"; return MPCGPU_ERR_ARG; }
    if (n == 0) return MPCGPU_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return MPCGPU_ERR_CUDA; }
    const int nit = h->S.D.nit;
    const bool traj = y || u;
    cudaStream_t s = h->stream;
    const size_t nT = (size_t)n * 2 * nit;
    const size_t nI = (size_t)n * (1 + NU + 1 + 1);
    const size_t nD = (size_t)n * (NY + NU + NY) + (traj ? 2 * nT : 0) + (size_t)NY * nit + (size_t)NX * nit + 8;
    int *dI = nullptr; double *dD = nullptr;
    if (cudaMalloc((void **)&dI, sizeof(int) * nI) != cudaSuccess || cudaMalloc((void **)&dD, sizeof(double) * nD) != cudaSuccess) {
        cudaFree(dI); h->err = "cudaMalloc failed"; return MPCGPU_ERR_CUDA;
    }
    int *dN = dI, *dNu = dN + n, *dSt = dNu + (size_t)n * NU, *dOrd = dSt + n;
    std::vector<int> order(n);
    for (int c = 0; c < n; ++c) order[c] = c;
    auto work = [&](int c) { int z = 0; for (int j = 0; j < NU; ++j) z += Nu[(size_t)c * NU + j]; return z; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return N[a] != N[b] ? N[a] > N[b] : work(a) > work(b); });
    double *q_ = dD;
    double *dQ = q_; q_ += (size_t)n * NY; double *dW = q_; q_ += (size_t)n * NU; double *dCost = q_; q_ += (size_t)n * NY;
    double *dY = nullptr, *dU = nullptr;
    if (traj) { dY = q_; q_ += nT; dU = q_; q_ += nT; }
    double *dRo = q_; q_ += (size_t)NY * nit; double *dNz = q_; q_ += (size_t)NX * nit;
    unsigned long long *dCnt = (unsigned long long *)q_;
    int rc = MPCGPU_OK;
    auto ck = [&](cudaError_t e2) { if (e2 != cudaSuccess && rc == MPCGPU_OK) { h->err = cudaGetErrorString(e2); rc = MPCGPU_ERR_CUDA; } };
    ck(cudaMemcpyAsync(dN, N, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dNu, Nu, sizeof(int) * n * NU, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dOrd, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dQ, Q, sizeof(double) * n * NY, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dW, W, sizeof(double) * n * NU, cudaMemcpyHostToDevice, s));
    if (r_override) ck(cudaMemcpyAsync(dRo, r_override, sizeof(double) * NY * nit, cudaMemcpyHostToDevice, s));
    if (noise) ck(cudaMemcpyAsync(dNz, noise, sizeof(double) * NX * nit, cudaMemcpyHostToDevice, s));
    ck(cudaMemsetAsync(dSt, 0, sizeof(int) * n, s));
    ck(cudaMemsetAsync(dCnt, 0, sizeof(unsigned long long) * 2, s));
    if (rc == MPCGPU_OK) {
        SsArgs A{dN, dNu, dQ, dW, r_override ? dRo : h->dR, noise ? dNz : nullptr, dCost, dY, dU, dSt, dCnt};
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        k_ssnmpc<<<(n + SS_THREADS - 1) / SS_THREADS, SS_THREADS, 0, s>>>(h->S, n, dOrd, A);
        ck(cudaGetLastError());
        cudaEventRecord(e1, s);
        unsigned long long cnt[2] = {0, 0};
        if (cost) ck(cudaMemcpyAsync(cost, dCost, sizeof(double) * n * NY, cudaMemcpyDeviceToHost, s));
        if (y) ck(cudaMemcpyAsync(y, dY, sizeof(double) * nT, cudaMemcpyDeviceToHost, s));
        if (u) ck(cudaMemcpyAsync(u, dU, sizeof(double) * nT, cudaMemcpyDeviceToHost, s));
        if (status) ck(cudaMemcpyAsync(status, dSt, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
        ck(cudaMemcpyAsync(cnt, dCnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
        ck(cudaStreamSynchronize(s));
        h->cnt.candidates += n; h->cnt.closed_loops += (uint64_t)n; h->cnt.qp_solves += cnt[0]; h->cnt.as_iterations += cnt[1];
        h->cnt.kernel_launches += 1;
        float ms = 0.f;
        if (rc == MPCGPU_OK && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) { h->cnt.last_sim_ms = ms; h->cnt.last_total_ms = ms; }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(dI); cudaFree(dD);
    return rc;
}

extern "C" int mpcgpu_ssnmpc_get_counters(mpcgpu_ssnmpc_handle *h, mpcgpu_counters *out) {
    if (!h || !out) return MPCGPU_ERR_ARG;
    *out = h->cnt;
    return MPCGPU_OK;
}
