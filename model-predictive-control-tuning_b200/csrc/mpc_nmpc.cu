// mpc_nmpc.cu -- nonlinear path (BASELINE.json configs[4]): batched closedloop_toolbox_nmpc for the Van de Vusse
// CSTR (closedloop_toolbox_nmpc.m:36-97, vandevusse_model.m:39-77), GAM / VNS costs fused (GAM_fun.m:110-115,
// VNS2.m:147-195 nonlinear branch).  C ABI: include/mpcgpu.h, NMPC section.
//
// One THREAD per closed-loop run (candidate x VNS run): the work of a run is a long serial chain (59 controller
// calls, each a few SQP iterations of p*nsub RK4 steps on a 3-state model) with no parallelism worth a warp, and a
// population supplies tens of thousands of independent runs.  Per controller call (nlmpcmove restated as N1-N4,
// DESIGN.md section 2 / include/mpcgpu.h):
//   rollout with forward sensitivities (RK4 stage Jacobians chained: [A|B] per sample, X = dx/dv carried),
//   Gauss-Newton model  H = sum S'Wy^2 S + D'Wdu^2 D,  g,  accumulated on the fly (S never stored),
//   exact box-constrained QP step (primal active set on the MV bounds, Cholesky of the free block),
//   backtracking on the true cost;  stop when the scaled step is < 1e-10.
// Per-thread state (H: nz^2 <= 900 doubles) lives in local memory, interleaved across the warp by the hardware.
// HBM traffic per run: 48 B in, 8-16 B out (+ trajectories on request): compute/latency bound, fp64.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"

#include "mpc_nmpc_core.h"
#ifdef NM_GLOBAL_WORK
#define NM_WORK_ALWAYS 1
#else
#define NM_WORK_ALWAYS 0
#endif

struct NmpcArgs {
    const int *N, *Nu;
    const double *delta, *lambda;
    const double *r, *yref;      // ny x nit
    double *cost, *part;         // GAM: n x ny; VNS partial: n x runs
    double *y, *u, *yopt, *uopt; // optional n x 2 x nit
    int *status;
    unsigned long long *counters;   // [0] controller calls, [1] SQP iterations
    double *work;                // per run: 2 * nz_max^2 doubles
};

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


// =================================================================================================
// Warp-per-run form of the same algorithm (the default): lane a owns decision variable a (nz = 2*m <= 32):
// its plan entry, its column of the sensitivity matrix X = dx/dv, its row of the Gauss-Newton Hessian (in
// registers) and its gradient entry.  The 3-state rollout and its stage Jacobians are computed redundantly by all
// lanes (uniform, no divergence); the box-QP is solved cooperatively in shared memory (packed Cholesky of the free
// block, column-oriented substitutions, warp reductions for the ratio test and the multiplier test).
// Measured on 16384 Van de Vusse candidates: 1.56 s against 1.47 s for the thread-per-run kernel above -- the run time
// is the serial chain of RK4 stages (three exp() per right-hand side), which neither mapping shortens; thread-per-run
// stays the default because it needs no shared memory and packs 32 runs into a warp.
// =================================================================================================
#define NMW_WARPS 4
#define NMW_LD 32
#define NMW_FULL 0xffffffffu
#define NMW_DOUBLES (2 * NMW_LD * NMW_LD + 8 * 32)

struct NmwSm { double *H, *Lc, *v, *S, *g, *d, *t, *lo, *hi, *vt; };

__device__ __forceinline__ double nmw_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NMW_FULL, v, o);
    return v;
}
__device__ __forceinline__ double nmw_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(NMW_FULL, v, o));
    return v;
}
// minimum value over the warp and the smallest lane index holding it (i < 0: lane does not take part)
__device__ __forceinline__ void nmw_argmin(double &v, int &i) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long key = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(NMW_FULL, i >= 0 ? hi : 0xffffffffu);
    const unsigned mlo = __reduce_min_sync(NMW_FULL, (i >= 0 && hi == mhi) ? lo : 0xffffffffu);
    const unsigned mi = __reduce_min_sync(NMW_FULL, (i >= 0 && hi == mhi && lo == mlo) ? (unsigned)i : 0xffffffffu);
    const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
    const unsigned long long mb = (mk >> 63) ? (mk & 0x7fffffffffffffffull) : ~mk;
    v = __longlong_as_double((long long)mb);
    i = mi == 0xffffffffu ? -1 : (int)mi;
}

// Predicted cost of the plan  clamp(v + alpha d)  for THIS LANE's step length alpha (v, d in shared memory, nz entries):
// the six step lengths of the backtracking line search are evaluated at once, one per lane group, instead of one rollout
// after the other (the rollout is a serial chain of RK4 stages; lanes are free).  alpha = 0: the cost of v itself.
__device__ double w_plan_cost(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m,
                              const double *wy2, const double *wu2, const double *vs, const double *ds, double alpha) {
    double x[NX] = {x0[0], x0[1], x0[2]};
    double J = 0.0;
    double up[NU] = {uprev[0], uprev[1]};
    for (int i = 0; i < p; ++i) {
        const int c = i < m ? i : m - 1;
        double u[NU];
#pragma unroll
        for (int j = 0; j < NU; ++j) u[j] = fmin(fmax(fma(alpha, ds[NU * c + j], vs[NU * c + j]), D.umin[j]), D.umax[j]);
        if (i < m) {
#pragma unroll
            for (int j = 0; j < NU; ++j) { const double du = u[j] - up[j]; J = fma(wu2[j] * du, du, J); up[j] = u[j]; }
        }
        rk4_sample(D, x, u, nullptr);
#pragma unroll
        for (int j = 0; j < NY; ++j) { const double e = r[j] - x[1 + j]; J = fma(wy2[j] * e, e, J); }
    }
    return J;
}

// One sample (nsub RK4 steps) with the sensitivities [A|B] = [dx+/dx | dx+/du] (3 x 5) computed LANE-PARALLEL: lane e < 15
// owns entry (r, c) = (e / 5, e % 5) of every 3 x 5 matrix of the chain (stage sensitivities Dk, the sub-step transition,
// the running [A|B]); a 3 x 3 by 3 x 5 product is three shuffles and three FMAs per lane instead of 45 serial FMAs.  The
// state x and the stage derivatives (two exp each) are uniform.  Leaves [A|B] in ABs[0..15) (shared) after a warp barrier.
__device__ void w_rk4_sens(const NmpcDev &D, double *x, const double *u, int lane, double *ABs) {
    const int e = lane % 15, r = e / 5, c = e - 5 * r;
    const double h = D.Ts / D.nsub, diag = r == c ? 1.0 : 0.0;
    double ABe = diag;
    for (int s = 0; s < D.nsub; ++s) {
        double k[4][NX], xs[NX], Jm[15], Dk[4];
        const double ca[4] = {0.0, 0.5, 0.5, 1.0};
#pragma unroll
        for (int st = 0; st < 4; ++st) {
#pragma unroll
            for (int i = 0; i < NX; ++i) xs[i] = st == 0 ? x[i] : x[i] + ca[st] * h * k[st - 1][i];
            vdv_rhs(xs, u, k[st], Jm);
            // this lane's row of J: entries q < 3 and (for c >= 3) the input column c
            const double j0 = r == 0 ? Jm[0] : (r == 1 ? Jm[5] : Jm[10]), j1 = r == 0 ? Jm[1] : (r == 1 ? Jm[6] : Jm[11]);
            const double j2 = r == 0 ? Jm[2] : (r == 1 ? Jm[7] : Jm[12]);
            const double ju = c == 3 ? (r == 0 ? Jm[3] : (r == 1 ? Jm[8] : Jm[13])) : (c == 4 ? (r == 0 ? Jm[4] : (r == 1 ? Jm[9] : Jm[14])) : 0.0);
            const double Dxe = (st == 0 ? 0.0 : ca[st] * h * Dk[st - 1]) + diag;   // d xs / d(x_sub, u)
            double acc = ju;
            acc = fma(j0, __shfl_sync(NMW_FULL, Dxe, c), acc);
            acc = fma(j1, __shfl_sync(NMW_FULL, Dxe, 5 + c), acc);
            acc = fma(j2, __shfl_sync(NMW_FULL, Dxe, 10 + c), acc);
            Dk[st] = acc;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] += (h / 6.0) * (k[0][i] + 2.0 * k[1][i] + 2.0 * k[2][i] + k[3][i]);
        const double Phie = (h / 6.0) * (Dk[0] + 2.0 * Dk[1] + 2.0 * Dk[2] + Dk[3]) + diag;   // transition of this sub-step
        double acc = c >= NX ? Phie : 0.0;
#pragma unroll
        for (int q = 0; q < NX; ++q) acc = fma(__shfl_sync(NMW_FULL, Phie, 5 * r + q), __shfl_sync(NMW_FULL, ABe, 5 * q + c), acc);
        ABe = acc;
    }
    __syncwarp();
    if (lane < 15) ABs[lane] = ABe;
    __syncwarp();
}

// exact  min 1/2 d'Hd + g'd,  lo <= d <= hi  (lo <= 0 <= hi), H SPD in sm.H (leading dimension NMW_LD); result in sm.d
__device__ int w_box_qp(int nz, const NmwSm &sm, int lane) {
    const bool var = lane < nz;
    const double ga = var ? sm.g[lane] : 0.0, loa = var ? sm.lo[lane] : 0.0, hia = var ? sm.hi[lane] : 0.0;
    int fixed = !var ? 2 : ((ga > 0.0 && loa >= 0.0) ? -1 : ((ga < 0.0 && hia <= 0.0) ? 1 : 0));
    double da = 0.0;
    const double gscale = nmw_max(fabs(ga));
    for (int it = 0; it < 6 * nz + 20; ++it) {
        sm.d[lane] = da;
        const unsigned fmask = __ballot_sync(NMW_FULL, fixed == 0);
        const int nf = __popc(fmask), fi = __popc(fmask & ((1u << lane) - 1u));
        __syncwarp();
        if (fixed == 0) {
            // packed row fi of H_FF and the right-hand side -(g_F + H_FA d_A)
            double rhs = -ga;
            int fj = 0;
            for (int b = 0; b < nz; ++b) {
                const double hab = sm.H[lane * NMW_LD + b];
                if ((fmask >> b) & 1u) { if (fj <= fi) sm.Lc[fi * NMW_LD + fj] = hab; fj++; }
                else rhs = fma(-hab, sm.d[b], rhs);
            }
            sm.t[fi] = rhs;
        }
        __syncwarp();
        // Cholesky of the packed block, lanes = rows
        for (int k = 0; k < nf; ++k) {
            const double dkk = sm.Lc[k * NMW_LD + k];
            if (!(dkk > 0.0)) return 3;
            const double ckk = sqrt(dkk);
            double lrk = 0.0;
            if (lane >= k && lane < nf) { lrk = lane == k ? ckk : sm.Lc[lane * NMW_LD + k] / ckk; sm.Lc[lane * NMW_LD + k] = lrk; }
            __syncwarp();
            if (lane > k && lane < nf)
                for (int c = k + 1; c <= lane; ++c) sm.Lc[lane * NMW_LD + c] = fma(-lrk, sm.Lc[c * NMW_LD + k], sm.Lc[lane * NMW_LD + c]);
            __syncwarp();
        }
        for (int k = 0; k < nf; ++k) {   // L y = t
            const double yk = sm.t[k] / sm.Lc[k * NMW_LD + k];
            __syncwarp();
            if (lane == k) sm.t[k] = yk;
            else if (lane > k && lane < nf) sm.t[lane] = fma(-sm.Lc[lane * NMW_LD + k], yk, sm.t[lane]);
            __syncwarp();
        }
        for (int k = nf - 1; k >= 0; --k) {   // L' x = y
            const double xk = sm.t[k] / sm.Lc[k * NMW_LD + k];
            __syncwarp();
            if (lane == k) sm.t[k] = xk;
            else if (lane < k) sm.t[lane] = fma(-sm.Lc[k * NMW_LD + lane], xk, sm.t[lane]);
            __syncwarp();
        }
        // longest feasible step toward the Newton point of the face
        double alpha = 1.0;
        int blk = -1, side = 0;
        double step = 0.0;
        if (fixed == 0) {
            step = sm.t[fi] - da;
            if (step > 0.0 && da + step > hia) { alpha = (hia - da) / step; blk = lane; side = 1; }
            if (step < 0.0 && da + step < loa) { alpha = (loa - da) / step; blk = lane; side = -1; }
        }
        double amin = alpha;
        int bl = blk;
        nmw_argmin(amin, bl);
        if (bl < 0) amin = 1.0;
        if (fixed == 0) da += amin * step;
        if (bl >= 0) {
            if (lane == bl) { da = side > 0 ? hia : loa; fixed = side; }
            continue;
        }
        // minimiser of the face: release the bound with the most wrong-signed multiplier
        sm.d[lane] = da;
        __syncwarp();
        double viol = 0.0;
        int cand = -1;
        if (fixed == 1 || fixed == -1) {
            double gi = ga;
            for (int b = 0; b < nz; ++b) gi = fma(sm.H[lane * NMW_LD + b], sm.d[b], gi);
            viol = fixed < 0 ? -gi : gi;
            if (viol > 0.0) cand = lane;
        }
        double key = -viol;
        nmw_argmin(key, cand);
        if (cand < 0 || -key <= 1e-14 * gscale) { sm.d[lane] = da; __syncwarp(); return 0; }
        if (lane == cand) fixed = 0;
    }
    sm.d[lane] = da;
    __syncwarp();
    return 2;
}

// one nlmpcmove by one warp: plan in sm.v (in: start, out: optimum)
__device__ int w_nlmpcmove(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m,
                           const double *wy2, const double *wu2, const NmwSm &sm, int lane, unsigned *n_sqp) {
    const int nz = NU * m;
    const bool var = lane < nz;
    const int ja = lane % NU, ca = lane / NU;
    const double umn = D.umin[ja], umx = D.umax[ja], sua = D.su[ja];
    double va = var ? fmin(fmax(sm.v[lane], umn), umx) : 0.0;
    sm.v[lane] = va;
    __syncwarp();
    sm.d[lane] = 0.0;
    __syncwarp();
    double Jcur = w_plan_cost(D, x0, uprev, r, p, m, wy2, wu2, sm.v, sm.d, 0.0);
    int status = 0;
    for (int it = 0; it < D.max_sqp; ++it) {
        *n_sqp += 1;
        double Hrow[NM_MAXZ];
#pragma unroll
        for (int b = 0; b < NM_MAXZ; ++b) Hrow[b] = 0.0;
        double g = 0.0, X0 = 0.0, X1 = 0.0, X2 = 0.0;
        double x[NX] = {x0[0], x0[1], x0[2]};
        for (int i = 0; i < p; ++i) {
            const int c = i < m ? i : m - 1;
            const double u[NU] = {sm.v[NU * c], sm.v[NU * c + 1]};
            w_rk4_sens(D, x, u, lane, sm.t);
            {   // X_a <- A X_a + B e_(c, j)
                const double *AB = sm.t;
                const double a0 = X0, a1 = X1, a2 = X2;
                X0 = AB[0] * a0 + AB[1] * a1 + AB[2] * a2;
                X1 = AB[5] * a0 + AB[6] * a1 + AB[7] * a2;
                X2 = AB[10] * a0 + AB[11] * a1 + AB[12] * a2;
                if (var && ca == c) { X0 += AB[NX + ja]; X1 += AB[5 + NX + ja]; X2 += AB[10 + NX + ja]; }
            }
#pragma unroll
            for (int j = 0; j < NY; ++j) {
                const double Sa = var ? (j == 0 ? X1 : X2) : 0.0;   // dy_j/dv_a
                sm.S[lane] = Sa;
                __syncwarp();
                const double e = r[j] - x[1 + j];
                const double wa = wy2[j] * Sa;
                g = fma(-wa, e, g);
#pragma unroll
                for (int b = 0; b < NM_MAXZ; ++b) Hrow[b] = fma(wa, sm.S[b], Hrow[b]);
                __syncwarp();
            }
        }
        // move-suppression terms (lane-local rows of D'Wdu^2 D)
        if (var) {
            const double w = wu2[ja];
            const double du = va - (ca == 0 ? uprev[ja] : sm.v[lane - NU]);
            g = fma(w, du, g);
            const bool has_next = ca + 1 < m;
            if (has_next) g = fma(-w, sm.v[lane + NU] - va, g);
#pragma unroll
            for (int b = 0; b < NM_MAXZ; ++b) {
                if (b == lane) Hrow[b] += has_next ? 2.0 * w : w;
                if (b + NU == lane) Hrow[b] -= w;
                if (b == lane + NU && has_next) Hrow[b] -= w;
            }
        }
#pragma unroll
        for (int b = 0; b < NM_MAXZ; ++b) sm.H[lane * NMW_LD + b] = Hrow[b];
        sm.g[lane] = g; sm.lo[lane] = umn - va; sm.hi[lane] = umx - va;
        __syncwarp();
        const int rc = w_box_qp(nz, sm, lane);
        if (rc) { status = rc; break; }
        const double d = var ? sm.d[lane] : 0.0;
        const double dmax = nmw_max(fabs(d) / sua);
        if (dmax < 1e-10) break;
        // ---- backtracking on the true cost: the six step lengths 1, 1/2, .. 1/32 at once, one per lane (lanes >= 6 repeat
        // the last one); the first (largest) that decreases the cost is taken, as the sequential search would ----
        const int bt = (lane & 7) < 5 ? (lane & 7) : 5;
        const double alpha_l = 1.0 / (double)(1 << bt);
        const double Jl = w_plan_cost(D, x0, uprev, r, p, m, wy2, wu2, sm.v, sm.d, alpha_l);
        const unsigned okm = __ballot_sync(NMW_FULL, lane < 6 && Jl < Jcur);
        if (!okm) break;   // no descent at this resolution: converged to rounding
        const int win = __ffs((int)okm) - 1;
        const double Jn = __shfl_sync(NMW_FULL, Jl, win);
        const double alpha = 1.0 / (double)(1 << win);
        va = var ? fmin(fmax(va + alpha * d, umn), umx) : 0.0;
        __syncwarp();
        sm.v[lane] = va;
        __syncwarp();
        Jcur = Jn;
    }
    return status;
}

// mode 0 RAW, 1 GAM, 2 VNS.  One warp per (candidate, run).
__global__ void __launch_bounds__(32 * NMW_WARPS) k_nmpc_w(const NmpcDev D, int n, int runs, int mode, const int *order, NmpcArgs A) {
    extern __shared__ double smem_n[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * NMW_WARPS + warp;
    if (item >= n * runs) return;
    const int c = order[item / runs], run = item - (item / runs) * runs;   // longest horizons first
    const int p = A.N[c], m = A.Nu[c], nit = D.nit;
    if (p < 2 || p > D.pmax || m < 1 || m > D.mmax || m >= p) {
        if (lane == 0) {
            A.status[c] = MPCGPU_CAND_INVALID;
            if (mode == 1) for (int j = 0; j < NY; ++j) A.cost[(size_t)c * NY + j] = NAN;
            if (mode == 2) A.part[(size_t)c * runs + run] = NAN;
        }
        return;
    }
    NmwSm sm;
    {
        double *q_ = smem_n + (size_t)warp * NMW_DOUBLES;
        sm.H = q_; q_ += NMW_LD * NMW_LD; sm.Lc = q_; q_ += NMW_LD * NMW_LD;
        sm.v = q_; q_ += 32; sm.S = q_; q_ += 32; sm.g = q_; q_ += 32; sm.d = q_; q_ += 32; sm.t = q_; q_ += 32;
        sm.lo = q_; q_ += 32; sm.hi = q_; q_ += 32; sm.vt = q_;
    }
    const int sel = mode == 2 ? run : -1;
    double wy2[NY], wu2[NU];
    for (int j = 0; j < NY; ++j) { const double w = A.delta[(size_t)c * NY + j] / D.sy[j]; wy2[j] = w * w; }
    for (int j = 0; j < NU; ++j) { const double w = A.lambda[(size_t)c * NU + j] / D.su[j]; wu2[j] = w * w; }
    double rr[NY];
    auto ref_at = [&](int k, double *out) {
        for (int j = 0; j < NY; ++j) out[j] = (sel < 0 || sel == j) ? A.r[(size_t)j * nit + k] : 0.0;
    };
    unsigned n_sqp = 0, n_calls = 0;
    int status = 0;
    const bool want_ol = mode != 1 || A.yopt || A.uopt;
    double jnu = 0.0, cost_acc[NY] = {0.0, 0.0}, vns_acc = 0.0;
    double xo[NX] = {D.x0[0], D.x0[1], D.x0[2]};
    double vopt_a = 0.0;   // lane a: entry a of the open-loop plan
    if (want_ol) {
        sm.v[lane] = D.u0[lane % NU];
        __syncwarp();
        ref_at(nit - 1, rr);
        const int rc = w_nlmpcmove(D, D.x0, D.u0, rr, p, m, wy2, wu2, sm, lane, &n_sqp);
        n_calls++;
        if (rc) status = rc;
        vopt_a = sm.v[lane];
        if (mode == 2) {   // Jnu (VNS2.m:183-191) on input `sel`
            const int j = sel;
            const double u0a = fabs(sm.v[j]);
            for (int cc = 0; cc + 1 < m && cc + 1 < nit; ++cc) {
                const double xn = u0a / fabs(sm.v[NU * (cc + 1) + j] - sm.v[NU * cc + j]);
                if (fabs(xn) <= 1.7976931348623157e308) jnu += xn * xn;
            }
        }
        __syncwarp();
    }
    double x[NX] = {D.x0[0], D.x0[1], D.x0[2]}, uprev[NU] = {D.u0[0], D.u0[1]};
    sm.v[lane] = D.u0[lane % NU];
    __syncwarp();
    for (int k = 0; k < nit; ++k) {
        double uo[NU] = {0.0, 0.0};
        if (want_ol) {
            const int cc = k < m ? k : m - 1;
            uo[0] = __shfl_sync(NMW_FULL, vopt_a, NU * cc);
            uo[1] = __shfl_sync(NMW_FULL, vopt_a, NU * cc + 1);
        }
        if (k > 0) {
            ref_at(k, rr);
            const int rc = w_nlmpcmove(D, x, uprev, rr, p, m, wy2, wu2, sm, lane, &n_sqp);   // warm start = previous plan
            n_calls++;
            if (rc) status = rc;
            uprev[0] = sm.v[0]; uprev[1] = sm.v[1];
            rk4_sample(D, x, uprev, nullptr);
            for (int i = 0; i < NX; ++i)
                if (x[i] < D.xmin[i] - 1e-9 || x[i] > D.xmax[i] + 1e-9) { if (!status) status = 5; }
            if (want_ol) rk4_sample(D, xo, uo, nullptr);
        }
        for (int j = 0; j < NY; ++j) {
            const bool mine = sel < 0 || sel == j;
            const double yj = x[1 + j], yoj = xo[1 + j], yr = A.yref[(size_t)j * nit + k];
            if (mine) {
                if (lane == 0) {
                    if (A.y) A.y[((size_t)c * NY + j) * nit + k] = yj;
                    if (A.u) A.u[((size_t)c * NU + j) * nit + k] = uprev[j];
                    if (want_ol && A.yopt) A.yopt[((size_t)c * NY + j) * nit + k] = yoj;
                    if (want_ol && A.uopt) A.uopt[((size_t)c * NU + j) * nit + k] = uo[j];
                }
                if (mode == 1) cost_acc[j] += (yj - yr) * (yj - yr);
                if (mode == 2 && k >= D.inK - 1) vns_acc += (yj - yoj) * (yj - yoj) + (yj - yr) * (yj - yr);
            }
        }
    }
    if (lane == 0) {
        if (mode == 1) for (int j = 0; j < NY; ++j) A.cost[(size_t)c * NY + j] = (status == 0 || status == 5) ? cost_acc[j] : NAN;
        if (mode == 2) A.part[(size_t)c * runs + run] = (status == 0 || status == 5) ? vns_acc + jnu : NAN;
        if (status) atomicMax(A.status + c, status);
        atomicAdd(A.counters + 0, (unsigned long long)n_calls);
        atomicAdd(A.counters + 1, (unsigned long long)n_sqp);
    }
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
    cudaFuncSetAttribute(k_nmpc_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * NMW_DOUBLES * NMW_WARPS));
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
    // default: one warp per run (lane-parallel sensitivities, all line-search steps at once); MPCGPU_NMPC_THREAD_PER_RUN=1: one thread per run
    const bool warp_per_run = getenv("MPCGPU_NMPC_THREAD_PER_RUN") == nullptr;
    const size_t nD = (size_t)n * (2 * NY + 2 * NU) + (size_t)n * runs + (traj ? 4 * nT : 0) + (size_t)NY * nit +
                      ((warp_per_run || NM_WORK_ALWAYS) ? (size_t)n * runs * 2 * NM_MAXZ * NM_MAXZ : 0) + 8;   // (the thread-per-run kernel keeps H thread-local)
    int *dI = nullptr; double *dD = nullptr;
    if (cudaMalloc((void **)&dI, sizeof(int) * nI) != cudaSuccess || cudaMalloc((void **)&dD, sizeof(double) * nD) != cudaSuccess) {
        cudaFree(dI); h->err = "cudaMalloc failed"; return MPCGPU_ERR_CUDA;
    }
    int *dN = dI, *dNu = dN + n, *dSt = dNu + n, *dOrd = dSt + n;
    // candidates binned by (Nu, N): the 32 runs of a warp then share their loop trip counts
    std::vector<int> order(n);
    for (int c = 0; c < n; ++c) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return Nu[a] != Nu[b] ? Nu[a] > Nu[b] : N[a] > N[b]; });
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
        if (warp_per_run) {   // alternative mapping, same algorithm (measured: no faster, see header)
            const size_t smem = sizeof(double) * NMW_DOUBLES * NMW_WARPS;
            k_nmpc_w<<<(items + NMW_WARPS - 1) / NMW_WARPS, 32 * NMW_WARPS, smem, s>>>(h->D, n, runs, cost_mode, dOrd, A);
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
