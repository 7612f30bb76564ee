// mpc_soft.cuh -- closed loop for plants WITH SOFT OUTPUT CONSTRAINTS (Shell7x5 band control,
// Shell7x5.m:155-165,186-189): one CTA (SOFT_THREADS threads) per closed-loop run.
//
// Same closed loop, costs and deviation-coordinate controller map as mpc_sim.cuh (closedloop_toolbox.m:50-100),
// but the QP has one more variable, the slack eps (cost rho_eps * eps^2, eps >= 0), and 2*ny*p more rows
//     ymin_i - eps*Vmin_i*sy_i <= y_i(k+t) <= ymax_i + eps*Vmax_i*sy_i ,   t = 1..p
// whose normals are rows of the dynamic matrix (MatG.m structure: G[(t,i),(c,j)] = s_ij(t-c)).  They are never
// stored: a row is regenerated from the step-response table when it is needed, the predicted outputs
//     y(k+t) = yfree(t) + G z
// are re-evaluated by the whole CTA after every step of the active-set method, and the free response comes from
// the channel recursion run through the dead time and closed-form (a^n table) beyond it.
// The solver is the dual active-set method of mpc_sim.cuh in block-wide form (Schur complement S = N'WN,
// inverse factor Li, V = W N Li', Givens removal, warm start on the carried set) with dense normals.
// Everything lives in shared memory; phases are separated by __syncthreads() (no warp-level assumptions), which
// is also what lets tests/host_emulation run this source thread-per-thread on a CPU.
#pragma once
#include <math.h>

#include "mpc_layout.h"
#include "mpc_sim.cuh"

#ifndef SOFT_THREADS
#define SOFT_THREADS 128
#endif
#define SOFT_SYNC() __syncthreads()
// |d2|^2 is a plain sum of squares here (no cancellation), so linear dependence is tested at the same level as the
// reference restatement (oracle DEP_TOL); the Schur form of mpc_sim.cuh needs the coarser SIM_DEP_TOL.
#ifndef SOFT_DEP_TOL
#define SOFT_DEP_TOL 1e-24   /* far below eps on purpose: see DEP_TOL in the reference restatement (band rows differ by the slack direction only) */
#endif
#ifndef SOFT_ITMAX_FACTOR
#define SOFT_ITMAX_FACTOR 400
#endif
#ifndef SOFT_WARM_START
#define SOFT_WARM_START 0   /* 1: carry the active set across samples.  Measured (round 2, tools/soft_probe.py, gpurun_out/probe4.log, probe5.log):
                               2.6 x fewer iterations and 2.6 x less time (203 vs 541 ms for 256 full-range candidates), but 70 of 256
                               candidates that the oracle resolves to 1e-7 then differ by more than 1e-6, with or without a rebuild
                               of the carried factor at every QP (so it is not factor drift; the cause is not established -- suspected:
                               the degenerate band rows, equal normals inside the dead time, make the multipliers of a carried set
                               non-unique).  Off: parity first. */
#endif
#ifndef SOFT_PREFER_LAST
#define SOFT_PREFER_LAST 1  /* cold start, but the constraints of the previous sample's final active set are tried first as pivots
                               (any violated constraint is a valid Goldfarb-Idnani pivot: the optimum is the same, the path
                               shorter, and such a pivot needs neither the predicted outputs of the whole horizon nor the
                               arg-min over all 2 ny p + 4 nu m rows) */
#endif
#ifndef SOFT_POLISH
#define SOFT_POLISH 1   /* at exit recompute the optimum from the final active set alone, z = z_unc + J1 R^-T (b_A - N_A' z_unc), instead of
                           keeping the sum of the steps that led there: the result then depends on the set, not on the path */
#endif
#ifndef SOFT_TCH
#define SOFT_TCH 8   /* prediction steps per thread in the evaluation of y = yfree + G z (SOFT_THREADS / ny threads x SOFT_TCH >= pmax) */
#endif
#ifndef SOFT_REFRESH_ROT
#define SOFT_REFRESH_ROT 0   /* rotations after which J is rebuilt from H^-1 (cold-start mode).  0 = at every constrained QP: J drifts
                               under the Givens rotations (cond(H) reaches 1e12 here), and a drifting J made the result depend on
                               the rounding of the build -- with FMA contraction a well-posed candidate left the oracle's trajectory
                               by 7e-9 at one sample and by 1e-3 at the end of the run, without it (round 1's -fmad=false) it did
                               not.  Refreshed at every QP both builds agree with the oracle and with each other (+4 % time;
                               tools/soft_probe.py, gpurun_out/probe2.log) */
#endif
#ifdef MPC_SIMT_EMULATION
#include <cstdio>
extern double *g_soft_dump;
extern int g_soft_dump_k, g_soft_k;
#endif

// Plant-model mismatch validation run (Shell3x3.m:271-286: options.Model = plant): the real process and the controller's
// state estimator (restated Toolbox default, mpcgpu/estimator.py E1-E4).  Device-resident, set by mpcgpu_set_mismatch.
struct MpcEst {
    double a[MPC_MAXY * MPC_MAXW], b0[MPC_MAXY * MPC_MAXW], b1[MPC_MAXY * MPC_MAXW];   // the real plant's channels
    int d[MPC_MAXY * MPC_MAXW];
    int hl;      // delay-line states per MV in the gain's state order
    int hlp;     // length of the real plant's input histories (its longest delay + 2)
    const double *gain;   // (ny*nw + nu*hl + ny) x ny, row-major: x_c(k|k) = x_c(k|k-1) + gain * (y - C x_c)
};
// extra shared memory of the validation kernel: xp, histp, xod, e, gvec, wy2
static MPC_HD size_t soft_est_doubles(const MpcLayout &L, int nu, int P, int hlp) {
    return (size_t)L.ny * L.nw + (size_t)L.nw * hlp + 3 * (size_t)L.ny + (size_t)nu * P + 2;
}
static MPC_HD size_t soft_smem_doubles(const MpcLayout &L, int nu, int P) {
    const int R = nu * P, NV = R + 1, QM = NV, nch = L.ny * L.nw, HL = sim_hl(L), nrow = L.ny * L.pmax;
    size_t n = 0;
    n += (L.nst + 1) & ~1;                    // st
    n += 2 * nch + (size_t)L.nw * HL;         // x, xol, hist
    n += 4 * nch;                             // cha, chb0, chb1, chg
    n += (size_t)SIM_CH * (2 * L.ny + L.nd);  // sig
    n += R + 4 * nu + nu;                     // uopt, bnd, ucur
    n += (size_t)nch * HL + 2 * nch;          // xfh, base, dev
    n += 7 * (size_t)NV;                      // z, lvl, w, wsc, dir, nrm, zu
    n += 4 * (size_t)QM;                      // g, l, rr, mu
    n += (size_t)QM * NV + (size_t)QM * QM;   // V, Li
    n += 2 * (size_t)nrow;                    // yfree, ypred
    n += SOFT_THREADS + 8;                    // reduction values, scalars
    n += (3 * nch + L.nst + R + 3 * QM + SOFT_THREADS + 16 + (nrow + 3) / 4 + 1) / 2 + 2;   // ints
    return n;
}

struct SoftSm {
    double *st, *x, *xol, *hist, *cha, *chb0, *chb1, *chg, *sig, *uopt, *bnd, *ucur, *xfh, *base, *dev;
    double *z, *lvl, *w, *wsc, *dir, *nrm, *zu, *g, *l, *rr, *mu, *V, *Li, *yfree, *ypred, *red, *sc;
    int *chd, *chj, *cht0, *role, *amask, *act, *dflag, *redi, *misc, *pref;
    unsigned char *ovmask;
};

// deterministic block reductions (fixed order: shuffle tree inside each warp, then the warps' partial results in warp order):
// every thread returns the result.  Two block barriers each; the first version (a seven-level tree through shared memory,
// eight barriers) was a fifth of an active-set iteration.
static_assert(SOFT_THREADS % 32 == 0 && SOFT_THREADS >= 32 && SOFT_THREADS <= 256, "whole warps");
__device__ __forceinline__ double soft_sum(double v, const SoftSm &sm) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) sm.red[tid >> 5] = v;
    SOFT_SYNC();
    double r = sm.red[0];
#pragma unroll
    for (int w = 1; w < SOFT_THREADS / 32; ++w) r += sm.red[w];
    SOFT_SYNC();
    return r;
}
// minimum value and, among equal values, the smallest non-negative index
__device__ __forceinline__ void soft_argmin(double &v, int &i, const SoftSm &sm) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov < v || (ov == v && oi >= 0 && (i < 0 || oi < i))) { v = ov; i = oi; }
    }
    if ((tid & 31) == 0) { sm.red[tid >> 5] = v; sm.redi[tid >> 5] = i; }
    SOFT_SYNC();
    v = sm.red[0]; i = sm.redi[0];
#pragma unroll
    for (int w = 1; w < SOFT_THREADS / 32; ++w) {
        const double ov = sm.red[w];
        const int oi = sm.redi[w];
        if (ov < v || (ov == v && oi >= 0 && (i < 0 || oi < i))) { v = ov; i = oi; }
    }
    SOFT_SYNC();
}

// Constraint id: type | index << 3.  types 0..3: MV rate low / rate high / level low / level high on row r = j*P + c;
// 4: ymax row, 5: ymin row (row = (t-1)*ny + i);  6: eps >= 0.
//
// Factorisation (Goldfarb & Idnani 1983, square-root form).  With H = L L', the solver carries
//     J  (NV x NE, NE = nu*m + 1 live variables):  J J' = H^-1,  J = L^-T Q
//     Li (q x q lower triangular) = R^-T, where  J' N = [R ; 0]  for the active normals N
// so that J1 (first q columns) = H^-1 N R^-1 spans the active normals and J2 is an orthonormal (in the H metric)
// basis of their complement.  For a candidate constraint n:  d = J' n,  step direction = J2 d2,  curvature
// rho = |d2|^2 (a sum of squares -- forming the Schur complement N'H^-1N explicitly, as the MV-only kernel does,
// would lose the slack direction here: its weight rho_eps is ~1e7 times the move weights), multiplier direction
// r = R^-1 d1 = Li' d1.  Adding n rotates d2 onto its first component (Givens on J2's columns), dropping a
// constraint rotates Li's rows / J1's columns (mpc_sim.cuh remove_at) and hands the freed column back to J2.
template <int NU, int P>
struct SoftQP {
    static constexpr int R = NU * P;
    static constexpr int NV = R + 1;
    static constexpr int QM = NV;
    const MpcLayout &L;
    const MpcTables &T;
    SoftSm sm;
    const double *W;   // global: R x R inverse Hessian (padded layout)
    int p, m, q, tid, n_rot, ne, npref;
    double *j0g;     // per-run global scratch for J0 (NV x ne doubles) or nullptr
    int j0_valid;
    unsigned long long n_con, n_it;
    int qmax;

    __device__ __forceinline__ SoftQP(const MpcLayout &L_, const MpcTables &T_) : L(L_), T(T_) {}

    __device__ __forceinline__ double S(int i, int j, int n) const { return __ldg(T.ST + (size_t)(i * NU + j) * T.st_stride + n); }
    __device__ __forceinline__ int rmap(int e) const { return e >= NU * m ? R : (e / m) * P + (e % m); }      // live variable -> row
    __device__ __forceinline__ int emap(int r) const { return r == R ? NU * m : ((r % P) < m ? (r / P) * m + (r % P) : -1); }
    // G_row(row)' x over the move rows
    __device__ __forceinline__ double ov_dot(int row, const double *x) const {
        const int t = row / L.ny + 1, i = row - (t - 1) * L.ny;
        const int cm = t < m ? t : m;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NU; ++j) {
            const double *s = T.ST + (size_t)(i * NU + j) * T.st_stride + t;
            double a = 0.0;
            for (int c = 0; c < cm; ++c) a = fma(__ldg(s - c), x[j * P + c], a);
            acc += a;
        }
        return acc;
    }
    // slack of constraint cid at the published iterate (sm.z, sm.lvl); OV rows evaluate their own prediction
    __device__ __forceinline__ double slack_of(int cid) const {
        const int type = cid & 7, k = cid >> 3;
        if (type < 4) {
            const double *b = sm.bnd + 4 * (k / P);
            switch (type) {
                case 0: return sm.z[k] - b[0];
                case 1: return b[1] - sm.z[k];
                case 2: return sm.lvl[k] - b[2];
                default: return b[3] - sm.lvl[k];
            }
        }
        if (type == 6) return sm.z[R];
        const int i = k % L.ny;
        const double y = sm.yfree[k] + ov_dot(k, sm.z);
        return type == 4 ? L.ymax[i] + sm.z[R] * L.emax[i] - y : y - L.ymin[i] + sm.z[R] * L.emin[i];
    }
    // per-input running sums over the horizon index: dst[j*P+c] = add_j + sum_{c'<=c} src[j*P+c']
    __device__ __forceinline__ void scan(const double *src, double *dst, const double *add) const {
        if (tid < NU) {
            double acc = add ? add[tid] : 0.0;
            for (int c = 0; c < P; ++c) { acc += src[tid * P + c]; dst[tid * P + c] = acc; }
        }
    }
    // dense normal of constraint cid into sm.nrm (NV entries), then d = J' n into sm.g (NE entries); returns |d|^2
    __device__ __forceinline__ double project(int cid) const {
        const int type = cid & 7, k = cid >> 3;
        for (int r = tid; r < NV; r += SOFT_THREADS) {
            double v = 0.0;
            const int j = r / P, c = r - j * P;
            if (type < 2) v = (r == k) ? ((type & 1) ? -1.0 : 1.0) : 0.0;
            else if (type < 4) v = (r < R && j == k / P && c <= k % P && c < m) ? ((type & 1) ? -1.0 : 1.0) : 0.0;
            else if (type == 6) v = (r == R) ? 1.0 : 0.0;
            else {
                const int t = k / L.ny + 1, i = k - (t - 1) * L.ny;
                if (r == R) v = type == 4 ? L.emax[i] : L.emin[i];
                else v = (c < m && c < t) ? (type == 4 ? -S(i, j, t - c) : S(i, j, t - c)) : 0.0;
            }
            sm.nrm[r] = v;
        }
        SOFT_SYNC();
        double part = 0.0;
        for (int e = tid; e < ne; e += SOFT_THREADS) {
            const double *col = sm.V + (size_t)e * NV;
            double a0 = 0.0, a1 = 0.0;
            int r = 0;
            for (; r + 1 < NV; r += 2) { a0 = fma(col[r], sm.nrm[r], a0); a1 = fma(col[r + 1], sm.nrm[r + 1], a1); }
            a0 = fma(col[r], sm.nrm[r], a0);   // NV is odd
            const double dd = a0 + a1;
            sm.g[e] = dd;
            part += dd * dd;
        }
        return soft_sum(part, sm);
    }
    // l = Li g over the first q entries (thread a owns row a); rr/out = Li' l
    __device__ __forceinline__ void tri_lower() const {
        for (int a = tid; a < q; a += SOFT_THREADS) {
            const double *row = sm.Li + (size_t)a * QM;
            double a0 = 0.0, a1 = 0.0;
            int b = 0;
            for (; b + 1 <= a; b += 2) { a0 = fma(row[b], sm.g[b], a0); a1 = fma(row[b + 1], sm.g[b + 1], a1); }
            if (b <= a) a0 = fma(row[b], sm.g[b], a0);
            sm.l[a] = a0 + a1;
        }
        SOFT_SYNC();
    }
    __device__ __forceinline__ void tri_upper(const double *lv, double *out) const {
        for (int a = tid; a < q; a += SOFT_THREADS) {
            double acc = 0.0;
            for (int b = a; b < q; ++b) acc = fma(sm.Li[(size_t)b * QM + a], lv[b], acc);
            out[a] = acc;
        }
        SOFT_SYNC();
    }
    // x += sign * J[:, e0:e1) coef[e0:e1)
    __device__ __forceinline__ void add_J(int e0, int e1, const double *coef, double sign, double *x) const {
        for (int r = tid; r < NV; r += SOFT_THREADS) {
            double acc = 0.0;
            for (int e = e0; e < e1; ++e) acc = fma(coef[e], sm.V[(size_t)e * NV + r], acc);
            x[r] += sign * acc;
        }
        SOFT_SYNC();
    }
    __device__ __forceinline__ void set_mask(int cid, bool on) const {
        if (tid == 0) {
            const int type = cid & 7, k = cid >> 3;
            if (type < 4) { if (on) sm.amask[k] |= (1 << type); else sm.amask[k] &= ~(1 << type); }
            else if (type == 6) { sm.misc[4] = on ? 1 : 0; }
            else { const unsigned char b = type == 4 ? 1 : 2; if (on) sm.ovmask[k] |= b; else sm.ovmask[k] &= (unsigned char)~b; }
        }
    }
    // J0 = Cholesky factor of H^-1 over the live variables (W is H^-1 of the moves; the slack is decoupled)
    __device__ __forceinline__ int factor_init() {
        const int nzu = NU * m;
        if (j0g && j0_valid) {   // J0 depends on the candidate only: computed once per run, afterwards a copy (ncu r2: the
            // Cholesky with its runtime index maps was 20 % of the kernel's instructions once J is refreshed at every QP)
            for (int idx = tid; idx < NV * ne; idx += SOFT_THREADS) sm.V[idx] = j0g[idx];
            SOFT_SYNC();
            q = 0;
            n_rot = 0;
            return 0;
        }
        for (int idx = tid; idx < NV * ne; idx += SOFT_THREADS) {
            const int e2 = idx / NV, r = idx - e2 * NV;
            double v = 0.0;
            if (r == R) v = (e2 == nzu) ? 1.0 / L.rho_ecr : 0.0;
            else if (e2 < nzu && emap(r) >= 0) v = __ldg(W + (size_t)r * R + rmap(e2));
            sm.V[idx] = v;
        }
        SOFT_SYNC();
        for (int k = 0; k < ne; ++k) {
            const double dkk = sm.V[(size_t)k * NV + rmap(k)];
            SOFT_SYNC();
            if (!(dkk > 0.0)) return 3;
            const double ckk = sqrt(dkk);
            for (int r = tid; r < NV; r += SOFT_THREADS) {
                const int e = emap(r);
                double *x = sm.V + (size_t)k * NV + r;
                *x = (e < k) ? 0.0 : (e == k ? ckk : *x / ckk);
            }
            SOFT_SYNC();
            const int nt = ne - k - 1;
            for (int idx = tid; idx < nt * NV; idx += SOFT_THREADS) {
                const int e2 = k + 1 + idx / NV, r = idx - (idx / NV) * NV;
                if (emap(r) >= e2) sm.V[(size_t)e2 * NV + r] -= sm.V[(size_t)k * NV + r] * sm.V[(size_t)k * NV + rmap(e2)];
            }
            SOFT_SYNC();
        }
        if (j0g) {
            for (int idx = tid; idx < NV * ne; idx += SOFT_THREADS) j0g[idx] = sm.V[idx];
            j0_valid = 1;
            SOFT_SYNC();
        }
        q = 0;
        n_rot = 0;
        return 0;
    }

#ifdef MPC_SIMT_EMULATION
    // debug self-check (emulation only): J J' = H^-1 and J2' N = 0
    __device__ __forceinline__ void self_check(const char *tag) {
        SOFT_SYNC();
        if (tid == 0 && g_soft_dump && g_soft_dump_k == g_soft_k) {
            double e1 = 0.0, e2 = 0.0, e3 = 0.0;
            for (int r = 0; r < NV; ++r)
                for (int r2 = 0; r2 < NV; ++r2) {
                    double acc = 0.0;
                    for (int e = 0; e < ne; ++e) acc += sm.V[(size_t)e * NV + r] * sm.V[(size_t)e * NV + r2];
                    double ref = 0.0;
                    if (r == R || r2 == R) ref = (r == R && r2 == R) ? 1.0 / L.rho_ecr : 0.0;
                    else if (emap(r) >= 0 && emap(r2) >= 0) ref = W[(size_t)r * R + r2];
                    e1 = fmax(e1, fabs(acc - ref) / (fabs(W[0]) + 1.0 / L.rho_ecr));
                }
            for (int a = 0; a < q; ++a) {
                const int cid = sm.act[a], type = cid & 7, k = cid >> 3;
                double nv[NV];
                for (int r = 0; r < NV; ++r) {
                    double v = 0.0; const int j = r / P, c = r - j * P;
                    if (type < 2) v = (r == k) ? ((type & 1) ? -1.0 : 1.0) : 0.0;
                    else if (type < 4) v = (r < R && j == k / P && c <= k % P && c < m) ? ((type & 1) ? -1.0 : 1.0) : 0.0;
                    else if (type == 6) v = (r == R) ? 1.0 : 0.0;
                    else { const int t = k / L.ny + 1, i = k - (t - 1) * L.ny;
                           if (r == R) v = type == 4 ? L.emax[i] : L.emin[i]; else v = (c < m && c < t) ? (type == 4 ? -S(i, j, t - c) : S(i, j, t - c)) : 0.0; }
                    nv[r] = v;
                }
                double dn = 0.0;
                double dd[NV];
                for (int e = 0; e < ne; ++e) { double acc = 0.0; for (int r = 0; r < NV; ++r) acc += sm.V[(size_t)e * NV + r] * nv[r]; dd[e] = acc; dn += acc * acc; }
                for (int e = q; e < ne; ++e) e2 = fmax(e2, fabs(dd[e]) / sqrt(dn));
                // Li (J1'n_a) should be e_a
                for (int c2 = 0; c2 < q; ++c2) { double acc = 0.0; for (int b = c2; b < q; ++b) acc += sm.Li[(size_t)b * QM + c2] * dd[b]; e3 = fmax(e3, fabs(acc - (c2 == a ? 1.0 : 0.0))); }
            }
            // (valid when z_unc = 0, i.e. zero tracking weights) z must lie in span(J1): least-squares residual
            double e4 = 0.0, e5 = 0.0;
            if (q > 0) {
                static double Gm[64 * 65];
                for (int a = 0; a < q; ++a) {
                    for (int b = 0; b < q; ++b) { double acc = 0.0; for (int r = 0; r < NV; ++r) acc += sm.V[(size_t)a * NV + r] * sm.V[(size_t)b * NV + r]; Gm[a * 65 + b] = acc; }
                    double acc = 0.0; for (int r = 0; r < NV; ++r) acc += sm.V[(size_t)a * NV + r] * sm.z[r]; Gm[a * 65 + q] = acc;
                }
                for (int c2 = 0; c2 < q; ++c2) {
                    int pv_ = c2; for (int r2 = c2 + 1; r2 < q; ++r2) if (fabs(Gm[r2 * 65 + c2]) > fabs(Gm[pv_ * 65 + c2])) pv_ = r2;
                    for (int k2 = 0; k2 <= q; ++k2) { double t_ = Gm[c2 * 65 + k2]; Gm[c2 * 65 + k2] = Gm[pv_ * 65 + k2]; Gm[pv_ * 65 + k2] = t_; }
                    for (int r2 = 0; r2 < q; ++r2) if (r2 != c2) { double f_ = Gm[r2 * 65 + c2] / Gm[c2 * 65 + c2]; for (int k2 = c2; k2 <= q; ++k2) Gm[r2 * 65 + k2] -= f_ * Gm[c2 * 65 + k2]; }
                }
                double zn = 0.0;
                for (int r = 0; r < NV; ++r) { double acc = 0.0; for (int a = 0; a < q; ++a) acc += sm.V[(size_t)a * NV + r] * Gm[a * 65 + q] / Gm[a * 65 + a]; e4 = fmax(e4, fabs(sm.z[r] - acc)); zn = fmax(zn, fabs(sm.z[r])); }
                e4 /= (zn + 1e-300);
                for (int a = 0; a < q; ++a) e5 = fmax(e5, fabs(slack_of(sm.act[a])));
            }
            printf("  check[%s] q=%d  |JJ'-W| %.2e  |J2'N| %.2e  |Li J1'N - I| %.2e  z off span(J1) %.2e  active slack %.2e\n", tag, q, e1, e2, e3, e4, e5);
        }
        SOFT_SYNC();
    }
#else
    __device__ __forceinline__ void self_check(const char *) {}
#endif
    // constraint cid becomes active: d = J'n is in sm.g, rr = Li' d1 in sm.rr.  Rotates d2 onto its first component.
    __device__ __forceinline__ void commit(int cid, double rho, double mu_new) {
        const int nt = ne - q;   // length of d2
        // suffix norms sigma_e = |d[e:]|, e = q..ne-1, in sm.w (indexed e - q)
        for (int t = tid; t < nt; t += SOFT_THREADS) {
            double ss = 0.0;
            for (int e = q + t; e < ne; ++e) ss = fma(sm.g[e], sm.g[e], ss);
            sm.w[t] = sqrt(ss);
        }
        SOFT_SYNC();
        // the rotation (c_j, s_j) that folds component j into j-1 is the same for every row of J: computed once (it was two fp64
        // divisions per row and step: 9 % of the kernel's instructions)
        for (int t = 1 + tid; t < nt; t += SOFT_THREADS) {
            const double sj1 = sm.w[t - 1], sj = sm.w[t];
            double c = 1.0, s = 0.0;
            if (sj1 > 0.0) { c = sm.g[q + t - 1] / sj1; s = sj / sj1; }
            sm.wsc[t] = c; sm.nrm[t] = s;
        }
        SOFT_SYNC();
        for (int r = tid; r < NV; r += SOFT_THREADS) {
            // the carried component must be +sigma: flip the last column if its d entry is negative (J J' is unchanged)
            double carry = sm.g[ne - 1] < 0.0 ? -sm.V[(size_t)(ne - 1) * NV + r] : sm.V[(size_t)(ne - 1) * NV + r];
            for (int j = ne - 1; j > q; --j) {
                const double c = sm.wsc[j - q], s = sm.nrm[j - q];
                const double o = sm.V[(size_t)(j - 1) * NV + r];
                sm.V[(size_t)j * NV + r] = c * carry - s * o;
                carry = c * o + s * carry;
            }
            sm.V[(size_t)q * NV + r] = carry;
        }
        const double isr = 1.0 / sqrt(rho);
        for (int b = tid; b < q; b += SOFT_THREADS) sm.Li[(size_t)q * QM + b] = -sm.rr[b] * isr;
        if (tid == 0) { sm.Li[(size_t)q * QM + q] = isr; sm.act[q] = cid; sm.mu[q] = mu_new; }
        set_mask(cid, true);
        SOFT_SYNC();
        q += 1;
        n_rot += 1;
    }
    // Givens removal of the constraint at list position a (derivation: mpc_sim.cuh remove_at); the freed column
    // goes back to the complement basis J2 at position q-1.
    __device__ __forceinline__ void remove_at(int a) {
        const int nrot = q - 1 - a;
        set_mask(sm.act[a], false);
        if (nrot > 0) {
            for (int t = tid; t < nrot; t += SOFT_THREADS) {
                double ss = 0.0;
                for (int i = a; i <= a + t; ++i) { const double v = sm.Li[(size_t)i * QM + a]; ss = fma(v, v, ss); }
                const double y = sm.Li[(size_t)(a + t + 1) * QM + a];
                const double inv = 1.0 / sqrt(fma(y, y, ss));
                sm.w[t] = sqrt(ss) * inv;
                sm.wsc[t] = y * inv;
            }
            SOFT_SYNC();
            for (int r = tid; r < NV; r += SOFT_THREADS) {
                double cv = sm.V[(size_t)a * NV + r];
                for (int t = 0; t < nrot; ++t) {
                    const int k = a + t;
                    const double o = sm.V[(size_t)(k + 1) * NV + r];
                    sm.V[(size_t)k * NV + r] = sm.w[t] * o - sm.wsc[t] * cv;
                    cv = sm.w[t] * cv + sm.wsc[t] * o;
                }
                sm.V[(size_t)(q - 1) * NV + r] = cv;
            }
            const int j = tid;
            double cy = (j < a) ? sm.Li[(size_t)a * QM + j] : 0.0;
            for (int t = 0; t < nrot; ++t) {
                const int k = a + t;
                SOFT_SYNC();
                if (j < q && j != a && j <= k + 1) {
                    const double o = sm.Li[(size_t)(k + 1) * QM + j];
                    sm.Li[(size_t)k * QM + (j < a ? j : j - 1)] = sm.w[t] * o - sm.wsc[t] * cy;
                    cy = sm.w[t] * cy + sm.wsc[t] * o;
                }
            }
            SOFT_SYNC();
            int ca = 0; double cm = 0.0;
            const int pos = a + 1 + tid;
            if (pos < q) { ca = sm.act[pos]; cm = sm.mu[pos]; }
            SOFT_SYNC();
            if (pos < q) { sm.act[pos - 1] = ca; sm.mu[pos - 1] = cm; }
        }
        SOFT_SYNC();
        q -= 1;
        n_rot += 1;
    }
    __device__ __forceinline__ void drop_flagged() {
        for (int a = q - 1; a >= 0; --a) {
            const int fl = sm.dflag[a];
            SOFT_SYNC();
            if (fl) remove_at(a);
        }
    }
    // rebuild the factorisation of the carried set from H^-1 (drift control after many rotations)
    __device__ __forceinline__ int rebuild() {
        const int qn = q;
        for (int a = tid; a < qn; a += SOFT_THREADS) { sm.dir[a] = sm.mu[a]; sm.dflag[a] = sm.act[a]; }   // parked copies
        SOFT_SYNC();
        const int rc = factor_init();
        if (rc) return rc;
        for (int a = 0; a < qn; ++a) {
            const int cid = sm.dflag[a];
            const double mua = sm.dir[a];
            SOFT_SYNC();
            const double dn2 = project(cid);
            double part = 0.0;
            for (int e = q + tid; e < ne; e += SOFT_THREADS) part += sm.g[e] * sm.g[e];
            const double rho = soft_sum(part, sm);
            if (rho > SOFT_DEP_TOL * dn2) {
                tri_upper(sm.g, sm.rr);
                commit(cid, rho, mua);
            } else {
                set_mask(cid, false);
                SOFT_SYNC();
            }
        }
        return 0;
    }
    // levels and predicted outputs of the current iterate
    __device__ __forceinline__ void evaluate(bool with_outputs) const {
        scan(sm.z, sm.lvl, sm.ucur);
        if (with_outputs) {
            // y = yfree + G z for every prediction row.  A thread owns output i and SOFT_TCH consecutive prediction steps
            // t0 .. t0+TCH-1; every step-response value s(n) it needs is loaded ONCE and feeds all the rows it belongs to
            // (row t takes s(t - c) x_c): TCH + P - 1 table loads and TCH x P multiply-adds per input with every index a
            // compile-time constant (ncu r2: the row-by-row form, ov_dot, was 19 % of the kernel's instructions, one load per
            // multiply-add).  The values are visited in descending n, i.e. ascending c for every row, and zeros stand where
            // ov_dot's loop ends (c >= min(t, m)): the same products in the same order, bit-identical sums.
            constexpr int TCH = SOFT_TCH;
            const int ny = L.ny;
            const int i = tid % ny, t0 = 1 + (tid / ny) * TCH;
            if (tid < ny * (SOFT_THREADS / ny) && t0 <= p) {
                double acc[TCH];
#pragma unroll 1
                for (int j = 0; j < NU; ++j) {
                    const double *s = T.ST + (size_t)(i * NU + j) * T.st_stride;
                    double xr[P], a[TCH];
#pragma unroll
                    for (int c = 0; c < P; ++c) xr[c] = c < m ? sm.z[j * P + c] : 0.0;
#pragma unroll
                    for (int k = 0; k < TCH; ++k) a[k] = 0.0;
#pragma unroll
                    for (int c0 = -(TCH - 1); c0 < P; ++c0) {            // n = t0 - c0 descends as c0 rises
                        const int n = t0 - c0;
                        const double sv = n >= 1 ? __ldg(s + n) : 0.0;
#pragma unroll
                        for (int k = 0; k < TCH; ++k)
                            if (k + c0 >= 0 && k + c0 < P) a[k] = fma(sv, xr[k + c0], a[k]);   // row t0 + k, move c = k + c0
                    }
#pragma unroll
                    for (int k = 0; k < TCH; ++k) acc[k] = j == 0 ? a[k] : acc[k] + a[k];
                }
#pragma unroll
                for (int k = 0; k < TCH; ++k)
                    if (t0 + k <= p) { const int row = (t0 + k - 1) * ny + i; sm.ypred[row] = sm.yfree[row] + acc[k]; }
            }
        }
        SOFT_SYNC();
    }

    // z (shared) in: unconstrained optimum (slack row 0); out: constrained optimum.  0 ok, 1 infeasible, 2 iteration cap.
    __device__ __forceinline__ int solve() {
        int it = 0;
        const int itmax = SOFT_ITMAX_FACTOR * (NU * m + 10);
#if !SOFT_WARM_START
        // Cold start, like the reference restatement: with the band constraints of Shell7x5 several output rows
        // share (almost) one normal -- [0 .. 0, e_i] inside the dead time -- and a carried set that holds two of
        // them makes R nearly singular; the warm-started optimum then loses ~8 digits (measured), the cold one
        // follows the oracle's pivot sequence.  J J' = H^-1 holds for any rotation of J in exact arithmetic; in fp64 it
        // drifts, so J is refreshed from H^-1 (SOFT_REFRESH_ROT).
        for (int a = q - 1; a >= 0; --a) { set_mask(sm.act[a], false); SOFT_SYNC(); }
        q = 0;
        if (n_rot > SOFT_REFRESH_ROT) { const int rc = factor_init(); if (rc) return rc; }
#else
        if (n_rot > SOFT_REFRESH_ROT) { const int rc = q > 0 ? rebuild() : factor_init(); if (rc) return rc; }
#endif
        // ---- warm start on the carried set: mu = R^-1 R^-T (b_A - N_A' z_unc), shed negative multipliers ----
        while (q > 0) {
            evaluate(false);
            for (int a = tid; a < q; a += SOFT_THREADS) sm.g[a] = -slack_of(sm.act[a]);
            SOFT_SYNC();
            tri_lower();
            tri_upper(sm.l, sm.mu);
            double mumax = 0.0;
            for (int a = 0; a < q; ++a) mumax = fmax(mumax, fabs(sm.mu[a]));
            int nd_ = 0;
            for (int a = 0; a < q; ++a) nd_ |= (sm.mu[a] < -1e-12 * mumax);
            SOFT_SYNC();
            for (int a = tid; a < q; a += SOFT_THREADS) sm.dflag[a] = sm.mu[a] < -1e-12 * mumax;
            SOFT_SYNC();
            if (!nd_) break;
            it += 1;
            drop_flagged();
        }
        if (q > 0) {
            add_J(0, q, sm.l, 1.0, sm.z);
            for (int a = tid; a < q; a += SOFT_THREADS) if (sm.mu[a] < 0.0) sm.mu[a] = 0.0;
            SOFT_SYNC();
        }
#if SOFT_POLISH
        for (int r = tid; r < NV; r += SOFT_THREADS) sm.zu[r] = sm.z[r];
        SOFT_SYNC();
#endif
        // ---- Goldfarb-Idnani iterations ----
        int pptr = 0;
        // Control horizon close to the prediction horizon (p < 1.5 m): few prediction rows per move, the band QPs are at their
        // most degenerate and both the most-violated rule and the previous set churn (hundreds of add/drop pivots per QP: the
        // heaviest runs of a Shell7x5 population).  Walking the violated constraints in horizon order -- MV limits by move
        // index, then output rows by prediction step -- needs a third to a quarter of the iterations there (measured on the
        // oracle's pivot rule 1: 591 k against 787 k / 906 k iterations for the 73 such candidates of 2048; it is far worse
        // everywhere else, 38 M against 4.3 M for N >= 4 Nu).  Any violated constraint is a valid pivot: same optimum.
        const bool index_order = 2 * p < 3 * m;
        for (;;) {
            double bv = -SIM_VIOL_TOL;
            int bi = -1;
#if SOFT_PREFER_LAST
            if (pptr < npref && !index_order) {   // next constraint of the previous sample's set that is violated and not active (uniform)
                evaluate(false);
                while (pptr < npref) {
                    const int cid = sm.pref[pptr++];
                    const int type = cid & 7, k = cid >> 3;
                    const int isact = type < 4 ? (sm.amask[k] >> type) & 1 : (type == 6 ? sm.misc[4] : (sm.ovmask[k] >> (type == 4 ? 0 : 1)) & 1);
                    if (isact) continue;
                    const double sl = slack_of(cid);
                    if (sl < -SIM_VIOL_TOL) { bi = cid; bv = sl; break; }
                }
                SOFT_SYNC();
            }
            if (bi < 0) {
#endif
            evaluate(true);
            int bkey = 0x7fffffff;   // index_order: smallest horizon-order key among the violated constraints
            auto take = [&](double s, int id) {
                if (index_order) {
                    if (s < -SIM_VIOL_TOL) {
                        const int type = id & 7, k = id >> 3;
                        const int key = type < 4 ? (((k % P) * NU + k / P) * 4 + type) : (type == 6 ? 0x7ffffffe : 4 * R + 2 * k + (type - 4));
                        if (key < bkey) { bkey = key; bv = s; bi = id; }
                    }
                } else if (s < bv || (s == bv && bi >= 0 && id < bi)) { bv = s; bi = id; }
            };
            for (int r = tid; r < R; r += SOFT_THREADS) {
                const int j = r / P, c = r - j * P;
                if (c >= m) continue;
                const double *b = sm.bnd + 4 * j;
                const int am = sm.amask[r];
                if (!(am & 1)) take(sm.z[r] - b[0], 0 | (r << 3));
                if (!(am & 2)) take(b[1] - sm.z[r], 1 | (r << 3));
                if (!(am & 4)) take(sm.lvl[r] - b[2], 2 | (r << 3));
                if (!(am & 8)) take(b[3] - sm.lvl[r], 3 | (r << 3));
            }
            const double eps = sm.z[R];
            for (int row = tid; row < L.ny * p; row += SOFT_THREADS) {
                const int i = row % L.ny;
                const int om = sm.ovmask[row];
                const double y = sm.ypred[row];
                if (!(om & 1)) take(L.ymax[i] + eps * L.emax[i] - y, 4 | (row << 3));   // an infinite bound never wins
                if (!(om & 2)) take(y - L.ymin[i] + eps * L.emin[i], 5 | (row << 3));
            }
            if (tid == 0 && !sm.misc[4]) take(eps, 6);
            if (index_order) {   // block-wide minimum of the key; the winner's slack rides along
                double kv = (double)bkey;
                int ki = bi;
                soft_argmin(kv, ki, sm);
                const int winner = ki;
                double wv_ = (bi == winner && winner >= 0) ? bv : 0.0;
                bv = soft_sum(wv_, sm);   // exactly one thread holds the winner
                bi = winner;
            } else
            soft_argmin(bv, bi, sm);
#if SOFT_PREFER_LAST
            }
#endif
            if (bi < 0) break;
            const int pv = bi;
            double sp = bv, mu_p = 0.0;
            for (;;) {
                if (++it > itmax) { n_it += it; return 2; }
                const double dn2 = project(pv);            // d = J'n in sm.g; d1 = R^-T N'H^-1 n
                double part = 0.0;
                for (int e = q + tid; e < ne; e += SOFT_THREADS) part += sm.g[e] * sm.g[e];
                const double rho = soft_sum(part, sm);     // |d2|^2 = curvature along the step direction J2 d2
                tri_upper(sm.g, sm.rr);                    // r = R^-1 d1
                const int dependent = !(rho > SOFT_DEP_TOL * dn2);
                double t1 = SIM_INF;
                int l1 = -1;
                for (int a = tid; a < q; a += SOFT_THREADS) {
                    const double ra = sm.rr[a];
                    if (ra > 0.0) { t1 = sm.mu[a] / ra; l1 = a; }
                }
                soft_argmin(t1, l1, sm);
                const double t2 = dependent ? SIM_INF : -sp / rho;
                const double t = t1 < t2 ? t1 : t2;
                if (!(t < SIM_INF)) { n_it += it; return 1; }
                const int full = !(dependent || t1 < t2);
                if (!dependent) {
                    for (int e = tid; e < ne; e += SOFT_THREADS) sm.dir[e] = t * sm.g[e];
                    SOFT_SYNC();
                    add_J(q, ne, sm.dir, 1.0, sm.z);
                    sp += t * rho;
                }
                for (int a = tid; a < q; a += SOFT_THREADS) {
                    sm.mu[a] -= t * sm.rr[a];
                    sm.dflag[a] = (a == l1) && !full;
                }
                mu_p += t;
                SOFT_SYNC();
                if (full) { commit(pv, rho, mu_p); self_check("add"); break; }
                drop_flagged();
                self_check("drop");
            }
        }
#if SOFT_POLISH
        if (q > 0) {
            for (int r = tid; r < NV; r += SOFT_THREADS) sm.z[r] = sm.zu[r];
            SOFT_SYNC();
            evaluate(false);
            for (int a = tid; a < q; a += SOFT_THREADS) sm.g[a] = -slack_of(sm.act[a]);
            SOFT_SYNC();
            tri_lower();
            add_J(0, q, sm.l, 1.0, sm.z);
            evaluate(false);
        }
#endif
        // ---- one Newton correction on the active constraints: z += J1 R^-T (-slack_A) ----
        if (q > 0) {
            double worst = 0.0;
            for (int a = tid; a < q; a += SOFT_THREADS) {
                const double sl = slack_of(sm.act[a]);
                sm.g[a] = -sl;
                worst = fmax(worst, fabs(sl));
            }
            int dummy = 0;
            worst = -worst;
            soft_argmin(worst, dummy, sm);
            if (-worst > 1e-13) {
                tri_lower();
                add_J(0, q, sm.l, 1.0, sm.z);
            }
        }
        n_it += it;
        if (q > qmax) qmax = q;
#if SOFT_PREFER_LAST
        SOFT_SYNC();
        for (int a = tid; a < q; a += SOFT_THREADS) sm.pref[a] = sm.act[a];
        npref = q;
        SOFT_SYNC();
#endif
        return 0;
    }
};

// ------------------------------------------------------------------------------------------------
// One closed-loop run by one CTA.  Arguments as sim_run (mpc_sim.cuh) plus the prediction horizon p.
// ------------------------------------------------------------------------------------------------
// EST: the validation run against a real plant that differs from the model (MpcEst): the controller corrects its state
// with the estimator gain at every sample and the unconstrained optimum is formed from the corrected free response,
// z_unc = W G'Wy^2 (r - yfree), as the oracle does (the deviation-coordinate map M st has no column for a corrected
// u(k-1) slot).  GAM / RAW only, no open-loop pass.
template <int NU, int P, bool EST = false>
__device__ __forceinline__ int soft_run(const MpcLayout &L, const MpcTables &T, int p, int m, const double *__restrict__ Mg,
                                        const double *__restrict__ Wg, int mode, int sel, double *smem, const MpcRunOut &out,
                                        const MpcEst *E = nullptr, const double *delta = nullptr, double *j0g = nullptr) {
    constexpr int R = NU * P;
    constexpr int NV = R + 1;
    constexpr int QM = NV;
    const int tid = threadIdx.x;
    const int ny = L.ny, nd = L.nd, nw = L.nw, nch = ny * nw, nst = L.nst, nit = L.nit;
    const int HL = sim_hl(L);
    const int nsig = 2 * ny + nd;
    const int nrow = ny * L.pmax;
    SoftQP<NU, P> qp(L, T);
    SoftSm &sm = qp.sm;
    {
        double *q_ = smem;
        sm.st = q_; q_ += (nst + 1) & ~1;
        sm.x = q_; q_ += nch; sm.xol = q_; q_ += nch;
        sm.hist = q_; q_ += (size_t)nw * HL;
        sm.cha = q_; q_ += nch; sm.chb0 = q_; q_ += nch; sm.chb1 = q_; q_ += nch; sm.chg = q_; q_ += nch;
        sm.sig = q_; q_ += (size_t)SIM_CH * nsig;
        sm.uopt = q_; q_ += R; sm.bnd = q_; q_ += 4 * NU; sm.ucur = q_; q_ += NU;
        sm.xfh = q_; q_ += (size_t)nch * HL; sm.base = q_; q_ += nch; sm.dev = q_; q_ += nch;
        sm.z = q_; q_ += NV; sm.lvl = q_; q_ += NV; sm.w = q_; q_ += NV; sm.wsc = q_; q_ += NV; sm.dir = q_; q_ += NV; sm.nrm = q_; q_ += NV; sm.zu = q_; q_ += NV;
        sm.g = q_; q_ += QM; sm.l = q_; q_ += QM; sm.rr = q_; q_ += QM; sm.mu = q_; q_ += QM;
        sm.V = q_; q_ += (size_t)QM * NV; sm.Li = q_; q_ += (size_t)QM * QM;
        sm.yfree = q_; q_ += nrow; sm.ypred = q_; q_ += nrow;
        sm.red = q_; q_ += SOFT_THREADS; sm.sc = q_; q_ += 8;
        int *ip = (int *)q_;
        sm.chd = ip; ip += nch; sm.chj = ip; ip += nch; sm.cht0 = ip; ip += nch; sm.role = ip; ip += nst;
        sm.amask = ip; ip += R; sm.act = ip; ip += QM; sm.dflag = ip; ip += QM; sm.redi = ip; ip += SOFT_THREADS;
        sm.misc = ip; ip += 16; sm.pref = ip; ip += QM;
        sm.ovmask = (unsigned char *)ip;
    }
    // validation run: the real plant's state and the estimator's extra state, behind the base layout
    double *xp = nullptr, *histp = nullptr, *xod = nullptr, *einn = nullptr, *gvec = nullptr, *wy2 = nullptr;
    int HLP = 0, headp = 0;
    if (EST) {
        HLP = E->hlp;
        double *q_ = smem + soft_smem_doubles(L, NU, P);
        xp = q_; q_ += nch; histp = q_; q_ += (size_t)nw * HLP; xod = q_; q_ += ny; einn = q_; q_ += ny; wy2 = q_; q_ += ny; gvec = q_;
        for (int i = tid; i < nch; i += SOFT_THREADS) xp[i] = 0.0;
        for (int i = tid; i < nw * HLP; i += SOFT_THREADS) histp[i] = 0.0;
        for (int i = tid; i < ny; i += SOFT_THREADS) { xod[i] = 0.0; einn[i] = 0.0; const double w = delta[i] / L.sy[i]; wy2[i] = w * w; }
    }
    qp.W = Wg; qp.p = p; qp.m = m; qp.q = 0; qp.tid = tid; qp.npref = 0; qp.j0g = j0g; qp.j0_valid = 0; qp.n_rot = 0; qp.n_con = 0; qp.n_it = 0; qp.qmax = 0;
    qp.ne = NU * m + 1;
    // ---- one-time staging ----
    for (int ch = tid; ch < nch; ch += SOFT_THREADS) {
        sm.cha[ch] = L.a[ch]; sm.chb0[ch] = L.b0[ch]; sm.chb1[ch] = L.b1[ch]; sm.chg[ch] = L.gain[ch];
        sm.chd[ch] = L.d[ch]; sm.chj[ch] = ch % nw;
        sm.cht0[ch] = (L.d[ch] + 1 < p) ? L.d[ch] + 1 : p;
        sm.x[ch] = 0.0; sm.xol[ch] = 0.0;
    }
    for (int i = tid; i < nw * HL; i += SOFT_THREADS) sm.hist[i] = 0.0;
    for (int i = tid; i < R; i += SOFT_THREADS) { sm.uopt[i] = 0.0; sm.amask[i] = 0; }
    for (int i = tid; i < nrow; i += SOFT_THREADS) sm.ovmask[i] = 0;
    for (int j = tid; j < NU; j += SOFT_THREADS) {
        sm.bnd[4 * j + 0] = L.dumin[j]; sm.bnd[4 * j + 1] = L.dumax[j]; sm.bnd[4 * j + 2] = L.umin[j]; sm.bnd[4 * j + 3] = L.umax[j];
        sm.ucur[j] = 0.0;
    }
    if (tid == 0) sm.misc[4] = 0;
    for (int col = tid; col < nst; col += SOFT_THREADS) {
        int role;
        if (col < nch) role = 0 | (col << 2);
        else if (col >= L.stoff_e) role = 2 | ((col - L.stoff_e) << 2);
        else {
            int j = 0;
            while (j + 1 < nw && col >= L.stoff_h[j + 1]) ++j;
            role = 1 | (j << 2) | ((L.hq0[j] + (col - L.stoff_h[j])) << 12);
        }
        sm.role[col] = role;
    }
    SOFT_SYNC();
    int head = 0;
    int status = qp.factor_init();   // J0 J0' = H^-1 (3: H^-1 not positive definite)
    if (status) {
        if (out.cost && tid == 0) {
            if (mode == 1) for (int i = 0; i < ny; ++i) out.cost[i] = NAN;
            if (mode == 2) out.cost[0] = NAN;
        }
        return status;
    }
    const bool want_ol = EST ? false : ((mode != 1) || out.ys || out.uopt);
    double jnu = 0.0;
    double cost_acc = 0.0;   // thread holding column stoff_e + i accumulates output i

    auto setpoint = [&](int i, int k, double r_user) -> double {
        if (sel == -2) return r_user;
        return (sel == -1 || sel == i) ? (k >= L.inK - 1 ? 1.0 : 0.0) : 0.0;
    };
    auto held = [&](int j, const double *sigrow) -> double { return j < NU ? sm.ucur[j] : sigrow[2 * ny + (j - NU)]; };
    auto build_st = [&](const double *sigrow, int k, bool do_cost) {
        for (int col = tid; col < nst; col += SOFT_THREADS) {
            const int role = sm.role[col];
            const int kind = role & 3, a = (role >> 2) & 1023, b = role >> 12;
            double val;
            if (kind == 0) {
                val = sm.x[a] - sm.chg[a] * held(sm.chj[a], sigrow);
            } else if (kind == 1) {
                int pos = head + b;
                if (pos >= HL) pos -= HL;
                val = sm.hist[a * HL + pos] - held(a, sigrow);
            } else {
                const int i = a;
                double yi = 0.0, ysi = 0.0, gsum = 0.0;
                for (int j = 0; j < nw; ++j) {
                    yi += EST ? xp[i * nw + j] : sm.x[i * nw + j];   // validation run: the REAL plant's output is what is logged and costed
                    ysi += sm.xol[i * nw + j];
                    gsum += sm.chg[i * nw + j] * held(j, sigrow);
                }
                val = setpoint(i, k, sigrow[i]) - gsum;
                if (do_cost) {
                    const bool mine = (sel < 0 || sel == i);
                    if (out.y && mine) out.y[(size_t)i * nit + k] = yi;
                    if (out.ys && mine) out.ys[(size_t)i * nit + k] = ysi;
                    if (mode == 1) {
                        const double e = yi - sigrow[ny + i];
                        cost_acc += e * e;
                    } else if (mode == 2 && k >= L.inK - 1 && mine) {
                        const double e2 = yi - ysi, er = yi - sigrow[ny + i];
                        cost_acc += e2 * e2 + er * er;
                    }
                }
            }
            sm.st[col] = val;
        }
        SOFT_SYNC();
    };
    auto stage_signals = [&](int k0) {
        const int cnt = (nit - k0) < SIM_CH ? (nit - k0) : SIM_CH;
        for (int idx = tid; idx < cnt * nsig; idx += SOFT_THREADS) {
            const int kk = idx / nsig, c = idx - kk * nsig;
            double v;
            if (c < ny) v = T.r[(size_t)(k0 + kk) * ny + c];
            else if (c < 2 * ny) v = T.yref[(size_t)(c - ny) * nit + (k0 + kk)];
            else v = T.v[(size_t)(k0 + kk) * nd + (c - 2 * ny)];
            sm.sig[kk * nsig + c] = v;
        }
        SOFT_SYNC();
    };
    // z_unc = M st (M in global memory, [col][row]); slack row = 0.  Then the free response over the horizon.
    auto controller_move = [&](const double *sigrow) -> int {
        if (!EST)
        for (int r = tid; r < NV; r += SOFT_THREADS) {
            double a0 = 0.0, a1 = 0.0;
            if (r < R) {
                int sg = 0;
                for (; sg + 1 < nst; sg += 2) {
                    a0 = fma(__ldg(Mg + (size_t)sg * R + r), sm.st[sg], a0);
                    a1 = fma(__ldg(Mg + (size_t)(sg + 1) * R + r), sm.st[sg + 1], a1);
                }
                if (sg < nst) a0 = fma(__ldg(Mg + (size_t)sg * R + r), sm.st[sg], a0);
            }
            sm.z[r] = a0 + a1;
        }
        // free response: channel recursion through the dead time (ctx: inputs held at hv from k on) ...
        for (int ch = tid; ch < nch; ch += SOFT_THREADS) {
            const int j = sm.chj[ch], dd = sm.chd[ch], t0 = sm.cht0[ch];
            const double hv = held(j, sigrow);
            double xf = sm.x[ch];
            for (int t = 1; t <= t0; ++t) {
                const int r0 = t - dd, r1 = t - dd - 1;
                int p0 = head + (r0 >= 0 ? 0 : -r0 - 1); if (p0 >= HL) p0 -= HL;
                int p1 = head + (r1 >= 0 ? 0 : -r1 - 1); if (p1 >= HL) p1 -= HL;
                const double w0 = r0 >= 0 ? hv : sm.hist[j * HL + p0];
                const double w1 = r1 >= 0 ? hv : sm.hist[j * HL + p1];
                xf = sm.cha[ch] * xf + sm.chb0[ch] * w0 + sm.chb1[ch] * w1;
                sm.xfh[ch * HL + t - 1] = xf;
            }
            sm.base[ch] = sm.chg[ch] * hv;
            sm.dev[ch] = xf - sm.chg[ch] * hv;
        }
        SOFT_SYNC();
        // ... and closed form beyond it: x(t) = g hv + a^(t - t0) (x(t0) - g hv)
        for (int row = tid; row < ny * p; row += SOFT_THREADS) {
            const int t = row / ny + 1, i = row - (t - 1) * ny;
            double acc = 0.0;
            for (int j = 0; j < nw; ++j) {
                const int ch = i * nw + j, t0 = sm.cht0[ch];
                acc += (t <= t0) ? sm.xfh[ch * HL + t - 1]
                                 : fma(__ldg(T.PA + (size_t)ch * (L.pmax + 1) + (t - t0)), sm.dev[ch], sm.base[ch]);
            }
            sm.yfree[row] = EST ? acc + xod[i] : acc;   // + the estimated output disturbance (an integrator: constant over the horizon)
        }
        SOFT_SYNC();
        if (EST) {   // z_unc = W G'Wy^2 (r - yfree) from the corrected free response (oracle: ctx_move)
            for (int r = tid; r < R; r += SOFT_THREADS) {
                const int j = r / P, c = r - j * P;
                double acc = 0.0;
                if (c < m)
                    for (int i = 0; i < ny; ++i) {
                        const double *sp_ = T.ST + (size_t)(i * NU + j) * T.st_stride;
                        const double w = wy2[i], ri = sigrow[i];
                        double a = 0.0;
                        for (int t = c + 1; t <= p; ++t) a = fma(__ldg(sp_ + (t - c)), ri - sm.yfree[(t - 1) * ny + i], a);
                        acc = fma(w, a, acc);
                    }
                gvec[r] = acc;
            }
            SOFT_SYNC();
            for (int r = tid; r < NV; r += SOFT_THREADS) {
                double a0 = 0.0;
                if (r < R)
                    for (int r2 = 0; r2 < R; ++r2) a0 = fma(__ldg(Wg + (size_t)r * R + r2), gvec[r2], a0);
                sm.z[r] = a0;
            }
            SOFT_SYNC();
        }
        if (qp.q == 0) {   // fast exit: nothing carried and nothing violated
            qp.evaluate(true);
            int bad = 0;
            for (int r = tid; r < R; r += SOFT_THREADS) {
                const int j = r / P, c = r - j * P;
                if (c >= m) continue;
                const double *b = sm.bnd + 4 * j;
                bad |= (sm.z[r] - b[0] < -SIM_VIOL_TOL) | (b[1] - sm.z[r] < -SIM_VIOL_TOL) | (sm.lvl[r] - b[2] < -SIM_VIOL_TOL) |
                       (b[3] - sm.lvl[r] < -SIM_VIOL_TOL);
            }
            for (int row = tid; row < ny * p; row += SOFT_THREADS) {
                const int i = row % ny;
                const double y = sm.ypred[row];
                bad |= (L.ymax[i] - y < -SIM_VIOL_TOL) | (y - L.ymin[i] < -SIM_VIOL_TOL);
            }
            double bsum = soft_sum((double)bad, sm);
            if (bsum == 0.0) return 0;
        }
        qp.n_con += 1;
#ifdef MPC_SIMT_EMULATION
        if (g_soft_dump && tid == 0 && g_soft_dump_k == g_soft_k) {
            double *d = g_soft_dump;
            for (int i = 0; i < ny * p; ++i) *d++ = sm.yfree[i];
            for (int i = 0; i < NV; ++i) *d++ = sm.z[i];
        }
        SOFT_SYNC();
        const int rc_ = qp.solve();
        if (g_soft_dump && tid == 0 && g_soft_dump_k == g_soft_k) {
            double *d = g_soft_dump + ny * p + NV;
            for (int i = 0; i < NV; ++i) *d++ = sm.z[i];
            *d++ = (double)qp.q;
            for (int a = 0; a < qp.q; ++a) *d++ = (double)sm.act[a];
        }
        SOFT_SYNC();
        return rc_;
#else
        return qp.solve();
#endif
    };

    // ---------------- open-loop optimum (closedloop_toolbox.m:85-98) = pass k = -1, then the closed loop (:50) with
    // the open-loop rollout (:100) in lock-step: one loop, so that the controller and the solver inlined into it exist
    // once in the kernel image (instruction cache) ----------------
    for (int k = want_ol ? -1 : 0; k < nit; ++k) {
        const bool ol = k < 0;
        if (ol) {
            for (int c = tid; c < nsig; c += SOFT_THREADS) {
                double v;
                if (c < ny) v = T.r[(size_t)(nit - 1) * ny + c];
                else if (c < 2 * ny) v = 0.0;
                else v = T.v[(size_t)(nit - 1) * nd + (c - 2 * ny)];
                sm.sig[c] = v;
            }
            SOFT_SYNC();
        } else if ((k & (SIM_CH - 1)) == 0) {
            stage_signals(k);
        }
        const double *sigrow = ol ? sm.sig : sm.sig + (size_t)(k & (SIM_CH - 1)) * nsig;
        if (EST) {   // measurement y(k) of the real plant, innovation, correction x_c(k|k) = x_c(k|k-1) + gain * e  (E4)
            for (int i = tid; i < ny; i += SOFT_THREADS) {
                double yp = 0.0, yh = xod[i];
                for (int j = 0; j < nw; ++j) { yp += xp[i * nw + j]; yh += sm.x[i * nw + j]; }
                einn[i] = yp - yh;
            }
            SOFT_SYNC();
            const int hl = E->hl, hlm = hl < HL ? hl : HL;   // delay-line states beyond the model's longest delay never reach its outputs
            const double *Gn = E->gain;
            for (int s_ = tid; s_ < nch + NU * hlm + ny; s_ += SOFT_THREADS) {
                int grow;
                double *dst;
                if (s_ < nch) { grow = s_; dst = sm.x + s_; }
                else if (s_ < nch + NU * hlm) {
                    const int j = (s_ - nch) / hlm, q_ = (s_ - nch) - j * hlm;
                    int pos = head + q_; if (pos >= HL) pos -= HL;
                    grow = nch + j * hl + q_; dst = sm.hist + j * HL + pos;
                } else { const int o = s_ - nch - NU * hlm; grow = nch + NU * hl + o; dst = xod + o; }
                double acc = *dst;
                for (int i = 0; i < ny; ++i) acc = fma(__ldg(Gn + (size_t)grow * ny + i), einn[i], acc);
                *dst = acc;
            }
            SOFT_SYNC();
        }
        build_st(sigrow, ol ? nit - 1 : k, !ol);
        const unsigned long long it_before = qp.n_it;
#ifdef MPC_SIMT_EMULATION
        if (tid == 0) g_soft_k = k;
        SOFT_SYNC();
#endif
        const int rc = controller_move(sigrow);
        if (rc) status = rc;
        if (ol) {
            // sequential cumulative sum of the moves (exact zeros stay exact, VNS2.m:183-191)
            for (int j = tid; j < NU; j += SOFT_THREADS) {
                double lvl = 0.0;
                for (int c = 0; c < P; ++c) { lvl += (c < m ? sm.z[j * P + c] : 0.0); sm.uopt[j * P + c] = lvl; }
            }
            SOFT_SYNC();
            if (mode == 2) {
                double part = 0.0;
                for (int j = tid; j < NU; j += SOFT_THREADS) {
                    if (sel < 0 || sel == j) {
                        const double u0 = fabs(sm.uopt[j * P]);
                        for (int c = 0; c + 1 < m && c + 1 < nit; ++c) {
                            const double df = fabs(sm.uopt[j * P + c + 1] - sm.uopt[j * P + c]);
                            const double xn = u0 / df;
                            if (fabs(xn) <= 1.7976931348623157e308) part += xn * xn;
                        }
                    }
                }
                jnu = soft_sum(part, sm);
            }
            // the closed loop starts from an empty active set
            for (int a = qp.q - 1; a >= 0; --a) { qp.set_mask(sm.act[a], false); SOFT_SYNC(); }
            qp.npref = 0;
            qp.q = 0;   // every column of J is complement basis again (J J' = H^-1 holds for any rotation of it)
            SOFT_SYNC();
            continue;
        }
        if (out.trace && tid == 0) { out.trace[2 * k] = (int)(qp.n_it - it_before); out.trace[2 * k + 1] = qp.q; }
        SOFT_SYNC();
        if (tid < NU) {
            const int j = tid;
            sm.ucur[j] += sm.z[j * P];
            const bool mine = (sel < 0 || sel == j);
            if (out.u && mine) out.u[(size_t)j * nit + k] = sm.ucur[j];
            if (out.uopt && mine) out.uopt[(size_t)j * nit + k] = sm.uopt[j * P + (k < m ? k : m - 1)];
        }
        SOFT_SYNC();
        for (int ch = tid; ch < nch; ch += SOFT_THREADS) {
            const int j = sm.chj[ch], dd = sm.chd[ch];
            const double wk = held(j, sigrow);   // ucur is u(k) now
            int p1 = head + (dd > 0 ? dd - 1 : 0); if (p1 >= HL) p1 -= HL;
            int p0 = head + (dd > 1 ? dd - 2 : 0); if (p0 >= HL) p0 -= HL;
            const double w1 = dd == 0 ? wk : sm.hist[j * HL + p1];
            const double w0 = dd == 0 ? 0.0 : (dd == 1 ? wk : sm.hist[j * HL + p0]);
            sm.x[ch] = sm.cha[ch] * sm.x[ch] + sm.chb0[ch] * w0 + sm.chb1[ch] * w1;
            if (EST) {   // the real plant, from the TRUE input history (slot q of histp holds w(k-1-q))
                const int dp = E->d[ch];
                int q1 = headp + (dp > 0 ? dp - 1 : 0); if (q1 >= HLP) q1 -= HLP;
                int q0 = headp + (dp > 1 ? dp - 2 : 0); if (q0 >= HLP) q0 -= HLP;
                const double v1 = dp == 0 ? wk : histp[j * HLP + q1];
                const double v0 = dp == 0 ? 0.0 : (dp == 1 ? wk : histp[j * HLP + q0]);
                xp[ch] = E->a[ch] * xp[ch] + E->b0[ch] * v0 + E->b1[ch] * v1;
            }
            if (want_ol) {
                double o1, o0;
                if (j < NU) {
                    const int k1 = k - dd, k0 = k + 1 - dd;
                    o1 = k1 < 0 ? 0.0 : sm.uopt[j * P + (k1 < m ? k1 : m - 1)];
                    o0 = (dd == 0 || k0 < 0) ? 0.0 : sm.uopt[j * P + (k0 < m ? k0 : m - 1)];
                } else {
                    o1 = w1; o0 = w0;
                }
                sm.xol[ch] = sm.cha[ch] * sm.xol[ch] + sm.chb0[ch] * o0 + sm.chb1[ch] * o1;
            }
        }
        SOFT_SYNC();
        const int nhead = head == 0 ? HL - 1 : head - 1;
        if (tid < nw) sm.hist[tid * HL + nhead] = held(tid, sigrow);
        head = nhead;
        if (EST) {
            const int nhp = headp == 0 ? HLP - 1 : headp - 1;
            if (tid < nw) histp[tid * HLP + nhp] = held(tid, sigrow);
            headp = nhp;
        }
        SOFT_SYNC();
    }
    // ---------------- costs ----------------
    if (out.cost) {
        if (mode == 1) {
            for (int col = tid; col < nst; col += SOFT_THREADS)
                if (col >= L.stoff_e) out.cost[col - L.stoff_e] = status ? NAN : cost_acc;
        } else if (mode == 2) {
            const double tot = soft_sum(cost_acc, sm) + jnu;
            if (tid == 0) out.cost[0] = status ? NAN : tot;
        }
    }
    if (out.counters && tid == 0) {
        atomicAdd(out.counters + 0, qp.n_con);
        atomicAdd(out.counters + 1, qp.n_it);
    }
    if (out.diag && tid == 0) { out.diag[0] = qp.n_con; out.diag[1] = qp.n_it; out.diag[2] = (unsigned long long)qp.qmax; }
    return status;
}
