// mpc_core.cuh -- the two device routines of the hot path, written warp/CTA-cooperatively.
//
//   mpc_build_candidate : per-candidate prediction-matrix + Hessian builder (one CTA per candidate)
//                         H = G'QG + R from the prefix-Gram tables, Cholesky, [M | W] = H^-1 [-K | I]
//   mpc_sim_run         : one closed-loop simulation = closedloop_toolbox.m:50-100 (one warp per run)
//                         nit x { z = M s ; box/rate check ; dual active-set QP if violated ; plant step }
//                         fused with the open-loop optimum, its rollout and the GAM / VNS cost sums.
//
// SINGLE SOURCE: this header compiles for sm_100a (nvcc, the product) and, with -DMPC_HOST_EMULATION,
// as lane-serialised host code used ONLY by tests/ to debug the algorithm without a GPU
// (tests/host_emulation).  In emulation a "lane loop" runs all lanes in one thread, so a value
// that must cross lanes has to go through shared memory or one of the w* reductions below; per-lane
// registers must not be live across WSYNC().  The product never builds or loads the emulation.
#pragma once
#include <math.h>

#include "mpc_layout.h"

#ifdef MPC_HOST_EMULATION
#define MPC_FN static inline
#define LANE_FOR(i, n) for (int i = 0; i < (n); ++i)
#define TFOR(i, n) for (int i = 0; i < (n); ++i)
#define WSYNC() ((void)0)
#define TSYNC() ((void)0)
#define IS_LANE0 (true)
#define IS_T0 (true)
static inline double wsum(double v) { return v; }
static inline int wany(int v) { return v; }
static inline double wmax(double v) { return v; }
static inline void wargmin(double &, int &) {}
static inline int tany(int v, int *) { return v; }
#define MPC_LDG(p) (*(p))
#else
#define MPC_FN __device__ __forceinline__
#define LANE_FOR(i, n) for (int i = (int)(threadIdx.x & 31u); i < (n); i += 32)
#define TFOR(i, n) for (int i = (int)threadIdx.x; i < (n); i += (int)blockDim.x)
#define WSYNC() __syncwarp()
#define TSYNC() __syncthreads()
#define IS_LANE0 ((threadIdx.x & 31u) == 0u)
#define IS_T0 (threadIdx.x == 0u)
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int wany(int v) { return __any_sync(0xffffffffu, v); }
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// all lanes end with the minimum value and, among equal values, the smallest index
__device__ __forceinline__ void wargmin(double &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov < v || (ov == v && oi >= 0 && (i < 0 || oi < i))) { v = ov; i = oi; }
    }
}
__device__ __forceinline__ int tany(int v, int *) { return __syncthreads_or(v); }
#define MPC_LDG(p) __ldg(p)
#endif

#define MPC_VIOL_TOL 1e-10   /* same constants as oracle/mpc_oracle.c */
#define MPC_DEP_TOL 1e-13
#define MPC_INF (__builtin_huge_val())

// ---------------------------------------------------------------------------------------------
// Builder: one CTA per candidate.
// Shared memory (doubles): Hs[nz*nz] | B[(nst+nz)*nz]   (B column-major: B[col*nz + row])
// Outputs (global): Mg[nst*nz] = M in deviation coordinates, stored [col][e]; Wg[nz*nz] = H^-1.
// ---------------------------------------------------------------------------------------------
static MPC_HD size_t mpc_builder_smem_doubles(int nz, int nst) { return (size_t)nz * nz + (size_t)(nst + nz) * nz; }

MPC_FN int mpc_build_candidate(const MpcLayout &L, const MpcTables &T, int p, int m, const double *delta,
                               const double *lambda, double *smem, double *Mg, double *Wg, int *flag_smem) {
    const int ny = L.ny, nu = L.nu, nw = L.nw, nz = nu * m, ns = L.nst, nsig = L.nsig, ncol = ns + nz;
    double *Hs = smem;
    double *B = smem + (size_t)nz * nz;
    double wy2[MPC_MAXY], wu2[MPC_MAXU];
    for (int i = 0; i < ny; ++i) { double w = delta[i] / L.sy[i]; wy2[i] = w * w; }
    for (int j = 0; j < nu; ++j) { double w = lambda[j] / L.su[j]; wu2[j] = w * w; }
    if (IS_T0) *flag_smem = 0;
    // ---- H = G' Wy^2 G + Wdu^2 from the prefix-Gram table ----
    TFOR(idx, nz * nz) {
        const int e1 = idx / nz, e2 = idx - e1 * nz;
        const int c1 = e1 / nu, j1 = e1 - c1 * nu, c2 = e2 / nu, j2 = e2 - c2 * nu;
        // the move with the smaller horizon index carries the shifted step response
        const int ca = c1 <= c2 ? c1 : c2, cb = c1 <= c2 ? c2 : c1;
        const int ja = c1 <= c2 ? j1 : j2, jb = c1 <= c2 ? j2 : j1;
        const int D = cb - ca, Lh = p - cb;
        double acc = 0.0;
        if (Lh > 0)
            for (int i = 0; i < ny; ++i) acc += wy2[i] * MPC_LDG(T.TG + mpc_tg_index(L, i, D, Lh, ja, jb));
        if (e1 == e2) acc += wu2[j1];
        Hs[idx] = acc;
    }
    // ---- B = [-K | I], K columns in deviation-coordinate order (x block, history blocks, e block) ----
    TFOR(idx, nz * ny * nw) {
        const int e = idx / (ny * nw), sig = idx - e * (ny * nw);
        const int c = e / nu, j = e - c * nu, Lh = p - c;
        const int i = sig / nw;  // x_ij only reaches output i
        const double acc = wy2[i] * MPC_LDG(T.TK + mpc_tk_index(L, i, j, c, Lh) + sig);
        B[(size_t)sig * nz + e] = -acc;
    }
    for (int jw = 0; jw < nw; ++jw) {
        const int nq = L.hlen[jw] - L.hq0[jw];
        TFOR(idx, nz * nq) {
            const int e = idx / nq, qq = idx - e * nq;
            const int c = e / nu, j = e - c * nu, Lh = p - c;
            const int sig = L.hoff[jw] + L.hq0[jw] + qq;
            double acc = 0.0;
            for (int i = 0; i < ny; ++i) acc += wy2[i] * MPC_LDG(T.TK + mpc_tk_index(L, i, j, c, Lh) + sig);
            B[(size_t)(L.stoff_h[jw] + qq) * nz + e] = -acc;
        }
    }
    TFOR(idx, nz * ny) {
        const int e = idx / ny, i = idx - e * ny;
        const int c = e / nu, j = e - c * nu, Lh = p - c;
        B[(size_t)(L.stoff_e + i) * nz + e] = wy2[i] * MPC_LDG(T.S1 + mpc_s1_index(L, i, j, Lh));
    }
    TFOR(idx, nz * nz) {
        const int col = idx / nz, row = idx - col * nz;
        B[(size_t)(ns + col) * nz + row] = (row == col) ? 1.0 : 0.0;
    }
    TSYNC();
    // ---- Cholesky, right-looking, lower factor mirrored into the upper triangle ----
    for (int k = 0; k < nz; ++k) {
        if (IS_T0) {
            const double dkk = Hs[k * nz + k];
            if (!(dkk > 0.0)) *flag_smem = 1;
            Hs[k * nz + k] = sqrt(dkk > 0.0 ? dkk : 1.0);
        }
        TSYNC();
        const double inv = 1.0 / Hs[k * nz + k];
        TFOR(i2, nz - k - 1) {
            const int i = k + 1 + i2;
            const double l = Hs[i * nz + k] * inv;
            Hs[i * nz + k] = l;
            Hs[k * nz + i] = l;
        }
        TSYNC();
        const int rem = nz - k - 1;
        TFOR(idx, rem * rem) {
            const int a = idx / rem, b = idx - a * rem;
            if (b <= a) {
                const int i = k + 1 + a, j = k + 1 + b;
                Hs[i * nz + j] -= Hs[k * nz + i] * Hs[k * nz + j];
            }
        }
        TSYNC();
    }
    const int bad = *flag_smem;
    // ---- forward solve L Y = B (column sweep; all right-hand sides at once) ----
    for (int k = 0; k < nz; ++k) {
        const double inv = 1.0 / Hs[k * nz + k];
        TFOR(col, ncol) B[(size_t)col * nz + k] *= inv;
        TSYNC();
        const int rem = nz - k - 1;
        TFOR(idx, rem * ncol) {
            const int col = idx / rem, i = k + 1 + (idx - col * rem);
            B[(size_t)col * nz + i] -= Hs[k * nz + i] * B[(size_t)col * nz + k];
        }
        TSYNC();
    }
    // ---- backward solve L' X = Y ----
    for (int k = nz - 1; k >= 0; --k) {
        const double inv = 1.0 / Hs[k * nz + k];
        TFOR(col, ncol) B[(size_t)col * nz + k] *= inv;
        TSYNC();
        TFOR(idx, k * ncol) {
            const int col = idx / k, i = idx - col * k;
            B[(size_t)col * nz + i] -= Hs[k * nz + i] * B[(size_t)col * nz + k];
        }
        TSYNC();
    }
    // ---- write out ----
    TFOR(idx, ns * nz) Mg[idx] = B[idx];
    TFOR(idx, nz * nz) Wg[idx] = B[(size_t)ns * nz + idx];
    TSYNC();
    return bad ? 3 : 0;
}

// ---------------------------------------------------------------------------------------------
// Closed-loop run: one warp.
// ---------------------------------------------------------------------------------------------
struct MpcSimSmem {
    double *M;      // nst*nz  z_unc = M * st
    double *s;      // ns      physical state vector (mpc_layout.h)
    double *st;     // nst     deviation coordinates fed to M
    double *z;      // nz      current QP iterate
    double *dir;    // nz      scratch: W n_a while re-appending after a drop
    double *w;      // nz      W n_p
    double *vv;     // nz      n_p - N r
    double *g;      // nz      N' w
    double *l;      // nz      Linv g
    double *rr;     // nz      Linv' l
    double *mu;     // nz      multipliers of the active set
    double *Li;     // nz*(nz+1)/2 packed rows of Linv_c (inverse Cholesky factor of S = N'WN)
    double *xol;    // ny*nw   open-loop plant channel states
    double *uopt;   // nu*m    open-loop optimal MV levels, [j*m + c]
    double *uprev;  // nu
    double *ybuf;   // 2*ny    y(k), ys(k)
    double *cost;   // ny      running cost sums
    double *hbuf;   // hmax    scratch for the history shift
    int *act;       // nz      active constraint ids: type*nz + e
    int *amask;     // nz      4 bits per variable
    int *dflag;     // nz      drop flags (scratch)
};

static MPC_HD int mpc_hmax(const MpcLayout &L) {
    int h = 1;
    for (int j = 0; j < L.nw; ++j) h = L.hlen[j] > h ? L.hlen[j] : h;
    return h;
}

static MPC_HD size_t mpc_sim_smem_doubles(const MpcLayout &L, int m) {
    const int nz = L.nu * m;
    size_t n = (size_t)L.nst * nz + L.ns + L.nst + 8 * (size_t)nz + (size_t)nz * (nz + 1) / 2 + L.ny * L.nw + L.nu * m + L.nu +
               3 * L.ny + mpc_hmax(L);
    n += (3 * (size_t)nz + 1) / 2 + 1;  // three int arrays
    return n;
}

MPC_FN void mpc_sim_carve(const MpcLayout &L, int m, double *base, MpcSimSmem &sm) {
    const int nz = L.nu * m;
    double *p = base;
    sm.M = p; p += (size_t)L.nst * nz;
    sm.s = p; p += L.ns;
    sm.st = p; p += L.nst;
    sm.z = p; p += nz;
    sm.dir = p; p += nz;
    sm.w = p; p += nz;
    sm.vv = p; p += nz;
    sm.g = p; p += nz;
    sm.l = p; p += nz;
    sm.rr = p; p += nz;
    sm.mu = p; p += nz;
    sm.Li = p; p += (size_t)nz * (nz + 1) / 2;
    sm.xol = p; p += L.ny * L.nw;
    sm.uopt = p; p += L.nu * m;
    sm.uprev = p; p += L.nu;
    sm.ybuf = p; p += 2 * L.ny;
    sm.cost = p; p += L.ny;
    sm.hbuf = p; p += mpc_hmax(L);
    sm.act = (int *)p;
    sm.amask = sm.act + nz;
    sm.dflag = sm.amask + nz;
}

// ---------------------------------------------------------------------------------------------
// Constraint ids: cid = type | (e << 2), e = c*NU + j the move index.
//   type 0: dz_e >= dumin_j      1: dz_e <= dumax_j        (MV rate, n = +-e_e)
//   type 2: u_j(k+c) >= umin_j   3: u_j(k+c) <= umax_j     (MV level, n = +-sum_{c'<=c} e_(c',j))
// All are hard (MV ECR = 0, oracle T4).
// ---------------------------------------------------------------------------------------------
MPC_FN double mpc_slack(const MpcLayout &L, int type, int j, double ze, double lvl) {
    switch (type) {
        case 0: return ze - L.dumin[j];
        case 1: return L.dumax[j] - ze;
        case 2: return lvl - L.umin[j];
        default: return L.umax[j] - lvl;
    }
}

template <int NU>
MPC_FN double mpc_level(const MpcSimSmem &sm, int c, int j) {
    double lvl = sm.uprev[j];
    for (int c2 = 0; c2 <= c; ++c2) lvl += sm.z[c2 * NU + j];
    return lvl;
}

// n_a[e] for constraint cid (0 / +-1)
template <int NU>
MPC_FN double mpc_normal_entry(int cid, int e) {
    const int type = cid & 3, ea = cid >> 2;
    if (type < 2) return (e == ea) ? (type == 0 ? 1.0 : -1.0) : 0.0;
    const int ca = ea / NU, ja = ea - ca * NU;
    const int c = e / NU, j = e - c * NU;
    return (j == ja && c <= ca) ? (type == 2 ? 1.0 : -1.0) : 0.0;
}

// n_a' x for a dense vector x in shared memory
template <int NU>
MPC_FN double mpc_normal_dot(int cid, const double *x) {
    const int type = cid & 3, ea = cid >> 2;
    if (type < 2) return type == 0 ? x[ea] : -x[ea];
    const int ca = ea / NU, ja = ea - ca * NU;
    double acc = 0.0;
    for (int c = 0; c <= ca; ++c) acc += x[c * NU + ja];
    return type == 2 ? acc : -acc;
}

// w = W n  (W symmetric, global memory, read coalesced as rows)
template <int NU>
MPC_FN void mpc_w_times_normal(int cid, int nz, const double *W, double *w) {
    const int type = cid & 3, ea = cid >> 2;
    if (type < 2) {
        const double sg = type == 0 ? 1.0 : -1.0;
        LANE_FOR(e, nz) w[e] = sg * W[(size_t)ea * nz + e];
    } else {
        const int ca = ea / NU, ja = ea - ca * NU;
        const double sg = type == 2 ? 1.0 : -1.0;
        LANE_FOR(e, nz) {
            double acc = 0.0;
#pragma unroll 4
            for (int c = 0; c <= ca; ++c) acc += W[(size_t)(c * NU + ja) * nz + e];
            w[e] = sg * acc;
        }
    }
}

// z += scale * W v   (v in shared memory)
MPC_FN void mpc_add_W_times(int nz, const double *W, const double *v, double scale, double *z) {
    LANE_FOR(e, nz) {
        double a0 = 0.0, a1 = 0.0;
        int e2 = 0;
#pragma unroll 4
        for (; e2 + 1 < nz; e2 += 2) {
            a0 += W[(size_t)e2 * nz + e] * v[e2];
            a1 += W[(size_t)(e2 + 1) * nz + e] * v[e2 + 1];
        }
        if (e2 < nz) a0 += W[(size_t)e2 * nz + e] * v[e2];
        z[e] += scale * (a0 + a1);
    }
}

MPC_FN double mpc_Li(const double *Li, int a, int b) { return Li[(a * (a + 1)) / 2 + b]; }  // b <= a

// out = S^-1 rhs with S^-1 = Li' Li  (rhs, out: length q in shared memory; tmp = sm.l)
MPC_FN void mpc_schur_solve(int q, const MpcSimSmem &sm, const double *rhs, double *out) {
    LANE_FOR(a, q) {
        double acc = 0.0;
        for (int b = 0; b <= a; ++b) acc += mpc_Li(sm.Li, a, b) * rhs[b];
        sm.l[a] = acc;
    }
    WSYNC();
    LANE_FOR(a, q) {
        double acc = 0.0;
        for (int b = a; b < q; ++b) acc += mpc_Li(sm.Li, b, a) * sm.l[b];
        out[a] = acc;
    }
    WSYNC();
}

// Schur column of constraint cid against the first q active constraints; wv = W n_cid.
// Leaves g = N'wv, l = Li g, rr = Li' l in shared memory; returns rho = gamma - |l|^2.
template <int NU>
MPC_FN double mpc_schur_column(int cid, int q, const double *wv, MpcSimSmem &sm, double *gamma_out) {
    const double gamma = mpc_normal_dot<NU>(cid, wv);
    LANE_FOR(a, q) sm.g[a] = mpc_normal_dot<NU>(sm.act[a], wv);
    WSYNC();
    double part = 0.0;
    LANE_FOR(a, q) {
        double acc = 0.0;
        for (int b = 0; b <= a; ++b) acc += mpc_Li(sm.Li, a, b) * sm.g[b];
        sm.l[a] = acc;
        part += acc * acc;
    }
    const double l2 = wsum(part);
    WSYNC();
    LANE_FOR(a, q) {
        double acc = 0.0;
        for (int b = a; b < q; ++b) acc += mpc_Li(sm.Li, b, a) * sm.l[b];
        sm.rr[a] = acc;
    }
    WSYNC();
    *gamma_out = gamma;
    return gamma - l2;
}

// Commit constraint cid at position q: new row of Linv_c = [-r'/sqrt(rho), 1/sqrt(rho)]
MPC_FN void mpc_schur_commit(int cid, int q, double rho, double mu_new, MpcSimSmem &sm) {
    const double isr = 1.0 / sqrt(rho);
    double *row = sm.Li + (q * (q + 1)) / 2;
    LANE_FOR(a, q) row[a] = -sm.rr[a] * isr;
    if (IS_LANE0) {
        row[q] = isr;
        sm.act[q] = cid;
        sm.mu[q] = mu_new;
        sm.amask[cid >> 2] |= (1 << (cid & 3));
    }
    WSYNC();
}

// Remove every active constraint whose drop flag (sm.dflag[a] != 0) is set, keeping the order of the
// rest, and rebuild the factor rows from the first removed position on.  Returns the new q.
template <int NU>
MPC_FN int mpc_drop_flagged(int q, int nz, const double *W, MpcSimSmem &sm) {
    int first = -1, qn = 0;
    for (int a = 0; a < q; ++a) {  // uniform, sequential: q <= nz is small
        const int fl = sm.dflag[a];
        const int cid = sm.act[a];
        const double mua = sm.mu[a];
        WSYNC();
        if (fl) {
            if (first < 0) first = a;
            if (IS_LANE0) sm.amask[cid >> 2] &= ~(1 << (cid & 3));
        } else {
            if (IS_LANE0) { sm.act[qn] = cid; sm.mu[qn] = mua; }
            qn += 1;
        }
        WSYNC();
    }
    if (first < 0) return q;
    for (int a = first; a < qn; ++a) {
        const int cid = sm.act[a];
        const double mua = sm.mu[a];
        mpc_w_times_normal<NU>(cid, nz, W, sm.dir);
        WSYNC();
        double gam2;
        const double rho2 = mpc_schur_column<NU>(cid, a, sm.dir, sm, &gam2);
        mpc_schur_commit(cid, a, rho2 > 0.0 ? rho2 : MPC_DEP_TOL * gam2, mua, sm);
    }
    return qn;
}

// Dual active-set QP (Goldfarb-Idnani step logic on the Schur complement S = N'WN of the active
// normals in the metric W = H^-1), WARM-STARTED: the active set and its factor survive from the
// previous sample (the Toolbox solver does the same: Optimizer.ActiveSetOptions.UseWarmStart = 1 in the
// reference's saved objects).  A warm start only changes the path, never the optimum (strictly convex).
//   in : sm.z = unconstrained optimum z_unc, sm.uprev, q_io = carried active-set size
//   out: sm.z = constrained optimum, q_io = final active-set size.  Returns status.
template <int NU>
MPC_FN int mpc_qp_active_set(const MpcLayout &L, int m, const double *W, MpcSimSmem &sm, int &q_io, int *iters_out) {
    const int nz = NU * m;
    int q = q_io, it = 0;
    const int itmax = 20 * (nz + 10);
    // ---- warm start: solve the equality-constrained problem on the carried set, shed negative multipliers
    while (q > 0) {
        LANE_FOR(a, q) {
            const int cid = sm.act[a];
            const int e = cid >> 2, c = e / NU, j = e - c * NU;
            sm.g[a] = -mpc_slack(L, cid & 3, j, sm.z[e], mpc_level<NU>(sm, c, j));
        }
        WSYNC();
        mpc_schur_solve(q, sm, sm.g, sm.mu);
        double mumax = 0.0;
        LANE_FOR(a, q) mumax = fmax(mumax, fabs(sm.mu[a]));
        mumax = wmax(mumax);
        int ndrop = 0;
        LANE_FOR(a, q) {
            const int fl = sm.mu[a] < -1e-12 * mumax;
            sm.dflag[a] = fl;
            ndrop |= fl;
        }
        ndrop = wany(ndrop);
        WSYNC();
        if (!ndrop) break;
        it += 1;
        q = mpc_drop_flagged<NU>(q, nz, W, sm);
    }
    if (q > 0) {  // z = z_unc + W N mu
        LANE_FOR(a, q) if (sm.mu[a] < 0.0) sm.mu[a] = 0.0;
        WSYNC();
        LANE_FOR(e, nz) {
            double acc = 0.0;
            for (int a = 0; a < q; ++a) acc += sm.mu[a] * mpc_normal_entry<NU>(sm.act[a], e);
            sm.vv[e] = acc;
        }
        WSYNC();
        mpc_add_W_times(nz, W, sm.vv, 1.0, sm.z);
        WSYNC();
    }
    // ---- Goldfarb-Idnani iterations from the S-pair (z, A) ----
    for (;;) {
        double bv = -MPC_VIOL_TOL;
        int bi = -1;
        LANE_FOR(e, nz) {
            const int c = e / NU, j = e - c * NU;
            const double lvl = mpc_level<NU>(sm, c, j);
            const double ze = sm.z[e];
            const int am = sm.amask[e];
            for (int type = 0; type < 4; ++type) {
                if (am & (1 << type)) continue;
                const double sl = mpc_slack(L, type, j, ze, lvl);
                const int id = type | (e << 2);
                if (sl < bv || (sl == bv && bi >= 0 && id < bi)) { bv = sl; bi = id; }
            }
        }
        wargmin(bv, bi);
        if (bi < 0) break;
        const int p = bi;
        double sp = bv, mu_p = 0.0;
        mpc_w_times_normal<NU>(p, nz, W, sm.w);
        WSYNC();
        for (;;) {
            if (++it > itmax) { *iters_out = it; q_io = q; return 2; }
            double gamma;
            const double rho = mpc_schur_column<NU>(p, q, sm.w, sm, &gamma);
            const int dependent = !(rho > MPC_DEP_TOL * gamma);
            double t1 = MPC_INF;
            int l1 = -1;
            LANE_FOR(a, q) {
                const double ra = sm.rr[a];
                if (ra > 0.0) {
                    const double t = sm.mu[a] / ra;
                    if (t < t1 || (t == t1 && l1 >= 0 && a < l1)) { t1 = t; l1 = a; }
                }
            }
            wargmin(t1, l1);
            const double t2 = dependent ? MPC_INF : -sp / rho;
            const double t = t1 < t2 ? t1 : t2;
            if (!(t < MPC_INF)) { *iters_out = it; q_io = q; return 1; }
            const int full = !(dependent || t1 < t2);
            if (!dependent) {  // primal step along dir = W (n_p - N r)
                LANE_FOR(e, nz) {
                    double acc = mpc_normal_entry<NU>(p, e);
                    for (int a = 0; a < q; ++a) acc -= sm.rr[a] * mpc_normal_entry<NU>(sm.act[a], e);
                    sm.vv[e] = acc;
                }
                WSYNC();
                mpc_add_W_times(nz, W, sm.vv, t, sm.z);
                sp += t * rho;
            }
            LANE_FOR(a, q) {
                sm.mu[a] -= t * sm.rr[a];
                sm.dflag[a] = (a == l1) && !full;
            }
            mu_p += t;
            WSYNC();
            if (full) {
                mpc_schur_commit(p, q, rho, mu_p, sm);
                q += 1;
                break;
            }
            q = mpc_drop_flagged<NU>(q, nz, W, sm);
        }
    }
    // ---- one Newton correction on the active constraints if they drifted ----
    if (q > 0) {
        double worst = 0.0;
        LANE_FOR(a, q) {
            const int cid = sm.act[a];
            const int e = cid >> 2, c = e / NU, j = e - c * NU;
            const double sl = mpc_slack(L, cid & 3, j, sm.z[e], mpc_level<NU>(sm, c, j));
            sm.g[a] = -sl;
            worst = fmax(worst, fabs(sl));
        }
        worst = wmax(worst);
        WSYNC();
        if (worst > 1e-13) {
            mpc_schur_solve(q, sm, sm.g, sm.rr);  // delta mu
            LANE_FOR(e, nz) {
                double acc = 0.0;
                for (int a = 0; a < q; ++a) acc += sm.rr[a] * mpc_normal_entry<NU>(sm.act[a], e);
                sm.vv[e] = acc;
            }
            WSYNC();
            mpc_add_W_times(nz, W, sm.vv, 1.0, sm.z);
            WSYNC();
        }
    }
    *iters_out = it;
    q_io = q;
    return 0;
}

// z = M st, then the box / rate check; runs the active-set QP when the check fails or when an active
// set is carried over from the previous sample.
template <int NU>
MPC_FN int mpc_controller_move(const MpcLayout &L, int m, const double *W, MpcSimSmem &sm, int &q_io, int *constrained,
                               int *iters) {
    const int ny = L.ny, nw = L.nw, nz = NU * m, ns = L.nst;
    // deviation coordinates (mpc_layout.h): everything relative to the steady state of the held inputs
    LANE_FOR(ch, ny * nw) {
        const int j = ch % nw;
        const double hv = j < NU ? sm.s[L.hoff[j]] : sm.s[L.off_v + (j - NU)];
        sm.st[ch] = sm.s[ch] - L.gain[ch] * hv;
    }
    for (int j = 0; j < nw; ++j) {
        const int q0 = L.hq0[j], nq = L.hlen[j] - q0;
        const double hv = j < NU ? sm.s[L.hoff[j]] : sm.s[L.off_v + (j - NU)];
        LANE_FOR(qq, nq) sm.st[L.stoff_h[j] + qq] = sm.s[L.hoff[j] + q0 + qq] - hv;
    }
    LANE_FOR(i, ny) {
        double acc = sm.s[L.off_r + i];
        for (int j = 0; j < nw; ++j) acc -= L.gain[i * nw + j] * (j < NU ? sm.s[L.hoff[j]] : sm.s[L.off_v + (j - NU)]);
        sm.st[L.stoff_e + i] = acc;
    }
    WSYNC();
    LANE_FOR(e, nz) {
        double a0 = 0.0, a1 = 0.0;
        int sg = 0;
#pragma unroll 4
        for (; sg + 1 < ns; sg += 2) {
            a0 += sm.M[(size_t)sg * nz + e] * sm.st[sg];
            a1 += sm.M[(size_t)(sg + 1) * nz + e] * sm.st[sg + 1];
        }
        if (sg < ns) a0 += sm.M[(size_t)sg * nz + e] * sm.st[sg];
        sm.z[e] = a0 + a1;
    }
    WSYNC();
    *iters = 0;
    if (q_io == 0) {
        int bad = 0;
        LANE_FOR(e, nz) {
            const int c = e / NU, j = e - c * NU;
            const double lvl = mpc_level<NU>(sm, c, j);
            const double ze = sm.z[e];
            bad |= (ze - L.dumin[j] < -MPC_VIOL_TOL) | (L.dumax[j] - ze < -MPC_VIOL_TOL) |
                   (lvl - L.umin[j] < -MPC_VIOL_TOL) | (L.umax[j] - lvl < -MPC_VIOL_TOL);
        }
        bad = wany(bad);
        *constrained = bad;
        if (!bad) return 0;
    } else {
        *constrained = 1;
    }
    return mpc_qp_active_set<NU>(L, m, W, sm, q_io, iters);
}

// One plant sample for channel ch=(i,j): x(k+1) = a x(k) + b0 w(k+1-d) + b1 w(k-d)
//   hist[q] = w(k-1-q) (q < hlen), wk = w(k).
MPC_FN double mpc_channel_step(const MpcLayout &L, int ch, double x, const double *hist, double wk) {
    const int dd = L.d[ch];
    const double w1 = dd == 0 ? wk : hist[dd - 1];
    const double w0 = dd == 0 ? 0.0 : (dd == 1 ? wk : hist[dd - 2]);
    return L.a[ch] * x + L.b0[ch] * w0 + L.b1[ch] * w1;
}

struct MpcRunOut {
    double *cost;   // per-run partial cost slots (ny doubles for GAM; 1 for VNS) or NULL
    double *y, *u, *ys, *uopt;  // trajectories for this candidate (signals x nit) or NULL
    unsigned long long *counters;  // [0] constrained QPs, [1] active-set iterations (atomic), or NULL
};

// sel: -2 user set-point (GAM / RAW);  -1 VNS step on every output;  i>=0 VNS step on output i only.
// mode: 0 RAW, 1 GAM, 2 VNS.  Returns status.
template <int NU>
MPC_FN int mpc_sim_run(const MpcLayout &L, const MpcTables &T, int p, int m, const double *Mg, const double *W,
                       int mode, int sel, double *smem_base, const MpcRunOut &out) {
    const int ny = L.ny, nu = NU, nd = L.nd, nw = L.nw, nz = NU * m, ns = L.ns, nit = L.nit;
    MpcSimSmem sm;
    mpc_sim_carve(L, m, smem_base, sm);
    double *cost_local = sm.cost;
    LANE_FOR(i, L.nst * nz) sm.M[i] = Mg[i];
    LANE_FOR(i, ns) sm.s[i] = 0.0;
    LANE_FOR(i, ny * nw) sm.xol[i] = 0.0;
    LANE_FOR(j, nu) sm.uprev[j] = 0.0;
    LANE_FOR(i, nu * m) sm.uopt[i] = 0.0;
    LANE_FOR(e, nz) sm.amask[e] = 0;
    WSYNC();
    int status = 0;
    int qact = 0;  // carried active-set size (warm start)
    unsigned long long n_con = 0, n_it = 0;
    const bool want_ol = (mode != 1) || out.ys || out.uopt;
    double jnu = 0.0;  // uniform
    // ---------------- open-loop optimum (closedloop_toolbox.m:85-98) ----------------
    if (want_ol) {
        LANE_FOR(i, ny) {
            double rv;
            if (sel == -2) rv = T.r[(size_t)(nit - 1) * ny + i];
            else rv = (sel == -1 || sel == i) ? ((nit - 1) >= (L.inK - 1) ? 1.0 : 0.0) : 0.0;
            sm.s[L.off_r + i] = rv;
        }
        LANE_FOR(j, nd) sm.s[L.off_v + j] = T.v[(size_t)(nit - 1) * nd + j];
        WSYNC();
        int con, its;
        const int rc = mpc_controller_move<NU>(L, m, W, sm, qact, &con, &its);
        if (rc) status = rc;
        n_con += con ? 1 : 0; n_it += its;
        LANE_FOR(j, nu) {
            double lvl = 0.0;
            for (int c = 0; c < m; ++c) { lvl += sm.z[c * nu + j]; sm.uopt[j * m + c] = lvl; }
        }
        WSYNC();
        // Jnu (VNS2.m:183-191): (|uopt_j(0)| / |uopt_j(k+1)-uopt_j(k)|)^2, inf/nan -> 0; only the first
        // m-1 differences can be non-zero because rows m.. repeat row m-1 (T6).
        if (mode == 2) {
            double part = 0.0;
            LANE_FOR(j, nu) {
                if (sel < 0 || sel == j) {
                    const double u0 = fabs(sm.uopt[j * m]);
                    for (int c = 0; c + 1 < m && c + 1 < nit; ++c) {
                        const double df = fabs(sm.uopt[j * m + c + 1] - sm.uopt[j * m + c]);
                        const double xn = u0 / df;
                        if (fabs(xn) <= 1.7976931348623157e308) part += xn * xn;  // inf / nan -> 0 (VNS2.m:186)
                    }
                }
            }
            jnu = wsum(part);
        }
        LANE_FOR(i, ns) sm.s[i] = 0.0;
        LANE_FOR(e, nz) sm.amask[e] = 0;   // the closed loop starts from an empty active set
        qact = 0;
        WSYNC();
    }
    LANE_FOR(i, ny) cost_local[i] = 0.0;
    WSYNC();
    // ---------------- closed loop (closedloop_toolbox.m:50) + open-loop rollout (:100) in lock-step ----
    for (int k = 0; k < nit; ++k) {
        // outputs y(k), ys(k)
        LANE_FOR(i, ny) {
            double acc = 0.0, acc2 = 0.0;
            for (int j = 0; j < nw; ++j) { acc += sm.s[i * nw + j]; acc2 += sm.xol[i * nw + j]; }
            sm.ybuf[i] = acc;
            sm.ybuf[ny + i] = acc2;
            if (out.y && (sel < 0 || sel == i)) out.y[(size_t)i * nit + k] = acc;
            if (out.ys && (sel < 0 || sel == i)) out.ys[(size_t)i * nit + k] = acc2;
            if (mode == 1) {
                const double e = acc - T.yref[(size_t)i * nit + k];
                cost_local[i] += e * e;
            } else if (mode == 2 && k >= L.inK - 1 && (sel < 0 || sel == i)) {
                const double e2 = acc - acc2, er = acc - T.yref[(size_t)i * nit + k];
                cost_local[i] += e2 * e2 + er * er;
            }
            double rv;
            if (sel == -2) rv = T.r[(size_t)k * ny + i];
            else rv = (sel == -1 || sel == i) ? (k >= L.inK - 1 ? 1.0 : 0.0) : 0.0;
            sm.s[L.off_r + i] = rv;
        }
        LANE_FOR(j, nd) sm.s[L.off_v + j] = T.v[(size_t)k * nd + j];
        WSYNC();
        int con, its;
        const int rc = mpc_controller_move<NU>(L, m, W, sm, qact, &con, &its);
        if (rc) status = rc;
        n_con += con ? 1 : 0; n_it += its;
        // apply the first move
        LANE_FOR(j, nu) {
            const double un = sm.uprev[j] + sm.z[j];
            sm.uprev[j] = un;
            if (out.u && (sel < 0 || sel == j)) out.u[(size_t)j * nit + k] = un;
            if (out.uopt && (sel < 0 || sel == j)) out.uopt[(size_t)j * nit + k] = sm.uopt[j * m + (k < m ? k : m - 1)];
        }
        WSYNC();
        // plant step, closed loop (history in s) and open loop (history from uopt / v directly)
        LANE_FOR(ch, ny * nw) {
            const int j = ch % nw;
            const double wk = j < nu ? sm.uprev[j] : sm.s[L.off_v + (j - nu)];
            const double xn = mpc_channel_step(L, ch, sm.s[ch], sm.s + L.hoff[j], wk);
            if (want_ol) {
                const int dd = L.d[ch];
                double w1, w0;
                if (j < nu) {
                    const int k1 = k - dd, k0 = k + 1 - dd;
                    w1 = k1 < 0 ? 0.0 : sm.uopt[j * m + (k1 < m ? k1 : m - 1)];
                    w0 = (dd == 0 || k0 < 0) ? 0.0 : sm.uopt[j * m + (k0 < m ? k0 : m - 1)];
                } else {
                    const int k1 = k - dd, k0 = k + 1 - dd;
                    w1 = k1 < 0 ? 0.0 : T.v[(size_t)k1 * nd + (j - nu)];
                    w0 = (dd == 0 || k0 < 0) ? 0.0 : T.v[(size_t)k0 * nd + (j - nu)];
                }
                sm.xol[ch] = L.a[ch] * sm.xol[ch] + L.b0[ch] * w0 + L.b1[ch] * w1;
            }
            sm.s[ch] = xn;
        }
        WSYNC();
        // push w(k) into the histories (shift by one, newest in slot 0)
        for (int j = 0; j < nw; ++j) {
            const int hl = L.hlen[j];
            if (hl == 0) continue;
            double *h = sm.s + L.hoff[j];
            const double wk = j < nu ? sm.uprev[j] : sm.s[L.off_v + (j - nu)];
            // read phase
            LANE_FOR(qh, hl) sm.hbuf[qh] = (qh == 0) ? wk : h[qh - 1];
            WSYNC();
            LANE_FOR(qh, hl) h[qh] = sm.hbuf[qh];
            WSYNC();
        }
    }
    // ---------------- costs ----------------
    if (out.cost) {
        if (mode == 1) {
            LANE_FOR(i, ny) out.cost[i] = status ? NAN : cost_local[i];
        } else if (mode == 2) {
            double part = 0.0;
            LANE_FOR(i, ny) part += cost_local[i];
            const double tot = wsum(part) + jnu;
            if (IS_LANE0) out.cost[0] = status ? NAN : tot;
        }
    }
#ifndef MPC_HOST_EMULATION
    if (out.counters && IS_LANE0) {
        atomicAdd(out.counters + 0, n_con);
        atomicAdd(out.counters + 1, n_it);
    }
#else
    if (out.counters) { out.counters[0] += n_con; out.counters[1] += n_it; }
#endif
    return status;
}
