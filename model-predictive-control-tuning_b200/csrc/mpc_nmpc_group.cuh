// mpc_nmpc_group.cuh -- the NMPC closed loop with G LANES PER RUN (G = 16 by default: two runs per warp; 8 and 32 also built).
//
// Why not a warp per run: a run is a serial chain of RK4 stages (two exp() per right-hand side) on a 3-state model; the
// chain itself is uniform work, so a warp that owns one run executes it 32 times over.  With 2048+ runs in flight the
// kernel was bound by the FP64 pipe doing that redundant work (ncu r2: k_nmpc_w, fp64 pipe 60 % busy, 260 ms for 2048
// candidates), not by latency.  A thread per run has no redundancy but nothing to hide the chain's latency with, and
// its per-run Hessian lived in local memory (1.6 s).  A group of lanes per run keeps the parts that ARE parallel busy --
//   * the five columns of the stage sensitivities [d k/d x | d k/d u] (lane c < 5 owns column c: a 3x3 by 3x5 product is
//     9 FMAs per lane and needs no exchange at all, the composition of two sub-steps 9 shuffles),
//   * the columns of X = dx/dv, the entries of the Gauss-Newton Hessian (packed lower triangle in shared memory, entry t
//     owned by lane t mod G), the rows of the Cholesky factor of the box-QP, the ratio / multiplier tests --
// and cut the redundant work.  The groups of a warp run independent closed loops: all exchanges are group-wide
// (__shfl_sync / __syncwarp with the group's lane mask), so groups may diverge; the host sorts the runs by horizon, but the
// SQP iteration counts still differ (ncu r2, G = 8: 14.5 of 32 lanes active per instruction), which is why two runs per warp
// measure better than four.  What bounds the kernel is the latency of ONE run: the heaviest of a population is a chain of
// ~40 M dependent-issue instructions (240 ms; 44 ms on a host core), whatever the population size up to one wave.
//
// The SQP iteration also needs one rollout instead of two: the line search tries the full step first by evaluating the
// MODEL (cost, gradient, Hessian) at v + d -- if the cost went down, which is the rule near convergence, the next
// iteration's model is already there; only a rejected full step falls back to cost-only rollouts of the shorter steps
// (all five at once, one per lane).  Same iterates as mpc_nmpc_core.h (the CPU port / thread-per-run kernel), which
// evaluates the step lengths 1, 1/2 .. 1/32 in that order and takes the first that decreases the cost.
#pragma once
#include "mpc_nmpc_core.h"

struct NmpcArgs {
    const int *N, *Nu;
    const double *delta, *lambda;
    const double *r, *yref;      // ny x nit
    double *cost, *part;         // GAM: n x ny; VNS partial: n x runs
    double *y, *u, *yopt, *uopt; // optional n x 2 x nit
    int *status;
    unsigned long long *counters;   // [0] controller calls, [1] SQP iterations
    double *work;                // thread-per-run kernel with NM_GLOBAL_WORK only: per run 2 * nz_max^2 doubles
};

// per-run shared memory (doubles) for plans of up to maxz variables: H, Lc packed lower triangles; v, vo, g, d, t, S0, S1
// (maxz each); AB (16); ints fixed, fl, finv.  The host launches the population in bins of the control horizon, each with
// the shared memory its largest plan needs: residency is bound by shared memory (9.6 KB per run at maxz = 30).
#ifndef NMG_MINB
#define NMG_MINB 1
#endif
static inline __host__ __device__ int nmg_doubles(int maxz) {
    const int n = maxz * (maxz + 1) + 7 * maxz + 16 + (3 * maxz + 1) / 2 + 1;
    return (n + 1) & ~1;
}

struct NmgSm { double *H, *Lc, *v, *vo, *g, *d, *t, *S0, *S1, *AB; int *fixed, *fl, *finv; };

template <int G>
struct NmGroup {
    int gl;          // lane within the group
    unsigned mask;   // the group's lanes
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ double bcast(double v, int src) const { return __shfl_sync(mask, v, src, G); }
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
        return v;
    }
    __device__ __forceinline__ double max(double v) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o, G));
        return v;
    }
    // minimum value over the group and the smallest index holding it (i < 0: does not take part)
    __device__ __forceinline__ void argmin(double &v, int &i) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(mask, v, o, G);
            const int oi = __shfl_xor_sync(mask, i, o, G);
            if (oi >= 0 && (i < 0 || ov < v || (ov == v && oi < i))) { v = ov; i = oi; }
        }
    }
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) & mask) >> (__ffs((int)mask) - 1); }
};

__device__ __forceinline__ int nmg_tri(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i
__device__ __forceinline__ double nmg_sym(const double *H, int a, int b) { return a >= b ? H[nmg_tri(a, b)] : H[nmg_tri(b, a)]; }
__device__ __forceinline__ double nmg_clamp(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// One sample (nsub RK4 steps) with [A|B] = [dx+/dx | dx+/du] (3 x 5): lane c < 5 owns column c of every 3 x 5 matrix of the
// chain (lanes >= 5 shadow column 4).  Leaves [A|B] in ABs[0..15) after a group barrier; x is advanced in place.
template <int G>
__device__ __forceinline__ void nmg_rk4_sens(const NmpcDev &D, double *x, const double *u, const NmGroup<G> &gp, double *ABs) {
    const int c = gp.gl < 5 ? gp.gl : 4;
    const double h = D.Ts / D.nsub;
    const double e0 = c == 0 ? 1.0 : 0.0, e1 = c == 1 ? 1.0 : 0.0, e2 = c == 2 ? 1.0 : 0.0;
    double A0 = e0, A1 = e1, A2 = e2;   // column c of the running [A|B]
    for (int s = 0; s < D.nsub; ++s) {
        double k[4][NX], xs[NX], Jm[15], Dk[4][NX];
        const double ca[4] = {0.0, 0.5, 0.5, 1.0};
#pragma unroll
        for (int st = 0; st < 4; ++st) {
#pragma unroll
            for (int i = 0; i < NX; ++i) xs[i] = st == 0 ? x[i] : x[i] + ca[st] * h * k[st - 1][i];
            vdv_rhs(xs, u, k[st], Jm);
            // column c of d xs / d(x_sub, u) = [I|0] + ca h Dk[st-1]
            const double d0 = st == 0 ? e0 : fma(ca[st] * h, Dk[st - 1][0], e0);
            const double d1 = st == 0 ? e1 : fma(ca[st] * h, Dk[st - 1][1], e1);
            const double d2 = st == 0 ? e2 : fma(ca[st] * h, Dk[st - 1][2], e2);
#pragma unroll
            for (int r = 0; r < NX; ++r) {
                double acc = c == 3 ? Jm[r * 5 + 3] : (c == 4 ? Jm[r * 5 + 4] : 0.0);
                acc = fma(Jm[r * 5 + 0], d0, acc);
                acc = fma(Jm[r * 5 + 1], d1, acc);
                acc = fma(Jm[r * 5 + 2], d2, acc);
                Dk[st][r] = acc;
            }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] += (h / 6.0) * (k[0][i] + 2.0 * k[1][i] + 2.0 * k[2][i] + k[3][i]);
        // transition of this sub-step, column c; its first three columns (Phi_x) come from lanes 0..2
        double Pc[NX];
#pragma unroll
        for (int r = 0; r < NX; ++r) Pc[r] = (h / 6.0) * (Dk[0][r] + 2.0 * Dk[1][r] + 2.0 * Dk[2][r] + Dk[3][r]) + (r == c ? 1.0 : 0.0);
        double n[NX];
#pragma unroll
        for (int r = 0; r < NX; ++r) {
            double acc = c >= NX ? Pc[r] : 0.0;
            acc = fma(gp.bcast(Pc[r], 0), A0, acc);
            acc = fma(gp.bcast(Pc[r], 1), A1, acc);
            acc = fma(gp.bcast(Pc[r], 2), A2, acc);
            n[r] = acc;
        }
        A0 = n[0]; A1 = n[1]; A2 = n[2];
    }
    gp.sync();   // the previous sample's readers are done with ABs
    if (gp.gl < 5) { ABs[c] = A0; ABs[5 + c] = A1; ABs[10 + c] = A2; }
    gp.sync();
}

// Gauss-Newton model of the plan  clamp(v + alpha d)  (v, d in shared memory): returns its cost J, leaves the gradient in
// sm.g and the Hessian in sm.H (packed lower triangle).  alpha = 0: the plan v itself.
template <int G>
__device__ __noinline__ double nmg_model(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m, const double *wy2,
                            const double *wu2, const NmgSm &sm, const NmGroup<G> &gp, double alpha) {
    constexpr int NC = (NM_MAXZ + G - 1) / G;
    const int gl = gp.gl, nz = NU * m;
    for (int t = gl; t < nz * (nz + 1) / 2; t += G) sm.H[t] = 0.0;
    double gacc[NC], X[NC][NX];
#pragma unroll
    for (int k = 0; k < NC; ++k) { gacc[k] = 0.0; X[k][0] = X[k][1] = X[k][2] = 0.0; }
    double x[NX] = {x0[0], x0[1], x0[2]}, up[NU] = {uprev[0], uprev[1]};
    double J = 0.0;
    auto plan = [&](int a) { return nmg_clamp(fma(alpha, sm.d[a], sm.v[a]), D.umin[a & 1], D.umax[a & 1]); };
    for (int i = 0; i < p; ++i) {
        const int c = i < m ? i : m - 1;
        const double u[NU] = {plan(NU * c), plan(NU * c + 1)};
        if (i < m) {
#pragma unroll
            for (int j = 0; j < NU; ++j) { const double du = u[j] - up[j]; J = fma(wu2[j] * du, du, J); up[j] = u[j]; }
        }
        nmg_rk4_sens<G>(D, x, u, gp, sm.AB);
        const int ncol = NU * (c + 1);   // only the columns of moves 0..c can be non-zero
        const double e0 = r[0] - x[1], e1 = r[1] - x[2];
        J = fma(wy2[0] * e0, e0, J);
        J = fma(wy2[1] * e1, e1, J);
        {
            const double *AB = sm.AB;
            const double a00 = AB[0], a01 = AB[1], a02 = AB[2], a10 = AB[5], a11 = AB[6], a12 = AB[7], a20 = AB[10], a21 = AB[11], a22 = AB[12];
#pragma unroll
            for (int k = 0; k < NC; ++k) {
                const int a = gl + G * k;
                if (a < ncol) {
                    const double q0 = X[k][0], q1 = X[k][1], q2 = X[k][2];
                    double n0 = a00 * q0 + a01 * q1 + a02 * q2, n1 = a10 * q0 + a11 * q1 + a12 * q2, n2 = a20 * q0 + a21 * q1 + a22 * q2;
                    if ((a >> 1) == c) { n0 += AB[NX + (a & 1)]; n1 += AB[5 + NX + (a & 1)]; n2 += AB[10 + NX + (a & 1)]; }
                    X[k][0] = n0; X[k][1] = n1; X[k][2] = n2;
                    sm.S0[a] = n1; sm.S1[a] = n2;   // dy_0/dv_a, dy_1/dv_a
                    gacc[k] = fma(-wy2[0] * n1, e0, gacc[k]);
                    gacc[k] = fma(-wy2[1] * n2, e1, gacc[k]);
                }
            }
        }
        gp.sync();
        {   // H += wy2_0 S0 S0' + wy2_1 S1 S1' on the packed triangle: entry t = a (a + 1) / 2 + b belongs to lane t mod G
            const int ntc = ncol * (ncol + 1) / 2;
            int a = 0, b = gl;
            while (b > a) { b -= a + 1; ++a; }
            for (int t = gl; t < ntc; t += G) {
                sm.H[t] = fma(wy2[0] * sm.S0[a], sm.S0[b], fma(wy2[1] * sm.S1[a], sm.S1[b], sm.H[t]));
                b += G;
                while (b > a) { b -= a + 1; ++a; }
            }
        }
        // (the next sample's barrier inside nmg_rk4_sens orders these reads of S0 / S1 before the next writes)
    }
    gp.sync();   // the last sample's accumulation into H is dealt by entry, the terms below by variable: other lanes' entries
                 // (found by the host emulation, where lanes are threads; a converged warp hides the race, it does not remove it)
    // move-suppression terms: rows of D'Wdu^2 D
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int a = gl + G * k;
        if (a < nz) {
            const int j = a & 1, cc = a >> 1;
            const double w = wu2[j], va = plan(a);
            const double du = va - (cc == 0 ? uprev[j] : plan(a - NU));
            const bool has_next = cc + 1 < m;
            double ga = fma(w, du, gacc[k]);
            if (has_next) ga = fma(-w, plan(a + NU) - va, ga);
            sm.g[a] = ga;
            sm.H[nmg_tri(a, a)] += has_next ? 2.0 * w : w;
            if (cc > 0) sm.H[nmg_tri(a, a - NU)] -= w;
        }
    }
    gp.sync();
    return J;
}

// one plant / model sample without sensitivities: ONE out-of-line copy for the line search and the closed loop (the kernel
// image is what the instruction cache has to hold: 24 k SASS instructions with everything inlined, 8 k like this)
__device__ __noinline__ void nmg_rk4_plain(const NmpcDev &D, double *x, const double *u) { rk4_sample(D, x, u, nullptr); }

// cost of the plan clamp(v + alpha d) for THIS LANE's alpha (no sensitivities): the shorter steps of the line search
template <int G>
__device__ __noinline__ double nmg_plan_cost(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m, const double *wy2,
                                const double *wu2, const NmgSm &sm, double alpha) {
    double x[NX] = {x0[0], x0[1], x0[2]}, up[NU] = {uprev[0], uprev[1]};
    double J = 0.0;
    for (int i = 0; i < p; ++i) {
        const int c = i < m ? i : m - 1;
        double u[NU];
#pragma unroll
        for (int j = 0; j < NU; ++j) u[j] = nmg_clamp(fma(alpha, sm.d[NU * c + j], sm.v[NU * c + j]), D.umin[j], D.umax[j]);
        if (i < m) {
#pragma unroll
            for (int j = 0; j < NU; ++j) { const double du = u[j] - up[j]; J = fma(wu2[j] * du, du, J); up[j] = u[j]; }
        }
        nmg_rk4_plain(D, x, u);
#pragma unroll
        for (int j = 0; j < NY; ++j) { const double e = r[j] - x[1 + j]; J = fma(wy2[j] * e, e, J); }
    }
    return J;
}

// exact  min 1/2 d'Hd + g'd,  umin - v <= d <= umax - v,  H SPD (sm.H packed), by the primal active-set method of box_qp
// (mpc_nmpc_core.h): result in sm.d.  0 ok, 2 iteration cap, 3 H not positive definite.
template <int G>
__device__ __noinline__ int nmg_box_qp(const NmpcDev &D, int nz, const NmgSm &sm, const NmGroup<G> &gp) {
    constexpr int NC = (NM_MAXZ + G - 1) / G;
    const int gl = gp.gl;
    double lo[NC], hi[NC], ga[NC];
    double gmx = 0.0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int a = gl + G * k;
        lo[k] = hi[k] = ga[k] = 0.0;
        if (a < nz) {
            ga[k] = sm.g[a]; lo[k] = D.umin[a & 1] - sm.v[a]; hi[k] = D.umax[a & 1] - sm.v[a];
            sm.fixed[a] = (ga[k] > 0.0 && lo[k] >= 0.0) ? -1 : ((ga[k] < 0.0 && hi[k] <= 0.0) ? 1 : 0);
            sm.d[a] = 0.0;
            gmx = fmax(gmx, fabs(ga[k]));
        }
    }
    const double gscale = gp.max(gmx);
    for (int it = 0; it < 6 * nz + 20; ++it) {
        gp.sync();
        int nf = 0;
        for (int b = 0; b < nz; ++b)
            if (sm.fixed[b] == 0) { if (gl == 0) { sm.fl[nf] = b; sm.finv[b] = nf; } nf++; }
        gp.sync();
        // packed H_FF and the right-hand side -(g_F + H_FA d_A)
        {
            const int ntf = nf * (nf + 1) / 2;
            int i = 0, j = gl;
            while (j > i) { j -= i + 1; ++i; }
            for (int t = gl; t < ntf; t += G) {
                sm.Lc[t] = nmg_sym(sm.H, sm.fl[i], sm.fl[j]);
                j += G;
                while (j > i) { j -= i + 1; ++i; }
            }
            for (int k = gl; k < nf; k += G) {
                const int a = sm.fl[k];
                double rhs = -sm.g[a];
                for (int b = 0; b < nz; ++b)
                    if (sm.fixed[b] != 0) rhs = fma(-nmg_sym(sm.H, a, b), sm.d[b], rhs);
                sm.t[k] = rhs;
            }
        }
        gp.sync();
        // Cholesky, right-looking; rows below the pivot are dealt to the lanes
        for (int k = 0; k < nf; ++k) {
            const double dkk = sm.Lc[nmg_tri(k, k)];
            if (!(dkk > 0.0)) return 3;
            const double ckk = sqrt(dkk);
            for (int i = k + 1 + gl; i < nf; i += G) sm.Lc[nmg_tri(i, k)] /= ckk;
            gp.sync();
            if (gl == 0) sm.Lc[nmg_tri(k, k)] = ckk;
            for (int i = k + 1 + gl; i < nf; i += G) {
                const double lik = sm.Lc[nmg_tri(i, k)];
                double *row = sm.Lc + nmg_tri(i, 0);
                for (int j = k + 1; j <= i; ++j) row[j] = fma(-lik, sm.Lc[nmg_tri(j, k)], row[j]);
            }
            gp.sync();
        }
        for (int k = 0; k < nf; ++k) {   // L y = t, y in S0
            const double yk = sm.t[k] / sm.Lc[nmg_tri(k, k)];
            if (gl == 0) sm.S0[k] = yk;
            for (int i = k + 1 + gl; i < nf; i += G) sm.t[i] = fma(-sm.Lc[nmg_tri(i, k)], yk, sm.t[i]);
            gp.sync();
        }
        for (int k = nf - 1; k >= 0; --k) {   // L' x = y, x in S1
            const double xk = sm.S0[k] / sm.Lc[nmg_tri(k, k)];
            if (gl == 0) sm.S1[k] = xk;
            for (int i = gl; i < k; i += G) sm.S0[i] = fma(-sm.Lc[nmg_tri(k, i)], xk, sm.S0[i]);
            gp.sync();
        }
        // longest feasible step toward the Newton point of the face
        double step[NC];
        double alpha = 1.0;
        int blk = -1;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int a = gl + G * k;
            step[k] = 0.0;
            if (a < nz && sm.fixed[a] == 0) {
                const double da = sm.d[a];
                step[k] = sm.S1[sm.finv[a]] - da;
                if (step[k] > 0.0 && da + step[k] > hi[k]) { const double al = (hi[k] - da) / step[k]; if (al < alpha) { alpha = al; blk = a; } }
                if (step[k] < 0.0 && da + step[k] < lo[k]) { const double al = (lo[k] - da) / step[k]; if (al < alpha) { alpha = al; blk = a; } }
            }
        }
        gp.argmin(alpha, blk);
        if (blk < 0) alpha = 1.0;
        gp.sync();
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int a = gl + G * k;
            if (a < nz && sm.fixed[a] == 0) {
                if (a == blk) { const int side = step[k] > 0.0 ? 1 : -1; sm.d[a] = side > 0 ? hi[k] : lo[k]; sm.fixed[a] = side; }
                else sm.d[a] = fma(alpha, step[k], sm.d[a]);
            }
        }
        if (blk >= 0) continue;
        gp.sync();
        // minimiser of the face: release the bound with the most wrong-signed multiplier
        double key = 0.0;
        int cand = -1;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const int a = gl + G * k;
            if (a < nz && sm.fixed[a] != 0) {
                double gi = ga[k];
                for (int b = 0; b < nz; ++b) gi = fma(nmg_sym(sm.H, a, b), sm.d[b], gi);
                const double viol = sm.fixed[a] < 0 ? -gi : gi;
                if (viol > 0.0 && (cand < 0 || -viol < key)) { key = -viol; cand = a; }
            }
        }
        gp.argmin(key, cand);
        if (cand < 0 || -key <= 1e-14 * gscale) { gp.sync(); return 0; }
        if ((cand % G) == gl) sm.fixed[cand] = 0;
    }
    gp.sync();
    return 2;
}

// one nlmpcmove by one group: plan in sm.v (in: start, out: optimum)
template <int G>
__device__ __noinline__ int nmg_nlmpcmove(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m, const double *wy2,
                             const double *wu2, const NmgSm &sm, const NmGroup<G> &gp, unsigned *n_sqp) {
    const int nz = NU * m, gl = gp.gl;
    gp.sync();   // every lane has read what it wanted from the previous plan (the caller takes u(k) = v[0..1] from it)
    for (int a = gl; a < nz; a += G) { sm.v[a] = nmg_clamp(sm.v[a], D.umin[a & 1], D.umax[a & 1]); sm.d[a] = 0.0; }
    gp.sync();
    double Jcur = nmg_model<G>(D, x0, uprev, r, p, m, wy2, wu2, sm, gp, 0.0);
    for (int it = 0; it < D.max_sqp; ++it) {
        *n_sqp += 1;
        const int rc = nmg_box_qp<G>(D, nz, sm, gp);
        if (rc) return rc;
        double dmx = 0.0;
        for (int a = gl; a < nz; a += G) dmx = fmax(dmx, fabs(sm.d[a]) / D.su[a & 1]);
        if (gp.max(dmx) < 1e-10) break;
        // full step first, through the model: accepted, it is the next iteration's model
        double Jn = nmg_model<G>(D, x0, uprev, r, p, m, wy2, wu2, sm, gp, 1.0);
        double alpha = 1.0;
        const bool full = Jn < Jcur;
        if (!full) {
            // the shorter steps 1/2 .. 1/32 at once, one per lane; the first (largest) that decreases the cost is taken
            const int bt = gl < 4 ? gl + 1 : 5;
            const double al = 1.0 / (double)(1 << bt);
            const double Jl = nmg_plan_cost<G>(D, x0, uprev, r, p, m, wy2, wu2, sm, al);
            const unsigned okm = gp.ballot(gl < 5 && Jl < Jcur);
            if (!okm) break;   // no descent at this resolution: converged to rounding
            const int win = __ffs((int)okm) - 1;
            alpha = 1.0 / (double)(1 << (win + 1));
        }
        gp.sync();
        for (int a = gl; a < nz; a += G) { sm.v[a] = nmg_clamp(fma(alpha, sm.d[a], sm.v[a]), D.umin[a & 1], D.umax[a & 1]); sm.d[a] = 0.0; }
        gp.sync();
        if (!full) Jn = nmg_model<G>(D, x0, uprev, r, p, m, wy2, wu2, sm, gp, 0.0);
        Jcur = Jn;
    }
    return 0;
}

// mode 0 RAW, 1 GAM, 2 VNS.  One group of G lanes per (candidate, run); 32 / G runs per warp, one warp per CTA.  maxz: the
// largest plan (2 Nu) among the items of this launch.
// (the body is a device function so that tests/host_emulation can run this source on a CPU, one warp of host threads per CTA)
template <int G>
__device__ __forceinline__ void nmg_run(const NmpcDev &D, int item0, int item1, int runs, int mode, const int *order, const NmpcArgs &A,
                                        int maxz, double *smem_g, int block) {
    const int lane = threadIdx.x & 31, grp = lane / G;
    const int item = item0 + block * (32 / G) + grp;   // items [item0, item1) of the sorted order
    if (item >= item1) return;
    NmGroup<G> gp;
    gp.gl = lane % G;
    gp.mask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (grp * G);
    const int gl = gp.gl;
    const int c = order[item / runs], run = item - (item / runs) * runs;   // longest horizons first
    const int p = A.N[c], m = A.Nu[c], nit = D.nit;
    if (p < 2 || p > D.pmax || m < 1 || m > D.mmax || m >= p) {
        if (gl == 0) {
            A.status[c] = MPCGPU_CAND_INVALID;
            if (mode == 1) for (int j = 0; j < NY; ++j) A.cost[(size_t)c * NY + j] = NAN;
            if (mode == 2) A.part[(size_t)c * runs + run] = NAN;
        }
        return;
    }
    NmgSm sm;
    {
        const int tri = maxz * (maxz + 1) / 2;
        double *q_ = smem_g + (size_t)grp * nmg_doubles(maxz);
        sm.H = q_; q_ += tri; sm.Lc = q_; q_ += tri;
        sm.v = q_; q_ += maxz; sm.vo = q_; q_ += maxz; sm.g = q_; q_ += maxz; sm.d = q_; q_ += maxz; sm.t = q_; q_ += maxz;
        sm.S0 = q_; q_ += maxz; sm.S1 = q_; q_ += maxz; sm.AB = q_; q_ += 16;
        sm.fixed = (int *)q_; sm.fl = sm.fixed + maxz; sm.finv = sm.fl + maxz;
    }
    const int nz = NU * m;
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
    for (int a = gl; a < nz; a += G) { sm.v[a] = D.u0[a & 1]; sm.vo[a] = D.u0[a & 1]; }
    gp.sync();
    if (want_ol) {   // open-loop optimum from (x0, u0) toward r(:, end)  (closedloop_toolbox_nmpc.m:79-94)
        ref_at(nit - 1, rr);
        const int rc = nmg_nlmpcmove<G>(D, D.x0, D.u0, rr, p, m, wy2, wu2, sm, gp, &n_sqp);
        n_calls++;
        if (rc) status = rc;
        gp.sync();
        for (int a = gl; a < nz; a += G) sm.vo[a] = sm.v[a];
        gp.sync();
        if (mode == 2) {   // Jnu (VNS2.m:183-191) on input `sel`
            const int j = sel;
            const double u0a = fabs(sm.vo[j]);
            for (int cc = 0; cc + 1 < m && cc + 1 < nit; ++cc) {
                const double xn = u0a / fabs(sm.vo[NU * (cc + 1) + j] - sm.vo[NU * cc + j]);
                if (fabs(xn) <= 1.7976931348623157e308) jnu += xn * xn;
            }
        }
        for (int a = gl; a < nz; a += G) sm.v[a] = D.u0[a & 1];
        gp.sync();
    }
    double x[NX] = {D.x0[0], D.x0[1], D.x0[2]}, uprev[NU] = {D.u0[0], D.u0[1]};
    for (int k = 0; k < nit; ++k) {
        const int cc = k < m ? k : m - 1;
        const double uo[NU] = {sm.vo[NU * cc], sm.vo[NU * cc + 1]};
        if (k > 0) {
            ref_at(k, rr);
            const int rc = nmg_nlmpcmove<G>(D, x, uprev, rr, p, m, wy2, wu2, sm, gp, &n_sqp);   // warm start = previous plan
            n_calls++;
            if (rc) status = rc;
            gp.sync();
            uprev[0] = sm.v[0]; uprev[1] = sm.v[1];
            nmg_rk4_plain(D, x, uprev);
            for (int i = 0; i < NX; ++i)
                if (x[i] < D.xmin[i] - 1e-9 || x[i] > D.xmax[i] + 1e-9) { if (!status) status = 5; }
            if (want_ol) nmg_rk4_plain(D, xo, uo);
        }
        for (int j = 0; j < NY; ++j) {
            const bool mine = sel < 0 || sel == j;
            const double yj = x[1 + j], yoj = xo[1 + j], yr = A.yref[(size_t)j * nit + k];
            if (mine) {
                if (gl == 0) {
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
    if (gl == 0) {
        if (mode == 1) for (int j = 0; j < NY; ++j) A.cost[(size_t)c * NY + j] = (status == 0 || status == 5) ? cost_acc[j] : NAN;
        if (mode == 2) A.part[(size_t)c * runs + run] = (status == 0 || status == 5) ? vns_acc + jnu : NAN;
        if (status) atomicMax(A.status + c, status);
        atomicAdd(A.counters + 0, (unsigned long long)n_calls);
        atomicAdd(A.counters + 1, (unsigned long long)n_sqp);
    }
}

#ifndef MPC_SIMT_EMULATION
template <int G>
__global__ void __launch_bounds__(32, NMG_MINB) k_nmpc_g(const NmpcDev D, int item0, int item1, int runs, int mode, const int *order, NmpcArgs A, int maxz) {
    extern __shared__ double smem_g[];
    nmg_run<G>(D, item0, item1, runs, mode, order, A, maxz, smem_g, (int)blockIdx.x);
}
#endif
