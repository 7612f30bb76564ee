// mpc_ssnmpc_core.h -- the second nonlinear formulation of the reference: the single-shooting NMPC
// of `Explicit NMPC/` (NMPC_Controller.m:1-141 controller, ClosedLoopNMPC.m:1-110 loop, plant_model.m:1-56 = the Van de Vusse
// right-hand side of mpc_nmpc_core.h).  Plain per-run code like mpc_nmpc_core.h: compiled by nvcc into k_ssnmpc (one thread per
// run) and, unchanged, by g++ into oracle/nmpc_port (host tests, CPU baseline).  What differs from nlmpcmove (N1-N4):
//   S1  decision X: per input j a block of Nu_j OFFSETS from u_j(k-1) -- u_j(k+i) = u_j(k-1) + X_j[min(i, Nu_j-1)]
//       (NMPC_Controller.m:87-96: not cumulative increments), each input with its OWN control horizon, started at X = 0 (:26)
//   S2  cost  sum_j Q_j sum_{i=1..N} (r_j(k) - (y_j(k+i) + n_j))^2 + sum_j W_j sum_c X_jc^2   (:128-138; weights NOT squared,
//       no scale factors: Qm = Q_j I, Wm = W_j I, ClosedLoopNMPC.m:29-39)
//   S3  n_j = x_cj(k) - [model step from x(k) under u(k-1)]_cj: the "model deviation" of :106-123 (computed from the CURRENT
//       state, so it is minus the one-step change of the model, not zero, also without mismatch -- reproduced as written)
//   S4  bounds lb - u(k-1) <= X <= ub - u(k-1) only (:14-23); no state / output bounds in this formulation
//   S5  u(k) = u(k-1) + X_j[0] (ClosedLoopNMPC.m:97-105); plant step, then the state perturbation noise(:,k) the caller
//       supplies (the reference draws 0.01*randn per sample, :88-90: pass the draws to reproduce a run, nullptr for none)
// Integrator: RK4 with nsub sub-steps where the reference uses ode23t (prediction) / ode45 (plant) -- the same stated
// deviation as N4.  Solver: Gauss-Newton on S2 with the exact box-QP of mpc_nmpc_core.h and backtracking on the true cost
// (plus a step that removes the slowest linearly converging mode, and plain steps once the cost no longer resolves them),
// where the reference calls fmincon-SQP (TolX 1e-6, TolFun 1e-7): the same minimiser to the tolerance stated in tests/.
#pragma once
#include "mpc_nmpc_core.h"

struct SsnmpcDev {
    NmpcDev D;        // nit, inK, nsub, max_sqp, Ts, x0, u0, umin = lb, umax = ub, su = ub - lb (step scaling of the stop test)
    int xc[NY];       // x_control, 0-based (main.m:75: [2 3] -> {1, 2})
    int pmax;         // N <= pmax
};

// cost S2 of plan X at (x, uprev); bias = n_j of S3
NM_FN double ss_cost(const SsnmpcDev &S, const double *x0, const double *uprev, const double *r, const double *bias, int p,
                     const int *nuj, const int *off, const double *Q, const double *W, const double *X) {
    double x[NX] = {x0[0], x0[1], x0[2]}, uf[NU];
    double J = 0.0;
    for (int i = 0; i < p; ++i) {
        for (int j = 0; j < NU; ++j) uf[j] = uprev[j] + X[off[j] + (i < nuj[j] ? i : nuj[j] - 1)];
        rk4_sample(S.D, x, uf, nullptr);
        for (int j = 0; j < NY; ++j) { const double e = r[j] - (x[S.xc[j]] + bias[j]); J = fma(Q[j] * e, e, J); }
    }
    for (int j = 0; j < NU; ++j)
        for (int c = 0; c < nuj[j]; ++c) J = fma(W[j] * X[off[j] + c], X[off[j] + c], J);
    return J;
}

// one NMPC_Controller call: X out (nz = sum Nu_j).  H, Lc: NM_LD x NM_LD scratch.  Returns 0 ok, 2 / 3 from box_qp.
NM_FN int ss_controller(const SsnmpcDev &S, const double *x0, const double *uprev, const double *r, int p, const int *nuj,
                        const double *Q, const double *W, double *X, double *H, double *Lc, unsigned *n_sqp) {
    int off[NU];
    int nz = 0;
    for (int j = 0; j < NU; ++j) { off[j] = nz; nz += nuj[j]; }
    double g[NM_MAXZ], d[NM_MAXZ], dp[NM_MAXZ], lo[NM_MAXZ], hi[NM_MAXZ], tmp[NM_MAXZ], Xt[NM_MAXZ], Xx[NM_MAXZ], Sx[NX * NM_MAXZ], AB[15], bias[NY], uf[NU];
    int fixed[NM_MAXZ];
    {   // S3
        double xs[NX] = {x0[0], x0[1], x0[2]};
        rk4_sample(S.D, xs, uprev, nullptr);
        for (int j = 0; j < NY; ++j) bias[j] = x0[S.xc[j]] - xs[S.xc[j]];
    }
    for (int i = 0; i < nz; ++i) X[i] = 0.0;                                   // :26
    double Jcur = ss_cost(S, x0, uprev, r, bias, p, nuj, off, Q, W, X);
    int status = 0;
    double dprev = INFINITY;
    int prev_plain = 0;
    for (int it = 0; it < S.D.max_sqp; ++it) {
        *n_sqp += 1;
        for (int a = 0; a < nz; ++a)
            for (int b = 0; b <= a; ++b) H[a * NM_LD + b] = 0.0;
        for (int i = 0; i < nz; ++i) g[i] = 0.0;
        for (int i = 0; i < NX * nz; ++i) Sx[i] = 0.0;
        double x[NX] = {x0[0], x0[1], x0[2]};
        for (int i = 0; i < p; ++i) {
            int col[NU];
            for (int j = 0; j < NU; ++j) { col[j] = off[j] + (i < nuj[j] ? i : nuj[j] - 1); uf[j] = uprev[j] + X[col[j]]; }
            rk4_sample(S.D, x, uf, AB);
            for (int c = 0; c < nz; ++c) {                                     // Sx <- A Sx + B E
                const double a0 = Sx[0 * nz + c], a1 = Sx[1 * nz + c], a2 = Sx[2 * nz + c];
                for (int rr = 0; rr < NX; ++rr) Sx[rr * nz + c] = AB[rr * 5 + 0] * a0 + AB[rr * 5 + 1] * a1 + AB[rr * 5 + 2] * a2;
            }
            for (int rr = 0; rr < NX; ++rr)
                for (int j = 0; j < NU; ++j) Sx[rr * nz + col[j]] += AB[rr * 5 + NX + j];
            for (int j = 0; j < NY; ++j) {
                const double *Sr = Sx + S.xc[j] * nz;
                const double e = r[j] - (x[S.xc[j]] + bias[j]);
                for (int a = 0; a < nz; ++a) {
                    const double wa = Q[j] * Sr[a];
                    g[a] = fma(-wa, e, g[a]);
                    for (int b = 0; b <= a; ++b) H[a * NM_LD + b] = fma(wa, Sr[b], H[a * NM_LD + b]);
                }
            }
        }
        for (int j = 0; j < NU; ++j)
            for (int c = 0; c < nuj[j]; ++c) { const int a = off[j] + c; g[a] = fma(W[j], X[a], g[a]); H[a * NM_LD + a] += W[j]; }
        for (int a = 0; a < nz; ++a)
            for (int b = a + 1; b < nz; ++b) H[a * NM_LD + b] = H[b * NM_LD + a];
        for (int j = 0; j < NU; ++j)
            for (int c = 0; c < nuj[j]; ++c) {                                 // S4
                const int a = off[j] + c;
                lo[a] = fmin(S.D.umin[j] - uprev[j] - X[a], 0.0); hi[a] = fmax(S.D.umax[j] - uprev[j] - X[a], 0.0);
            }
        const int rc = box_qp(nz, H, g, lo, hi, d, Lc, tmp, fixed);
        if (rc) { status = rc; break; }
        double dmax = 0.0;
        for (int j = 0; j < NU; ++j)
            for (int c = 0; c < nuj[j]; ++c) dmax = fmax(dmax, fabs(d[off[j] + c]) / S.D.su[j]);
        if (dmax < 1e-10) break;
        double alpha = 1.0, Jn = 0.0;
        int acc_ = 0;
        // Gauss-Newton converges LINEARLY here after a set-point jump (large residual, W ~ 1e-4): the error contracts by a constant
        // factor rho per full step along one direction -- measured rho ~ -0.95, the steps alternate in sign and shrink by 5 %
        // (the Gauss-Newton Hessian underestimates the curvature there and every step overshoots to the other side).  When
        // two successive full steps show that pattern (parallel or anti-parallel), rho = <d, d_prev> / <d_prev, d_prev> and the
        // step that removes the mode, alpha = 1 / (1 - rho), is tried first; it is kept only if it lowers the true cost below
        // the plain step's (it changes how fast the minimiser is reached, not which one)
        double Jx = INFINITY;
        if (prev_plain) {
            double dot = 0.0, n1 = 0.0, n2 = 0.0;
            for (int i = 0; i < nz; ++i) { dot = fma(d[i], dp[i], dot); n1 = fma(d[i], d[i], n1); n2 = fma(dp[i], dp[i], n2); }
            const double rho = n2 > 0.0 ? dot / n2 : 0.0;
            if (fabs(dot) > 0.95 * sqrt(n1 * n2) && fabs(rho) > 0.3 && fabs(rho) < 0.995) {
                const double ax = fmin(1.0 / (1.0 - rho), 32.0);
                for (int j = 0; j < NU; ++j)
                    for (int c = 0; c < nuj[j]; ++c) {
                        const int a = off[j] + c;
                        Xx[a] = fmin(fmax(X[a] + ax * d[a], S.D.umin[j] - uprev[j]), S.D.umax[j] - uprev[j]);
                    }
                Jx = ss_cost(S, x0, uprev, r, bias, p, nuj, off, Q, W, Xx);
            }
        }
        for (int i = 0; i < nz; ++i) dp[i] = d[i];
        for (int bt = 0; bt < 6; ++bt) {
            for (int j = 0; j < NU; ++j)
                for (int c = 0; c < nuj[j]; ++c) {
                    const int a = off[j] + c;
                    Xt[a] = fmin(fmax(X[a] + alpha * d[a], S.D.umin[j] - uprev[j]), S.D.umax[j] - uprev[j]);
                }
            Jn = ss_cost(S, x0, uprev, r, bias, p, nuj, off, Q, W, Xt);
            if (Jn < Jcur) { acc_ = 1; break; }
            alpha *= 0.5;
        }
        prev_plain = acc_ && alpha == 1.0;
        if (Jx < Jcur && (!acc_ || Jx < Jn)) {     // the mode-removing step wins
            for (int i = 0; i < nz; ++i) Xt[i] = Xx[i];
            Jn = Jx; acc_ = 1; prev_plain = 0;     // no pattern to read off the step that follows it
        }
        if (!acc_) {
            // the cost no longer resolves the step (|dJ| < 1e-16 J leaves X open to ~1e-8 along the flat directions, and the
            // closed loop amplifies that): finish with plain Gauss-Newton steps for as long as they contract
            if (dmax >= 1e-6 || dmax >= dprev) break;
            for (int j = 0; j < NU; ++j)
                for (int c = 0; c < nuj[j]; ++c) {
                    const int a = off[j] + c;
                    Xt[a] = fmin(fmax(X[a] + d[a], S.D.umin[j] - uprev[j]), S.D.umax[j] - uprev[j]);
                }
            Jn = ss_cost(S, x0, uprev, r, bias, p, nuj, off, Q, W, Xt);
        }
        dprev = dmax;
        for (int i = 0; i < nz; ++i) X[i] = Xt[i];
        Jcur = Jn;
    }
    return status;
}

// One closed loop of ClosedLoopNMPC.m:61-108.  r: ny x nit; noise: nx x nit or nullptr; y, u: ny|nu x nit blocks or nullptr;
// cost (if not nullptr): ny sums of (y_j(k) - r_j(k))^2 over k = inK..nit (1-based) -- the sweep objective of this library
// (the reference's demo has no tuning cost; the window is the simulated part of the run, main.m:80-86).  Returns status.
NM_FN int ssnmpc_run(const SsnmpcDev &S, int p, const int *nuj, const double *Q, const double *W, const double *r,
                     const double *noise, double *y, double *u, double *cost, double *H, double *Lc, unsigned *n_calls_out,
                     unsigned *n_sqp_out) {
    const int nit = S.D.nit, k0 = S.D.inK - 1;                                 // 0-based first simulated sample
    double x[NX] = {S.D.x0[0], S.D.x0[1], S.D.x0[2]}, uprev[NU] = {S.D.u0[0], S.D.u0[1]}, X[NM_MAXZ], rr[NY], acc[NY] = {0.0, 0.0};
    unsigned n_sqp = 0, n_calls = 0;
    int status = 0;
    for (int k = 0; k < nit; ++k) {
        if (k >= k0) {
            rk4_sample(S.D, x, uprev, nullptr);                               // :82-86 plant under u(:, k-1)
            if (noise) for (int i = 0; i < NX; ++i) x[i] += noise[(size_t)i * nit + k];
            for (int j = 0; j < NY; ++j) rr[j] = r[(size_t)j * nit + k];
            const int rc = ss_controller(S, x, uprev, rr, p, nuj, Q, W, X, H, Lc, &n_sqp);
            n_calls++;
            if (rc && !status) status = rc;
            int o = 0;
            for (int j = 0; j < NU; ++j) { uprev[j] += X[o]; o += nuj[j]; }    // S5
        }
        for (int j = 0; j < NY; ++j) {
            const double yj = x[S.xc[j]];                                      // y(:, k) is measured BEFORE u(:, k) acts (:93)
            if (y) y[(size_t)j * nit + k] = yj;
            if (k >= k0) { const double e = yj - r[(size_t)j * nit + k]; acc[j] = fma(e, e, acc[j]); }
        }
        if (u) for (int j = 0; j < NU; ++j) u[(size_t)j * nit + k] = uprev[j];
    }
    if (cost) for (int j = 0; j < NY; ++j) cost[j] = status == 0 ? acc[j] : NAN;
    *n_calls_out = n_calls; *n_sqp_out = n_sqp;
    return status;
}

// ------------------------------------------------------------------------------------------------
// One item of a population: what a thread of k_ssnmpc does (validation of the horizons, the run, the stores).  Shared with the
// host build so that the indexing and the edge cases the host tests exercise are the kernel's own.
// ------------------------------------------------------------------------------------------------
struct SsArgs {
    const int *N, *Nu;           // n, n x NU
    const double *Q, *W;         // n x NY, n x NU
    const double *r, *noise;     // NY x nit, NX x nit or nullptr
    double *cost, *y, *u;        // n x NY; n x NY x nit, n x NU x nit or nullptr
    int *status;
    unsigned long long *counters;   // [0] controller calls, [1] Gauss-Newton iterations
};
#ifdef __CUDACC__
#define SS_COUNT(p, v) atomicAdd((p), (unsigned long long)(v))
#else
#define SS_COUNT(p, v) __atomic_fetch_add((p), (unsigned long long)(v), __ATOMIC_RELAXED)
#endif

NM_FN void ss_item(const SsnmpcDev &S, const int *order, const SsArgs &A, int item, double *H, double *Lc) {
    const int c = order ? order[item] : item, nit = S.D.nit;
    const int p = A.N[c];
    int nuj[NU], nz = 0;
    bool ok = p >= 1 && p <= S.pmax;
    for (int j = 0; j < NU; ++j) { nuj[j] = A.Nu[(size_t)c * NU + j]; ok = ok && nuj[j] >= 1 && nuj[j] <= p; nz += nuj[j]; }
    if (!ok || nz > NM_MAXZ) {
        A.status[c] = 4;   /* MPCGPU_CAND_INVALID */
        if (A.cost) for (int j = 0; j < NY; ++j) A.cost[(size_t)c * NY + j] = NAN;
        return;
    }
    unsigned n_calls = 0, n_sqp = 0;
    A.status[c] = ssnmpc_run(S, p, nuj, A.Q + (size_t)c * NY, A.W + (size_t)c * NU, A.r, A.noise,
                             A.y ? A.y + (size_t)c * NY * nit : nullptr, A.u ? A.u + (size_t)c * NU * nit : nullptr,
                             A.cost ? A.cost + (size_t)c * NY : nullptr, H, Lc, &n_calls, &n_sqp);
    if (A.counters) { SS_COUNT(A.counters + 0, n_calls); SS_COUNT(A.counters + 1, n_sqp); }
}
