// mpc_nmpc_core.h -- the per-run algorithm of the nonlinear path (one nlmpcmove = Gauss-Newton SQP on the restated NLP,
// N1-N4 of include/mpcgpu.h): Van de Vusse right-hand side (vandevusse_model.m:39-77), RK4 sample with sensitivities,
// plan cost, exact box-QP, the SQP loop.  Plain C-style code on per-run data: compiled by nvcc into k_nmpc (one THREAD per
// run) and, unchanged, by g++ into oracle/nmpc_port (the CPU baseline of bench.py --config vdv: same algorithm, OpenMP over
// candidates).  The scipy-based checker of the tests stays the independent reference for parity.
#pragma once
#include <math.h>
#ifndef NM_FN   /* a second translation unit of the library defines it `static __device__` (mpc_ssnmpc.cu) */
#ifdef __CUDACC__
#define NM_FN __device__
#else
#define NM_FN static inline
#endif
#endif

#define NX 3
#define NU 2
#define NY 2
#define NM_MAXM 15
#define NM_MAXZ (NU * NM_MAXM)
#define NM_THREADS 64
#define NM_LD NM_MAXZ   /* leading dimension of H and of its factor: the same for every run, so that equal (row, col) means equal address offset across a warp */

struct NmpcDev {
    int nit, pmax, mmax, inK, nsub, max_sqp;
    double Ts;
    double x0[NX], u0[NU], umin[NU], umax[NU], xmin[NX], xmax[NX], su[NU], sy[NY];
};

// vandevusse_model.m:42-57
#define VDV_K10 1.287e12
#define VDV_K20 1.287e12
#define VDV_K30 9.043e9
#define VDV_E1 (-9758.3)
#define VDV_E2 (-9758.3)
#define VDV_E3 (-8560.0)
#define VDV_DAB (-4.20)
#define VDV_DBC 11.00
#define VDV_DAD 41.85
#define VDV_RHO 0.9342
#define VDV_CP 3.01
#define VDV_KW 4032.0
#define VDV_AR 0.215
#define VDV_V 10.0
#define VDV_T0 130.00
#define VDV_CA0 5.10

// f(x,u) and, if J != nullptr, J = [df/dx | df/du] (3 x 5, row-major)
NM_FN void vdv_rhs(const double *x, const double *u, double *f, double *J) {
    const double fov = u[0], Tk = u[1], ca = x[0], cb = x[1], T = x[2];
    const double Tk_ = T + 273.15, iT = 1.0 / Tk_;
    const double k1 = VDV_K10 * exp(VDV_E1 * iT), k3 = VDV_K30 * exp(VDV_E3 * iT);
    const double k2 = (VDV_K20 == VDV_K10 && VDV_E2 == VDV_E1) ? k1 : VDV_K20 * exp(VDV_E2 * iT);   // the reference's k20, E2 equal k10, E1
    const double irc = 1.0 / (VDV_RHO * VDV_CP), beta = VDV_KW * VDV_AR / (VDV_RHO * VDV_CP * VDV_V);
    f[0] = fov * (VDV_CA0 - ca) - k1 * ca - k3 * ca * ca;
    f[1] = -fov * cb + k1 * ca - k2 * cb;
    f[2] = irc * (k1 * ca * VDV_DAB + k2 * cb * VDV_DBC + k3 * ca * ca * VDV_DAD) + fov * (VDV_T0 - T) + beta * (Tk - T);
    if (J) {
        const double d1 = k1 * (-VDV_E1) * iT * iT, d2 = k2 * (-VDV_E2) * iT * iT, d3 = k3 * (-VDV_E3) * iT * iT;   // dk/dT
        J[0] = -fov - k1 - 2.0 * k3 * ca; J[1] = 0.0; J[2] = -d1 * ca - d3 * ca * ca; J[3] = VDV_CA0 - ca; J[4] = 0.0;
        J[5] = k1; J[6] = -fov - k2; J[7] = d1 * ca - d2 * cb; J[8] = -cb; J[9] = 0.0;
        J[10] = irc * (k1 * VDV_DAB + 2.0 * k3 * ca * VDV_DAD); J[11] = irc * k2 * VDV_DBC;
        J[12] = irc * (d1 * ca * VDV_DAB + d2 * cb * VDV_DBC + d3 * ca * ca * VDV_DAD) - fov - beta;
        J[13] = VDV_T0 - T; J[14] = beta;
    }
}

// one sample (nsub RK4 steps); if AB != nullptr also [A|B] = [dx+/dx | dx+/du] (3 x 5)
NM_FN void rk4_sample(const NmpcDev &D, double *x, const double *u, double *AB) {
    const double h = D.Ts / D.nsub;
    if (AB) {
        for (int i = 0; i < 15; ++i) AB[i] = 0.0;
        AB[0] = AB[6] = AB[12] = 1.0;
    }
    for (int s = 0; s < D.nsub; ++s) {
        double k[4][NX], xs[NX], Jm[15], Dk[4][15], Dx[15];
        const double ca[4] = {0.0, 0.5, 0.5, 1.0};
        for (int st = 0; st < 4; ++st) {
            for (int i = 0; i < NX; ++i) xs[i] = st == 0 ? x[i] : x[i] + ca[st] * h * k[st - 1][i];
            vdv_rhs(xs, u, k[st], AB ? Jm : nullptr);
            if (AB) {
                // Dx = d xs / d(x_sub, u) = [I|0] + ca*h*Dk[st-1];  Dk = Jx Dx + [0|Ju]
                for (int i = 0; i < 15; ++i) Dx[i] = st == 0 ? 0.0 : ca[st] * h * Dk[st - 1][i];
                Dx[0] += 1.0; Dx[6] += 1.0; Dx[12] += 1.0;
                for (int r = 0; r < NX; ++r)
                    for (int c = 0; c < 5; ++c) {
                        double acc = c >= NX ? Jm[r * 5 + c] : 0.0;
                        for (int q = 0; q < NX; ++q) acc = fma(Jm[r * 5 + q], Dx[q * 5 + c], acc);
                        Dk[st][r * 5 + c] = acc;
                    }
            }
        }
        for (int i = 0; i < NX; ++i) x[i] += (h / 6.0) * (k[0][i] + 2.0 * k[1][i] + 2.0 * k[2][i] + k[3][i]);
        if (AB) {
            double Phi[15], nAB[15];   // transition of this sub-step, then composition with what came before
            for (int i = 0; i < 15; ++i) Phi[i] = (h / 6.0) * (Dk[0][i] + 2.0 * Dk[1][i] + 2.0 * Dk[2][i] + Dk[3][i]);
            Phi[0] += 1.0; Phi[6] += 1.0; Phi[12] += 1.0;
            for (int r = 0; r < NX; ++r)
                for (int c = 0; c < 5; ++c) {
                    double acc = c >= NX ? Phi[r * 5 + c] : 0.0;
                    for (int q = 0; q < NX; ++q) acc = fma(Phi[r * 5 + q], AB[q * 5 + c], acc);
                    nAB[r * 5 + c] = acc;
                }
            for (int i = 0; i < 15; ++i) AB[i] = nAB[i];
        }
    }
}

// predicted cost of plan v (N2)
NM_FN double plan_cost(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m,
                            const double *wy2, const double *wu2, const double *v) {
    double x[NX] = {x0[0], x0[1], x0[2]};
    double J = 0.0;
    for (int i = 0; i < p; ++i) {
        const int c = i < m ? i : m - 1;
        rk4_sample(D, x, v + NU * c, nullptr);
        for (int j = 0; j < NY; ++j) { const double e = r[j] - x[1 + j]; J = fma(wy2[j] * e, e, J); }
    }
    for (int c = 0; c < m; ++c)
        for (int j = 0; j < NU; ++j) {
            const double du = v[NU * c + j] - (c == 0 ? uprev[j] : v[NU * (c - 1) + j]);
            J = fma(wu2[j] * du, du, J);
        }
    return J;
}

// exact solution of  min 1/2 d'Hd + g'd,  lo <= d <= hi  (lo <= 0 <= hi), H SPD (nz x nz, row-major), by a primal
// active-set method started at d = 0.  Lc: scratch nz*nz.  Returns 0 ok, 2 iteration cap, 3 H not PD.
NM_FN int box_qp(int nz, const double *H, const double *g, const double *lo, const double *hi, double *d, double *Lc,
                      double *tmp, int *fixed) {
    for (int i = 0; i < nz; ++i) {
        d[i] = 0.0;
        // bound is binding at the start when the gradient pushes against a bound of zero width in that direction
        fixed[i] = (g[i] > 0.0 && lo[i] >= 0.0) ? -1 : ((g[i] < 0.0 && hi[i] <= 0.0) ? 1 : 0);
    }
    for (int it = 0; it < 6 * nz + 20; ++it) {
        // reduced Newton point on the free variables: H_FF t_F = -(g_F + H_FA d_A)
        int nf = 0;
        for (int i = 0; i < nz; ++i) if (!fixed[i]) nf++;
        // Cholesky of H_FF (in the index space of the free variables, stored packed in Lc[nz*nz])
        int fi = 0;
        for (int i = 0; i < nz; ++i) {
            if (fixed[i]) continue;
            int fj = 0;
            for (int j = 0; j <= i; ++j) {
                if (fixed[j]) continue;
                double acc = H[i * NM_LD + j];
                for (int k = 0; k < fj; ++k) acc -= Lc[fi * NM_LD + k] * Lc[fj * NM_LD + k];
                if (j == i) {
                    if (!(acc > 0.0)) return 3;
                    Lc[fi * NM_LD + fj] = sqrt(acc);
                } else {
                    Lc[fi * NM_LD + fj] = acc / Lc[fj * NM_LD + fj];
                }
                fj++;
            }
            fi++;
        }
        fi = 0;
        for (int i = 0; i < nz; ++i) {
            if (fixed[i]) continue;
            double acc = -g[i];
            for (int j = 0; j < nz; ++j) if (fixed[j]) acc -= H[i * NM_LD + j] * d[j];
            for (int k = 0; k < fi; ++k) acc -= Lc[fi * NM_LD + k] * tmp[k];
            tmp[fi] = acc / Lc[fi * NM_LD + fi];
            fi++;
        }
        for (int a = nf - 1; a >= 0; --a) {
            double acc = tmp[a];
            for (int k = a + 1; k < nf; ++k) acc -= Lc[k * NM_LD + a] * tmp[k];
            tmp[a] = acc / Lc[a * NM_LD + a];
        }
        // longest feasible step from d toward the Newton point
        double alpha = 1.0;
        int blk = -1, side = 0;
        fi = 0;
        for (int i = 0; i < nz; ++i) {
            if (fixed[i]) continue;
            const double step = tmp[fi] - d[i];
            if (step > 0.0 && d[i] + step > hi[i]) { const double a = (hi[i] - d[i]) / step; if (a < alpha) { alpha = a; blk = i; side = 1; } }
            if (step < 0.0 && d[i] + step < lo[i]) { const double a = (lo[i] - d[i]) / step; if (a < alpha) { alpha = a; blk = i; side = -1; } }
            fi++;
        }
        fi = 0;
        for (int i = 0; i < nz; ++i) {
            if (fixed[i]) continue;
            d[i] += alpha * (tmp[fi] - d[i]);
            fi++;
        }
        if (blk >= 0) { d[blk] = side > 0 ? hi[blk] : lo[blk]; fixed[blk] = side; continue; }
        // at the minimiser of the current face: release the bound with the most wrong-signed multiplier
        double worst = 0.0;
        int rel = -1;
        for (int i = 0; i < nz; ++i) {
            if (!fixed[i]) continue;
            double gi = g[i];
            for (int j = 0; j < nz; ++j) gi = fma(H[i * NM_LD + j], d[j], gi);
            const double viol = fixed[i] < 0 ? -gi : gi;   // lower bound needs gi >= 0, upper bound gi <= 0
            if (viol > worst) { worst = viol; rel = i; }
        }
        if (rel < 0) return 0;
        double scale = 0.0;
        for (int i = 0; i < nz; ++i) scale = fmax(scale, fabs(g[i]));
        if (worst <= 1e-14 * scale) return 0;
        fixed[rel] = 0;
    }
    return 2;
}

// one nlmpcmove: plan v (in: start, out: optimum).  Returns status.
NM_FN int nlmpcmove(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m,
                         const double *wy2, const double *wu2, double *v, double *H, double *Lc, unsigned *n_sqp) {
    const int nz = NU * m;
    double g[NM_MAXZ], d[NM_MAXZ], lo[NM_MAXZ], hi[NM_MAXZ], tmp[NM_MAXZ], vt[NM_MAXZ], X[NX * NM_MAXZ], AB[15];
    int fixed[NM_MAXZ];
    for (int i = 0; i < nz; ++i) v[i] = fmin(fmax(v[i], D.umin[i % NU]), D.umax[i % NU]);
    double Jcur = plan_cost(D, x0, uprev, r, p, m, wy2, wu2, v);
    int status = 0;
    for (int it = 0; it < D.max_sqp; ++it) {
        *n_sqp += 1;
        // ---- Gauss-Newton model at v ----
        for (int a = 0; a < nz; ++a)
            for (int b = 0; b <= a; ++b) H[a * NM_LD + b] = 0.0;
        for (int i = 0; i < nz; ++i) g[i] = 0.0;
        for (int i = 0; i < NX * nz; ++i) X[i] = 0.0;
        double x[NX] = {x0[0], x0[1], x0[2]};
        for (int i = 0; i < p; ++i) {
            const int c = i < m ? i : m - 1;
            rk4_sample(D, x, v + NU * c, AB);
            // X <- A X + B E_c   (only the columns of moves 0..c can be non-zero)
            const int ncol = NU * (c + 1);
            for (int col = 0; col < ncol; ++col) {
                const double a0 = X[0 * nz + col], a1 = X[1 * nz + col], a2 = X[2 * nz + col];
                for (int rr = 0; rr < NX; ++rr) X[rr * nz + col] = AB[rr * 5 + 0] * a0 + AB[rr * 5 + 1] * a1 + AB[rr * 5 + 2] * a2;
            }
            for (int rr = 0; rr < NX; ++rr)
                for (int j = 0; j < NU; ++j) X[rr * nz + NU * c + j] += AB[rr * 5 + NX + j];
            for (int j = 0; j < NY; ++j) {
                const double *S = X + (1 + j) * nz;   // dy_j/dv
                const double e = r[j] - x[1 + j];
                for (int a = 0; a < ncol; ++a) {
                    const double wa = wy2[j] * S[a];
                    g[a] = fma(-wa, e, g[a]);
                    for (int b = 0; b <= a; ++b) H[a * NM_LD + b] = fma(wa, S[b], H[a * NM_LD + b]);
                }
            }
        }
        for (int c = 0; c < m; ++c)
            for (int j = 0; j < NU; ++j) {
                const int a = NU * c + j;
                const double du = v[a] - (c == 0 ? uprev[j] : v[a - NU]);
                g[a] = fma(wu2[j], du, g[a]);
                H[a * NM_LD + a] += wu2[j];
                if (c > 0) { g[a - NU] = fma(-wu2[j], du, g[a - NU]); H[(a - NU) * NM_LD + (a - NU)] += wu2[j]; H[a * NM_LD + (a - NU)] -= wu2[j]; }
            }
        for (int a = 0; a < nz; ++a)
            for (int b = a + 1; b < nz; ++b) H[a * NM_LD + b] = H[b * NM_LD + a];
        for (int i = 0; i < nz; ++i) { lo[i] = D.umin[i % NU] - v[i]; hi[i] = D.umax[i % NU] - v[i]; }
        const int rc = box_qp(nz, H, g, lo, hi, d, Lc, tmp, fixed);
        if (rc) { status = rc; break; }
        double dmax = 0.0;
        for (int i = 0; i < nz; ++i) dmax = fmax(dmax, fabs(d[i]) / D.su[i % NU]);
        if (dmax < 1e-10) break;
        // ---- backtracking on the true cost ----
        double alpha = 1.0, Jn = 0.0;
        int acc_ = 0;
        for (int bt = 0; bt < 6; ++bt) {
            for (int i = 0; i < nz; ++i) vt[i] = fmin(fmax(v[i] + alpha * d[i], D.umin[i % NU]), D.umax[i % NU]);
            Jn = plan_cost(D, x0, uprev, r, p, m, wy2, wu2, vt);
            if (Jn < Jcur) { acc_ = 1; break; }
            alpha *= 0.5;
        }
        if (!acc_) break;   // no descent at this resolution: converged to rounding
        for (int i = 0; i < nz; ++i) v[i] = vt[i];
        Jcur = Jn;
    }
    return status;
}


// ------------------------------------------------------------------------------------------------
// One closed-loop run of closedloop_toolbox_nmpc.m:36-97 with the GAM / VNS sums fused (GAM_fun.m:110-115, VNS2.m:147-195
// nonlinear branch).  mode 0 RAW, 1 GAM, 2 VNS; sel: -1 user set-point, j >= 0 only set-point j kept (VNS2.m:150-155).
// r, yref: ny x nit.  y, u, yopt, uopt: this candidate's ny|nu x nit blocks or nullptr.  cost: GAM 2 slots / VNS 1 slot.
// H, Lc: NM_LD x NM_LD scratch.  Returns the status (0 ok, 5: a state bound was crossed, cost still valid).
// ------------------------------------------------------------------------------------------------
NM_FN int nmpc_run(const NmpcDev &D, int p, int m, int mode, int sel, const double *delta, const double *lambda, const double *r,
                   const double *yref, double *y, double *u, double *yopt, double *uopt, double *cost, double *H, double *Lc,
                   unsigned *n_calls_out, unsigned *n_sqp_out) {
    const int nit = D.nit;
    double wy2[NY], wu2[NU];
    for (int j = 0; j < NY; ++j) { const double w = delta[j] / D.sy[j]; wy2[j] = w * w; }
    for (int j = 0; j < NU; ++j) { const double w = lambda[j] / D.su[j]; wu2[j] = w * w; }
    double v[NM_MAXZ], rr[NY];
    unsigned n_sqp = 0, n_calls = 0;
    int status = 0;
    const bool want_ol = mode != 1 || yopt || uopt;
    double jnu = 0.0, cost_acc[NY] = {0.0, 0.0}, vns_acc = 0.0;
    // ---------------- open-loop optimum from (x0, u0) toward r(:, end)  (:79-94) ----------------
    double xo[NX] = {D.x0[0], D.x0[1], D.x0[2]};
    double vopt[NM_MAXZ];
    if (want_ol) {
        for (int i = 0; i < NU * m; ++i) vopt[i] = D.u0[i % NU];
        for (int j = 0; j < NY; ++j) rr[j] = (sel < 0 || sel == j) ? r[(size_t)j * nit + (nit - 1)] : 0.0;
        const int rc = nlmpcmove(D, D.x0, D.u0, rr, p, m, wy2, wu2, vopt, H, Lc, &n_sqp);
        n_calls++;
        if (rc) status = rc;
        if (mode == 2) {   // Jnu over the whole padded uopt row (VNS2.m:183-191): only the first m-1 differences can be non-zero
            const int j = sel;
            const double u0a = fabs(vopt[j]);
            for (int cc = 0; cc + 1 < m && cc + 1 < nit; ++cc) {
                const double xn = u0a / fabs(vopt[NU * (cc + 1) + j] - vopt[NU * cc + j]);
                if (fabs(xn) <= 1.7976931348623157e308) jnu += xn * xn;
            }
        }
    }
    // ---------------- closed loop (:61-74) in lock-step with the open-loop rollout ----------------
    double x[NX] = {D.x0[0], D.x0[1], D.x0[2]}, uprev[NU] = {D.u0[0], D.u0[1]};
    for (int i = 0; i < NU * m; ++i) v[i] = D.u0[i % NU];
    for (int k = 0; k < nit; ++k) {
        if (k > 0) {
            for (int j = 0; j < NY; ++j) rr[j] = (sel < 0 || sel == j) ? r[(size_t)j * nit + k] : 0.0;
            const int rc = nlmpcmove(D, x, uprev, rr, p, m, wy2, wu2, v, H, Lc, &n_sqp);   // v: warm start = previous plan
            n_calls++;
            if (rc) status = rc;
            uprev[0] = v[0]; uprev[1] = v[1];
            rk4_sample(D, x, uprev, nullptr);
            for (int i = 0; i < NX; ++i)
                if (x[i] < D.xmin[i] - 1e-9 || x[i] > D.xmax[i] + 1e-9) { if (!status) status = 5; }
            if (want_ol) {
                const int cc = k < m ? k : m - 1;   // uopt(:, k): MVopt row k (rows m.. repeat row m-1)
                rk4_sample(D, xo, vopt + NU * cc, nullptr);
            }
        }
        for (int j = 0; j < NY; ++j) {
            const bool mine = sel < 0 || sel == j;
            const double yj = x[1 + j], yoj = xo[1 + j], yr = yref[(size_t)j * nit + k];
            if (mine) {
                if (y) y[(size_t)j * nit + k] = yj;
                if (u) u[(size_t)j * nit + k] = uprev[j];
                if (want_ol && yopt) yopt[(size_t)j * nit + k] = yoj;
                if (want_ol && uopt) uopt[(size_t)j * nit + k] = vopt[NU * (k < m ? k : m - 1) + j];
                if (mode == 1) cost_acc[j] += (yj - yr) * (yj - yr);
                if (mode == 2 && k >= D.inK - 1) vns_acc += (yj - yoj) * (yj - yoj) + (yj - yr) * (yj - yr);
            }
        }
    }
    const bool okc = status == 0 || status == 5;
    if (mode == 1) for (int j = 0; j < NY; ++j) cost[j] = okc ? cost_acc[j] : NAN;
    if (mode == 2) cost[0] = okc ? vns_acc + jnu : NAN;
    *n_calls_out = n_calls; *n_sqp_out = n_sqp;
    return status;
}
