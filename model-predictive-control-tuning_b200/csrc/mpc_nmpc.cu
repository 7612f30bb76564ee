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

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"

#define NX 3
#define NU 2
#define NY 2
#define NM_MAXM 15
#define NM_MAXZ (NU * NM_MAXM)
#define NM_THREADS 64

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
__device__ __forceinline__ void vdv_rhs(const double *x, const double *u, double *f, double *J) {
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
__device__ void rk4_sample(const NmpcDev &D, double *x, const double *u, double *AB) {
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
__device__ double plan_cost(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m,
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
__device__ int box_qp(int nz, const double *H, const double *g, const double *lo, const double *hi, double *d, double *Lc,
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
                double acc = H[i * nz + j];
                for (int k = 0; k < fj; ++k) acc -= Lc[fi * nz + k] * Lc[fj * nz + k];
                if (j == i) {
                    if (!(acc > 0.0)) return 3;
                    Lc[fi * nz + fj] = sqrt(acc);
                } else {
                    Lc[fi * nz + fj] = acc / Lc[fj * nz + fj];
                }
                fj++;
            }
            fi++;
        }
        fi = 0;
        for (int i = 0; i < nz; ++i) {
            if (fixed[i]) continue;
            double acc = -g[i];
            for (int j = 0; j < nz; ++j) if (fixed[j]) acc -= H[i * nz + j] * d[j];
            for (int k = 0; k < fi; ++k) acc -= Lc[fi * nz + k] * tmp[k];
            tmp[fi] = acc / Lc[fi * nz + fi];
            fi++;
        }
        for (int a = nf - 1; a >= 0; --a) {
            double acc = tmp[a];
            for (int k = a + 1; k < nf; ++k) acc -= Lc[k * nz + a] * tmp[k];
            tmp[a] = acc / Lc[a * nz + a];
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
            for (int j = 0; j < nz; ++j) gi = fma(H[i * nz + j], d[j], gi);
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
__device__ int nlmpcmove(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m,
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
        for (int i = 0; i < nz * nz; ++i) H[i] = 0.0;
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
                    for (int b = 0; b <= a; ++b) H[a * nz + b] = fma(wa, S[b], H[a * nz + b]);
                }
            }
        }
        for (int c = 0; c < m; ++c)
            for (int j = 0; j < NU; ++j) {
                const int a = NU * c + j;
                const double du = v[a] - (c == 0 ? uprev[j] : v[a - NU]);
                g[a] = fma(wu2[j], du, g[a]);
                H[a * nz + a] += wu2[j];
                if (c > 0) { g[a - NU] = fma(-wu2[j], du, g[a - NU]); H[(a - NU) * nz + (a - NU)] += wu2[j]; H[a * nz + (a - NU)] -= wu2[j]; }
            }
        for (int a = 0; a < nz; ++a)
            for (int b = a + 1; b < nz; ++b) H[a * nz + b] = H[b * nz + a];
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

// mode 0 RAW, 1 GAM, 2 VNS.  One thread per (candidate, run).
__global__ void __launch_bounds__(NM_THREADS) k_nmpc(const NmpcDev D, int n, int runs, int mode, NmpcArgs A) {
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n * runs) return;
    const int c = item / runs, run = item - c * runs;
    const int p = A.N[c], m = A.Nu[c], nit = D.nit;
    if (p < 2 || p > D.pmax || m < 1 || m > D.mmax || m >= p) {
        A.status[c] = MPCGPU_CAND_INVALID;
        if (mode == 1) for (int j = 0; j < NY; ++j) A.cost[(size_t)c * NY + j] = NAN;
        if (mode == 2) A.part[(size_t)c * runs + run] = NAN;
        return;
    }
    const int sel = mode == 2 ? run : -1;   // VNS: only set-point `sel` is kept, the others are ZEROED (VNS2.m:150-155)
    double wy2[NY], wu2[NU];
    for (int j = 0; j < NY; ++j) { const double w = A.delta[(size_t)c * NY + j] / D.sy[j]; wy2[j] = w * w; }
    for (int j = 0; j < NU; ++j) { const double w = A.lambda[(size_t)c * NU + j] / D.su[j]; wu2[j] = w * w; }
    double *H = A.work + (size_t)item * 2 * NM_MAXZ * NM_MAXZ, *Lc = H + NM_MAXZ * NM_MAXZ;
    double v[NM_MAXZ], rr[NY];
    auto ref_at = [&](int k, double *out) {
        for (int j = 0; j < NY; ++j) out[j] = (sel < 0 || sel == j) ? A.r[(size_t)j * nit + k] : 0.0;
    };
    unsigned n_sqp = 0, n_calls = 0;
    int status = 0;
    const bool want_ol = mode != 1 || A.yopt || A.uopt;
    double jnu = 0.0, cost_acc[NY] = {0.0, 0.0}, vns_acc = 0.0;
    // ---------------- open-loop optimum from (x0, u0) toward r(:, end)  (:79-94) ----------------
    double xo[NX] = {D.x0[0], D.x0[1], D.x0[2]};
    double vopt[NM_MAXZ];
    if (want_ol) {
        for (int i = 0; i < NU * m; ++i) vopt[i] = D.u0[i % NU];
        ref_at(nit - 1, rr);
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
            ref_at(k, rr);
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
            const double yj = x[1 + j], yoj = xo[1 + j], yr = A.yref[(size_t)j * nit + k];
            if (mine) {
                if (A.y) A.y[((size_t)c * NY + j) * nit + k] = yj;
                if (A.u) A.u[((size_t)c * NU + j) * nit + k] = uprev[j];
                if (want_ol && A.yopt) A.yopt[((size_t)c * NY + j) * nit + k] = yoj;
                if (want_ol && A.uopt) A.uopt[((size_t)c * NU + j) * nit + k] = vopt[NU * (k < m ? k : m - 1) + j];
                if (mode == 1) cost_acc[j] += (yj - yr) * (yj - yr);
                if (mode == 2 && k >= D.inK - 1) vns_acc += (yj - yoj) * (yj - yoj) + (yj - yr) * (yj - yr);
            }
        }
    }
    if (mode == 1) for (int j = 0; j < NY; ++j) A.cost[(size_t)c * NY + j] = (status == 0 || status == 5) ? cost_acc[j] : NAN;
    if (mode == 2) A.part[(size_t)c * runs + run] = (status == 0 || status == 5) ? vns_acc + jnu : NAN;
    if (status) atomicMax(A.status + c, status);
    atomicAdd(A.counters + 0, (unsigned long long)n_calls);
    atomicAdd(A.counters + 1, (unsigned long long)n_sqp);
}


// =================================================================================================
// Warp-per-run form of the same algorithm (MPCGPU_NMPC_WARP_PER_RUN=1): lane a owns decision variable a (nz = 2*m <= 32):
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

// predicted cost of the plan in `vs` (shared, nz entries); every lane returns the same value
__device__ double w_plan_cost(const NmpcDev &D, const double *x0, const double *uprev, const double *r, int p, int m,
                              const double *wy2, const double *wu2, const double *vs, int lane) {
    double x[NX] = {x0[0], x0[1], x0[2]};
    double J = 0.0;
    for (int i = 0; i < p; ++i) {
        const int c = i < m ? i : m - 1;
        const double u[NU] = {vs[NU * c], vs[NU * c + 1]};
        rk4_sample(D, x, u, nullptr);
        for (int j = 0; j < NY; ++j) { const double e = r[j] - x[1 + j]; J = fma(wy2[j] * e, e, J); }
    }
    double part = 0.0;
    if (lane < NU * m) {
        const int j = lane % NU;
        const double du = vs[lane] - (lane < NU ? uprev[j] : vs[lane - NU]);
        part = wu2[j] * du * du;
    }
    return J + nmw_sum(part);
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
    double Jcur = w_plan_cost(D, x0, uprev, r, p, m, wy2, wu2, sm.v, lane);
    int status = 0;
    for (int it = 0; it < D.max_sqp; ++it) {
        *n_sqp += 1;
        double Hrow[NM_MAXZ], AB[15];
#pragma unroll
        for (int b = 0; b < NM_MAXZ; ++b) Hrow[b] = 0.0;
        double g = 0.0, X0 = 0.0, X1 = 0.0, X2 = 0.0;
        double x[NX] = {x0[0], x0[1], x0[2]};
        for (int i = 0; i < p; ++i) {
            const int c = i < m ? i : m - 1;
            const double u[NU] = {sm.v[NU * c], sm.v[NU * c + 1]};
            rk4_sample(D, x, u, AB);
            {   // X_a <- A X_a + B e_(c, j)
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
        double alpha = 1.0, Jn = 0.0;
        int acc_ = 0;
        for (int bt = 0; bt < 6; ++bt) {
            sm.vt[lane] = var ? fmin(fmax(va + alpha * d, umn), umx) : 0.0;
            __syncwarp();
            Jn = w_plan_cost(D, x0, uprev, r, p, m, wy2, wu2, sm.vt, lane);
            if (Jn < Jcur) { acc_ = 1; break; }
            alpha *= 0.5;
            __syncwarp();
        }
        if (!acc_) break;
        va = sm.vt[lane];
        __syncwarp();
        sm.v[lane] = va;
        __syncwarp();
        Jcur = Jn;
    }
    return status;
}

// mode 0 RAW, 1 GAM, 2 VNS.  One warp per (candidate, run).
__global__ void __launch_bounds__(32 * NMW_WARPS) k_nmpc_w(const NmpcDev D, int n, int runs, int mode, NmpcArgs A) {
    extern __shared__ double smem_n[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * NMW_WARPS + warp;
    if (item >= n * runs) return;
    const int c = item / runs, run = item - c * runs;
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
    const size_t nI = (size_t)n * 3, nT = (size_t)n * 2 * nit;
    const size_t nD = (size_t)n * (2 * NY + 2 * NU) + (size_t)n * runs + (traj ? 4 * nT : 0) + (size_t)NY * nit +
                      (size_t)n * runs * 2 * NM_MAXZ * NM_MAXZ + 8;
    int *dI = nullptr; double *dD = nullptr;
    if (cudaMalloc((void **)&dI, sizeof(int) * nI) != cudaSuccess || cudaMalloc((void **)&dD, sizeof(double) * nD) != cudaSuccess) {
        cudaFree(dI); h->err = "cudaMalloc failed"; return MPCGPU_ERR_CUDA;
    }
    int *dN = dI, *dNu = dN + n, *dSt = dNu + n;
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
    ck(cudaMemcpyAsync(dDl, delta, sizeof(double) * n * NY, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(dLm, lambda, sizeof(double) * n * NU, cudaMemcpyHostToDevice, s));
    if (r_override) ck(cudaMemcpyAsync(dRo, r_override, sizeof(double) * NY * nit, cudaMemcpyHostToDevice, s));
    ck(cudaMemsetAsync(dSt, 0, sizeof(int) * n, s));
    ck(cudaMemsetAsync(dCnt, 0, sizeof(unsigned long long) * 2, s));
    if (traj) ck(cudaMemsetAsync(dY, 0, sizeof(double) * 4 * nT, s));
    if (rc == MPCGPU_OK) {
        NmpcArgs A{dN, dNu, dDl, dLm, r_override ? dRo : h->dR, h->dYref, dCost, dPart, dY, dU, dYo, dUo, dSt, dCnt, dWork};
        const int items = n * runs;
        if (getenv("MPCGPU_NMPC_WARP_PER_RUN")) {   // alternative mapping, same algorithm (measured: no faster, see header)
            const size_t smem = sizeof(double) * NMW_DOUBLES * NMW_WARPS;
            k_nmpc_w<<<(items + NMW_WARPS - 1) / NMW_WARPS, 32 * NMW_WARPS, smem, s>>>(h->D, n, runs, cost_mode, A);
        } else {
            k_nmpc<<<(items + NM_THREADS - 1) / NM_THREADS, NM_THREADS, 0, s>>>(h->D, n, runs, cost_mode, A);
        }
        ck(cudaGetLastError());
        if (cost_mode == MPCGPU_COST_VNS) {
            k_nmpc_finish<<<(n + 127) / 128, 128, 0, s>>>(n, runs, dN, dPart, dSt, dCost);
            ck(cudaGetLastError());
        }
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
    }
    cudaFree(dI); cudaFree(dD);
    return rc;
}

extern "C" int mpcgpu_nmpc_get_counters(mpcgpu_nmpc_handle *h, mpcgpu_counters *out) {
    if (!h || !out) return MPCGPU_ERR_ARG;
    *out = h->cnt;
    return MPCGPU_OK;
}
