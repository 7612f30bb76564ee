/*
 * mpc_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY: may be imported / linked / executed only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
 * (libmpcgpu.so) never calls into this file and has no CPU fallback.
 *
 * What it restates (fp64, scalar C, one candidate at a time):
 *   closedloop_toolbox      /root/reference/MPC-Tuning/MPC_Tuning/closedloop_toolbox.m:29-107
 *   GAM objective           /root/reference/MPC-Tuning/MPC_Tuning/GAM_fun.m:54-115
 *   VNS objective           /root/reference/MPC-Tuning/MPC_Tuning/VNS2.m:58-61,147-195
 *
 * PARITY UNPINNED.  The arithmetic of `sim`, `mpcmove` and the KWIK QP solver lives in the closed-source
 * MathWorks Model Predictive Control Toolbox (object version '7.1', MATLAB R2021a per the reference's
 * .mat files), which is neither under /root/reference nor installable here (no MATLAB/Octave).  The
 * reference holds no per-candidate cost or trajectory fixture.  The modelling conventions below are
 * therefore restated from the published Toolbox documentation ("Optimization Problem", "QP Matrices",
 * "QP Solvers", "Controller State Estimation") and isolated in this one file:
 *
 *   T1  cost  J = sum_{t=1..p} sum_i (wy_i/sy_i * (r_i - y_i(k+t)))^2
 *                + sum_{c=0..m-1} sum_j (wdu_j/su_j * du_j(k+c))^2 + rho_eps * eps^2
 *       weights enter SQUARED, scale factors divide, Weights.MV = 0, MV target inactive.
 *   T2  scalar control horizon m: moves k..k+m-1 free, zero afterwards (closedloop_toolbox.m:38-40
 *       collapses the per-input Nu vector to max(Nu)).
 *   T3  no reference / MD look-ahead (sim defaults): r(k+t) = r(k), v(k+t) = v(k).
 *   T4  MV bounds and MV-rate bounds hard (ECR 0); OV bounds soft:
 *         ymin_i - eps*Vmin_i*sy_i <= y_i(k+t) <= ymax_i + eps*Vmax_i*sy_i ,  t = 1..p,  eps >= 0.
 *   T5  nominal noise-free `sim` with plant == prediction model from the nominal (zero) state: the
 *       default output-disturbance estimator has zero innovation for all k, so the controller state
 *       is the true plant state and the estimator is a no-op.
 *   T6  Info.Uopt is (p+1) x nu, row c = u(k+c), rows m..p repeat row m-1 (last row duplicates row p).
 *   T7  the QP is strictly convex (wdu >= 1e-5 > 0, MPCTuning.m:302), so its optimum is unique and any
 *       exact solver returns the Toolbox's answer up to the Toolbox's own ConstraintTolerance (1e-6).
 *       The solver here is a Goldfarb-Idnani dual active-set method (the family KWIK belongs to),
 *       written from the 1983 paper; its optimum is cross-checked against scipy in tests/.
 *
 * What *is* pinned by reference fixtures: the scaled discrete plants and limits fed to this file
 * (tests/test_plant_kats.py against tests/golden/fixture_kats.json).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int ny, nu, nd, nit;
    const double *a, *b0, *b1; /* ny x nw row-major, nw = nu + nd (MVs then MDs) */
    const int *d;              /* y(k) = a y(k-1) + b0 w(k-d) + b1 w(k-d-1)      */
    const double *umin, *umax, *dumin, *dumax;   /* nu, +-inf allowed */
    const double *ymin, *ymax, *ecr_min, *ecr_max; /* ny */
    const double *su, *sy;
    double rho_ecr;
    const double *r; /* nit x ny, time-major */
    const double *v; /* nit x nd */
} orc_problem;

#define VIOL_TOL 1e-10
/* linear dependence of a new normal on the active ones: |d2|^2 <= DEP_TOL |d|^2.  |d2|^2 is a plain sum of squares (no
 * cancellation), accurate to ~n eps^2 |d|^2 = 2e-30 |d|^2, so the test can sit far below eps: with soft output limits and
 * zero tracking weights (Shell7x5) the slack direction is all that separates two band rows and its share of |d|^2 is
 * (V/sqrt(rho_eps))^2 / (g/lambda)^2 ~ 1e-15 at lambda = 1e-4.  A test at 1e-15 declared such rows dependent and the QP
 * infeasible, which a QP with a slack can never be (70 of 256 Shell7x5 candidates; 0 with this value). */
#define DEP_TOL 1e-24

/* ------------------------------------------------------------------------------------------ */
/* Goldfarb-Idnani dual active-set QP:  min 1/2 x'Hx + f'x  s.t.  n_i'x >= b_i                */
/* H enters through J0 = L^-T (H = L L').  Constraints are supplied by callbacks so that the  */
/* caller can keep them structured.                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int n;           /* variables */
    int mc;          /* constraints */
    void *ctx;
    void (*slacks)(void *ctx, const double *x, double *s); /* s_i = n_i'x - b_i, +inf if absent */
    void (*normal)(void *ctx, int i, double *nv);          /* dense normal of constraint i      */
} gi_constraints;

typedef struct {
    double *J, *R, *d, *z, *rv, *nv, *s, *u;
    int *A;
    int *pref, npref;   /* final active set of the previous solve with this work area (pivot rule 2) */
} gi_work;

static gi_work *gi_alloc(int n, int mc) {
    gi_work *w = (gi_work *)calloc(1, sizeof(gi_work));
    w->J = (double *)malloc(sizeof(double) * n * n);
    w->R = (double *)malloc(sizeof(double) * n * n);
    w->d = (double *)malloc(sizeof(double) * n);
    w->z = (double *)malloc(sizeof(double) * n);
    w->rv = (double *)malloc(sizeof(double) * n);
    w->nv = (double *)malloc(sizeof(double) * n);
    w->s = (double *)malloc(sizeof(double) * (mc + 1));
    w->u = (double *)malloc(sizeof(double) * (n + 1));
    w->A = (int *)malloc(sizeof(int) * (n + 1));
    w->pref = (int *)malloc(sizeof(int) * (n + 1));
    w->npref = 0;
    return w;
}
static void gi_free(gi_work *w) {
    free(w->J); free(w->R); free(w->d); free(w->z); free(w->rv); free(w->nv); free(w->s); free(w->u); free(w->A); free(w->pref);
    free(w);
}

static void gi_drop(gi_work *w, int n, int *q, int l) {
    double *J = w->J, *R = w->R;
    int qq = *q;
    for (int i = l; i < qq - 1; ++i) {
        w->A[i] = w->A[i + 1];
        w->u[i] = w->u[i + 1];
        for (int k = 0; k <= i + 1; ++k) R[k * n + i] = R[k * n + i + 1];
    }
    w->u[qq - 1] = w->u[qq]; /* slot of the incoming constraint moves down with the list */
    qq -= 1;
    for (int j = l; j < qq; ++j) {
        double cc = R[j * n + j], ss = R[(j + 1) * n + j];
        double h = hypot(cc, ss);
        if (h == 0.0) continue;
        cc /= h; ss /= h;
        R[j * n + j] = h; R[(j + 1) * n + j] = 0.0;
        for (int k = j + 1; k < qq; ++k) {
            double t1 = R[j * n + k], t2 = R[(j + 1) * n + k];
            R[j * n + k] = cc * t1 + ss * t2;
            R[(j + 1) * n + k] = -ss * t1 + cc * t2;
        }
        for (int k = 0; k < n; ++k) {
            double t1 = J[k * n + j], t2 = J[k * n + j + 1];
            J[k * n + j] = cc * t1 + ss * t2;
            J[k * n + j + 1] = -ss * t1 + cc * t2;
        }
    }
    *q = qq;
}

/* x: in = unconstrained optimum -H^-1 f, out = constrained optimum.  J0 = L^-T (n x n, row-major).
 * Returns 0 ok, 1 infeasible, 2 iteration cap.  iters counts constraint additions + drops. */
/* Pivot rule of the outer loop: 0 = most violated constraint (the default), 1 = first violated constraint in index order,
 * 2 = the violated constraints of the previous sample's final active set first, then the most violated (what the
 * soft-constraint kernel does, SOFT_PREFER_LAST in csrc/mpc_soft.cuh).
 * All are valid Goldfarb-Idnani pivots and reach the same optimum in exact arithmetic; the spread between the two runs of
 * the SAME oracle is the yard-stick tests/ and bench.py use for what fp64 can resolve on a candidate (oracle/parity.py). */
static int g_pivot_rule = 0;
void orc_set_pivot_rule(int rule) { g_pivot_rule = rule; }

static int gi_solve(const gi_constraints *cs, const double *J0, double *x, gi_work *w, int *iters,
                    int *nact_out) {
    const int n = cs->n, mc = cs->mc;
    double *J = w->J, *R = w->R, *d = w->d, *z = w->z, *rv = w->rv, *nv = w->nv, *s = w->s, *u = w->u;
    int *A = w->A;
    int q = 0, it = 0;
    const int itmax = 400 * (n + 10);   /* degenerate band QPs need hundreds of pivots; 20 (n + 10) cut 14 % of the Shell7x5 candidates short */
    int J_init = 0;
    for (;;) {
        cs->slacks(cs->ctx, x, s);
        int p = -1;
        double smin = -VIOL_TOL;
        if (g_pivot_rule == 2) {
            for (int e = 0; e < w->npref && p < 0; ++e) {
                const int i = w->pref[e];
                int act = 0;
                for (int k = 0; k < q; ++k) if (A[k] == i) { act = 1; break; }
                if (!act && s[i] < -VIOL_TOL) { smin = s[i]; p = i; }
            }
        }
        if (p < 0)
        for (int i = 0; i < mc; ++i) {
            int act = 0;
            for (int k = 0; k < q; ++k) if (A[k] == i) { act = 1; break; }
            if (!act && s[i] < smin) { smin = s[i]; p = i; if (g_pivot_rule == 1) break; }
        }
        if (p < 0) break;
        if (!J_init) { memcpy(J, J0, sizeof(double) * n * n); J_init = 1; }
        cs->normal(cs->ctx, p, nv);
        u[q] = 0.0;
        double sp = s[p];
        for (;;) { /* step 2 of the paper: (partial) steps until constraint p becomes active */
            if (++it > itmax) { *iters = it; return 2; }
            double dn2 = 0.0;
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int k = 0; k < n; ++k) acc += J[k * n + j] * nv[k];
                d[j] = acc; dn2 += acc * acc;
            }
            double zn = 0.0; /* z'n+ = |d2|^2 */
            for (int j = q; j < n; ++j) zn += d[j] * d[j];
            for (int k = 0; k < n; ++k) {
                double acc = 0.0;
                for (int j = q; j < n; ++j) acc += J[k * n + j] * d[j];
                z[k] = acc;
            }
            for (int i = q - 1; i >= 0; --i) { /* r = R^-1 d1 */
                double acc = d[i];
                for (int k = i + 1; k < q; ++k) acc -= R[i * n + k] * rv[k];
                rv[i] = acc / R[i * n + i];
            }
            int dependent = !(zn > DEP_TOL * dn2);
            double t1 = INFINITY, t2 = INFINITY;
            int l = -1;
            for (int k = 0; k < q; ++k)
                if (rv[k] > 0.0) { double t = u[k] / rv[k]; if (t < t1) { t1 = t; l = k; } }
            if (!dependent) t2 = -sp / zn;
            double t = t1 < t2 ? t1 : t2;
            if (!(t < INFINITY)) { *iters = it; return 1; }
            if (dependent || t1 < t2) {
                /* dual (or partial primal+dual) step, then drop constraint l */
                if (!dependent) { for (int k = 0; k < n; ++k) x[k] += t * z[k]; sp += t * zn; }
                for (int k = 0; k < q; ++k) u[k] -= t * rv[k];
                u[q] += t;
                gi_drop(w, n, &q, l);
                continue;
            }
            /* full step: constraint p becomes active */
            for (int k = 0; k < n; ++k) x[k] += t * z[k];
            for (int k = 0; k < q; ++k) u[k] -= t * rv[k];
            u[q] += t;
            /* add: rotate d2 into its first component, same rotations on columns of J */
            for (int j = n - 1; j > q; --j) {
                double cc = d[j - 1], ss = d[j];
                double h = hypot(cc, ss);
                if (h == 0.0) continue;
                cc /= h; ss /= h;
                d[j - 1] = h; d[j] = 0.0;
                for (int k = 0; k < n; ++k) {
                    double a1 = J[k * n + j - 1], a2 = J[k * n + j];
                    J[k * n + j - 1] = cc * a1 + ss * a2;
                    J[k * n + j] = -ss * a1 + cc * a2;
                }
            }
            for (int k = 0; k <= q; ++k) R[k * n + q] = d[k];
            A[q] = p;
            q += 1;
            break;
        }
    }
    *iters = it;
    if (nact_out) *nact_out = q;
    for (int k = 0; k < q; ++k) w->pref[k] = A[k];
    w->npref = q;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Candidate context: prediction matrices for (p, m, delta, lambda)                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const orc_problem *pb;
    int p, m, nz, nzu, has_eps;
    int nw;
    double *S;     /* step responses s_ij(n), n = 0..p : [(i*nu+j)*(p+1) + n]                 */
    double *G;     /* (ny*p) x nzu, row (t-1)*ny+i, col c*nu+j                               */
    double *wy2;   /* ny: (delta_i/sy_i)^2                                                    */
    double *wu2;   /* nu: (lambda_j/su_j)^2                                                   */
    double *H, *Lc, *J0; /* nz x nz */
    double *yfree; /* ny*p */
    const double *uprev; /* nu, set per step */
    const double *yoff;  /* ny or NULL: estimated output disturbance, constant over the horizon (orc_closedloop_est) */
    double *f, *x;
} cand_ctx;

static void ctx_slacks(void *vctx, const double *x, double *s) {
    cand_ctx *c = (cand_ctx *)vctx;
    const orc_problem *pb = c->pb;
    const int nu = pb->nu, ny = pb->ny, m = c->m, p = c->p, nzu = c->nzu;
    const double eps = c->has_eps ? x[nzu] : 0.0;
    for (int j = 0; j < nu; ++j) {
        double lvl = c->uprev[j];
        for (int cc = 0; cc < m; ++cc) {
            int e = cc * nu + j;
            lvl += x[e];
            s[e] = isfinite(pb->dumin[j]) ? x[e] - pb->dumin[j] : INFINITY;
            s[nzu + e] = isfinite(pb->dumax[j]) ? pb->dumax[j] - x[e] : INFINITY;
            s[2 * nzu + e] = isfinite(pb->umin[j]) ? lvl - pb->umin[j] : INFINITY;
            s[3 * nzu + e] = isfinite(pb->umax[j]) ? pb->umax[j] - lvl : INFINITY;
        }
    }
    if (c->has_eps) {
        int base = 4 * nzu;
        for (int t = 0; t < p; ++t)
            for (int i = 0; i < ny; ++i) {
                int row = t * ny + i;
                double yp = c->yfree[row];
                const double *g = c->G + (size_t)row * nzu;
                for (int e = 0; e < nzu; ++e) yp += g[e] * x[e];
                s[base + row] = isfinite(pb->ymax[i]) ? pb->ymax[i] + eps * pb->ecr_max[i] * pb->sy[i] - yp : INFINITY;
                s[base + ny * p + row] = isfinite(pb->ymin[i]) ? yp - (pb->ymin[i] - eps * pb->ecr_min[i] * pb->sy[i]) : INFINITY;
            }
        s[base + 2 * ny * p] = eps;
    }
}

static void ctx_normal(void *vctx, int idx, double *nv) {
    cand_ctx *c = (cand_ctx *)vctx;
    const orc_problem *pb = c->pb;
    const int nu = pb->nu, ny = pb->ny, p = c->p, nzu = c->nzu, nz = c->nz;
    for (int e = 0; e < nz; ++e) nv[e] = 0.0;
    if (idx < nzu) { nv[idx] = 1.0; return; }
    if (idx < 2 * nzu) { nv[idx - nzu] = -1.0; return; }
    if (idx < 4 * nzu) {
        int hi = idx >= 3 * nzu;
        int e = idx - (hi ? 3 : 2) * nzu;
        int cc = e / nu, j = e % nu;
        for (int c2 = 0; c2 <= cc; ++c2) nv[c2 * nu + j] = hi ? -1.0 : 1.0;
        return;
    }
    int k = idx - 4 * nzu;
    if (k < ny * p) { /* ymax row */
        int i = k % ny;
        const double *g = c->G + (size_t)k * nzu;
        for (int e = 0; e < nzu; ++e) nv[e] = -g[e];
        nv[nzu] = pb->ecr_max[i] * pb->sy[i];
    } else if (k < 2 * ny * p) {
        k -= ny * p;
        int i = k % ny;
        const double *g = c->G + (size_t)k * nzu;
        for (int e = 0; e < nzu; ++e) nv[e] = g[e];
        nv[nzu] = pb->ecr_min[i] * pb->sy[i];
    } else {
        nv[nzu] = 1.0;
    }
}

static int ctx_build(cand_ctx *c, const orc_problem *pb, int p, int m, const double *delta, const double *lambda) {
    const int ny = pb->ny, nu = pb->nu, nw = pb->nu + pb->nd;
    c->pb = pb; c->p = p; c->m = m; c->nw = nw;
    c->nzu = nu * m;
    c->has_eps = 0;
    for (int i = 0; i < ny; ++i) if (isfinite(pb->ymin[i]) || isfinite(pb->ymax[i])) c->has_eps = 1;
    c->nz = c->nzu + c->has_eps;
    const int nz = c->nz, nzu = c->nzu;
    c->S = (double *)calloc((size_t)ny * nu * (p + 1), sizeof(double));
    c->G = (double *)calloc((size_t)ny * p * nzu, sizeof(double));
    c->wy2 = (double *)malloc(sizeof(double) * ny);
    c->wu2 = (double *)malloc(sizeof(double) * nu);
    c->H = (double *)calloc((size_t)nz * nz, sizeof(double));
    c->Lc = (double *)calloc((size_t)nz * nz, sizeof(double));
    c->J0 = (double *)calloc((size_t)nz * nz, sizeof(double));
    c->yfree = (double *)calloc((size_t)ny * p, sizeof(double));
    c->f = (double *)calloc(nz, sizeof(double));
    c->x = (double *)calloc(nz, sizeof(double));
    for (int i = 0; i < ny; ++i) { double w = delta[i] / pb->sy[i]; c->wy2[i] = w * w; }   /* T1 */
    for (int j = 0; j < nu; ++j) { double w = lambda[j] / pb->su[j]; c->wu2[j] = w * w; }
    /* step response of every MV channel: unit step applied at n = 0 */
    for (int i = 0; i < ny; ++i)
        for (int j = 0; j < nu; ++j) {
            const int ch = i * nw + j, dd = pb->d[ch];
            double *s = c->S + (size_t)(i * nu + j) * (p + 1);
            s[0] = 0.0;
            for (int n = 1; n <= p; ++n)
                s[n] = pb->a[ch] * s[n - 1] + (n - dd >= 0 ? pb->b0[ch] : 0.0) + (n - dd - 1 >= 0 ? pb->b1[ch] : 0.0);
        }
    /* dynamic matrix: a move at horizon index c reaches y(k+t) with s(t-c) */
    for (int t = 1; t <= p; ++t)
        for (int i = 0; i < ny; ++i)
            for (int cc = 0; cc < m; ++cc)
                for (int j = 0; j < nu; ++j)
                    if (t - cc >= 1)
                        c->G[(size_t)((t - 1) * ny + i) * nzu + cc * nu + j] = c->S[(size_t)(i * nu + j) * (p + 1) + (t - cc)];
    /* H = G' Wy^2 G + Wdu^2  (+ rho for the slack) */
    for (int e1 = 0; e1 < nzu; ++e1)
        for (int e2 = 0; e2 <= e1; ++e2) {
            double acc = 0.0;
            for (int t = 0; t < p; ++t)
                for (int i = 0; i < ny; ++i) {
                    const double *g = c->G + (size_t)(t * ny + i) * nzu;
                    acc += c->wy2[i] * g[e1] * g[e2];
                }
            if (e1 == e2) acc += c->wu2[e1 % nu];
            c->H[e1 * nz + e2] = acc; c->H[e2 * nz + e1] = acc;
        }
    if (c->has_eps) c->H[nzu * nz + nzu] = pb->rho_ecr;
    /* Cholesky (lower) */
    for (int jj = 0; jj < nz; ++jj) {
        double dsum = c->H[jj * nz + jj];
        for (int k = 0; k < jj; ++k) dsum -= c->Lc[jj * nz + k] * c->Lc[jj * nz + k];
        if (!(dsum > 0.0)) return 3;
        double djj = sqrt(dsum);
        c->Lc[jj * nz + jj] = djj;
        for (int i = jj + 1; i < nz; ++i) {
            double acc = c->H[i * nz + jj];
            for (int k = 0; k < jj; ++k) acc -= c->Lc[i * nz + k] * c->Lc[jj * nz + k];
            c->Lc[i * nz + jj] = acc / djj;
        }
    }
    /* J0 = L^-T : column j of L^-1 by forward substitution, stored transposed */
    for (int col = 0; col < nz; ++col) {
        double *tmp = c->f; /* scratch */
        for (int i = 0; i < nz; ++i) {
            double acc = (i == col) ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) acc -= c->Lc[i * nz + k] * tmp[k];
            tmp[i] = acc / c->Lc[i * nz + i];
        }
        for (int i = 0; i < nz; ++i) c->J0[col * nz + i] = tmp[i]; /* J0[col][i] = Linv[i][col] */
    }
    return 0;
}
static void ctx_free(cand_ctx *c) {
    free(c->S); free(c->G); free(c->wy2); free(c->wu2); free(c->H); free(c->Lc); free(c->J0);
    free(c->yfree); free(c->f); free(c->x);
}

/* One controller move (mpcmove): given channel states xs (ny x nw) at time k, the input history
 * whist (value of input j at time k-1-q is whist[j*hl + q], q = 0..hl-1), the held values hv
 * (u(k-1) for MVs, v(k) for MDs) and the constant reference rk, solve the QP; result in c->x. */
static int ctx_move(cand_ctx *c, gi_work *gw, const double *xs, const double *whist, int hl, const double *hv,
                    const double *rk, const double *uprev, int *iters, int *nact) {
    const orc_problem *pb = c->pb;
    const int ny = pb->ny, nu = pb->nu, nw = c->nw, p = c->p, nz = c->nz, nzu = c->nzu;
    /* free response (T3): future inputs held at hv */
    for (int i = 0; i < ny * p; ++i) c->yfree[i] = 0.0;
    for (int i = 0; i < ny; ++i)
        for (int j = 0; j < nw; ++j) {
            const int ch = i * nw + j, dd = pb->d[ch];
            double xf = xs[ch];
            for (int t = 1; t <= p; ++t) {
                /* input seen at relative time t-dd and t-dd-1 (relative to k); >=0 -> held value */
                int r0 = t - dd, r1 = t - dd - 1;
                double w0 = r0 >= 0 ? hv[j] : whist[j * hl + (-r0 - 1)];
                double w1 = r1 >= 0 ? hv[j] : whist[j * hl + (-r1 - 1)];
                xf = pb->a[ch] * xf + pb->b0[ch] * w0 + pb->b1[ch] * w1;
                c->yfree[(t - 1) * ny + i] += xf;
            }
        }
    if (c->yoff)
        for (int t = 0; t < p; ++t)
            for (int i = 0; i < ny; ++i) c->yfree[t * ny + i] += c->yoff[i];
    /* f = -G' Wy^2 (r - yfree) */
    for (int e = 0; e < nz; ++e) c->f[e] = 0.0;
    for (int t = 0; t < p; ++t)
        for (int i = 0; i < ny; ++i) {
            double err = c->wy2[i] * (rk[i] - c->yfree[t * ny + i]);
            if (err == 0.0) continue;
            const double *g = c->G + (size_t)(t * ny + i) * nzu;
            for (int e = 0; e < nzu; ++e) c->f[e] -= g[e] * err;
        }
    /* x = -H^-1 f */
    double *x = c->x;
    for (int i = 0; i < nz; ++i) {
        double acc = -c->f[i];
        for (int k = 0; k < i; ++k) acc -= c->Lc[i * nz + k] * x[k];
        x[i] = acc / c->Lc[i * nz + i];
    }
    for (int i = nz - 1; i >= 0; --i) {
        double acc = x[i];
        for (int k = i + 1; k < nz; ++k) acc -= c->Lc[k * nz + i] * x[k];
        x[i] = acc / c->Lc[i * nz + i];
    }
    c->uprev = uprev;
    gi_constraints cs;
    cs.n = nz; cs.mc = 4 * nzu + (c->has_eps ? 2 * ny * p + 1 : 0);
    cs.ctx = c; cs.slacks = ctx_slacks; cs.normal = ctx_normal;
    return gi_solve(&cs, c->J0, x, gw, iters, nact);
}

/* ------------------------------------------------------------------------------------------ */
/* closedloop_toolbox.m:29-107                                                                */
/* Outputs are signals x time (closedloop_toolbox.m:103-107): y[i*nit+k], u[j*nit+k], ...      */
/* stats[0] = QP solves, stats[1] = active-set iterations, stats[2] = QPs with active set > 0  */
/* ------------------------------------------------------------------------------------------ */
int orc_closedloop(const orc_problem *pb, int p, int m, const double *delta, const double *lambda,
                   double *y, double *u, double *ys, double *uopt, long long *stats) {
    const int ny = pb->ny, nu = pb->nu, nd = pb->nd, nw = nu + nd, nit = pb->nit;
    cand_ctx c;
    memset(&c, 0, sizeof(c));
    int rc = ctx_build(&c, pb, p, m, delta, lambda);
    if (rc) { ctx_free(&c); return rc; }
    gi_work *gw = gi_alloc(c.nz, 4 * c.nzu + 2 * ny * p + 1);
    int dmax = 0;
    for (int i = 0; i < ny * nw; ++i) if (pb->d[i] > dmax) dmax = pb->d[i];
    const int hl = dmax + 2;
    double *xs = (double *)calloc(ny * nw, sizeof(double));
    double *wh = (double *)calloc((size_t)nw * hl, sizeof(double));
    double *hv = (double *)calloc(nw, sizeof(double));
    double *up = (double *)calloc(nu, sizeof(double));
    double *wk = (double *)calloc(nw, sizeof(double));
    long long nqp = 0, nit_as = 0, nqp_act = 0;
    int status = 0;
    /* ---- closed loop: [y,t,u] = sim(mpc,nit,r,v)  (closedloop_toolbox.m:50), T5 ---- */
    for (int k = 0; k < nit; ++k) {
        for (int i = 0; i < ny; ++i) {
            double acc = 0.0;
            for (int j = 0; j < nw; ++j) acc += xs[i * nw + j];
            y[i * nit + k] = acc;
        }
        for (int j = 0; j < nu; ++j) hv[j] = up[j];
        for (int j = 0; j < nd; ++j) hv[nu + j] = pb->v[k * nd + j];
        int iters = 0, nact = 0;
        rc = ctx_move(&c, gw, xs, wh, hl, hv, pb->r + (size_t)k * ny, up, &iters, &nact);
        if (rc) status = rc;
        nqp++; nit_as += iters; if (nact > 0) nqp_act++;
        for (int j = 0; j < nu; ++j) { up[j] += c.x[j]; u[j * nit + k] = up[j]; wk[j] = up[j]; }
        for (int j = 0; j < nd; ++j) wk[nu + j] = pb->v[k * nd + j];
        /* plant step: x(k+1) = a x(k) + b0 w(k+1-d) + b1 w(k-d), with w(k) = wk now known */
        for (int j = 0; j < nw; ++j) { /* push w(k) into the history: slot q=0 is time k afterwards */
            for (int qh = hl - 1; qh > 0; --qh) wh[j * hl + qh] = wh[j * hl + qh - 1];
            wh[j * hl] = wk[j];
        }
        for (int i = 0; i < ny; ++i)
            for (int j = 0; j < nw; ++j) {
                const int ch = i * nw + j, dd = pb->d[ch];
                /* history is now relative to k+1: slot q holds w(k-q) = w((k+1)-1-q) */
                double w0 = dd >= 1 ? wh[j * hl + dd - 1] : 0.0; /* w(k+1-d); d=0 has b0=0 */
                double w1 = wh[j * hl + dd];                      /* w(k-d) */
                xs[ch] = pb->a[ch] * xs[ch] + pb->b0[ch] * w0 + pb->b1[ch] * w1;
            }
    }
    /* ---- open-loop optimum from the fresh state toward the last set-point row (:85-98), T6 ---- */
    if (ys && uopt) {
        memset(xs, 0, sizeof(double) * ny * nw);
        memset(wh, 0, sizeof(double) * nw * hl);
        memset(up, 0, sizeof(double) * nu);
        for (int j = 0; j < nu; ++j) hv[j] = 0.0;
        for (int j = 0; j < nd; ++j) hv[nu + j] = pb->v[(size_t)(nit - 1) * nd + j];
        int iters = 0, nact = 0;
        rc = ctx_move(&c, gw, xs, wh, hl, hv, pb->r + (size_t)(nit - 1) * ny, up, &iters, &nact);
        if (rc) status = rc;
        nqp++; nit_as += iters; if (nact > 0) nqp_act++;
        for (int j = 0; j < nu; ++j) {
            double lvl = 0.0;
            for (int k = 0; k < nit; ++k) {
                if (k < m) lvl += c.x[k * nu + j];   /* rows m..p and the padding repeat row m-1 */
                uopt[j * nit + k] = lvl;
            }
        }
        /* ys = lsim(Pz,[uopt v],t)  (:100) */
        for (int k = 0; k < nit; ++k) {
            for (int i = 0; i < ny; ++i) {
                double acc = 0.0;
                for (int j = 0; j < nw; ++j) acc += xs[i * nw + j];
                ys[i * nit + k] = acc;
            }
            for (int j = 0; j < nu; ++j) wk[j] = uopt[j * nit + k];
            for (int j = 0; j < nd; ++j) wk[nu + j] = pb->v[k * nd + j];
            for (int j = 0; j < nw; ++j) {
                for (int qh = hl - 1; qh > 0; --qh) wh[j * hl + qh] = wh[j * hl + qh - 1];
                wh[j * hl] = wk[j];
            }
            for (int i = 0; i < ny; ++i)
                for (int j = 0; j < nw; ++j) {
                    const int ch = i * nw + j, dd = pb->d[ch];
                    double w0 = dd >= 1 ? wh[j * hl + dd - 1] : 0.0;
                    double w1 = wh[j * hl + dd];
                    xs[ch] = pb->a[ch] * xs[ch] + pb->b0[ch] * w0 + pb->b1[ch] * w1;
                }
        }
    }
    if (stats) { stats[0] += nqp; stats[1] += nit_as; stats[2] += nqp_act; }
    free(xs); free(wh); free(hv); free(up); free(wk);
    gi_free(gw);
    ctx_free(&c);
    return status;
}

/* ------------------------------------------------------------------------------------------ */
/* Plant-model mismatch validation run (Shell3x3.m:271-286, WoodBerry.m:263-278, Shell7x5.m:293-306):   */
/*   options = mpcsimopt(mpc); options.Model = plant; sim(mpc, nit, r, [], options)                     */
/* The controller no longer sees the plant's state: it runs the Toolbox's state estimator.  Restated    */
/* (PARITY UNPINNED, "Controller State Estimation" in the Toolbox documentation):                       */
/*   E1  controller state x_c = [model channel states; MV delay-line states; output-disturbance states] */
/*   E2  default output-disturbance model: one discrete integrator per measured output, driven by       */
/*       unit-variance white noise; default measurement-noise model: unit-variance white noise          */
/*   E3  unit-variance white noise added to every MV (the Toolbox's robustness term); MDs are known     */
/*   E4  steady-state Kalman filter of that model; per sample                                           */
/*         x_c(k|k) = x_c(k|k-1) + M (y(k) - C x_c(k|k-1)),  mpcmove from x_c(k|k),                      */
/*         x_c(k+1|k) = A x_c(k|k) + B u(k)                                                             */
/* The gain M is an INPUT here, (nch + nu*hl + ny) x ny row-major in the state order of E1 (delay-line  */
/* state q of input j = w_j(k-1-q)); tests compute it from E2-E3 with scipy's DARE (tests/, mpcgpu.estimator). */
/* plant: the real process, same channel structure as the model (pa, pb0, pb1, pd).                     */
/* ------------------------------------------------------------------------------------------ */
int orc_closedloop_est(const orc_problem *pb, const double *pa, const double *pb0, const double *pb1, const int *pd,
                       const double *M, int p, int m, const double *delta, const double *lambda, double *y, double *u,
                       long long *stats) {
    const int ny = pb->ny, nu = pb->nu, nd = pb->nd, nw = nu + nd, nit = pb->nit, nch = ny * nw;
    cand_ctx c;
    memset(&c, 0, sizeof(c));
    int rc = ctx_build(&c, pb, p, m, delta, lambda);
    if (rc) { ctx_free(&c); return rc; }
    gi_work *gw = gi_alloc(c.nz, 4 * c.nzu + 2 * ny * p + 1);
    int dmax = 0;
    for (int i = 0; i < nch; ++i) { if (pb->d[i] > dmax) dmax = pb->d[i]; if (pd[i] > dmax) dmax = pd[i]; }
    const int hl = dmax + 2;
    double *xs = (double *)calloc(nch, sizeof(double));              /* controller: model channel states */
    double *wh = (double *)calloc((size_t)nw * hl, sizeof(double));  /* controller: input histories (MV rows estimated) */
    double *xod = (double *)calloc(ny, sizeof(double));              /* controller: output-disturbance states */
    double *xp = (double *)calloc(nch, sizeof(double));              /* plant channel states */
    double *whp = (double *)calloc((size_t)nw * hl, sizeof(double)); /* plant: true input histories */
    double *hv = (double *)calloc(nw, sizeof(double));
    double *up = (double *)calloc(nu, sizeof(double));
    double *wk = (double *)calloc(nw, sizeof(double));
    double *e = (double *)calloc(ny, sizeof(double));
    long long nqp = 0, nit_as = 0, nqp_act = 0;
    int status = 0;
    c.yoff = xod;
    for (int k = 0; k < nit; ++k) {
        /* measurement and correction (E4) */
        for (int i = 0; i < ny; ++i) {
            double yp = 0.0, yh = xod[i];
            for (int j = 0; j < nw; ++j) { yp += xp[i * nw + j]; yh += xs[i * nw + j]; }
            y[i * nit + k] = yp;
            e[i] = yp - yh;
        }
        for (int s = 0; s < nch; ++s)
            for (int i = 0; i < ny; ++i) xs[s] += M[(size_t)s * ny + i] * e[i];
        for (int j = 0; j < nu; ++j)
            for (int q = 0; q < hl; ++q)
                for (int i = 0; i < ny; ++i) wh[j * hl + q] += M[(size_t)(nch + j * hl + q) * ny + i] * e[i];
        for (int o = 0; o < ny; ++o)
            for (int i = 0; i < ny; ++i) xod[o] += M[(size_t)(nch + nu * hl + o) * ny + i] * e[i];
        /* controller move from the corrected state */
        for (int j = 0; j < nu; ++j) hv[j] = up[j];
        for (int j = 0; j < nd; ++j) hv[nu + j] = pb->v[k * nd + j];
        int iters = 0, nact = 0;
        rc = ctx_move(&c, gw, xs, wh, hl, hv, pb->r + (size_t)k * ny, up, &iters, &nact);
        if (rc) status = rc;
        nqp++; nit_as += iters; if (nact > 0) nqp_act++;
        for (int j = 0; j < nu; ++j) { up[j] += c.x[j]; u[j * nit + k] = up[j]; wk[j] = up[j]; }
        for (int j = 0; j < nd; ++j) wk[nu + j] = pb->v[k * nd + j];
        /* time update of the controller state (model) and step of the real plant, both with the true w(k) */
        for (int j = 0; j < nw; ++j) {
            for (int qh = hl - 1; qh > 0; --qh) { wh[j * hl + qh] = wh[j * hl + qh - 1]; whp[j * hl + qh] = whp[j * hl + qh - 1]; }
            wh[j * hl] = wk[j]; whp[j * hl] = wk[j];
        }
        for (int i = 0; i < ny; ++i)
            for (int j = 0; j < nw; ++j) {
                const int ch = i * nw + j;
                int dd = pb->d[ch];
                xs[ch] = pb->a[ch] * xs[ch] + pb->b0[ch] * (dd >= 1 ? wh[j * hl + dd - 1] : 0.0) + pb->b1[ch] * wh[j * hl + dd];
                dd = pd[ch];
                xp[ch] = pa[ch] * xp[ch] + pb0[ch] * (dd >= 1 ? whp[j * hl + dd - 1] : 0.0) + pb1[ch] * whp[j * hl + dd];
            }
    }
    if (stats) { stats[0] += nqp; stats[1] += nit_as; stats[2] += nqp_act; }
    free(xs); free(wh); free(xod); free(xp); free(whp); free(hv); free(up); free(wk); free(e);
    gi_free(gw);
    ctx_free(&c);
    return status;
}

/* One mpcmove from a given state, for QP-level tests (tests/test_oracle_qp.py): returns the
 * optimal z (nz), H (nz*nz), f (nz) so that an independent solver can be run on the same QP. */
int orc_single_qp(const orc_problem *pb, int p, int m, const double *delta, const double *lambda,
                  const double *xs, const double *whist, int hl, const double *hv, const double *rk,
                  const double *uprev, double *z_out, double *H_out, double *f_out, double *G_out,
                  double *yfree_out, int *nz_out, int *iters_out) {
    cand_ctx c;
    memset(&c, 0, sizeof(c));
    int rc = ctx_build(&c, pb, p, m, delta, lambda);
    if (rc) { ctx_free(&c); return rc; }
    gi_work *gw = gi_alloc(c.nz, 4 * c.nzu + 2 * pb->ny * p + 1);
    int iters = 0, nact = 0;
    rc = ctx_move(&c, gw, xs, whist, hl, hv, rk, uprev, &iters, &nact);
    memcpy(z_out, c.x, sizeof(double) * c.nz);
    memcpy(H_out, c.H, sizeof(double) * c.nz * c.nz);
    memcpy(f_out, c.f, sizeof(double) * c.nz);
    if (G_out) memcpy(G_out, c.G, sizeof(double) * pb->ny * p * c.nzu);
    if (yfree_out) memcpy(yfree_out, c.yfree, sizeof(double) * pb->ny * p);
    *nz_out = c.nz; *iters_out = iters;
    gi_free(gw);
    ctx_free(&c);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Objectives                                                                                 */
/* ------------------------------------------------------------------------------------------ */
/* GAM_fun.m:110-115: g_i = sum_k (y_i(k) - Yref_i(k))^2 over ALL samples. */
void orc_cost_gam(int ny, int nit, const double *y, const double *yref, double *g) {
    for (int i = 0; i < ny; ++i) {
        double acc = 0.0;
        for (int k = 0; k < nit; ++k) { double e = y[i * nit + k] - yref[i * nit + k]; acc += e * e; }
        g[i] = acc;
    }
}

/* VNS2.m:172-195 on assembled rows: F = sum_i(j21_i + j22_i) + N + sum_j Jnu_j.
 * Xy,Xyma are nrow_y x nit ; Xuma is nrow_u x nit ; inK is 1-based (VNS2.m:43). */
double orc_cost_vns(int nrow_y, int nrow_u, int nit, int inK, int N, const double *Xy, const double *Xyma,
                    const double *Xuma, const double *yref) {
    double F = 0.0;
    for (int i = 0; i < nrow_y; ++i)
        for (int k = inK - 1; k < nit; ++k) {
            double e2 = Xy[i * nit + k] - Xyma[i * nit + k];
            double er = Xy[i * nit + k] - yref[i * nit + k];
            F += e2 * e2 + er * er;
        }
    for (int j = 0; j < nrow_u; ++j) {
        double u0 = fabs(Xuma[j * nit + 0]);     /* VNS2.m:185 uses column 1, not inK */
        for (int k = 0; k + 1 < nit; ++k) {
            double df = fabs(Xuma[j * nit + k + 1] - Xuma[j * nit + k]);
            double xn = u0 / df;
            if (isinf(xn) || isnan(xn)) xn = 0.0; /* VNS2.m:186 */
            F += xn * xn;
        }
    }
    return F + (double)N;
}

/* Batch evaluation used for the CPU baseline timing and the parity tests.
 * mode 0: GAM (cost is n x ny, row-major per candidate), pb->r is the user's set-point.
 * mode 1: VNS (cost is n): square plants run ny single-set-point closed loops (VNS2.m:148-165),
 *         non-square plants one (VNS2.m:168); set-point = unit step from inK (VNS2.m:58-61).
 * status: n ints.  stats: 3 counters accumulated over the batch.  nthreads <= 0: all. */
int orc_eval_batch(const orc_problem *pb, int n, const int *N, const int *Nu, const double *delta,
                   const double *lambda, int mode, const double *yref, int inK, double *cost, int *status,
                   long long *stats, int nthreads) {
    const int ny = pb->ny, nu = pb->nu, nit = pb->nit;
    long long s0 = 0, s1 = 0, s2 = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : s0, s1, s2)
    for (int c = 0; c < n; ++c) {
        double *y = (double *)malloc(sizeof(double) * nit * ny);
        double *u = (double *)malloc(sizeof(double) * nit * nu);
        double *ys = (double *)malloc(sizeof(double) * nit * ny);
        double *uo = (double *)malloc(sizeof(double) * nit * nu);
        long long st[3] = {0, 0, 0};
        const double *dl = delta + (size_t)c * ny, *lm = lambda + (size_t)c * nu;
        int rc = 0;
        if (mode == 0) {
            rc = orc_closedloop(pb, N[c], Nu[c], dl, lm, y, u, NULL, NULL, st);
            orc_cost_gam(ny, nit, y, yref, cost + (size_t)c * ny);
        } else {
            orc_problem q = *pb;
            double *rr = (double *)calloc((size_t)nit * ny, sizeof(double));
            q.r = rr;
            if (ny == nu) {
                double *Xy = (double *)malloc(sizeof(double) * nit * ny);
                double *Xyma = (double *)malloc(sizeof(double) * nit * ny);
                double *Xuma = (double *)malloc(sizeof(double) * nit * nu);
                for (int i = 0; i < ny; ++i) {
                    memset(rr, 0, sizeof(double) * nit * ny);
                    for (int k = inK - 1; k < nit; ++k) rr[k * ny + i] = 1.0;
                    int r1 = orc_closedloop(&q, N[c], Nu[c], dl, lm, y, u, ys, uo, st);
                    if (r1) rc = r1;
                    memcpy(Xy + i * nit, y + i * nit, sizeof(double) * nit);
                    memcpy(Xyma + i * nit, ys + i * nit, sizeof(double) * nit);
                    memcpy(Xuma + i * nit, uo + i * nit, sizeof(double) * nit);
                }
                cost[c] = orc_cost_vns(ny, nu, nit, inK, N[c], Xy, Xyma, Xuma, yref);
                free(Xy); free(Xyma); free(Xuma);
            } else {
                for (int k = inK - 1; k < nit; ++k)
                    for (int i = 0; i < ny; ++i) rr[k * ny + i] = 1.0;
                rc = orc_closedloop(&q, N[c], Nu[c], dl, lm, y, u, ys, uo, st);
                cost[c] = orc_cost_vns(ny, nu, nit, inK, N[c], y, ys, uo, yref);
            }
            free(rr);
        }
        if (status) status[c] = rc;
        s0 += st[0]; s1 += st[1]; s2 += st[2];
        free(y); free(u); free(ys); free(uo);
    }
    if (stats) { stats[0] += s0; stats[1] += s1; stats[2] += s2; }
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
