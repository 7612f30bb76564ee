/*
 * mpcgpu_mex.c -- thin MEX gateway over libmpcgpu.so (include/mpcgpu.h).
 *
 * Build (on a machine with MATLAB):
 *   mex -I<repo>/include mpcgpu_mex.c -L<repo>/model-predictive-control-tuning_b200/csrc -lmpcgpu
 * The build image of this repository has neither MATLAB nor `mex`; the file is compiled here against a stub `mex.h`
 * (tests/mex_stub) and `mexFunction` is driven through create -> eval -> closedloop -> destroy by tests/test_mex.py
 * (CPU: argument handling and error paths; GPU box: results against the ctypes binding).
 *
 * Commands (first argument).  The signature-identical MATLAB wrappers are matlab/closedloop_toolbox.m and
 * matlab/closedloop_toolbox_nmpc.m (same argument lists and output counts as the reference's functions).
 *
 *   linear path (MPC_Tuning/closedloop_toolbox.m:1, GAM_fun.m:81, VNS2.m:153/168)
 *     h = mpcgpu_mex('create', P)                       P: struct with the fields of mpcgpu_problem
 *     h = mpcgpu_mex('create_multi', P, devices)        devices: int32 vector (empty: all) -- one evaluator per GPU
 *     [cost, status] = mpcgpu_mex('eval', h, N, Nu, delta, lambda, mode)        mode 'gam' | 'vns'; N, Nu n x 1 (any
 *                      numeric class), delta n x ny, lambda n x nu; cost n x ny | n x 1.  With one output a failed
 *                      candidate raises mpcgpu:candidate (the reference's try/catch blocks keep working).
 *     [cost, status] = mpcgpu_mex('eval_multi', hm, N, Nu, delta, lambda, mode) the same over all GPUs of the handle
 *     [y,u,t,ys,uopt] = mpcgpu_mex('closedloop', h, r, v, N, Nu, delta, lambda, nit)   closedloop_toolbox.m:1 -- r, v in
 *                      either orientation (row2col.m), N / Nu vectors -> max (:38-40), outputs signals x time (:103-107);
 *                      the handle's own signals (Par.Xsp, Par.Yref) are untouched.  Needs P.Ts for t.
 *     mpcgpu_mex('option', h, 'vns_legality', 0|1)      VNS2.m:135 inside the library
 *     mpcgpu_mex('mismatch', h, Pl, gain, hl)           validation run against a real plant that differs from the model
 *                      (Shell3x3.m:271-286  options.Model = plant): Pl struct with a, b0, b1, d (ny x nw, row-major like P's),
 *                      gain (ny*nw + nu*hl + ny) x ny MATLAB matrix of the state estimator (mpcgpu.h, mpcgpu_set_mismatch);
 *                      mpcgpu_mex('mismatch', h) switches back to the nominal evaluation
 *     mpcgpu_mex('destroy', h) / mpcgpu_mex('destroy_multi', hm)
 *   nonlinear path (closedloop_toolbox_nmpc.m:1)
 *     hn = mpcgpu_mex('nmpc_create', Pn)                Pn: struct with the fields of mpcgpu_nmpc_problem
 *     [cost, status] = mpcgpu_mex('nmpc_eval', hn, N, Nu, delta, lambda, mode)
 *     [y,u,yopt,uopt] = mpcgpu_mex('nmpc_closedloop', hn, r, N, Nu, delta, lambda, nit)   r: 2 x nit (either orientation)
 *     mpcgpu_mex('nmpc_destroy', hn)
 *   single-shooting NMPC (Explicit NMPC/ClosedLoopNMPC.m:1, NMPC_Controller.m:1)
 *     hs = mpcgpu_mex('ssnmpc_create', Ps)              Ps: struct with the fields of mpcgpu_ssnmpc_problem (x_control 1-based)
 *     [cost, status] = mpcgpu_mex('ssnmpc_eval', hs, N, Nu, Q, W)                 N n x 1, Nu / Q / W n x 2: a sweep
 *     [y, u] = mpcgpu_mex('ssnmpc_closedloop', hs, r, N, Nu, Q, W [, noise])      one run; r 2 x nit, noise 3 x nit (the randn
 *                      draws of ClosedLoopNMPC.m:89, already scaled), either orientation
 *     mpcgpu_mex('ssnmpc_destroy', hs)
 *   DTC-GPC sweep (DTC-GPC/DTC_GPC_WW.m:56-164)
 *     hd = mpcgpu_mex('dtc_create', Pd)                 Pd: struct with the fields of mpcgpu_dtc_problem
 *     [ise, status, y, u] = mpcgpu_mex('dtc_eval', hd, p, m, delta, lambda, fr_num, fr_den, fr_len)
 *                      p n x ny, m n x nu, fr_num / fr_den n x ny x MAXF, fr_len n x ny x 2 (MATLAB arrays, column-major)
 *     [ise, status, y, u] = mpcgpu_mex('dtc_eval_design', hd, p, m, delta, lambda, alfa, raio)   the same with the robustness
 *                      filter of every candidate designed on the device (mimofilter.m / filtro_siso.m), alfa, raio n x 1
 *     mpcgpu_mex('dtc_destroy', hd)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "mex.h"
#include "mpcgpu.h"

#define ARG(cond, ...) do { if (!(cond)) mexErrMsgIdAndTxt("mpcgpu:arg", __VA_ARGS__); } while (0)

static const mxArray *field(const mxArray *s, const char *name, int required) {
    const mxArray *f = mxIsStruct(s) ? mxGetField(s, 0, name) : NULL;
    if (!f && required) mexErrMsgIdAndTxt("mpcgpu:arg", "problem field %s missing", name);
    return f;
}
static double *f64(const mxArray *s, const char *name, int required) {
    const mxArray *f = field(s, name, required);
    if (!f) return NULL;
    if (!mxIsDouble(f)) mexErrMsgIdAndTxt("mpcgpu:arg", "problem field %s must be double", name);
    return mxGetNumberOfElements(f) ? mxGetPr(f) : NULL;
}
static int scalar_i(const mxArray *s, const char *name) { return (int)mxGetScalar(field(s, name, 1)); }
/* any numeric array -> freshly allocated int32 (MATLAB callers pass doubles: N = 24) */
static int32_t *as_i32(const mxArray *a, mwSize *n_out) {
    const mwSize n = mxGetNumberOfElements(a);
    int32_t *o = (int32_t *)mxMalloc(sizeof(int32_t) * (n ? n : 1));
    if (mxIsInt32(a)) memcpy(o, mxGetData(a), sizeof(int32_t) * n);
    else if (mxIsDouble(a)) { const double *p = mxGetPr(a); for (mwSize i = 0; i < n; ++i) o[i] = (int32_t)p[i]; }
    else mexErrMsgIdAndTxt("mpcgpu:arg", "integer arguments must be int32 or double");
    if (n_out) *n_out = n;
    return o;
}
static int32_t *i32_field(const mxArray *s, const char *name) { return as_i32(field(s, name, 1), NULL); }
static int32_t max_of(const mxArray *a) {   /* closedloop_toolbox.m:38-40: vectors collapse to their max */
    mwSize n; int32_t *v = as_i32(a, &n); int32_t m = n ? v[0] : 0;
    for (mwSize i = 1; i < n; ++i) m = v[i] > m ? v[i] : m;
    mxFree(v); return m;
}
static void *handle_of(const mxArray *a) {
    ARG(mxIsUint64(a) && mxGetNumberOfElements(a) == 1, "handle must be the uint64 scalar returned by create");
    return (void *)(uintptr_t)(*(uint64_t *)mxGetData(a));
}
static mxArray *handle_out(void *h) {
    mxArray *o = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *(uint64_t *)mxGetData(o) = (uint64_t)(uintptr_t)h;
    return o;
}
/* n x k column-major (MATLAB) -> candidate-major rows (ABI) */
static double *rows_of(const mxArray *a, mwSize n, mwSize k) {
    ARG(mxIsDouble(a) && mxGetNumberOfElements(a) == n * k, "expected a %d x %d double array", (int)n, (int)k);
    const double *src = mxGetPr(a);
    double *dst = (double *)mxMalloc(sizeof(double) * ((n * k) > 0 ? n * k : 1));
    for (mwSize c = 0; c < n; ++c)
        for (mwSize j = 0; j < k; ++j) dst[c * k + j] = src[j * n + c];
    return dst;
}
/* a signal given in either orientation (row2col.m:3-8) -> time-major nit x k rows */
static double *time_major(const mxArray *a, int nit, mwSize k) {
    ARG(mxIsDouble(a), "signals must be double");
    const mwSize m = mxGetM(a), n = mxGetN(a);
    const double *src = mxGetPr(a);
    double *dst = (double *)mxMalloc(sizeof(double) * (((mwSize)nit * k) > 0 ? (mwSize)nit * k : 1));
    if (m < n) {            /* signals x time: row2col transposes */
        ARG(m == k && n >= (mwSize)nit, "signal must be %d x >=%d", (int)k, nit);
        for (int t = 0; t < nit; ++t) for (mwSize j = 0; j < k; ++j) dst[t * k + j] = src[t * m + j];
    } else {                /* time x signals */
        ARG(n == k && m >= (mwSize)nit, "signal must be >=%d x %d", nit, (int)k);
        for (int t = 0; t < nit; ++t) for (mwSize j = 0; j < k; ++j) dst[t * k + j] = src[j * m + t];
    }
    return dst;
}
/* ABI block (rows x nit, row-major) -> MATLAB rows x nit matrix */
static mxArray *sig_out(const double *src, mwSize rows, int nit) {
    mxArray *o = mxCreateDoubleMatrix(rows, nit, mxREAL);
    for (mwSize i = 0; i < rows; ++i)
        for (int k = 0; k < nit; ++k) mxGetPr(o)[k * rows + i] = src[i * nit + k];
    return o;
}
static int cost_mode_of(const mxArray *a) {
    char mode[8];
    ARG(!mxGetString(a, mode, sizeof(mode)), "mode must be 'gam' or 'vns'");
    if (!strcmp(mode, "vns")) return MPCGPU_COST_VNS;
    ARG(!strcmp(mode, "gam"), "mode must be 'gam' or 'vns'");
    return MPCGPU_COST_GAM;
}
static void cost_out(int nlhs, mxArray *plhs[], const double *cost, const int32_t *st, mwSize n, mwSize kc) {
    plhs[0] = mxCreateDoubleMatrix(n, kc, mxREAL);
    for (mwSize c = 0; c < n; ++c)
        for (mwSize j = 0; j < kc; ++j) mxGetPr(plhs[0])[j * n + c] = cost[c * kc + j];
    if (nlhs > 1) {
        plhs[1] = mxCreateNumericMatrix(n, 1, mxINT32_CLASS, mxREAL);
        memcpy(mxGetData(plhs[1]), st, sizeof(int32_t) * n);
    } else {
        for (mwSize c = 0; c < n; ++c)
            if (st[c] && st[c] != MPCGPU_CAND_BOUND_CROSSED)
                mexErrMsgIdAndTxt("mpcgpu:candidate", "Error in closed-loop simulation (candidate %d, status %d)", (int)c + 1, (int)st[c]);
    }
}

/* the per-problem sampling time, kept for the `t` output of closedloop (closedloop_toolbox.m:102) */
#define MAX_HANDLES 64
static struct { void *h; double Ts; int ny, nu, nd; } g_lin[MAX_HANDLES];
static void remember(void *h, double Ts, int ny, int nu, int nd) {
    for (int i = 0; i < MAX_HANDLES; ++i) if (!g_lin[i].h) { g_lin[i].h = h; g_lin[i].Ts = Ts; g_lin[i].ny = ny; g_lin[i].nu = nu; g_lin[i].nd = nd; return; }
}
static int lookup(void *h) { for (int i = 0; i < MAX_HANDLES; ++i) if (g_lin[i].h == h) return i; return -1; }
static void forget(void *h) { const int i = lookup(h); if (i >= 0) g_lin[i].h = NULL; }
static int nit_of(void *h) { const int i = lookup(h); return i >= 0 ? g_lin[i].nd : 0; }   /* handles whose nd slot keeps nit */

static void fill_problem(const mxArray *P, mpcgpu_problem *pb, int32_t **d, int32_t **dmin) {
    memset(pb, 0, sizeof(*pb));
    pb->ny = scalar_i(P, "ny"); pb->nu = scalar_i(P, "nu"); pb->nd = scalar_i(P, "nd"); pb->nit = scalar_i(P, "nit");
    pb->pmax = scalar_i(P, "pmax"); pb->mmax = scalar_i(P, "mmax"); pb->inK = scalar_i(P, "inK");
    pb->a = f64(P, "a", 1); pb->b0 = f64(P, "b0", 1); pb->b1 = f64(P, "b1", 1);      /* row-major ny x (nu+nd) */
    *d = i32_field(P, "d"); pb->d = *d;
    pb->umin = f64(P, "umin", 1); pb->umax = f64(P, "umax", 1); pb->dumin = f64(P, "dumin", 1); pb->dumax = f64(P, "dumax", 1);
    pb->ymin = f64(P, "ymin", 0); pb->ymax = f64(P, "ymax", 0); pb->ecr_min = f64(P, "ecr_min", 0); pb->ecr_max = f64(P, "ecr_max", 0);
    pb->su = f64(P, "su", 1); pb->sy = f64(P, "sy", 1);
    pb->rho_ecr = field(P, "rho_ecr", 0) ? mxGetScalar(field(P, "rho_ecr", 0)) : 1e5;
    pb->r = f64(P, "r", 1); pb->v = pb->nd ? f64(P, "v", 1) : NULL; pb->yref = f64(P, "yref", 1);
    *dmin = i32_field(P, "dmin"); pb->dmin = *dmin;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    char cmd[32];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof(cmd))) mexErrMsgIdAndTxt("mpcgpu:arg", "first argument: command");
    /* ------------------------------------------------ linear path ------------------------------------------------ */
    if (!strcmp(cmd, "create") || !strcmp(cmd, "create_multi")) {
        ARG(nrhs >= 2 && mxIsStruct(prhs[1]), "create: problem struct expected");
        mpcgpu_problem pb; int32_t *d, *dmin;
        fill_problem(prhs[1], &pb, &d, &dmin);
        const double Ts = field(prhs[1], "Ts", 0) ? mxGetScalar(field(prhs[1], "Ts", 0)) : 1.0;
        void *h = NULL;
        int rc;
        if (!strcmp(cmd, "create")) {
            const int dev = nrhs > 2 ? (int)mxGetScalar(prhs[2]) : -1;
            rc = mpcgpu_create(&pb, dev, (mpcgpu_handle **)&h);
        } else {
            mwSize nd_ = 0; int32_t *devs = nrhs > 2 && mxGetNumberOfElements(prhs[2]) ? as_i32(prhs[2], &nd_) : NULL;
            const int ndev = devs ? (int)nd_ : mpcgpu_device_count();
            rc = mpcgpu_create_multi(&pb, devs, ndev, (mpcgpu_multi **)&h);
            if (devs) mxFree(devs);
        }
        mxFree(d); mxFree(dmin);
        if (rc != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:create", "%s", !strcmp(cmd, "create") ? mpcgpu_last_error(NULL) : mpcgpu_multi_last_error(NULL));
        remember(h, Ts, pb.ny, pb.nu, pb.nd);
        plhs[0] = handle_out(h);
    } else if (!strcmp(cmd, "eval") || !strcmp(cmd, "eval_multi")) {
        ARG(nrhs == 7, "eval: (h, N, Nu, delta, lambda, mode)");
        void *h = handle_of(prhs[1]);
        const int ix = lookup(h);
        ARG(ix >= 0, "unknown handle");
        mwSize n, n2;
        int32_t *N = as_i32(prhs[2], &n), *Nu = as_i32(prhs[3], &n2);
        ARG(n == n2, "N and Nu must have the same length");
        const mwSize ny = g_lin[ix].ny, nu = g_lin[ix].nu;
        const int cm = cost_mode_of(prhs[6]);
        double *dl = rows_of(prhs[4], n, ny), *lm = rows_of(prhs[5], n, nu);
        const mwSize kc = cm == MPCGPU_COST_GAM ? ny : 1;
        double *cost = (double *)mxMalloc(sizeof(double) * ((n * kc) > 0 ? n * kc : 1));
        int32_t *st = (int32_t *)mxMalloc(sizeof(int32_t) * (n ? n : 1));
        int rc;
        if (!strcmp(cmd, "eval")) {
            rc = mpcgpu_eval_batch((mpcgpu_handle *)h, (int)n, N, Nu, dl, lm, cm, cost, NULL, NULL, NULL, NULL, st);
            if (rc != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_last_error((mpcgpu_handle *)h));
        } else {
            rc = mpcgpu_multi_eval_batch((mpcgpu_multi *)h, (int)n, N, Nu, dl, lm, cm, cost, st);
            if (rc != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_multi_last_error((mpcgpu_multi *)h));
        }
        cost_out(nlhs, plhs, cost, st, n, kc);
        mxFree(N); mxFree(Nu); mxFree(dl); mxFree(lm); mxFree(cost); mxFree(st);
    } else if (!strcmp(cmd, "closedloop")) {
        /* [y,u,t,ys,uopt] = closedloop(h, r, v, N, Nu, delta, lambda, nit) -- closedloop_toolbox.m:1 */
        ARG(nrhs == 9, "closedloop: (h, r, v, N, Nu, delta, lambda, nit)");
        mpcgpu_handle *h = (mpcgpu_handle *)handle_of(prhs[1]);
        const int ix = lookup(h);
        ARG(ix >= 0, "unknown handle");
        const int nit = (int)mxGetScalar(prhs[8]);
        const mwSize ny = g_lin[ix].ny, nu = g_lin[ix].nu, nd = g_lin[ix].nd;
        ARG(mxGetNumberOfElements(prhs[6]) == ny && mxGetNumberOfElements(prhs[7]) == nu, "delta must have ny entries, lambda nu");
        double *r = time_major(prhs[2], nit, ny);
        double *v = nd ? time_major(prhs[3], nit, nd) : NULL;           /* mdv may be nit x 0 (MPCTuning.m:113) */
        const int32_t N = max_of(prhs[4]), Nu = max_of(prhs[5]);
        int32_t st = 0;
        double *y = (double *)mxMalloc(sizeof(double) * ny * nit), *u = (double *)mxMalloc(sizeof(double) * nu * nit);
        double *ys = (double *)mxMalloc(sizeof(double) * ny * nit), *uo = (double *)mxMalloc(sizeof(double) * nu * nit);
        const int rc = mpcgpu_closedloop(h, nit, r, v, N, Nu, mxGetPr(prhs[6]), mxGetPr(prhs[7]), y, u, ys, uo, &st);
        if (rc != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_last_error(h));
        if (st) mexErrMsgIdAndTxt("mpcgpu:candidate", "Error in closed-loop simulation (status %d)", (int)st);
        plhs[0] = sig_out(y, ny, nit);
        if (nlhs > 1) plhs[1] = sig_out(u, nu, nit);
        if (nlhs > 2) {                                                   /* t = 0:Ts:(nit-1)*Ts as a row (closedloop_toolbox.m:102,105) */
            plhs[2] = mxCreateDoubleMatrix(1, nit, mxREAL);
            for (int k = 0; k < nit; ++k) mxGetPr(plhs[2])[k] = k * g_lin[ix].Ts;
        }
        if (nlhs > 3) plhs[3] = sig_out(ys, ny, nit);
        if (nlhs > 4) plhs[4] = sig_out(uo, nu, nit);
        mxFree(r); if (v) mxFree(v); mxFree(y); mxFree(u); mxFree(ys); mxFree(uo);
    } else if (!strcmp(cmd, "option")) {
        ARG(nrhs == 4, "option: (h, name, value)");
        char name[32];
        ARG(!mxGetString(prhs[2], name, sizeof(name)) && !strcmp(name, "vns_legality"), "unknown option");
        if (mpcgpu_set_option((mpcgpu_handle *)handle_of(prhs[1]), MPCGPU_OPT_VNS_LEGALITY, (int)mxGetScalar(prhs[3])) != MPCGPU_OK)
            mexErrMsgIdAndTxt("mpcgpu:arg", "set_option failed");
    } else if (!strcmp(cmd, "mismatch")) {
        ARG(nrhs == 2 || nrhs == 5, "mismatch: (h) or (h, plant struct, gain, hl)");
        mpcgpu_handle *h = (mpcgpu_handle *)handle_of(prhs[1]);
        if (nrhs == 2) {
            if (mpcgpu_set_mismatch(h, NULL, NULL, NULL, NULL, NULL, 0) != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:arg", "%s", mpcgpu_last_error(h));
        } else {
            const int ix = lookup(h);
            ARG(ix >= 0 && mxIsStruct(prhs[2]), "mismatch: unknown handle / plant struct expected");
            const int hl = (int)mxGetScalar(prhs[4]);
            const mwSize ns = (mwSize)(g_lin[ix].ny * (g_lin[ix].nu + g_lin[ix].nd) + g_lin[ix].nu * hl + g_lin[ix].ny);
            double *gain = rows_of(prhs[3], ns, (mwSize)g_lin[ix].ny);
            int32_t *d = i32_field(prhs[2], "d");
            const int rc = mpcgpu_set_mismatch(h, f64(prhs[2], "a", 1), f64(prhs[2], "b0", 1), f64(prhs[2], "b1", 1), d, gain, hl);
            mxFree(gain); mxFree(d);
            if (rc != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:arg", "%s", mpcgpu_last_error(h));
        }
    } else if (!strcmp(cmd, "destroy")) {
        void *h = handle_of(prhs[1]); forget(h); mpcgpu_destroy((mpcgpu_handle *)h);
    } else if (!strcmp(cmd, "destroy_multi")) {
        void *h = handle_of(prhs[1]); forget(h); mpcgpu_destroy_multi((mpcgpu_multi *)h);
    /* ------------------------------------------------ nonlinear path ------------------------------------------------ */
    } else if (!strcmp(cmd, "nmpc_create")) {
        ARG(nrhs >= 2 && mxIsStruct(prhs[1]), "nmpc_create: problem struct expected");
        const mxArray *P = prhs[1];
        mpcgpu_nmpc_problem pb;
        memset(&pb, 0, sizeof(pb));
        pb.nit = scalar_i(P, "nit"); pb.pmax = scalar_i(P, "pmax"); pb.mmax = scalar_i(P, "mmax"); pb.inK = scalar_i(P, "inK");
        pb.nsub = field(P, "nsub", 0) ? scalar_i(P, "nsub") : 4; pb.max_sqp = field(P, "max_sqp", 0) ? scalar_i(P, "max_sqp") : 30;
        pb.model = MPCGPU_MODEL_VANDEVUSSE; pb.Ts = mxGetScalar(field(P, "Ts", 1));
        pb.x0 = f64(P, "x0", 1); pb.u0 = f64(P, "u0", 1); pb.umin = f64(P, "umin", 1); pb.umax = f64(P, "umax", 1);
        pb.xmin = f64(P, "xmin", 0); pb.xmax = f64(P, "xmax", 0); pb.su = f64(P, "su", 1); pb.sy = f64(P, "sy", 1);
        pb.r = f64(P, "r", 1); pb.yref = f64(P, "yref", 1);              /* 2 x nit, row-major (signals x time) */
        mpcgpu_nmpc_handle *h = NULL;
        if (mpcgpu_nmpc_create(&pb, nrhs > 2 ? (int)mxGetScalar(prhs[2]) : -1, &h) != MPCGPU_OK)
            mexErrMsgIdAndTxt("mpcgpu:create", "%s", mpcgpu_nmpc_last_error(NULL));
        remember(h, pb.Ts, 2, 2, 0);
        plhs[0] = handle_out(h);
    } else if (!strcmp(cmd, "nmpc_eval")) {
        ARG(nrhs == 7, "nmpc_eval: (h, N, Nu, delta, lambda, mode)");
        mpcgpu_nmpc_handle *h = (mpcgpu_nmpc_handle *)handle_of(prhs[1]);
        mwSize n, n2;
        int32_t *N = as_i32(prhs[2], &n), *Nu = as_i32(prhs[3], &n2);
        ARG(n == n2, "N and Nu must have the same length");
        const int cm = cost_mode_of(prhs[6]);
        double *dl = rows_of(prhs[4], n, 2), *lm = rows_of(prhs[5], n, 2);
        const mwSize kc = cm == MPCGPU_COST_GAM ? 2 : 1;
        double *cost = (double *)mxMalloc(sizeof(double) * ((n * kc) > 0 ? n * kc : 1));
        int32_t *st = (int32_t *)mxMalloc(sizeof(int32_t) * (n ? n : 1));
        if (mpcgpu_nmpc_eval_batch(h, (int)n, N, Nu, dl, lm, cm, NULL, cost, NULL, NULL, NULL, NULL, st) != MPCGPU_OK)
            mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_nmpc_last_error(h));
        cost_out(nlhs, plhs, cost, st, n, kc);
        mxFree(N); mxFree(Nu); mxFree(dl); mxFree(lm); mxFree(cost); mxFree(st);
    } else if (!strcmp(cmd, "nmpc_closedloop")) {
        /* [y,u,yopt,uopt] = nmpc_closedloop(h, r, N, Nu, delta, lambda, nit) -- closedloop_toolbox_nmpc.m:1 */
        ARG(nrhs == 8, "nmpc_closedloop: (h, r, N, Nu, delta, lambda, nit)");
        mpcgpu_nmpc_handle *h = (mpcgpu_nmpc_handle *)handle_of(prhs[1]);
        const int nit = (int)mxGetScalar(prhs[7]);
        double *rt = time_major(prhs[2], nit, 2);                          /* nit x 2 */
        double *r = (double *)mxMalloc(sizeof(double) * 2 * nit);          /* ABI: 2 x nit */
        for (int k = 0; k < nit; ++k) { r[k] = rt[2 * k]; r[nit + k] = rt[2 * k + 1]; }
        const int32_t N = max_of(prhs[3]), Nu = max_of(prhs[4]);
        ARG(mxGetNumberOfElements(prhs[5]) == 2 && mxGetNumberOfElements(prhs[6]) == 2, "delta and lambda must have 2 entries");
        int32_t st = 0;
        double *b[4];
        for (int o = 0; o < 4; ++o) b[o] = (double *)mxMalloc(sizeof(double) * 2 * nit);
        if (mpcgpu_nmpc_eval_batch(h, 1, &N, &Nu, mxGetPr(prhs[5]), mxGetPr(prhs[6]), MPCGPU_COST_RAW, r, NULL, b[0], b[1], b[2], b[3], &st) != MPCGPU_OK)
            mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_nmpc_last_error(h));
        if (st && st != MPCGPU_CAND_BOUND_CROSSED) mexErrMsgIdAndTxt("mpcgpu:candidate", "Error in closed-loop simulation (status %d)", (int)st);
        for (int o = 0; o < 4 && o < (nlhs ? nlhs : 1); ++o) plhs[o] = sig_out(b[o], 2, nit);
        for (int o = 0; o < 4; ++o) mxFree(b[o]);
        mxFree(rt); mxFree(r);
    } else if (!strcmp(cmd, "nmpc_destroy")) {
        void *h = handle_of(prhs[1]); forget(h); mpcgpu_nmpc_destroy((mpcgpu_nmpc_handle *)h);
    /* ------------------------------------------------ single-shooting NMPC ------------------------------------------------ */
    } else if (!strcmp(cmd, "ssnmpc_create")) {
        ARG(nrhs >= 2 && mxIsStruct(prhs[1]), "ssnmpc_create: problem struct expected");
        const mxArray *P = prhs[1];
        mpcgpu_ssnmpc_problem pb;
        memset(&pb, 0, sizeof(pb));
        pb.nit = scalar_i(P, "nit"); pb.pmax = field(P, "pmax", 0) ? scalar_i(P, "pmax") : 31; pb.inK = scalar_i(P, "inK");
        pb.nsub = field(P, "nsub", 0) ? scalar_i(P, "nsub") : 4; pb.max_sqp = field(P, "max_sqp", 0) ? scalar_i(P, "max_sqp") : 400;
        pb.model = MPCGPU_MODEL_VANDEVUSSE; pb.Ts = mxGetScalar(field(P, "Ts", 1));
        {
            mwSize nxc; int32_t *xc = as_i32(field(P, "x_control", 1), &nxc);
            ARG(nxc == 2, "x_control must have 2 entries");
            pb.x_control[0] = xc[0] - 1; pb.x_control[1] = xc[1] - 1;      /* MATLAB indices */
            mxFree(xc);
        }
        pb.x0 = f64(P, "x0", 1); pb.u0 = f64(P, "u0", 1); pb.lb = f64(P, "lb", 1); pb.ub = f64(P, "ub", 1);
        pb.r = f64(P, "r", 1);                                             /* 2 x nit, row-major (signals x time) */
        mpcgpu_ssnmpc_handle *h = NULL;
        if (mpcgpu_ssnmpc_create(&pb, nrhs > 2 ? (int)mxGetScalar(prhs[2]) : -1, &h) != MPCGPU_OK)
            mexErrMsgIdAndTxt("mpcgpu:create", "%s", mpcgpu_ssnmpc_last_error(NULL));
        remember(h, pb.Ts, 2, 2, pb.nit);
        plhs[0] = handle_out(h);
    } else if (!strcmp(cmd, "ssnmpc_eval")) {
        ARG(nrhs == 6, "ssnmpc_eval: (h, N, Nu, Q, W)");
        mpcgpu_ssnmpc_handle *h = (mpcgpu_ssnmpc_handle *)handle_of(prhs[1]);
        mwSize n, n2;
        int32_t *N = as_i32(prhs[2], &n), *NuC = as_i32(prhs[3], &n2);
        ARG(n2 == 2 * n, "Nu must be n x 2");
        int32_t *Nu = (int32_t *)mxMalloc(sizeof(int32_t) * (n2 ? n2 : 1));
        for (mwSize c = 0; c < n; ++c) { Nu[2 * c] = NuC[c]; Nu[2 * c + 1] = NuC[n + c]; }   /* column-major -> candidate rows */
        double *Q = rows_of(prhs[4], n, 2), *W = rows_of(prhs[5], n, 2);
        double *cost = (double *)mxMalloc(sizeof(double) * ((n * 2) > 0 ? n * 2 : 1));
        int32_t *st = (int32_t *)mxMalloc(sizeof(int32_t) * (n ? n : 1));
        if (mpcgpu_ssnmpc_eval_batch(h, (int)n, N, Nu, Q, W, NULL, NULL, cost, NULL, NULL, st) != MPCGPU_OK)
            mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_ssnmpc_last_error(h));
        cost_out(nlhs, plhs, cost, st, n, 2);
        mxFree(N); mxFree(NuC); mxFree(Nu); mxFree(Q); mxFree(W); mxFree(cost); mxFree(st);
    } else if (!strcmp(cmd, "ssnmpc_closedloop")) {
        /* [y, u] = ssnmpc_closedloop(h, r, N, Nu, Q, W [, noise]) -- ClosedLoopNMPC.m:1 */
        ARG(nrhs == 7 || nrhs == 8, "ssnmpc_closedloop: (h, r, N, Nu, Q, W [, noise])");
        mpcgpu_ssnmpc_handle *h = (mpcgpu_ssnmpc_handle *)handle_of(prhs[1]);
        const int nit = nit_of(h);
        ARG(nit > 0, "unknown handle");
        double *rt = time_major(prhs[2], nit, 2);                          /* nit x 2 */
        double *r = (double *)mxMalloc(sizeof(double) * 2 * nit);          /* ABI: 2 x nit */
        for (int k = 0; k < nit; ++k) { r[k] = rt[2 * k]; r[nit + k] = rt[2 * k + 1]; }
        double *nz = NULL;
        if (nrhs == 8 && mxGetNumberOfElements(prhs[7])) {
            double *nt = time_major(prhs[7], nit, 3);
            nz = (double *)mxMalloc(sizeof(double) * 3 * nit);
            for (int k = 0; k < nit; ++k) for (int i = 0; i < 3; ++i) nz[i * nit + k] = nt[3 * k + i];
            mxFree(nt);
        }
        mwSize nn; int32_t *Nv = as_i32(prhs[3], &nn);
        ARG(nn >= 1, "N missing");
        const int32_t N = Nv[0];                                           /* Q(i)*eye(N(1)), ClosedLoopNMPC.m:37 */
        mwSize nnu; int32_t *Nu = as_i32(prhs[4], &nnu);
        ARG(nnu == 2 && mxGetNumberOfElements(prhs[5]) == 2 && mxGetNumberOfElements(prhs[6]) == 2, "Nu, Q and W must have 2 entries");
        int32_t st = 0;
        double *y = (double *)mxMalloc(sizeof(double) * 2 * nit), *u = (double *)mxMalloc(sizeof(double) * 2 * nit);
        if (mpcgpu_ssnmpc_eval_batch(h, 1, &N, Nu, mxGetPr(prhs[5]), mxGetPr(prhs[6]), r, nz, NULL, y, u, &st) != MPCGPU_OK)
            mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_ssnmpc_last_error(h));
        if (st) mexErrMsgIdAndTxt("mpcgpu:candidate", "Error in closed-loop simulation (status %d)", (int)st);
        plhs[0] = sig_out(y, 2, nit);
        if (nlhs > 1) plhs[1] = sig_out(u, 2, nit);
        mxFree(rt); mxFree(r); if (nz) mxFree(nz); mxFree(Nv); mxFree(Nu); mxFree(y); mxFree(u);
    } else if (!strcmp(cmd, "ssnmpc_destroy")) {
        void *h = handle_of(prhs[1]); forget(h); mpcgpu_ssnmpc_destroy((mpcgpu_ssnmpc_handle *)h);
    /* ------------------------------------------------ DTC-GPC sweep ------------------------------------------------ */
    } else if (!strcmp(cmd, "dtc_create")) {
        ARG(nrhs >= 2 && mxIsStruct(prhs[1]), "dtc_create: problem struct expected");
        const mxArray *P = prhs[1];
        mpcgpu_dtc_problem pb;
        memset(&pb, 0, sizeof(pb));
        pb.ny = scalar_i(P, "ny"); pb.nu = scalar_i(P, "nu"); pb.nq = scalar_i(P, "nq"); pb.nit = scalar_i(P, "nit");
        pb.pmax = scalar_i(P, "pmax"); pb.mmax = scalar_i(P, "mmax"); pb.k_start = field(P, "k_start", 0) ? scalar_i(P, "k_start") : 4;
        int32_t *md = i32_field(P, "md"), *pd = i32_field(P, "pd"), *qd = i32_field(P, "qd");
        pb.ma = f64(P, "ma", 1); pb.mb0 = f64(P, "mb0", 1); pb.mb1 = f64(P, "mb1", 1); pb.md = md;
        pb.pa = f64(P, "pa", 1); pb.pb0 = f64(P, "pb0", 1); pb.pb1 = f64(P, "pb1", 1); pb.pd = pd;
        pb.qa = f64(P, "qa", 1); pb.qb0 = f64(P, "qb0", 1); pb.qb1 = f64(P, "qb1", 1); pb.qd = qd;
        pb.L = f64(P, "L", 1); pb.R = f64(P, "R", 1); pb.r = f64(P, "r", 1); pb.q = f64(P, "q", 1);
        mpcgpu_dtc_handle *h = NULL;
        const int rc = mpcgpu_dtc_create(&pb, nrhs > 2 ? (int)mxGetScalar(prhs[2]) : -1, &h);
        mxFree(md); mxFree(pd); mxFree(qd);
        if (rc != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:create", "%s", mpcgpu_dtc_last_error(NULL));
        remember(h, 1.0, pb.ny, pb.nu, pb.nit);          /* nd slot keeps nit for dtc_eval's optional trajectories */
        plhs[0] = handle_out(h);
    } else if (!strcmp(cmd, "dtc_eval")) {
        ARG(nrhs == 9, "dtc_eval: (h, p, m, delta, lambda, fr_num, fr_den, fr_len)");
        mpcgpu_dtc_handle *h = (mpcgpu_dtc_handle *)handle_of(prhs[1]);
        const int ix = lookup(h);
        ARG(ix >= 0, "unknown handle");
        const mwSize ny = g_lin[ix].ny, nu = g_lin[ix].nu, n = mxGetM(prhs[2]);
        const int nit = g_lin[ix].nd;
        /* integer matrices n x k column-major -> candidate-major */
        mwSize t_;
        int32_t *pc = as_i32(prhs[2], &t_), *mc = as_i32(prhs[3], &t_), *lc = as_i32(prhs[8], &t_);
        int32_t *p = (int32_t *)mxMalloc(sizeof(int32_t) * n * ny), *m = (int32_t *)mxMalloc(sizeof(int32_t) * n * nu);
        int32_t *fl = (int32_t *)mxMalloc(sizeof(int32_t) * n * ny * 2);
        for (mwSize c = 0; c < n; ++c) {
            for (mwSize i = 0; i < ny; ++i) p[c * ny + i] = pc[i * n + c];
            for (mwSize j = 0; j < nu; ++j) m[c * nu + j] = mc[j * n + c];
            for (mwSize i = 0; i < ny; ++i) for (int e = 0; e < 2; ++e) fl[(c * ny + i) * 2 + e] = lc[(e * ny + i) * n + c];
        }
        double *dl = rows_of(prhs[4], n, ny), *lm = rows_of(prhs[5], n, nu);
        double *fn = rows_of(prhs[6], n, ny * MPCGPU_DTC_MAXF), *fd = rows_of(prhs[7], n, ny * MPCGPU_DTC_MAXF);
        /* rows_of yields [c][f*ny + i] (MATLAB n x ny x MAXF); the ABI wants [c][i][f] */
        double *fn2 = (double *)mxMalloc(sizeof(double) * n * ny * MPCGPU_DTC_MAXF), *fd2 = (double *)mxMalloc(sizeof(double) * n * ny * MPCGPU_DTC_MAXF);
        for (mwSize c = 0; c < n; ++c)
            for (mwSize i = 0; i < ny; ++i)
                for (int f = 0; f < MPCGPU_DTC_MAXF; ++f) {
                    fn2[(c * ny + i) * MPCGPU_DTC_MAXF + f] = fn[c * ny * MPCGPU_DTC_MAXF + f * ny + i];
                    fd2[(c * ny + i) * MPCGPU_DTC_MAXF + f] = fd[c * ny * MPCGPU_DTC_MAXF + f * ny + i];
                }
        double *ise = (double *)mxMalloc(sizeof(double) * n * ny);
        int32_t *st = (int32_t *)mxMalloc(sizeof(int32_t) * (n ? n : 1));
        double *y = nlhs > 2 ? (double *)mxMalloc(sizeof(double) * n * ny * nit) : NULL;
        double *u = nlhs > 3 ? (double *)mxMalloc(sizeof(double) * n * nu * nit) : NULL;
        if (mpcgpu_dtc_eval_batch(h, (int)n, p, m, dl, lm, fn2, fd2, fl, ise, y, u, st) != MPCGPU_OK)
            mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_dtc_last_error(h));
        cost_out(nlhs > 1 ? 2 : 1, plhs, ise, st, n, ny);
        if (y) { plhs[2] = sig_out(y, n * ny, nit); mxFree(y); }          /* (n*ny) x nit, candidate-major blocks */
        if (u) { plhs[3] = sig_out(u, n * nu, nit); mxFree(u); }
        mxFree(pc); mxFree(mc); mxFree(lc); mxFree(p); mxFree(m); mxFree(fl); mxFree(dl); mxFree(lm);
        mxFree(fn); mxFree(fd); mxFree(fn2); mxFree(fd2); mxFree(ise); mxFree(st);
    } else if (!strcmp(cmd, "dtc_eval_design")) {   /* the sweep with the robustness filter designed on the device from (alfa, raio) */
        ARG(nrhs == 8, "dtc_eval_design: (h, p, m, delta, lambda, alfa, raio)");
        mpcgpu_dtc_handle *h = (mpcgpu_dtc_handle *)handle_of(prhs[1]);
        const int ix = lookup(h);
        ARG(ix >= 0, "unknown handle");
        const mwSize ny = g_lin[ix].ny, nu = g_lin[ix].nu, n = mxGetM(prhs[2]);
        const int nit = g_lin[ix].nd;
        mwSize t_;
        int32_t *pc = as_i32(prhs[2], &t_), *mc = as_i32(prhs[3], &t_);
        int32_t *p = (int32_t *)mxMalloc(sizeof(int32_t) * n * ny), *m = (int32_t *)mxMalloc(sizeof(int32_t) * n * nu);
        for (mwSize c = 0; c < n; ++c) {
            for (mwSize i = 0; i < ny; ++i) p[c * ny + i] = pc[i * n + c];
            for (mwSize j = 0; j < nu; ++j) m[c * nu + j] = mc[j * n + c];
        }
        double *dl = rows_of(prhs[4], n, ny), *lm = rows_of(prhs[5], n, nu);
        ARG(mxIsDouble(prhs[6]) && mxIsDouble(prhs[7]) && mxGetNumberOfElements(prhs[6]) == n && mxGetNumberOfElements(prhs[7]) == n, "alfa, raio: n doubles each");
        double *ise = (double *)mxMalloc(sizeof(double) * n * ny);
        int32_t *st = (int32_t *)mxMalloc(sizeof(int32_t) * (n ? n : 1));
        double *y = nlhs > 2 ? (double *)mxMalloc(sizeof(double) * n * ny * nit) : NULL;
        double *u = nlhs > 3 ? (double *)mxMalloc(sizeof(double) * n * nu * nit) : NULL;
        if (mpcgpu_dtc_eval_batch_design(h, (int)n, p, m, dl, lm, mxGetPr(prhs[6]), mxGetPr(prhs[7]), ise, y, u, st) != MPCGPU_OK)
            mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_dtc_last_error(h));
        cost_out(nlhs > 1 ? 2 : 1, plhs, ise, st, n, ny);
        if (y) { plhs[2] = sig_out(y, n * ny, nit); mxFree(y); }
        if (u) { plhs[3] = sig_out(u, n * nu, nit); mxFree(u); }
        mxFree(pc); mxFree(mc); mxFree(p); mxFree(m); mxFree(dl); mxFree(lm); mxFree(ise); mxFree(st);
    } else if (!strcmp(cmd, "dtc_destroy")) {
        void *h = handle_of(prhs[1]); forget(h); mpcgpu_dtc_destroy((mpcgpu_dtc_handle *)h);
    } else {
        mexErrMsgIdAndTxt("mpcgpu:arg", "unknown command %s", cmd);
    }
}
