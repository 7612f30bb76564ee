/*
 * mpcgpu_mex.c -- thin MEX gateway over libmpcgpu.so (include/mpcgpu.h).
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has neither MATLAB nor `mex`.  It is the
 * binding a maintainer of sergioacg/Model-Predictive-Control-Tuning adds to switch the tuner's evaluator
 * to the GPU:   mex -I<repo>/include mpcgpu_mex.c -L<repo>/model-predictive-control-tuning_b200/csrc -lmpcgpu
 *
 * MATLAB side (drop-in for MPC_Tuning/closedloop_toolbox.m:1, GAM_fun.m:81, VNS2.m:153/168):
 *   h      = mpcgpu_mex('create', P)                          P: struct, fields of mpcgpu_problem (scaled plant ...)
 *   cost   = mpcgpu_mex('eval', h, N, Nu, delta, lambda, mode) mode: 'gam' | 'vns';  N,Nu int32 n x 1,
 *                                                              delta n x ny, lambda n x nu (MATLAB column-major is
 *                                                              transposed here into the ABI's candidate-major rows)
 *   [y,u,ys,uopt] = mpcgpu_mex('closedloop', h, r, v, N, Nu, delta, lambda, nit)   one candidate, signals x time
 *            mpcgpu_mex('destroy', h)
 * A failed candidate raises mpcgpu:candidate so the reference's try/catch blocks (GAM_fun.m:80-91,
 * VNS2.m:151-163) behave as with the Toolbox.
 */
#include <string.h>

#include "mex.h"
#include "mpcgpu.h"

static double *f64(const mxArray *s, const char *name) {
    const mxArray *f = mxGetField(s, 0, name);
    if (!f || !mxIsDouble(f)) mexErrMsgIdAndTxt("mpcgpu:arg", "problem field %s missing or not double", name);
    return mxGetPr(f);
}
static int32_t *i32(const mxArray *s, const char *name) {
    const mxArray *f = mxGetField(s, 0, name);
    if (!f || !mxIsInt32(f)) mexErrMsgIdAndTxt("mpcgpu:arg", "problem field %s missing or not int32", name);
    return (int32_t *)mxGetData(f);
}
static int scalar_i(const mxArray *s, const char *name) {
    const mxArray *f = mxGetField(s, 0, name);
    if (!f) mexErrMsgIdAndTxt("mpcgpu:arg", "problem field %s missing", name);
    return (int)mxGetScalar(f);
}
static mpcgpu_handle *handle_of(const mxArray *a) { return (mpcgpu_handle *)(uintptr_t)(*(uint64_t *)mxGetData(a)); }

/* n x k column-major (MATLAB) -> candidate-major rows (ABI) */
static double *rows_of(const mxArray *a, mwSize n, mwSize k) {
    const double *src = mxGetPr(a);
    double *dst = (double *)mxMalloc(sizeof(double) * n * k);
    for (mwSize c = 0; c < n; ++c)
        for (mwSize j = 0; j < k; ++j) dst[c * k + j] = src[j * n + c];
    return dst;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    char cmd[32];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof(cmd))) mexErrMsgIdAndTxt("mpcgpu:arg", "first argument: command");
    if (!strcmp(cmd, "create")) {
        const mxArray *P = prhs[1];
        mpcgpu_problem pb;
        memset(&pb, 0, sizeof(pb));
        pb.ny = scalar_i(P, "ny"); pb.nu = scalar_i(P, "nu"); pb.nd = scalar_i(P, "nd"); pb.nit = scalar_i(P, "nit");
        pb.pmax = scalar_i(P, "pmax"); pb.mmax = scalar_i(P, "mmax"); pb.inK = scalar_i(P, "inK");
        pb.a = f64(P, "a"); pb.b0 = f64(P, "b0"); pb.b1 = f64(P, "b1"); pb.d = i32(P, "d");   /* row-major ny x (nu+nd) */
        pb.umin = f64(P, "umin"); pb.umax = f64(P, "umax"); pb.dumin = f64(P, "dumin"); pb.dumax = f64(P, "dumax");
        pb.ymin = f64(P, "ymin"); pb.ymax = f64(P, "ymax"); pb.ecr_min = f64(P, "ecr_min"); pb.ecr_max = f64(P, "ecr_max");
        pb.su = f64(P, "su"); pb.sy = f64(P, "sy"); pb.rho_ecr = mxGetScalar(mxGetField(P, 0, "rho_ecr"));
        pb.r = f64(P, "r"); pb.v = pb.nd ? f64(P, "v") : NULL; pb.yref = f64(P, "yref"); pb.dmin = i32(P, "dmin");
        mpcgpu_handle *h = NULL;
        if (mpcgpu_create(&pb, -1, &h) != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:create", "%s", mpcgpu_last_error(NULL));
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t *)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)h;
    } else if (!strcmp(cmd, "eval")) {
        mpcgpu_handle *h = handle_of(prhs[1]);
        const mwSize n = mxGetNumberOfElements(prhs[2]);
        const mwSize ny = mxGetN(prhs[4]), nu = mxGetN(prhs[5]);
        char mode[8];
        mxGetString(prhs[6], mode, sizeof(mode));
        const int cm = !strcmp(mode, "vns") ? MPCGPU_COST_VNS : MPCGPU_COST_GAM;
        double *dl = rows_of(prhs[4], n, ny), *lm = rows_of(prhs[5], n, nu);
        const mwSize kc = cm == MPCGPU_COST_GAM ? ny : 1;
        double *cost = (double *)mxMalloc(sizeof(double) * n * kc);
        int32_t *st = (int32_t *)mxMalloc(sizeof(int32_t) * n);
        int rc = mpcgpu_eval_batch(h, (int)n, (const int32_t *)mxGetData(prhs[2]), (const int32_t *)mxGetData(prhs[3]), dl, lm, cm,
                                   cost, NULL, NULL, NULL, NULL, st);
        if (rc != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_last_error(h));
        plhs[0] = mxCreateDoubleMatrix(n, kc, mxREAL);
        for (mwSize c = 0; c < n; ++c)
            for (mwSize j = 0; j < kc; ++j) mxGetPr(plhs[0])[j * n + c] = cost[c * kc + j];
        if (nlhs > 1) {
            plhs[1] = mxCreateNumericMatrix(n, 1, mxINT32_CLASS, mxREAL);
            memcpy(mxGetData(plhs[1]), st, sizeof(int32_t) * n);
        } else {
            for (mwSize c = 0; c < n; ++c)
                if (st[c]) mexErrMsgIdAndTxt("mpcgpu:candidate", "Error in closed-loop simulation (candidate %d, status %d)", (int)c + 1, st[c]);
        }
        mxFree(dl); mxFree(lm); mxFree(cost); mxFree(st);
    } else if (!strcmp(cmd, "closedloop")) {
        /* [y,u,ys,uopt] = closedloop(h, r (nit x ny), v (nit x nd), N, Nu, delta, lambda, nit): closedloop_toolbox.m */
        mpcgpu_handle *h = handle_of(prhs[1]);
        const int nit = (int)mxGetScalar(prhs[8]);
        const mwSize ny = mxGetNumberOfElements(prhs[6]), nu = mxGetNumberOfElements(prhs[7]);
        double *r = rows_of(prhs[2], nit, ny);
        double *v = mxGetNumberOfElements(prhs[3]) ? rows_of(prhs[3], nit, mxGetN(prhs[3])) : NULL;
        if (mpcgpu_set_signals(h, nit, r, v, NULL) != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:signals", "%s", mpcgpu_last_error(h));
        int32_t N = (int32_t)mxGetScalar(prhs[4]), Nu = (int32_t)mxGetScalar(prhs[5]), st = 0;
        /* signals x time, row-major in the ABI == (time x signals) column-major: transpose on the way out */
        double *y = (double *)mxMalloc(sizeof(double) * ny * nit), *u = (double *)mxMalloc(sizeof(double) * nu * nit);
        double *ys = (double *)mxMalloc(sizeof(double) * ny * nit), *uo = (double *)mxMalloc(sizeof(double) * nu * nit);
        int rc = mpcgpu_eval_batch(h, 1, &N, &Nu, mxGetPr(prhs[6]), mxGetPr(prhs[7]), MPCGPU_COST_RAW, NULL, y, u, ys, uo, &st);
        if (rc != MPCGPU_OK) mexErrMsgIdAndTxt("mpcgpu:eval", "%s", mpcgpu_last_error(h));
        if (st) mexErrMsgIdAndTxt("mpcgpu:candidate", "Error in closed-loop simulation (status %d)", st);
        double *src[4] = {y, u, ys, uo};
        mwSize rows[4] = {ny, nu, ny, nu};
        for (int o = 0; o < 4 && o < (nlhs ? nlhs : 1); ++o) {
            plhs[o] = mxCreateDoubleMatrix(rows[o], nit, mxREAL);
            for (mwSize i = 0; i < rows[o]; ++i)
                for (int k = 0; k < nit; ++k) mxGetPr(plhs[o])[k * rows[o] + i] = src[o][i * nit + k];
        }
        mxFree(r); if (v) mxFree(v); mxFree(y); mxFree(u); mxFree(ys); mxFree(uo);
    } else if (!strcmp(cmd, "destroy")) {
        mpcgpu_destroy(handle_of(prhs[1]));
    } else {
        mexErrMsgIdAndTxt("mpcgpu:arg", "unknown command %s", cmd);
    }
}
