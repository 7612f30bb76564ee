/* driver.c -- TEST harness: implements the stub MEX API of mex.h and lets Python (ctypes) build mxArrays, call the
 * gateway's mexFunction and read the outputs.  mexErrMsgIdAndTxt unwinds to stub_call by longjmp, as MATLAB does. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mex.h"

struct mxArray_tag {
    mxClassID cls;
    mwSize m, n;
    void *data;
    int nfields;
    char names[48][24];
    mxArray *vals[48];
};
static jmp_buf g_jmp;
static char g_errid[128], g_errmsg[512];

static size_t elsize(mxClassID c) { return c == mxDOUBLE_CLASS ? 8 : (c == mxINT32_CLASS ? 4 : (c == mxUINT64_CLASS ? 8 : 1)); }
static mxArray *mk(mxClassID cls, mwSize m, mwSize n) {
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    a->cls = cls; a->m = m; a->n = n;
    a->data = calloc((m * n) > 0 ? m * n : 1, elsize(cls));
    return a;
}
const mxArray *mxGetField(const mxArray *s, mwSize idx, const char *name) {
    (void)idx;
    if (!s || s->cls != mxSTRUCT_CLASS) return NULL;
    for (int i = 0; i < s->nfields; ++i) if (!strcmp(s->names[i], name)) return s->vals[i];
    return NULL;
}
bool mxIsStruct(const mxArray *a) { return a && a->cls == mxSTRUCT_CLASS; }
bool mxIsDouble(const mxArray *a) { return a && a->cls == mxDOUBLE_CLASS; }
bool mxIsInt32(const mxArray *a) { return a && a->cls == mxINT32_CLASS; }
bool mxIsUint64(const mxArray *a) { return a && a->cls == mxUINT64_CLASS; }
double *mxGetPr(const mxArray *a) { return (double *)a->data; }
void *mxGetData(const mxArray *a) { return a->data; }
double mxGetScalar(const mxArray *a) {
    if (!a || a->m * a->n == 0) mexErrMsgIdAndTxt("stub:scalar", "mxGetScalar of an empty array");
    switch (a->cls) {
        case mxDOUBLE_CLASS: return *(double *)a->data;
        case mxINT32_CLASS: return (double)*(int32_t *)a->data;
        case mxUINT64_CLASS: return (double)*(uint64_t *)a->data;
        default: return 0.0;
    }
}
mwSize mxGetNumberOfElements(const mxArray *a) { return a ? a->m * a->n : 0; }
mwSize mxGetM(const mxArray *a) { return a->m; }
mwSize mxGetN(const mxArray *a) { return a->n; }
int mxGetString(const mxArray *a, char *buf, mwSize buflen) {
    if (!a || a->cls != mxCHAR_CLASS || a->m * a->n + 1 > buflen) return 1;
    memcpy(buf, a->data, a->m * a->n); buf[a->m * a->n] = 0;
    return 0;
}
void *mxMalloc(size_t n) { return malloc(n ? n : 1); }
void mxFree(void *p) { free(p); }
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { (void)c; return mk(mxDOUBLE_CLASS, m, n); }
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c) { (void)c; return mk(cls, m, n); }
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_errmsg, sizeof(g_errmsg), fmt, ap);
    va_end(ap);
    snprintf(g_errid, sizeof(g_errid), "%s", id);
    longjmp(g_jmp, 1);
}

/* ---- builders / accessors for the Python side ---- */
mxArray *stub_double(mwSize m, mwSize n, const double *src) { mxArray *a = mk(mxDOUBLE_CLASS, m, n); if (src) memcpy(a->data, src, 8 * m * n); return a; }
mxArray *stub_int32(mwSize m, mwSize n, const int32_t *src) { mxArray *a = mk(mxINT32_CLASS, m, n); if (src) memcpy(a->data, src, 4 * m * n); return a; }
mxArray *stub_string(const char *s) { mxArray *a = mk(mxCHAR_CLASS, 1, strlen(s)); memcpy(a->data, s, strlen(s)); return a; }
mxArray *stub_struct(void) { mxArray *a = mk(mxSTRUCT_CLASS, 1, 1); return a; }
void stub_set_field(mxArray *s, const char *name, mxArray *v) { snprintf(s->names[s->nfields], 24, "%s", name); s->vals[s->nfields++] = v; }
int stub_class(const mxArray *a) { return (int)a->cls; }
mwSize stub_m(const mxArray *a) { return a->m; }
mwSize stub_n(const mxArray *a) { return a->n; }
void *stub_data(const mxArray *a) { return a->data; }
const char *stub_errid(void) { return g_errid; }
const char *stub_errmsg(void) { return g_errmsg; }
int stub_call(int nlhs, mxArray **plhs, int nrhs, mxArray **prhs) {
    g_errid[0] = g_errmsg[0] = 0;
    if (setjmp(g_jmp)) return 1;
    mexFunction(nlhs, plhs, nrhs, (const mxArray **)prhs);
    return 0;
}
