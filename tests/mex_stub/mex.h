/* mex.h -- TEST STUB of the MATLAB MEX API, just enough to compile mex/mpcgpu_mex.c without MATLAB and to drive its
 * mexFunction from tests/test_mex.py (through driver.c).  Column-major arrays, value semantics, errors by longjmp. */
#ifndef MPCGPU_TEST_MEX_H
#define MPCGPU_TEST_MEX_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
typedef size_t mwSize;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxUINT64_CLASS = 15, mxCHAR_CLASS = 4, mxSTRUCT_CLASS = 2 } mxClassID;
typedef enum { mxREAL = 0 } mxComplexity;
typedef struct mxArray_tag mxArray;
const mxArray *mxGetField(const mxArray *s, mwSize idx, const char *name);
bool mxIsStruct(const mxArray *a);
bool mxIsDouble(const mxArray *a);
bool mxIsInt32(const mxArray *a);
bool mxIsUint64(const mxArray *a);
double *mxGetPr(const mxArray *a);
void *mxGetData(const mxArray *a);
double mxGetScalar(const mxArray *a);
mwSize mxGetNumberOfElements(const mxArray *a);
mwSize mxGetM(const mxArray *a);
mwSize mxGetN(const mxArray *a);
int mxGetString(const mxArray *a, char *buf, mwSize buflen);
void *mxMalloc(size_t n);
void mxFree(void *p);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#endif
