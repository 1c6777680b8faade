/*
 * mex.h - minimal stand-in for the MATLAB / Octave MEX API, TEST INFRASTRUCTURE ONLY.
 *
 * The build image has neither MATLAB nor Octave, so the gateway
 * krylov_robustness_b200/mex/kr_mex.c cannot be linked against a real interpreter here.  This header and
 * mex_stub.c implement just the subset of the documented mx* / mex* API the gateway uses (dense real
 * double arrays, real sparse double arrays in CSC form, char row vectors, logical and uint64 scalars), so that
 * the gateway is compiled by the CPU test tier and EXECUTED by the GPU test tier (tests/test_mex_gateway.py).
 */
#ifndef KR_TEST_MEX_H
#define KR_TEST_MEX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxLOGICAL_CLASS = 3, mxCHAR_CLASS = 4, mxUINT64_CLASS = 13 } mxClassID;
typedef struct mxArray_tag mxArray;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateLogicalScalar(int v);
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray* mxCreateSparse(mwSize m, mwSize n, mwSize nzmax, mxComplexity c);
mxArray* mxCreateString(const char* s);
void mxDestroyArray(mxArray* a);
double* mxGetPr(const mxArray* a);
void* mxGetData(const mxArray* a);
double mxGetScalar(const mxArray* a);
mwSize mxGetM(const mxArray* a);
mwSize mxGetN(const mxArray* a);
mwSize mxGetNumberOfElements(const mxArray* a);
mwIndex* mxGetJc(const mxArray* a);
mwIndex* mxGetIr(const mxArray* a);
int mxIsSparse(const mxArray* a);
int mxIsDouble(const mxArray* a);
int mxIsUint64(const mxArray* a);
int mxIsEmpty(const mxArray* a);
int mxGetString(const mxArray* a, char* buf, mwSize buflen);
void* mxMalloc(size_t n);
void mxFree(void* p);
void mexErrMsgTxt(const char* msg);
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
int mexAtExit(void (*fn)(void));
int mxGetClassID(const mxArray* a);
int mxIsLogical(const mxArray* a);
int mxIsChar(const mxArray* a);
/* not MEX API: entry for a host that wants gateway errors back as a message (see mex_stub.c) */
int kr_stub_call(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[], char* err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
