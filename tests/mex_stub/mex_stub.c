/* mex_stub.c - see mex.h: a few dozen lines of mxArray bookkeeping, test infrastructure only. */
#include "mex.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct mxArray_tag {
    mxClassID cls;
    mwSize m, n;
    int sparse;
    void* data;       /* doubles, uint64, logical (as unsigned char) or chars */
    mwIndex *ir, *jc;
};

static mxArray* make(mxClassID cls, mwSize m, mwSize n, size_t elem) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->cls = cls; a->m = m; a->n = n;
    a->data = calloc((m * n) != 0 ? m * n : 1, elem);
    return a;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { (void)c; return make(mxDOUBLE_CLASS, m, n, sizeof(double)); }
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = make(mxDOUBLE_CLASS, 1, 1, sizeof(double)); *(double*)a->data = v; return a; }
mxArray* mxCreateLogicalScalar(int v) { mxArray* a = make(mxLOGICAL_CLASS, 1, 1, 1); *(unsigned char*)a->data = v != 0; return a; }
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c) { (void)c; return make(cls, m, n, 8); }
mxArray* mxCreateSparse(mwSize m, mwSize n, mwSize nzmax, mxComplexity c) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    (void)c;
    a->cls = mxDOUBLE_CLASS; a->m = m; a->n = n; a->sparse = 1;
    a->data = calloc(nzmax ? nzmax : 1, sizeof(double));
    a->ir = (mwIndex*)calloc(nzmax ? nzmax : 1, sizeof(mwIndex));
    a->jc = (mwIndex*)calloc(n + 1, sizeof(mwIndex));
    return a;
}
mxArray* mxCreateString(const char* s) {
    mxArray* a = make(mxCHAR_CLASS, 1, strlen(s), 1);
    memcpy(a->data, s, strlen(s));
    return a;
}
void mxDestroyArray(mxArray* a) { if (a) { free(a->data); free(a->ir); free(a->jc); free(a); } }
double* mxGetPr(const mxArray* a) { return (double*)a->data; }
void* mxGetData(const mxArray* a) { return a->data; }
double mxGetScalar(const mxArray* a) {
    if (a->cls == mxLOGICAL_CLASS) return *(unsigned char*)a->data;
    if (a->cls == mxUINT64_CLASS) return (double)*(uint64_t*)a->data;
    return *(double*)a->data;
}
mwSize mxGetM(const mxArray* a) { return a->m; }
mwSize mxGetN(const mxArray* a) { return a->n; }
mwSize mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
mwIndex* mxGetJc(const mxArray* a) { return a->jc; }
mwIndex* mxGetIr(const mxArray* a) { return a->ir; }
int mxIsSparse(const mxArray* a) { return a->sparse; }
int mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
int mxIsUint64(const mxArray* a) { return a->cls == mxUINT64_CLASS; }
int mxIsEmpty(const mxArray* a) { return a->m * a->n == 0; }
int mxGetString(const mxArray* a, char* buf, mwSize buflen) {
    mwSize len = a->m * a->n;
    if (a->cls != mxCHAR_CLASS || len + 1 > buflen) return 1;
    memcpy(buf, a->data, len);
    buf[len] = 0;
    return 0;
}
void* mxMalloc(size_t n) { return malloc(n ? n : 1); }
void mxFree(void* p) { free(p); }
/* Errors: a real MEX host unwinds to the interpreter.  The harness executable simply exits; a host that calls the
 * gateway through kr_stub_call (the interpreter bridge, oracle/mlab/mexbridge.py) gets the message back instead. */
static jmp_buf g_jmp;
static int g_armed = 0;
static char g_errmsg[1024];
static void raise_error(void) {
    if (g_armed) longjmp(g_jmp, 1);
    fprintf(stderr, "MEX error: %s\n", g_errmsg);
    exit(3);
}
void mexErrMsgTxt(const char* msg) { snprintf(g_errmsg, sizeof g_errmsg, "%s", msg); raise_error(); }
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap;
    (void)id;
    va_start(ap, fmt);
    vsnprintf(g_errmsg, sizeof g_errmsg, fmt, ap);
    va_end(ap);
    raise_error();
}
int kr_stub_call(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[], char* err, size_t errlen) {
    g_armed = 1;
    if (setjmp(g_jmp)) {
        g_armed = 0;
        snprintf(err, errlen, "%s", g_errmsg);
        return 1;
    }
    mexFunction(nlhs, plhs, nrhs, prhs);
    g_armed = 0;
    return 0;
}
int mxGetClassID(const mxArray* a) { return (int)a->cls; }
int mxIsLogical(const mxArray* a) { return a->cls == mxLOGICAL_CLASS; }
int mxIsChar(const mxArray* a) { return a->cls == mxCHAR_CLASS; }
int mexAtExit(void (*fn)(void)) { return atexit(fn); }
