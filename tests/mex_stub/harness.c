/*
 * harness.c - drives krylov_robustness_b200/mex/kr_mex.c through the stub MEX runtime (mex.h / mex_stub.c):
 * the calls a MATLAB wrapper in krylov_robustness_b200/matlab/ would make, with MATLAB-shaped arguments
 * (CSC sparse A, column-major doubles, 1-based index lists as doubles, func2str strings).
 *
 *   harness <input.txt> <output.txt>
 * input : n nnz | jc[n+1] | ir[nnz] | pr[nnz] | k | X[n*k] | nE | E[nE*2] (column-major) | b tol it fun
 * output: one "name count v0 v1 ..." line per result.
 */
#include <stdio.h>
#include <stdlib.h>
#include "mex.h"

static void put(FILE* f, const char* name, const double* v, size_t cnt) {
    size_t i;
    fprintf(f, "%s %zu", name, cnt);
    for (i = 0; i < cnt; ++i) fprintf(f, " %.17g", v[i]);
    fprintf(f, "\n");
}

int main(int argc, char** argv) {
    FILE *in, *out;
    size_t n, nnz, k, nE, i;
    double b, tol, itmax;
    char fun[16];
    mxArray *A, *X, *E;
    if (argc < 3 || !(in = fopen(argv[1], "r")) || !(out = fopen(argv[2], "w"))) return 2;
    if (fscanf(in, "%zu %zu", &n, &nnz) != 2) return 2;
    A = mxCreateSparse(n, n, nnz, mxREAL);
    for (i = 0; i <= n; ++i) if (fscanf(in, "%zu", &mxGetJc(A)[i]) != 1) return 2;
    for (i = 0; i < nnz; ++i) if (fscanf(in, "%zu", &mxGetIr(A)[i]) != 1) return 2;
    for (i = 0; i < nnz; ++i) if (fscanf(in, "%lf", &mxGetPr(A)[i]) != 1) return 2;
    if (fscanf(in, "%zu", &k) != 1) return 2;
    X = mxCreateDoubleMatrix(n, k, mxREAL);
    for (i = 0; i < n * k; ++i) if (fscanf(in, "%lf", &mxGetPr(X)[i]) != 1) return 2;
    if (fscanf(in, "%zu", &nE) != 1) return 2;
    E = mxCreateDoubleMatrix(nE, 2, mxREAL);
    for (i = 0; i < nE * 2; ++i) if (fscanf(in, "%lf", &mxGetPr(E)[i]) != 1) return 2;
    if (fscanf(in, "%lf %lf %lf %15s", &b, &tol, &itmax, fun) != 4) return 2;

    {   /* Y = kr_mex('spmm', A, X) */
        mxArray* lhs[1] = {NULL};
        const mxArray* rhs[3];
        rhs[0] = mxCreateString("spmm"); rhs[1] = A; rhs[2] = X;
        mexFunction(1, lhs, 3, rhs);
        put(out, "spmm", mxGetPr(lhs[0]), n * k);
        mxDestroyArray(lhs[0]);
    }
    {   /* [Xm,iter,lucky] = kr_mex('trace_fun_update_edges', A, E, b, tol, it, 'exp') */
        mxArray* lhs[3] = {NULL, NULL, NULL};
        const mxArray* rhs[7];
        rhs[0] = mxCreateString("trace_fun_update_edges"); rhs[1] = A; rhs[2] = E; rhs[3] = mxCreateDoubleScalar(b);
        rhs[4] = mxCreateDoubleScalar(tol); rhs[5] = mxCreateDoubleScalar(itmax); rhs[6] = mxCreateString(fun);
        mexFunction(3, lhs, 7, rhs);
        put(out, "edges_x", mxGetPr(lhs[0]), nE);
        put(out, "edges_iter", mxGetPr(lhs[1]), nE);
        put(out, "edges_lucky", mxGetPr(lhs[2]), nE);
    }
    {   /* [e,cnt] = kr_mex('normest', A, 1e-2) */
        mxArray* lhs[2] = {NULL, NULL};
        const mxArray* rhs[3];
        double v[2];
        rhs[0] = mxCreateString("normest"); rhs[1] = A; rhs[2] = mxCreateDoubleScalar(1e-2);
        mexFunction(2, lhs, 3, rhs);
        v[0] = mxGetScalar(lhs[0]); v[1] = mxGetScalar(lhs[1]);
        put(out, "normest", v, 2);
    }
    {   /* [X,iter] = kr_mex('function_multiple_entries', A, omega, 'exp', tol, it) */
        mxArray* lhs[2] = {NULL, NULL};
        const mxArray* rhs[6];
        double it;
        rhs[0] = mxCreateString("function_multiple_entries"); rhs[1] = A; rhs[2] = E; rhs[3] = mxCreateString(fun);
        rhs[4] = mxCreateDoubleScalar(tol); rhs[5] = mxCreateDoubleScalar(itmax);
        mexFunction(2, lhs, 6, rhs);
        put(out, "entries", mxGetPr(lhs[0]), nE);
        it = mxGetScalar(lhs[1]);
        put(out, "entries_iter", &it, 1);
    }
    {   /* [f,s,m,mv,mvd,unA] = kr_mex('expmv', t, A, b, [], shift, full_term) */
        mxArray* lhs[6] = {NULL, NULL, NULL, NULL, NULL, NULL};
        const mxArray* rhs[7];
        double v[5];
        rhs[0] = mxCreateString("expmv"); rhs[1] = mxCreateDoubleScalar(1.0); rhs[2] = A; rhs[3] = X;
        rhs[4] = mxCreateDoubleMatrix(0, 0, mxREAL); rhs[5] = mxCreateDoubleScalar(1.0); rhs[6] = mxCreateDoubleScalar(0.0);
        mexFunction(6, lhs, 7, rhs);
        put(out, "expmv_f", mxGetPr(lhs[0]), n * k);
        for (i = 0; i < 5; ++i) v[i] = mxGetScalar(lhs[i + 1]);
        put(out, "expmv_info", v, 5);
    }
    {   /* [V,H,K,last,lucky,handle] = kr_mex('krylov_start', A, b, arnoldi); extend once; free */
        mxArray* lhs[6] = {NULL, NULL, NULL, NULL, NULL, NULL};
        mxArray* lhs2[6] = {NULL, NULL, NULL, NULL, NULL, NULL};
        const mxArray* rhs[4];
        const mxArray* rhs2[3];
        const mxArray* rhs3[2];
        rhs[0] = mxCreateString("krylov_start"); rhs[1] = A; rhs[2] = X; rhs[3] = mxCreateDoubleScalar(0.0);
        mexFunction(6, lhs, 4, rhs);
        rhs2[0] = mxCreateString("krylov_extend"); rhs2[1] = lhs[5]; rhs2[2] = mxCreateDoubleScalar((double)n);
        mexFunction(6, lhs2, 3, rhs2);
        put(out, "lanczos_H", mxGetPr(lhs2[1]), mxGetM(lhs2[1]) * mxGetN(lhs2[1]));
        {
            double d[2];
            d[0] = (double)mxGetM(lhs2[0]); d[1] = (double)mxGetN(lhs2[0]);
            put(out, "lanczos_Vdims", d, 2);
        }
        rhs3[0] = mxCreateString("krylov_free"); rhs3[1] = lhs2[5];
        mexFunction(0, lhs, 2, rhs3);
    }
    fclose(out);
    fclose(in);
    return 0;
}
