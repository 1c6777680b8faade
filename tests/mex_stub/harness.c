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
    {   /* [Xm,iter,lucky,Um] = kr_mex('fun_update', A, U, B, 'exp', tol, it, 1): U = first nE... unit columns e_{E(:,1)} */
        mxArray* lhs[4] = {NULL, NULL, NULL, NULL};
        const mxArray* rhs[8];
        size_t rk = nE < 4 ? nE : 4, q;
        mxArray* U = mxCreateDoubleMatrix(n, rk, mxREAL);
        mxArray* B = mxCreateDoubleMatrix(rk, rk, mxREAL);
        double v[2];
        for (q = 0; q < rk; ++q) mxGetPr(U)[q * n + (size_t)mxGetPr(E)[q] - 1] = 1.0;      /* distinct nodes E(q,1) */
        for (q = 0; q + 1 < rk; ++q) { mxGetPr(B)[q + (q + 1) * rk] = 0.1 * (q + 1); mxGetPr(B)[(q + 1) + q * rk] = 0.1 * (q + 1); }
        rhs[0] = mxCreateString("fun_update"); rhs[1] = A; rhs[2] = U; rhs[3] = B; rhs[4] = mxCreateString(fun);
        rhs[5] = mxCreateDoubleScalar(tol); rhs[6] = mxCreateDoubleScalar(itmax); rhs[7] = mxCreateDoubleScalar(1.0);
        mexFunction(4, lhs, 8, rhs);
        put(out, "fun_update_Xm", mxGetPr(lhs[0]), mxGetM(lhs[0]) * mxGetN(lhs[0]));
        v[0] = mxGetScalar(lhs[1]); v[1] = mxGetScalar(lhs[2]);
        put(out, "fun_update_info", v, 2);
        put(out, "fun_update_Um", mxGetPr(lhs[3]), mxGetM(lhs[3]) * mxGetN(lhs[3]));
    }
    {   /* [f,gr] = kr_mex('fun_and_grad', X, A, Omega, 'sinh', 'cosh', dfA, tol, it) and Hes = kr_mex('hessian', A, Omega, 'exp', tol, it) */
        mxArray* lhs[2] = {NULL, NULL};
        mxArray* lh[1] = {NULL};
        const mxArray* rhs[9];
        const mxArray* rh[6];
        mxArray* Xw = mxCreateDoubleMatrix(nE, 1, mxREAL);
        mxArray* dfA = mxCreateDoubleMatrix(nE, 1, mxREAL);
        double f;
        for (i = 0; i < nE; ++i) { mxGetPr(Xw)[i] = 0.01 * (double)(i + 1); mxGetPr(dfA)[i] = 0.5; }
        rhs[0] = mxCreateString("fun_and_grad"); rhs[1] = Xw; rhs[2] = A; rhs[3] = E; rhs[4] = mxCreateString("sinh");
        rhs[5] = mxCreateString("cosh"); rhs[6] = dfA; rhs[7] = mxCreateDoubleScalar(1e-8); rhs[8] = mxCreateDoubleScalar(itmax);
        mexFunction(2, lhs, 9, rhs);
        f = mxGetScalar(lhs[0]);
        put(out, "fg_f", &f, 1);
        put(out, "fg_gr", mxGetPr(lhs[1]), nE);
        rh[0] = mxCreateString("hessian"); rh[1] = A; rh[2] = E; rh[3] = mxCreateString(fun); rh[4] = mxCreateDoubleScalar(1e-10);
        rh[5] = mxCreateDoubleScalar(itmax);
        mexFunction(1, lh, 6, rh);
        put(out, "hessian", mxGetPr(lh[0]), nE * nE);
    }
    {   /* [tr,res,it] = kr_mex('mc_trace', A, 0, 1e-3, 30, probes): probes = X's columns repeated to n x 20 */
        mxArray* lhs[3] = {NULL, NULL, NULL};
        const mxArray* rhs[6];
        mxArray* P = mxCreateDoubleMatrix(n, 20, mxREAL);
        double v[3];
        size_t c;
        for (c = 0; c < 20; ++c) for (i = 0; i < n; ++i) mxGetPr(P)[c * n + i] = mxGetPr(X)[(c % k) * n + i] * ((i + c) % 3 == 0 ? -1.0 : 1.0);
        rhs[0] = mxCreateString("mc_trace"); rhs[1] = A; rhs[2] = mxCreateDoubleScalar(0.0); rhs[3] = mxCreateDoubleScalar(1e-3);
        rhs[4] = mxCreateDoubleScalar(30.0); rhs[5] = P;
        mexFunction(3, lhs, 6, rhs);
        v[0] = mxGetScalar(lhs[0]); v[1] = mxGetScalar(lhs[1]); v[2] = mxGetScalar(lhs[2]);
        put(out, "mc_trace", v, 3);
        put(out, "mc_probes", mxGetPr(P), n * 20);
    }
    {   /* the device-matrix cache must key on CONTENT: the same mxArray (same data pointer, same pattern, same nnz)
         * with different values - what MATLAB produces when a freed `A + XX + XX'` address is re-used */
        mxArray* lhs[1] = {NULL};
        const mxArray* rhs[3];
        for (i = 0; i < nnz; ++i) mxGetPr(A)[i] *= 2.0;
        rhs[0] = mxCreateString("spmm"); rhs[1] = A; rhs[2] = X;
        mexFunction(1, lhs, 3, rhs);
        put(out, "spmm_doubled_values", mxGetPr(lhs[0]), n * k);
        for (i = 0; i < nnz; ++i) mxGetPr(A)[i] *= 0.5;
    }
    {   /* h = kr_mex('matrix_create', A); kr_mex('matrix_set_edges', h, i, j, 1); scores through the handle; free twice */
        mxArray* lh[1] = {NULL};
        mxArray* lhs[3] = {NULL, NULL, NULL};
        const mxArray* r1[2];
        const mxArray* r2[5];
        const mxArray* rhs[8];
        const mxArray* r3[2];
        r1[0] = mxCreateString("matrix_create"); r1[1] = A;
        mexFunction(1, lh, 2, r1);
        r2[0] = mxCreateString("matrix_set_edges"); r2[1] = lh[0]; r2[2] = mxCreateDoubleScalar(mxGetPr(E)[0]);
        r2[3] = mxCreateDoubleScalar(mxGetPr(E)[nE]); r2[4] = mxCreateDoubleScalar(1.0);
        mexFunction(0, lhs, 5, r2);                                  /* insert the first candidate edge on the device copy */
        rhs[0] = mxCreateString("trace_fun_update_edges"); rhs[1] = lh[0]; rhs[2] = E; rhs[3] = mxCreateDoubleScalar(b);
        rhs[4] = mxCreateDoubleScalar(tol); rhs[5] = mxCreateDoubleScalar(itmax); rhs[6] = mxCreateString(fun);
        rhs[7] = mxCreateDoubleScalar(b);
        mexFunction(3, lhs, 8, rhs);
        put(out, "edges_x_after_insert", mxGetPr(lhs[0]), nE);
        {   /* [best,val] = kr_mex('greedy_round', h, E, b, tol, it, fun, b_self, mode, screen) on the edited matrix */
            mxArray* gl[3] = {NULL, NULL, NULL};
            const mxArray* gr[10];
            double gv[3];
            gr[0] = mxCreateString("greedy_round"); gr[1] = lh[0]; gr[2] = E; gr[3] = mxCreateDoubleScalar(b);
            gr[4] = mxCreateDoubleScalar(tol); gr[5] = mxCreateDoubleScalar(itmax); gr[6] = mxCreateString(fun);
            gr[7] = mxCreateDoubleScalar(b); gr[8] = mxCreateDoubleScalar(1.0); gr[9] = mxCreateDoubleScalar(0.0);
            mexFunction(3, gl, 10, gr);
            gv[0] = mxGetScalar(gl[0]); gv[1] = mxGetScalar(gl[1]); gv[2] = mxGetScalar(gl[2]);
            put(out, "greedy_round_make", gv, 3);
        }
        r3[0] = mxCreateString("matrix_free"); r3[1] = lh[0];
        mexFunction(0, lhs, 2, r3);
        mexFunction(0, lhs, 2, r3);                                  /* idempotent */
    }
    fclose(out);
    fclose(in);
    return 0;
}
