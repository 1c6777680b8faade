// smalldense.cuh - batched small symmetric eigenproblems in shared memory (one CTA per matrix).
//
// The Krylov projections of the path are tiny (2j x 2j with j ~ 3..11 for candidate edges, j x j
// for single-vector spaces; <= 200 x 200 at it = 100), and there are thousands of them per step:
// the reference calls eig() / expm() on each (functions/trace_fun_update.m:83-84,
// functions/function_multiple_entries.m:118).  Here every matrix gets one CTA running a cyclic
// two-sided Jacobi iteration with the round-robin parallel ordering; eigenvalues come out with
// absolute accuracy ~eps*||A||_F like LAPACK's dsyev.
#pragma once
#include "common.cuh"

namespace kr {

constexpr int JAC_THREADS = 128;
constexpr int JAC_MAX_N = 256;

struct JacobiShared {
    double c[JAC_MAX_N / 2 + 1];
    double s[JAC_MAX_N / 2 + 1];
    int p[JAC_MAX_N / 2 + 1];
    int q[JAC_MAX_N / 2 + 1];
    double red[JAC_THREADS / 32 * 2];
    int flag;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();                                   // protect red[] from the previous use
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < JAC_THREADS / 32; ++w) s += red[w];
    return s;
}

// In-place Jacobi eigen-decomposition of the symmetric n x n matrix A (column-major, leading dim lda,
// shared or global memory).  On exit diag(A) holds the eigenvalues (unsorted); if V != nullptr it must
// be initialised to the identity by the caller and ends up holding the eigenvectors in its columns.
// All JAC_THREADS threads of the CTA must call this.
__device__ void block_jacobi(double* A, int n, int lda, double* V, int ldv, JacobiShared* sh) {
    if (n <= 1) return;
    const int tid = threadIdx.x;
    const int ne = (n + 1) & ~1;              // padded to even; index n (if any) is a bye
    const int half = ne / 2;
    for (int sweep = 0; sweep < 40; ++sweep) {
        // convergence: off-diagonal mass against the Frobenius norm
        double off = 0.0, fro = 0.0;
        for (int e = tid; e < n * n; e += JAC_THREADS) {
            int i = e % n, j = e / n;
            double a = A[i + j * lda];
            fro += a * a;
            if (i != j) off += a * a;
        }
        off = block_sum(off, sh->red);
        fro = block_sum(fro, sh->red);
        if (off <= 4.930380657631324e-32 * fro * 1e-2 || fro == 0.0) break;   // off <= 0.1*eps*||A||_F
        for (int r = 0; r < ne - 1; ++r) {
            // rotation parameters for the half pairs of this round
            for (int k = tid; k < half; k += JAC_THREADS) {
                int p, q;
                if (k == 0) { p = ne - 1; q = r; }
                else { p = (r + k) % (ne - 1); q = (r - k + (ne - 1)) % (ne - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                if (q < n) {
                    double apq = A[p + q * lda];
                    if (apq != 0.0) {
                        double app = A[p + p * lda], aqq = A[q + q * lda];
                        double theta = (aqq - app) / (2.0 * apq);
                        double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                    }
                } else {
                    p = -1;                    // bye
                }
                sh->c[k] = c; sh->s[k] = s; sh->p[k] = p; sh->q[k] = q;
            }
            __syncthreads();
            // columns: A <- A J (and V <- V J)
            for (int e = tid; e < half * n; e += JAC_THREADS) {
                int k = e / n, i = e % n;
                int p = sh->p[k];
                if (p < 0) continue;
                int q = sh->q[k];
                double c = sh->c[k], s = sh->s[k];
                if (s == 0.0) continue;
                double aip = A[i + p * lda], aiq = A[i + q * lda];
                A[i + p * lda] = c * aip - s * aiq;
                A[i + q * lda] = s * aip + c * aiq;
                if (V) {
                    double vip = V[i + p * ldv], viq = V[i + q * ldv];
                    V[i + p * ldv] = c * vip - s * viq;
                    V[i + q * ldv] = s * vip + c * viq;
                }
            }
            __syncthreads();
            // rows: A <- J' A
            for (int e = tid; e < half * n; e += JAC_THREADS) {
                int k = e / n, j = e % n;
                int p = sh->p[k];
                if (p < 0) continue;
                int q = sh->q[k];
                double c = sh->c[k], s = sh->s[k];
                if (s == 0.0) continue;
                double apj = A[p + j * lda], aqj = A[q + j * lda];
                A[p + j * lda] = c * apj - s * aqj;
                A[q + j * lda] = s * apj + c * aqj;
            }
            __syncthreads();
            for (int k = tid; k < half; k += JAC_THREADS) {
                int p = sh->p[k];
                if (p >= 0 && sh->s[k] != 0.0) {
                    int q = sh->q[k];
                    A[p + q * lda] = 0.0;
                    A[q + p * lda] = 0.0;
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
}

// out[rank] = diag(A)[i], ascending (rank sort; ties broken by index).
__device__ void block_sorted_diag(const double* A, int n, int lda, double* out) {
    for (int i = threadIdx.x; i < n; i += JAC_THREADS) {
        double di = A[i + i * lda];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            double dj = A[j + j * lda];
            rank += (dj < di) || (dj == di && j < i);
        }
        out[rank] = di;
    }
    __syncthreads();
}

__device__ __forceinline__ double fun_eval(int fun, double x) {
    return fun == KR_FUN_EXP ? exp(x) : fun == KR_FUN_SINH ? sinh(x) : cosh(x);
}

// trace formula of functions/trace_fun_update.m:85-89 on two sorted spectra (thread 0 result valid
// in every thread after the reduction).
__device__ double block_trace_formula(int fun, const double* d1, const double* d2, int n, double* red) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += JAC_THREADS) {
        if (fun == KR_FUN_EXP) acc += exp(d1[i]) * (1.0 - exp(d2[i] - d1[i]));
        else acc += fun_eval(fun, d1[i]) - fun_eval(fun, d2[i]);
    }
    return block_sum(acc, red);
}

// ------------------------------------------------------------------ generic batched kernels
// evals[b][0..n) = sorted eigenvalues of mats[b] (n x n column-major, contiguous, symmetric).
// Dynamic shared memory: n * (n|1) doubles (the launcher falls back to a global scratch copy when
// that exceeds the opt-in limit).
__global__ void __launch_bounds__(JAC_THREADS)
eigvals_batched_kernel(const double* __restrict__ mats, int n, double* __restrict__ evals,
                       double* __restrict__ gscratch) {
    extern __shared__ double dyn[];
    __shared__ JacobiShared sh;
    const int lda = n | 1;
    double* A = gscratch ? gscratch + (size_t)blockIdx.x * n * lda : dyn;
    const double* M = mats + (size_t)blockIdx.x * n * n;
    for (int e = threadIdx.x; e < n * n; e += JAC_THREADS) {
        int i = e % n, j = e / n;
        A[i + j * lda] = 0.5 * (M[i + j * n] + M[j + i * n]);
    }
    __syncthreads();
    block_jacobi(A, n, lda, nullptr, 0, &sh);
    block_sorted_diag(A, n, lda, evals + (size_t)blockIdx.x * n);
}

// F[b] = V f(D) V' for symmetric mats[b]; also evals.  smem: 2 * n * (n|1) doubles.
__global__ void __launch_bounds__(JAC_THREADS)
symfun_batched_kernel(const double* __restrict__ mats, int n, int fun, double* __restrict__ F,
                      double* __restrict__ gscratch) {
    extern __shared__ double dyn[];
    __shared__ JacobiShared sh;
    const int lda = n | 1;
    double* A = gscratch ? gscratch + (size_t)blockIdx.x * 2 * n * lda : dyn;
    double* V = A + n * lda;
    const double* M = mats + (size_t)blockIdx.x * n * n;
    for (int e = threadIdx.x; e < n * n; e += JAC_THREADS) {
        int i = e % n, j = e / n;
        A[i + j * lda] = 0.5 * (M[i + j * n] + M[j + i * n]);
        V[i + j * lda] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    block_jacobi(A, n, lda, V, lda, &sh);
    // f(lambda) into the diagonal slot
    for (int i = threadIdx.x; i < n; i += JAC_THREADS) A[i + i * lda] = fun_eval(fun, A[i + i * lda]);
    __syncthreads();
    double* out = F + (size_t)blockIdx.x * n * n;
    for (int e = threadIdx.x; e < n * n; e += JAC_THREADS) {
        int i = e % n, j = e / n;
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += V[i + k * lda] * A[k + k * lda] * V[j + k * lda];
        out[i + j * n] = s;
    }
}

inline size_t jacobi_smem_bytes(int n, bool vectors) {
    return (size_t)n * (n | 1) * sizeof(double) * (vectors ? 2 : 1);
}
constexpr size_t JAC_SMEM_LIMIT = 200 * 1024;

}  // namespace kr
