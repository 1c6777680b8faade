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
    double red[64];                  // one slot per warp (CTAs of up to 1024 threads), twice over
    int flag;
};

// NT = threads of the calling CTA (all of them must call); fixed summation order
template <int NT = JAC_THREADS>
__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();                                   // protect red[] from the previous use
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < NT / 32; ++w) s += red[w];
    return s;
}

// In-place Jacobi eigen-decomposition of the symmetric n x n matrix A (column-major, leading dim lda,
// shared or global memory).  On exit diag(A) holds the eigenvalues (unsorted); if V != nullptr it must
// be initialised to the identity by the caller and ends up holding the eigenvectors in its columns.
// All NT threads of the CTA must call this.
template <int NT = JAC_THREADS>
__device__ void block_jacobi(double* A, int n, int lda, double* V, int ldv, JacobiShared* sh) {
    if (n <= 1) return;
    const int tid = threadIdx.x;
    const int ne = (n + 1) & ~1;              // padded to even; index n (if any) is a bye
    const int half = ne / 2;
    for (int sweep = 0; sweep < 40; ++sweep) {
        // convergence: off-diagonal mass against the Frobenius norm
        double off = 0.0, fro = 0.0;
        for (int e = tid; e < n * n; e += NT) {
            int i = e % n, j = e / n;
            double a = A[i + j * lda];
            fro += a * a;
            if (i != j) off += a * a;
        }
        off = block_sum<NT>(off, sh->red);
        fro = block_sum<NT>(fro, sh->red);
        if (off <= 4.930380657631324e-32 * fro * 1e-2 || fro == 0.0) break;   // off <= 0.1*eps*||A||_F
        for (int r = 0; r < ne - 1; ++r) {
            // rotation parameters for the half pairs of this round
            for (int k = tid; k < half; k += NT) {
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
            for (int e = tid; e < half * n; e += NT) {
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
            for (int e = tid; e < half * n; e += NT) {
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
            for (int k = tid; k < half; k += NT) {
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

// ---------------------------------------------------------------------------------------------------------------
// Eigenvalues only, LAPACK-style (tridiagonalise, then solve the tridiagonal problem), for the projections of the
// candidate-edge path where thousands of tiny problems sit on the critical path of one CTA each:
//   warp_tridiag      Householder tridiagonalisation (dsytd2, lower) by ONE warp, lanes own rows
//   sturm_multisect2  all eigenvalues of TWO tridiagonal matrices of the same order by parallel multisection on Sturm
//                     counts: every thread evaluates the count at one shift per round, every eigenvalue's bracket
//                     shrinks by (S + 1) per round, S = threads / n shifts per bracket
// Absolute accuracy ~ eps * ||A|| like dsyev (validated against numpy.linalg.eigvalsh on block tridiagonal matrices
// with zero / tiny couplings and repeated eigenvalues: 2.7e-15 * ||A|| worst case); eigenvalues come out ascending,
// i.e. already as sort(eig(.)) of functions/trace_fun_update.m:83-84.  The cyclic Jacobi iteration above needs
// ~8 sweeps of n/2 * (n-1) rotations with a barrier between rounds: 130 us per block-Lanczos step at n = 20 against
// ~15 us for this pair (profiles/r02n_*).
constexpr int MS_MAX_N = 64;

__device__ __forceinline__ double warp_sum_all(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// A: symmetric n x n, column-major, BOTH triangles stored (destroyed).  d[0..n), e[0..n-1): the tridiagonal matrix.
// v, w: n doubles of scratch each.  One warp, all 32 lanes must call.
__device__ void warp_tridiag(double* A, int n, int lda, double* d, double* e, double* v, double* w) {
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < n - 2; ++k) {
        const double alpha = A[(k + 1) + k * lda];
        double xn2 = 0.0;
        for (int i = k + 2 + lane; i < n; i += 32) {
            const double x = A[i + k * lda];
            xn2 += x * x;
        }
        xn2 = warp_sum_all(xn2);
        double tau = 0.0, beta = alpha, scale = 0.0;
        if (xn2 != 0.0) {                                   // dlarfg
            beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        if (lane == 0) {
            d[k] = A[k + k * lda];
            e[k] = beta;
        }
        if (tau != 0.0) {                                   // warp-uniform
            for (int i = k + 1 + lane; i < n; i += 32) v[i] = (i == k + 1) ? 1.0 : A[i + k * lda] * scale;
            __syncwarp();
            double dot = 0.0;
            for (int i = k + 1 + lane; i < n; i += 32) {    // p = tau * A22 v
                double acc = 0.0;
                for (int j = k + 1; j < n; ++j) acc += A[i + j * lda] * v[j];
                acc *= tau;
                w[i] = acc;
                dot += acc * v[i];
            }
            dot = warp_sum_all(dot);
            const double K = -0.5 * tau * dot;
            for (int i = k + 1 + lane; i < n; i += 32) w[i] += K * v[i];
            __syncwarp();
            for (int i = k + 1 + lane; i < n; i += 32) {    // A22 -= v w' + w v'
                const double vi = v[i], wi = w[i];
                for (int j = k + 1; j < n; ++j) A[i + j * lda] -= vi * w[j] + wi * v[j];
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        if (n >= 2) {
            d[n - 2] = A[(n - 2) + (n - 2) * lda];
            e[n - 2] = A[(n - 1) + (n - 2) * lda];
        }
        d[n - 1] = A[(n - 1) + (n - 1) * lda];
    }
    __syncwarp();
}

// number of eigenvalues of the tridiagonal (d, e2 = squared off-diagonals) below sigma: sign changes of the Sturm
// sequence p_i = (d_i - sigma) p_{i-1} - e2_{i-1} p_{i-2}, rescaled by exact powers of two; an exact zero takes the
// sign opposite to its predecessor (LAPACK dstebz: |q| < pivmin -> q = -pivmin)
__device__ __forceinline__ int sturm_count(const double* d, const double* e2, int n, double sigma) {
    double pm = 1.0, p = d[0] - sigma;
    if (p == 0.0) p = -0x1p-600;
    int sgn = p < 0.0 ? -1 : 1;
    int cnt = sgn < 0;
    for (int i = 1; i < n; ++i) {
        const double pn = fma(d[i] - sigma, p, -e2[i - 1] * pm);
        pm = p;
        p = pn;
        if ((i & 3) == 0) {                                 // |p| grows by < 2^60 per step for ||T|| < 2^29
            const double a = fmax(fabs(p), fabs(pm));
            if (a > 0x1p500) { p *= 0x1p-500; pm *= 0x1p-500; }
            else if (a < 0x1p-500) { p *= 0x1p500; pm *= 0x1p500; }
        }
        if (p == 0.0) p = (sgn > 0 ? -0x1p-600 : 0x1p-600) * fmax(fabs(pm), 0x1p-400);
        const int s2 = p < 0.0 ? -1 : 1;
        cnt += (s2 != sgn);
        sgn = s2;
    }
    return cnt;
}

// work[mtx] (mtx = 0, 1): 6 * MS_MAX_N doubles laid out as d | e | v | w | lo | hi, d / e filled by warp_tridiag.
// Threads [0, NT/2) serve matrix 0, the rest matrix 1; all NT threads must call.  out0 / out1: ascending eigenvalues.
template <int NT>
__device__ void sturm_multisect2(double (*work)[6 * MS_MAX_N], int n, double* out0, double* out1, int* cnts) {
    constexpr int HALF = NT / 2;
    const int tid = threadIdx.x, mtx = tid / HALF, t = tid % HALF;
    double* d = work[mtx];
    double* e = d + MS_MAX_N;
    double* gb = d + 2 * MS_MAX_N;           // Gershgorin bounds {gl, gu}
    double* lo = d + 4 * MS_MAX_N;
    double* hi = d + 5 * MS_MAX_N;
    if (t < 32) {
        double mn = 1e300, mx = -1e300;
        for (int i = t; i < n; i += 32) {
            const double r = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i + 1 < n ? fabs(e[i]) : 0.0);
            mn = fmin(mn, d[i] - r);
            mx = fmax(mx, d[i] + r);
        }
        for (int o = 16; o; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (t == 0) {
            const double tn = fmax(fmax(fabs(mn), fabs(mx)), 1e-300);
            const double pad = 4.0 * tn * 2.220446049250313e-16 * n + 1e-290;
            gb[0] = mn - pad;
            gb[1] = mx + pad;
        }
    }
    __syncthreads();
    if (t < n) {
        lo[t] = gb[0];
        hi[t] = gb[1];
        if (t + 1 < n) e[t] = e[t] * e[t];
    }
    int S = HALF / n;
    S = S < 1 ? 1 : (S > 16 ? 16 : S);
    int rounds = 1;                           // (S+1)^rounds >= 2^62
    {
        double f = 1.0;
        while (f < 0x1p62) { f *= (double)(S + 1); ++rounds; }
    }
    __syncthreads();
    const int m = t / S, sidx = t - m * S;
    const bool active = m < n;
    const double inv = 1.0 / (double)(S + 1);
    for (int r = 0; r < rounds; ++r) {
        if (active) {
            const double l = lo[m], h = hi[m];
            cnts[tid] = sturm_count(d, e, n, l + (h - l) * ((double)(sidx + 1) * inv));
        }
        __syncthreads();
        if (active && sidx == 0) {
            double l = lo[m], h = hi[m];
            const double l0 = l, w0 = h - l;
            for (int q = 0; q < S; ++q) {
                const double sg = l0 + w0 * ((double)(q + 1) * inv);
                if (cnts[tid + q] > m) { h = sg; break; }
                l = sg;
            }
            lo[m] = l;
            hi[m] = h;
        }
        __syncthreads();
    }
    if (t < n) (mtx == 0 ? out0 : out1)[t] = 0.5 * (lo[t] + hi[t]);
    __syncthreads();
}

// out[rank] = diag(A)[i], ascending (rank sort; ties broken by index).
template <int NT = JAC_THREADS>
__device__ void block_sorted_diag(const double* A, int n, int lda, double* out) {
    for (int i = threadIdx.x; i < n; i += NT) {
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
template <int NT = JAC_THREADS>
__device__ double block_trace_formula(int fun, const double* d1, const double* d2, int n, double* red) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += NT) {
        if (fun == KR_FUN_EXP) acc += exp(d1[i]) * (1.0 - exp(d2[i] - d1[i]));
        else acc += fun_eval(fun, d1[i]) - fun_eval(fun, d2[i]);
    }
    return block_sum<NT>(acc, red);
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
