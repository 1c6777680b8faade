// smallgemm.cuh - dense arithmetic on the PROJECTED matrices of the wide-block path (size j*k <= ~1000),
// entirely on the device: a batched column-major fp64 GEMM on the FP64 tensor cores (DMMA m8n8k4) and, built
// on it, f(S) for symmetric S with f = exp / sinh / cosh by scaling and squaring of a degree-18 Taylor
// polynomial (Paterson-Stockmeyer, 7 products + s squarings).  This is what the reference does with expm
// (functions/fun_update.m:43-59: exp -> expm(M); sinh/cosh -> (expm(M) -+ expm(-M))/2); it replaces round 1's
// cuSOLVER syevd + cuBLAS dgemm.  The number of squarings s is decided ON THE DEVICE from ||S||_1 (no host
// round trip): the host enqueues s_max products (s_max from a rigorous host-side bound), the ones beyond s
// degenerate to copies.
#pragma once
#include "tsdense.cuh"

namespace kr {

// C[b] = alpha * A[b] * B[b] + beta * C[b]; column-major, batch strides in elements.  64 x 64 tile per CTA,
// 8 warps as 2 (m) x 4 (n), warp tile 32 x 16, K slab 16.  If sq_count != nullptr and sq_index >= *sq_count the
// product is skipped and C = A (a squaring that is not needed).
constexpr int SG_T = 64, SG_K = 32;
constexpr int SG_AS = SG_T + 4;     // As[k][m]
constexpr int SG_BS = SG_K + 4;     // Bs[n][k]

__global__ void __launch_bounds__(256)
sgemm_kernel(int m, int n, int k, double alpha, const double* __restrict__ A, int lda, int64_t sa,
             const double* __restrict__ B, int ldb, int64_t sb, double beta, double* __restrict__ C, int ldc, int64_t sc,
             const int* __restrict__ sq_count, int sq_index) {
    __shared__ __align__(16) double As[SG_K * SG_AS];
    __shared__ __align__(16) double Bs[SG_T * SG_BS];
    A += blockIdx.z * sa;
    B += blockIdx.z * sb;
    C += blockIdx.z * sc;
    const int m0 = blockIdx.x * SG_T, n0 = blockIdx.y * SG_T;
    if (sq_count && sq_index >= *sq_count) {
        for (int e = threadIdx.x; e < SG_T * SG_T; e += 256) {
            const int i = m0 + e % SG_T, j = n0 + e / SG_T;
            if (i < m && j < n) C[i + (int64_t)j * ldc] = A[i + (int64_t)j * lda];
        }
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kk = lane & 3, idx = lane >> 2;
    const int wm = warp & 1, wn = warp >> 1;          // warp tile rows [32 wm, +32), cols [16 wn, +16)
    double acc[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int k0 = 0; k0 < k; k0 += SG_K) {
        __syncthreads();
        for (int e = threadIdx.x; e < SG_K * SG_T; e += 256) {
            const int i = e % SG_T, kq = e / SG_T;                    // A(m0+i, k0+kq): coalesced in i
            As[kq * SG_AS + i] = (m0 + i < m && k0 + kq < k) ? A[(m0 + i) + (int64_t)(k0 + kq) * lda] : 0.0;
        }
        for (int e = threadIdx.x; e < SG_K * SG_T; e += 256) {
            const int kq = e % SG_K, j = e / SG_K;                    // B(k0+kq, n0+j): coalesced in kq
            Bs[j * SG_BS + kq] = (n0 + j < n && k0 + kq < k) ? B[(k0 + kq) + (int64_t)(n0 + j) * ldb] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < SG_K; ks += 4) {
            double a[4], b[2];
#pragma unroll
            for (int t = 0; t < 4; ++t) a[t] = As[(ks + kk) * SG_AS + wm * 32 + t * 8 + idx];
#pragma unroll
            for (int t = 0; t < 2; ++t) b[t] = Bs[(wn * 16 + t * 8 + idx) * SG_BS + ks + kk];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int i = m0 + wm * 32 + mt * 8 + idx;
            const int j = n0 + wn * 16 + nt * 8 + 2 * kk;
            if (i < m) {
                if (j < n) {
                    double* c = C + i + (int64_t)j * ldc;
                    *c = alpha * acc[mt][nt][0] + (beta != 0.0 ? beta * *c : 0.0);
                }
                if (j + 1 < n) {
                    double* c = C + i + (int64_t)(j + 1) * ldc;
                    *c = alpha * acc[mt][nt][1] + (beta != 0.0 ? beta * *c : 0.0);
                }
            }
        }
}

inline void sgemm(kr_ctx* ctx, int m, int n, int k, double alpha, const double* A, int lda, int64_t sa, const double* B,
                  int ldb, int64_t sb, double beta, double* C, int ldc, int64_t sc, int batch,
                  const int* sq_count = nullptr, int sq_index = 0) {
    if (m <= 0 || n <= 0 || batch <= 0) return;
    dim3 grid((unsigned)ceil_div(m, SG_T), (unsigned)ceil_div(n, SG_T), (unsigned)batch);
    KR_LAUNCH(ctx, sgemm_kernel, grid, 256, 0, m, n, k, alpha, A, lda, sa, B, ldb, sb, beta, C, ldc, sc, sq_count, sq_index);
}

// ------------------------------------------------------------------------------------ f(S)
struct ExpmCtl {
    int s;              // squarings actually needed
    int clamped;        // sticky: some evaluation needed more squarings than were enqueued (the caller re-runs it)
    double scale;       // 2^-s
    double norm1;       // max over the batch of ||S||_1
};

// ||S||_1 = max row abs-sum (the matrices are symmetric) over the batch: row sums are coalesced; the maximum of
// non-negative doubles via atomicMax on the bit patterns is order independent (deterministic).
__global__ void __launch_bounds__(256)
expm_rowmax_kernel(const double* __restrict__ S, int nn, unsigned long long* __restrict__ bits) {
    const double* M = S + (int64_t)blockIdx.y * nn * nn;
    const int i = blockIdx.x * 256 + threadIdx.x;
    double s = 0.0;
    if (i < nn)
        for (int c = 0; c < nn; ++c) s += fabs(M[i + (int64_t)c * nn]);
    for (int off = 16; off; off >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, off));
    if ((threadIdx.x & 31) == 0) atomicMax(bits, (unsigned long long)__double_as_longlong(s));
}
__global__ void expm_decide_kernel(unsigned long long* __restrict__ bits, double theta, int s_max, ExpmCtl* ctl) {
    const double nrm = __longlong_as_double((long long)*bits);
    *bits = 0ull;                                  // ready for the next call
    int s = 0;
    if (nrm > theta) s = (int)ceil(log2(nrm / theta));
    if (s > s_max) { s = s_max; ctl->clamped = 1; }   // only possible with a hint-based s_max (ExpmWork::hint)
    ctl->s = s;
    ctl->scale = ldexp(1.0, -s);
    ctl->norm1 = nrm;
}

// X = sign * S * 2^-s for every matrix of the batch
__global__ void expm_scale_kernel(const double* __restrict__ S, double* __restrict__ X, int64_t total, double sign,
                                  const ExpmCtl* __restrict__ ctl) {
    const double f = sign * ctl->scale;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        X[e] = f * S[e];
}

// Taylor coefficients 1/i!, i = 0..18
__constant__ double KR_INV_FACT[19] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                                       1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0,
                                       1.0 / 87178291200.0, 1.0 / 1307674368000.0, 1.0 / 20922789888000.0,
                                       1.0 / 355687428096000.0, 1.0 / 6402373705728000.0};

// Paterson-Stockmeyer blocks of p(X) = sum_{i<=18} X^i / i! = B0 + X^6 (B1 + X^6 (B2 + c18 X^6)):
//   P2 = B2 + c18 X^6,  B1,  B0   with Bq = sum_{i<6} c_{6q+i} X^i.  pw: X^1..X^6 stored consecutively.
__global__ void expm_blocks_kernel(const double* __restrict__ pw, int nn, int64_t mat, int64_t total,
                                   double* __restrict__ P2, double* __restrict__ B1, double* __restrict__ B0) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t within = e % mat;
        const int i = (int)(within % nn), j = (int)(within / nn);
        const double id = (i == j) ? 1.0 : 0.0;
        double x[7];
        x[0] = id;
#pragma unroll
        for (int p = 1; p <= 6; ++p) x[p] = pw[(int64_t)(p - 1) * total + e];
        double b0 = 0.0, b1 = 0.0, b2 = 0.0;
#pragma unroll
        for (int p = 5; p >= 0; --p) {
            b0 += KR_INV_FACT[p] * x[p];
            b1 += KR_INV_FACT[6 + p] * x[p];
            b2 += KR_INV_FACT[12 + p] * x[p];
        }
        P2[e] = b2 + KR_INV_FACT[18] * x[6];
        B1[e] = b1;
        B0[e] = b0;
    }
}

// out = ca * A + cb * B   (B may be null)
__global__ void axpby_kernel(double* out, double ca, const double* A, double cb, const double* B, int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = ca * A[e] + (B ? cb * B[e] : 0.0);
}

constexpr double EXPM_THETA18 = 1.0908637192900361;     // theta_18 of the Taylor table (functions/theta_taylor.mat)

struct ExpmWork {
    DevBuf<double> pw, t0, t1, t2, neg;
    DevBuf<ExpmCtl> ctl;
    DevBuf<unsigned long long> bits;
    // The squarings are enqueued before ||S||_1 is known on the host.  The rigorous bound (sqrt(n) ||S||_2) over-
    // estimates by 4-6 squarings on the reference's graphs (each a wasted launch); callers that synchronise once per
    // Krylov step anyway pass the previous step's measured norm as a hint (the projected matrices only grow by a
    // block per step) and re-run the evaluation with the rigorous bound in the rare case the hint was too small.
    double hint = -1.0;             // last measured ||S||_1 (host), < 0: none
    ExpmCtl host_ctl;
    void begin(kr_ctx* ctx, bool rigorous) {
        if (!ctl.p) {
            ctl.reset(ctx, 1);
            bits.reset(ctx, 1);
            bits.zero();
        }
        ctl.zero();
        if (rigorous) hint = -1.0;
    }
    void enqueue_readback(kr_ctx* ctx) {
        KR_CUDA(cudaMemcpyAsync(&host_ctl, ctl.p, sizeof(ExpmCtl), cudaMemcpyDeviceToHost, ctx->stream));
    }
    bool accept() {                 // after the stream has been synchronised
        if (host_ctl.clamped) { hint = -1.0; return false; }
        hint = host_ctl.norm1;
        return true;
    }
};

inline int ew_grid(kr_ctx* ctx, int64_t total) {
    return (int)std::min<int64_t>((int64_t)ctx->num_sms * 8, std::max<int64_t>(1, ceil_div(total, 256)));
}

// E[b] = exp(sign * S[b]) for `batch` symmetric nn x nn matrices stored consecutively (column-major).
// norm_bound: a rigorous host-side bound on ||S[b]||_2 (bounds the squarings enqueued).
inline void expm_batched(kr_ctx* ctx, const double* S, int nn, int batch, double sign, double norm_bound, double* E,
                         ExpmWork& w) {
    const int64_t mat = (int64_t)nn * nn, total = mat * batch;
    if (total == 0) return;
    // ||S||_1 <= sqrt(nn) ||S||_2
    double b1 = std::sqrt((double)nn) * std::max(norm_bound, 0.0);
    if (w.hint >= 0.0) b1 = std::min(b1, 4.0 * w.hint + 1e-300);
    int s_max = b1 > EXPM_THETA18 ? (int)std::ceil(std::log2(b1 / EXPM_THETA18)) + 1 : 0;
    s_max = std::min(s_max, 60);
    if (w.pw.count < (size_t)(6 * total)) w.pw.reset(ctx, (size_t)(6 * total));
    if (w.t0.count < (size_t)total) w.t0.reset(ctx, (size_t)total);
    if (w.t1.count < (size_t)total) w.t1.reset(ctx, (size_t)total);
    if (w.t2.count < (size_t)total) w.t2.reset(ctx, (size_t)total);
    if (!w.ctl.p) w.begin(ctx, false);
    const int g = ew_grid(ctx, total);
    double* X1 = w.pw.p;
    auto P = [&](int p) { return w.pw.p + (int64_t)(p - 1) * total; };
    KR_LAUNCH(ctx, expm_rowmax_kernel, dim3((unsigned)ceil_div(nn, 256), (unsigned)batch), 256, 0, S, nn, w.bits.p);
    KR_LAUNCH(ctx, expm_decide_kernel, 1, 1, 0, w.bits.p, EXPM_THETA18, s_max, w.ctl.p);
    KR_LAUNCH(ctx, expm_scale_kernel, g, 256, 0, S, X1, total, sign, w.ctl.p);
    sgemm(ctx, nn, nn, nn, 1.0, P(1), nn, mat, P(1), nn, mat, 0.0, P(2), nn, mat, batch);      // X^2
    sgemm(ctx, nn, nn, nn, 1.0, P(2), nn, mat, P(1), nn, mat, 0.0, P(3), nn, mat, batch);      // X^3
    sgemm(ctx, nn, nn, nn, 1.0, P(2), nn, mat, P(2), nn, mat, 0.0, P(4), nn, mat, batch);      // X^4
    sgemm(ctx, nn, nn, nn, 1.0, P(3), nn, mat, P(2), nn, mat, 0.0, P(5), nn, mat, batch);      // X^5
    sgemm(ctx, nn, nn, nn, 1.0, P(3), nn, mat, P(3), nn, mat, 0.0, P(6), nn, mat, batch);      // X^6
    KR_LAUNCH(ctx, expm_blocks_kernel, g, 256, 0, w.pw.p, nn, mat, total, w.t0.p, w.t1.p, w.t2.p);  // P2, B1, B0
    sgemm(ctx, nn, nn, nn, 1.0, P(6), nn, mat, w.t0.p, nn, mat, 1.0, w.t1.p, nn, mat, batch);  // B1 + X^6 P2
    sgemm(ctx, nn, nn, nn, 1.0, P(6), nn, mat, w.t1.p, nn, mat, 1.0, w.t2.p, nn, mat, batch);  // B0 + X^6 (..)
    // squarings: ping-pong between t2 and t1; the result ends in `cur`
    double* cur = w.t2.p;
    double* oth = w.t1.p;
    for (int q = 0; q < s_max; ++q) {
        sgemm(ctx, nn, nn, nn, 1.0, cur, nn, mat, cur, nn, mat, 0.0, oth, nn, mat, batch, &w.ctl.p->s, q);
        std::swap(cur, oth);
    }
    KR_CUDA(cudaMemcpyAsync(E, cur, (size_t)total * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
}

// F[b] = f(S[b]), f = exp / sinh / cosh   (functions/fun_update.m:43-59)
inline void symfun_batched(kr_ctx* ctx, const double* S, int nn, int batch, int fun, double norm_bound, double* F,
                           ExpmWork& w) {
    const int64_t total = (int64_t)nn * nn * batch;
    if (total == 0) return;
    expm_batched(ctx, S, nn, batch, 1.0, norm_bound, F, w);
    if (fun == KR_FUN_EXP) return;
    if (w.neg.count < (size_t)total) w.neg.reset(ctx, (size_t)total);
    expm_batched(ctx, S, nn, batch, -1.0, norm_bound, w.neg.p, w);
    KR_LAUNCH(ctx, axpby_kernel, ew_grid(ctx, total), 256, 0, F, 0.5, F, fun == KR_FUN_SINH ? -0.5 : 0.5, w.neg.p, total);
}

// ------------------------------------------------------------------------------------ small reductions
// out[0] = trace(F1) - trace(F0) summed entry by entry as (F1_ii - F0_ii)   (one CTA, fixed order)
__global__ void __launch_bounds__(256)
trace_diff_kernel(const double* __restrict__ F1, const double* __restrict__ F0, int nn, double* __restrict__ out) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < nn; i += 256) s += F1[i + (int64_t)i * nn] - F0[i + (int64_t)i * nn];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

// ||D||_2 for a symmetric nn x nn matrix D (column-major) by Lanczos with full re-orthogonalisation from a
// fixed start vector (k <= 64 steps), Ritz values by bisection-free Jacobi on the small tridiagonal: one CTA.
// Used only for the stopping test of fun_update (fun_update.m:112-114) when the Frobenius norm does not
// already decide (||D||_F / sqrt(nn) <= ||D||_2 <= ||D||_F).
constexpr int NRM2_STEPS = 48;
__global__ void __launch_bounds__(256)
sym_norm2_kernel(const double* __restrict__ D, int nn, double* __restrict__ Q /* nn x (NRM2_STEPS+1) scratch */,
                 const int* __restrict__ need, double tol, double* __restrict__ out) {
    // Every Lanczos vector q_j is a unit vector, so ||D q_j|| <= ||D||_2: as soon as one of them reaches tol the
    // test "||D||_2 < tol" is decided (false) and the kernel stops - the usual case inside the Frobenius band.
    // Only a step that really has converged runs the full recurrence (once per fun_update call).
    if (need && !*need) return;
    __shared__ double red[256];
    __shared__ double al[NRM2_STEPS], be[NRM2_STEPS + 1], dl[NRM2_STEPS + 1];
    __shared__ double T[NRM2_STEPS * (NRM2_STEPS + 1)];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bsum = [&](double v) -> double {
        red[tid] = v;
        __syncthreads();
        for (int off = 128; off; off >>= 1) {
            if (tid < off) red[tid] += red[tid + off];
            __syncthreads();
        }
        const double r = red[0];
        __syncthreads();
        return r;
    };
    double nrm = 0.0;
    for (int i = tid; i < nn; i += 256) {
        const double v = 1.0 + 0.5 * sin(1.0 + 0.7 * i);       // fixed start vector (deterministic)
        Q[i] = v;
        nrm += v * v;
    }
    nrm = sqrt(bsum(nrm));
    for (int i = tid; i < nn; i += 256) Q[i] /= nrm;
    __syncthreads();
    const int steps = min(NRM2_STEPS, nn);
    int m = 0;
    for (int j = 0; j < steps; ++j) {
        const double* qj = Q + (int64_t)j * nn;
        double* w = Q + (int64_t)(j + 1) * nn;
        for (int i = tid; i < nn; i += 256) {                   // w = D qj (row i of the symmetric D = column i)
            double s = 0.0;
            for (int c = 0; c < nn; ++c) s += D[i + (int64_t)c * nn] * qj[c];
            w[i] = s;
        }
        __syncthreads();
        double a = 0.0;
        for (int pass = 0; pass < 2; ++pass) {                   // classical Gram-Schmidt twice against q_0..q_j
            for (int l = warp; l <= j; l += 8) {
                const double* ql = Q + (int64_t)l * nn;
                double d = 0.0;
                for (int i = lane; i < nn; i += 32) d += w[i] * ql[i];
                for (int off = 16; off; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
                if (lane == 0) dl[l] = d;
            }
            __syncthreads();
            if (pass == 0) a = dl[j];
            for (int i = tid; i < nn; i += 256) {
                double v = w[i];
                for (int l = 0; l <= j; ++l) v -= dl[l] * Q[(int64_t)l * nn + i];
                w[i] = v;
            }
            __syncthreads();
        }
        double b = 0.0;
        for (int i = tid; i < nn; i += 256) b += w[i] * w[i];
        b = sqrt(bsum(b));
        if (tid == 0) { al[j] = a; be[j + 1] = b; }
        m = j + 1;
        __syncthreads();
        {
            const double bp = j > 0 ? be[j] : 0.0;
            const double lb = sqrt(a * a + b * b + bp * bp);    // = ||D q_j|| (three-term recurrence)
            if (lb >= tol) {
                if (tid == 0) out[0] = lb;
                return;
            }
        }
        if (!(b > 1e-300) || b < 1e-15 * fabs(al[0])) break;
        for (int i = tid; i < nn; i += 256) w[i] /= b;
        __syncthreads();
    }
    // eigenvalues of the m x m tridiagonal by cyclic Jacobi (serial in thread 0: m <= 48)
    if (tid == 0) {
        const int ld = m + 1;
        for (int i = 0; i < m * ld; ++i) T[i] = 0.0;
        for (int i = 0; i < m; ++i) {
            T[i + i * ld] = al[i];
            if (i + 1 < m) { T[i + 1 + i * ld] = be[i + 1]; T[i + (i + 1) * ld] = be[i + 1]; }
        }
        for (int sweep = 0; sweep < 30; ++sweep) {
            double off = 0.0, dg = 0.0;
            for (int p = 0; p < m; ++p) {
                dg += T[p + p * ld] * T[p + p * ld];
                for (int q = p + 1; q < m; ++q) off += T[p + q * ld] * T[p + q * ld];
            }
            if (off <= 1e-34 * dg) break;
            for (int p = 0; p < m - 1; ++p)
                for (int q = p + 1; q < m; ++q) {
                    const double apq = T[p + q * ld];
                    if (apq == 0.0) continue;
                    const double theta = (T[q + q * ld] - T[p + p * ld]) / (2.0 * apq);
                    const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                    for (int i = 0; i < m; ++i) {
                        const double aip = T[i + p * ld], aiq = T[i + q * ld];
                        T[i + p * ld] = c * aip - sn * aiq;
                        T[i + q * ld] = sn * aip + c * aiq;
                    }
                    for (int i = 0; i < m; ++i) {
                        const double api = T[p + i * ld], aqi = T[q + i * ld];
                        T[p + i * ld] = c * api - sn * aqi;
                        T[q + i * ld] = sn * api + c * aqi;
                    }
                }
        }
        double mx = 0.0;
        for (int i = 0; i < m; ++i) mx = fmax(mx, fabs(T[i + i * ld]));
        out[0] = mx;
    }
}

}  // namespace kr
