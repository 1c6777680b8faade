// vecops.cuh - single-vector kernels on plain length-n device vectors: a real one-column SpMV (8 lanes per stored
// row, not a 1/16-filled panel SpMM) and fixed-order reductions.  Used by normest (functions/
// fun_and_grad_krylov_exp.m:26), the leading-eigenvector centrality (functions/compute_centrality.m:15-17) and the
// 1-norm power sequence of normAm (functions/normAm.m:17-23).
#pragma once
#include "csr.cuh"

namespace kr {

// y[row] = alpha * (sum_p val[p] * x[col[p]] - mu * x[row]); rows are walked in stored (length-sorted) order, 8 lanes
// per row, shuffle reduction in a fixed order.
__global__ void __launch_bounds__(256)
spmv_kernel(CsrDevView A, const double* __restrict__ x, double* __restrict__ y, double alpha, double mu) {
    const int g = (blockIdx.x * 256 + threadIdx.x) >> 3, sub = threadIdx.x & 7;
    if (g >= A.n) return;                    // whole 8-lane groups leave together
    const int p0 = A.row_ptr[g], p1 = A.row_ptr[g + 1];
    double s = 0.0;
    for (int p = p0 + sub; p < p1; p += 8) s += (A.val ? A.val[p] : 1.0) * x[A.col[p]];
    const unsigned m = __activemask();
    s += __shfl_xor_sync(m, s, 1);
    s += __shfl_xor_sync(m, s, 2);
    s += __shfl_xor_sync(m, s, 4);
    if (sub == 0) {
        const int row = A.row_order[g];
        if (!A.val) s *= A.uval;
        y[row] = alpha * (s - mu * x[row]);
    }
}

// ---- one-CTA products for the whole-loop kernels of small graphs (normest, power iteration): latency-bound, so the
// rows a lane group has in flight matter more than anything else.  Rows are stored by decreasing length:
// cta_spmv_split finds the first stored row with at most 32 nonzeros (every thread runs the same binary search),
// rows before it get a warp each, the rest four lanes each with four rows of a lane group in flight.
__device__ __forceinline__ int cta_spmv_split(const CsrDevView& S) {
    int lo = 0, hi = S.n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (S.row_ptr[mid + 1] - S.row_ptr[mid] > 32) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// out[row] = sum_p val[p] * in[col[p]] * (scale ? scale[col[p]] : 1); all NT threads call; ends with a CTA barrier
template <int NT>
__device__ __forceinline__ void cta_spmv(const CsrDevView& S, const double* in, double* out, const double* scale, int split) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool hv = S.val != nullptr;
    for (int g = warp; g < split; g += NT / 32) {
        const int p0 = S.row_ptr[g], p1 = S.row_ptr[g + 1];
        double s = 0.0;
#pragma unroll 4
        for (int p = p0 + lane; p < p1; p += 32) {
            const int c = S.col[p];
            s += (hv ? S.val[p] : 1.0) * (scale ? in[c] * scale[c] : in[c]);
        }
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) out[S.row_order[g]] = hv ? s : s * S.uval;
    }
    constexpr int G = NT / 4, ILP = 4;
    const int sub = lane & 3;
    for (int base = split + warp * 8; base < S.n; base += G * ILP) {
        int q0[ILP], q1[ILP];
        double acc[ILP];
        int maxlen = 0;
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const int r = base + u * G + (lane >> 2);
            q0[u] = q1[u] = 0;
            acc[u] = 0.0;
            if (r < S.n) { q0[u] = S.row_ptr[r]; q1[u] = S.row_ptr[r + 1]; }
            maxlen = max(maxlen, q1[u] - q0[u]);
        }
        for (int off = sub; off < maxlen; off += 4) {
#pragma unroll
            for (int u = 0; u < ILP; ++u) {
                const int p = q0[u] + off;
                if (p < q1[u]) {
                    const int c = S.col[p];
                    acc[u] += (hv ? S.val[p] : 1.0) * (scale ? in[c] * scale[c] : in[c]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const int r = base + u * G + (lane >> 2);
            double v = acc[u];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (sub == 0 && r < S.n) out[S.row_order[r]] = hv ? v : v * S.uval;
        }
    }
    __syncthreads();
}

inline void spmv(kr_ctx* ctx, const CsrDev& A, const double* x, double* y, double alpha = 1.0, double mu = 0.0) {
    if (A.n == 0) return;
    KR_LAUNCH(ctx, spmv_kernel, (int)ceil_div(A.n * 8, 256), 256, 0, A.view(), x, y, alpha, mu);
    ctx->counters[1] += 1;
    ctx->counters[2] += 1;
}

// mode 0: sum x*y   mode 1: max |x|   mode 2: sum |x|     -> partial[blockIdx.x]; vec_reduce_finish folds them in order
constexpr int VEC_RED_CTAS = 296;
template <int MODE>
__global__ void __launch_bounds__(256)
vec_reduce_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n, double* __restrict__ partial) {
    __shared__ double red[8];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        if (MODE == 0) s += x[i] * y[i];
        else if (MODE == 1) s = fmax(s, fabs(x[i]));
        else s += fabs(x[i]);
    }
    for (int off = 16; off; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, s, off);
        s = MODE == 1 ? fmax(s, o) : s + o;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) s = MODE == 1 ? fmax(s, red[w]) : s + red[w];
        partial[blockIdx.x] = s;
    }
}
template <int MODE>
__global__ void vec_reduce_finish_kernel(const double* __restrict__ partial, int count, double* __restrict__ out) {
    double s = 0.0;
    for (int i = 0; i < count; ++i) s = MODE == 1 ? fmax(s, partial[i]) : s + partial[i];
    out[0] = s;
}
template <int MODE>
inline void vec_reduce(kr_ctx* ctx, const double* x, const double* y, int64_t n, double* partial /* >= VEC_RED_CTAS */,
                       double* out) {
    const int ctas = (int)std::min<int64_t>(VEC_RED_CTAS, std::max<int64_t>(1, ceil_div(n, 256)));
    KR_LAUNCH(ctx, vec_reduce_kernel<MODE>, ctas, 256, 0, x, y, n, partial);
    KR_LAUNCH(ctx, vec_reduce_finish_kernel<MODE>, 1, 1, 0, partial, ctas, out);
}

// x[i] = x[i] * (*num_or_null ? ...): y = x / sqrt(*s2)   (normalisation by a device-resident squared norm)
__global__ void vec_scale_inv_sqrt_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n,
                                          const double* __restrict__ s2) {
    const double d = sqrt(*s2);
    const double f = d > 0.0 ? 1.0 / d : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = x[i] * f;
}

}  // namespace kr
