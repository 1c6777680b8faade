// blockkrylov.cuh - general-width block Lanczos / Arnoldi (functions/lanczos_krylov.m,
// functions/arnoldi_krylov.m) and the evaluators built on them with one wide block:
// trace_fun_update (rk > 2), fun_update, fun_and_grad_krylov_*, mc_trace, normest.
//
// Blocks live on the device column-major (n x bs, MATLAB layout) so that the tall-skinny Gram /
// update products are plain library GEMMs (cuBLAS dgemm -> FP64 DMMA on B200) and the thin QR is
// cuSOLVER's Householder geqrf/orgqr - the same factorisation (and sign / zero-column conventions)
// as the reference's qr(w,0).  The SpMM runs on the panel-major copy (two cheap transposes per
// step).  The small projected matrices (H, K, Cm, Gm) live on the host; their eigen-solves run on
// the device (Jacobi kernel for n <= 110, cuSOLVER syevd above).
#pragma once
#include <chrono>
#include <memory>

#include "dense.cuh"
#include "smalldense.cuh"

namespace kr {

struct CmMat {                       // owning column-major device matrix, ld == rows
    DevBuf<double> buf;
    int64_t rows = 0, cols = 0;
    CmMat() = default;
    CmMat(kr_ctx* ctx, int64_t r, int64_t c) { reset(ctx, r, c); }
    void reset(kr_ctx* ctx, int64_t r, int64_t c) {
        rows = r;
        cols = c;
        buf.reset(ctx, (size_t)std::max<int64_t>(r * c, 1));
    }
    double* p() const { return buf.p; }
    double* col(int64_t c) const { return buf.p + c * rows; }
};

struct HostMat {                     // small column-major host matrix
    int64_t rows = 0, cols = 0;
    std::vector<double> a;
    HostMat() = default;
    HostMat(int64_t r, int64_t c) : rows(r), cols(c), a((size_t)(r * c), 0.0) {}
    double& operator()(int64_t i, int64_t j) { return a[(size_t)(i + j * rows)]; }
    double operator()(int64_t i, int64_t j) const { return a[(size_t)(i + j * rows)]; }
    void grow(int64_t r, int64_t c) {     // zero-padded enlarge (H(end+bs, end+bs) = 0)
        HostMat n(r, c);
        for (int64_t j = 0; j < cols; ++j)
            for (int64_t i = 0; i < rows; ++i) n(i, j) = (*this)(i, j);
        *this = std::move(n);
    }
};

inline void upload_host_cm(kr_ctx* ctx, const double* host, int64_t ld, CmMat& dst) {
    if (dst.rows == 0 || dst.cols == 0) return;
    KR_CUDA(cudaMemcpy2DAsync(dst.p(), dst.rows * sizeof(double), host, ld * sizeof(double),
                              dst.rows * sizeof(double), dst.cols, cudaMemcpyHostToDevice, ctx->stream));
    ctx->counters[3] += dst.rows * dst.cols * (int64_t)sizeof(double);
    KR_CUDA(cudaStreamSynchronize(ctx->stream));
}
inline void download_cm_to_host(kr_ctx* ctx, const double* dev, int64_t rows, int64_t cols, double* host, int64_t ld) {
    if (rows == 0 || cols == 0) return;
    KR_CUDA(cudaMemcpy2DAsync(host, ld * sizeof(double), dev, rows * sizeof(double), rows * sizeof(double), cols,
                              cudaMemcpyDeviceToHost, ctx->stream));
    ctx->counters[4] += rows * cols * (int64_t)sizeof(double);
    KR_CUDA(cudaStreamSynchronize(ctx->stream));
}

// C(m x n) = alpha * op(A) * op(B) + beta * C   (device, column-major)
inline void gemm(kr_ctx* ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                 int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
    if (m == 0 || n == 0) return;
    KR_CUBLAS(cublasDgemm(ctx->cublas, ta ? CUBLAS_OP_T : CUBLAS_OP_N, tb ? CUBLAS_OP_T : CUBLAS_OP_N, (int)m, (int)n,
                          (int)k, &alpha, A, (int)lda, B, (int)ldb, &beta, C, (int)ldc));
    ctx->counters[0] += 1;
}

// ------------------------------------------------------------------ FP64 tensor-core Gram
// G = V' W for tall-skinny column-major blocks (V: n x c, W: n x b, b <= 128): the CGS2 / third-pass
// Gram of functions/lanczos_krylov.m:110-112 and functions/arnoldi_krylov.m:104,120-122 when the block
// width makes it a real dense contraction.  DMMA m8n8k4 (the only fp64 tensor shape family on sm_100;
// tcgen05 has no fp64 kind): the A fragment is a 8x4 tile of V' and the B fragment a 4x8 tile of W,
// both read straight from global memory - 4 consecutive rows of 8 columns = 8 fully used 32-byte
// sectors per warp load, so no shared-memory staging is needed.  grid = (row chunks, ceil(c/32)); a
// CTA owns a 32 x b tile of G for its chunk of rows, warp w owns the 8-column tiles w and w+8.
// Split-K partials land in a fixed layout and are summed in a fixed order (deterministic).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int GRAM_KCH = 4096;     // rows per CTA
__global__ void __launch_bounds__(256)
gram_dmma_kernel(const double* __restrict__ V, int64_t ldv, int c, const double* __restrict__ W, int64_t ldw,
                 int b, int64_t n, double* __restrict__ partial) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kk = lane & 3, idx = lane >> 2;
    const int m0 = blockIdx.y * 32;
    const int64_t r0 = (int64_t)blockIdx.x * GRAM_KCH, r1 = min(n, r0 + (int64_t)GRAM_KCH);
    const int ntiles = (b + 7) >> 3;
    double acc[4][2][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int s = 0; s < 2; ++s) acc[mt][s][0] = acc[mt][s][1] = 0.0;
    const bool has0 = warp < ntiles, has1 = warp + 8 < ntiles;     // warp-uniform
    if (has0) {
        for (int64_t k0 = r0; k0 < r1; k0 += 4) {
            const int64_t row = k0 + kk;
            const bool rok = row < r1;
            double a[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                const int col = m0 + mt * 8 + idx;
                a[mt] = (rok && col < c) ? __ldg(V + row + (int64_t)col * ldv) : 0.0;
            }
            {
                const int col = warp * 8 + idx;
                const double bb = (rok && col < b) ? __ldg(W + row + (int64_t)col * ldw) : 0.0;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) dmma884(acc[mt][0][0], acc[mt][0][1], a[mt], bb);
            }
            if (has1) {
                const int col = (warp + 8) * 8 + idx;
                const double bb = (rok && col < b) ? __ldg(W + row + (int64_t)col * ldw) : 0.0;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) dmma884(acc[mt][1][0], acc[mt][1][1], a[mt], bb);
            }
        }
    }
    // C fragment: lane holds C[idx][2*kk], C[idx][2*kk+1] of each 8x8 tile; partial is column-major c x b
    double* out = partial + (int64_t)blockIdx.x * c * b;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int nt = warp + s * 8;
        if (nt >= ntiles) continue;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int m = m0 + mt * 8 + idx;
            const int n0 = nt * 8 + 2 * kk;
            if (m < c) {
                if (n0 < b) out[m + (int64_t)n0 * c] = acc[mt][s][0];
                if (n0 + 1 < b) out[m + (int64_t)(n0 + 1) * c] = acc[mt][s][1];
            }
        }
    }
}

// G (c x b, column-major, device) = V' W.  Falls back to cuBLAS for b > 128.
inline void gemm(kr_ctx* ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                 int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc);
inline void gram_tn(kr_ctx* ctx, const double* V, int64_t ldv, int64_t c, const double* W, int64_t ldw, int64_t b,
                    int64_t n, double* G) {
    if (c == 0 || b == 0) return;
    static const bool force_cublas = getenv("KR_GRAM_CUBLAS") != nullptr;   // A/B switch for debugging
    if (b > 128 || force_cublas) {
        gemm(ctx, true, false, c, b, n, 1.0, V, ldv, W, ldw, 0.0, G, c);
        return;
    }
    const int chunks = (int)ceil_div(n, GRAM_KCH);
    DevBuf<double> partial(ctx, (size_t)chunks * c * b);
    dim3 grid((unsigned)chunks, (unsigned)ceil_div(c, 32));
    KR_LAUNCH(ctx, gram_dmma_kernel, grid, 256, 0, V, ldv, (int)c, W, ldw, (int)b, n, partial.p);
    sum_partials(ctx, partial.p, chunks, (int)(c * b), G);
}

// Y = A * X for column-major device blocks (n x k)
inline void spmm_cm(kr_ctx* ctx, const kr_matrix* M, const double* X, int64_t k, double* Y) {
    const int64_t n = M->dev.n;
    PanelBuf xb(ctx, n, (int)k), yb(ctx, n, (int)k);
    cm_to_panel(ctx, X, n, xb);
    EpiPlain epi{yb.p(), xb.p(), 1.0, 0.0};
    launch_spmm(ctx, M->dev, xb.p(), xb.panels, epi, nullptr, (int)k);
    panel_to_cm(ctx, yb, Y, n);
}

// Thin Householder QR in place: W (n x bs) <- Q, R (bs x bs, host, upper triangular).
inline void qr_thin(kr_ctx* ctx, double* W, int64_t n, int64_t bs, HostMat& R) {
    R = HostMat(bs, bs);
    if (n < bs) fail(KR_ERR_UNSUPPORTED, "thin QR needs n >= block size");
    int lwork1 = 0, lwork2 = 0;
    KR_CUSOLVER(cusolverDnDgeqrf_bufferSize(ctx->cusolver, (int)n, (int)bs, W, (int)n, &lwork1));
    DevBuf<double> tau(ctx, bs);
    KR_CUSOLVER(cusolverDnDorgqr_bufferSize(ctx->cusolver, (int)n, (int)bs, (int)bs, W, (int)n, tau.p, &lwork2));
    const int lwork = std::max(lwork1, lwork2);
    DevBuf<double> work(ctx, std::max(lwork, 1));
    DevBuf<int> info(ctx, 1);
    KR_CUSOLVER(cusolverDnDgeqrf(ctx->cusolver, (int)n, (int)bs, W, (int)n, tau.p, work.p, lwork, info.p));
    std::vector<double> top((size_t)bs * bs);
    // leading bs x bs of the factored W (leading dimension n): strided copy
    KR_CUDA(cudaMemcpy2DAsync(top.data(), bs * sizeof(double), W, n * sizeof(double), bs * sizeof(double), bs,
                              cudaMemcpyDeviceToHost, ctx->stream));
    KR_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t j = 0; j < bs; ++j)
        for (int64_t i = 0; i <= j; ++i) R(i, j) = top[(size_t)(i + j * bs)];
    KR_CUSOLVER(cusolverDnDorgqr(ctx->cusolver, (int)n, (int)bs, (int)bs, W, (int)n, tau.p, work.p, lwork, info.p));
    ctx->counters[0] += 2;
}

inline double fro_norm(const HostMat& R) {
    double s = 0;
    for (double v : R.a) s += v * v;
    return std::sqrt(s);
}

// ---- host-side phase timers of the wide-block path (KR_PROFILE_WIDE=1 prints them to stderr at the end of
// fun_update / trace_fun_update; they synchronise the stream, so they are a diagnostic, not a product path)
struct WideProf {
    bool on = getenv("KR_PROFILE_WIDE") != nullptr;
    double eig_small = 0, eig_large = 0, step = 0, other = 0;
    int n_small = 0, n_large = 0, n_step = 0, max_dim = 0;
    static double now() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    void report(const char* what) {
        if (!on) return;
        fprintf(stderr, "[kr wide] %s: krylov steps %d %.2f ms | eig(n<=110, smem Jacobi) %d %.2f ms | eig(syevd) %d %.2f ms "
                        "(largest n %d)\n", what, n_step, step, n_small, eig_small, n_large, eig_large, max_dim);
        eig_small = eig_large = step = other = 0;
        n_small = n_large = n_step = max_dim = 0;
    }
};
inline WideProf& wide_prof() { static WideProf p; return p; }

// Vf(:, k) = V(:, k) * f(w_k)  (column-major n x n), the first half of F = V f(D) V'
__global__ void scale_cols_fun_kernel(const double* __restrict__ V, const double* __restrict__ w, int n, int fun,
                                      double* __restrict__ Vf) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * n) return;
    const double x = w[e / n];
    const double fx = fun == KR_FUN_EXP ? exp(x) : fun == KR_FUN_SINH ? sinh(x) : cosh(x);
    Vf[e] = V[e] * fx;
}

inline int jacobi_max_dim() {
    static const int v = [] { const char* e = getenv("KR_JACOBI_MAX"); return e ? atoi(e) : 16; }();
    return v;
}

// ---- small symmetric eigen-solves on the device (host in / host out)
// evals (ascending); if F != nullptr also F = V f(D) V'.
inline void sym_eig_dev(kr_ctx* ctx, const HostMat& S, std::vector<double>& evals, int fun, HostMat* F) {
    const int n = (int)S.rows;
    evals.assign(n, 0.0);
    if (n == 0) return;
    WideProf& wp = wide_prof();
    const double t_begin = wp.on ? WideProf::now() : 0.0;
    struct Scope {
        WideProf& wp; double t0; int n; bool small;
        ~Scope() {
            if (!wp.on) return;
            const double dt = WideProf::now() - t0;
            if (small) { wp.eig_small += dt; wp.n_small++; } else { wp.eig_large += dt; wp.n_large++; }
            wp.max_dim = std::max(wp.max_dim, n);
        }
    } scope{wp, t_begin, n, n <= jacobi_max_dim()};
    DevBuf<double> dA(ctx, (size_t)n * n), dW(ctx, n);
    dA.upload(S.a.data(), (size_t)n * n);
    const bool vec = F != nullptr;
    // One-CTA cyclic Jacobi only for tiny projections: measured 12-15 ms per solve at n = 100-156 against
    // 1.7 ms for cuSOLVER syevd at n = 208 (profiles/r01_wide_block_phases.txt); it wins only below ~16-32, where
    // syevd's fixed cost of several launches dominates.
    if (n <= jacobi_max_dim() && jacobi_smem_bytes(n, vec) <= JAC_SMEM_LIMIT) {
        if (!vec) {
            static bool set1 = false;
            if (!set1) { KR_CUDA(cudaFuncSetAttribute(eigvals_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JAC_SMEM_LIMIT)); set1 = true; }
            KR_LAUNCH(ctx, eigvals_batched_kernel, 1, JAC_THREADS, jacobi_smem_bytes(n, false), dA.p, n, dW.p, (double*)nullptr);
            evals = dW.to_host();
        } else {
            static bool set2 = false;
            if (!set2) { KR_CUDA(cudaFuncSetAttribute(symfun_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JAC_SMEM_LIMIT)); set2 = true; }
            DevBuf<double> dF(ctx, (size_t)n * n);
            KR_LAUNCH(ctx, symfun_batched_kernel, 1, JAC_THREADS, jacobi_smem_bytes(n, true), dA.p, n, fun, dF.p, (double*)nullptr);
            *F = HostMat(n, n);
            dF.download(F->a.data(), (size_t)n * n);
        }
        return;
    }
    // large projection: cuSOLVER syevd (plain library call), then F = V f(D) V' by dgemm
    int lwork = 0;
    cusolverEigMode_t jobz = vec ? CUSOLVER_EIG_MODE_VECTOR : CUSOLVER_EIG_MODE_NOVECTOR;
    KR_CUSOLVER(cusolverDnDsyevd_bufferSize(ctx->cusolver, jobz, CUBLAS_FILL_MODE_LOWER, n, dA.p, n, dW.p, &lwork));
    DevBuf<double> work(ctx, std::max(lwork, 1));
    DevBuf<int> info(ctx, 1);
    KR_CUSOLVER(cusolverDnDsyevd(ctx->cusolver, jobz, CUBLAS_FILL_MODE_LOWER, n, dA.p, n, dW.p, work.p, lwork, info.p));
    ctx->counters[0] += 1;
    if (vec) {
        // F = (V f(D)) V' entirely on the device: one column-scaling kernel + one dgemm, one download
        DevBuf<double> dVf(ctx, (size_t)n * n), dF(ctx, (size_t)n * n);
        KR_LAUNCH(ctx, scale_cols_fun_kernel, (int)ceil_div((int64_t)n * n, 256), 256, 0, dA.p, dW.p, n, fun, dVf.p);
        gemm(ctx, false, true, n, n, n, 1.0, dVf.p, n, dA.p, n, 0.0, dF.p, n);
        *F = HostMat(n, n);
        KR_CUDA(cudaMemcpyAsync(evals.data(), dW.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        dF.download(F->a.data(), (size_t)n * n);
    } else {
        evals = dW.to_host();
    }
}

// spectral norm of a symmetric host matrix = max |eigenvalue| (device eigen-solve)
inline double sym_norm2_dev(kr_ctx* ctx, const HostMat& S) {
    std::vector<double> ev;
    sym_eig_dev(ctx, S, ev, KR_FUN_EXP, nullptr);
    double m = 0;
    for (double v : ev) m = std::max(m, std::abs(v));
    return m;
}

inline double trace_formula_host(int fun, const std::vector<double>& d1, const std::vector<double>& d2) {
    double s = 0;
    for (size_t i = 0; i < d1.size(); ++i) {
        if (fun == KR_FUN_EXP) s += std::exp(d1[i]) * (1.0 - std::exp(d2[i] - d1[i]));
        else if (fun == KR_FUN_SINH) s += std::sinh(d1[i]) - std::sinh(d2[i]);
        else s += std::cosh(d1[i]) - std::cosh(d2[i]);
    }
    return s;
}

}  // namespace kr

// ------------------------------------------------------------------------------------------------
// params + V + H (+K) of the reference, device-resident behind a handle
struct kr_krylov {
    kr_ctx* ctx = nullptr;
    const kr_matrix* A = nullptr;
    bool arnoldi = false;
    int64_t n = 0, bs = 0, steps = 0;
    kr::CmMat V;            // Lanczos: n x (bs or 2bs) window; Arnoldi: n x (steps+1)*bs (capacity grows)
    int64_t vcols = 0;
    kr::HostMat H, K;
    bool lucky = false;
};

namespace kr {

// one add_inf_pole step (lanczos_krylov.m:73-101 / arnoldi_krylov.m:78-111); the continuation
// block params.last is always the last bs columns of V.
inline void krylov_step(kr_krylov* st) {
    kr_ctx* ctx = st->ctx;
    const int64_t n = st->n, bs = st->bs, c = st->vcols;
    CmMat W(ctx, n, bs);
    spmm_cm(ctx, st->A, st->V.col(c - bs), bs, W.p());
    // CGS2 against V (n x c)
    DevBuf<double> dh(ctx, (size_t)c * bs), dh1(ctx, (size_t)c * bs);
    gram_tn(ctx, st->V.p(), n, c, W.p(), n, bs, n, dh.p);
    gemm(ctx, false, false, n, bs, c, -1.0, st->V.p(), n, dh.p, c, 1.0, W.p(), n);
    gram_tn(ctx, st->V.p(), n, c, W.p(), n, bs, n, dh1.p);
    gemm(ctx, false, false, n, bs, c, -1.0, st->V.p(), n, dh1.p, c, 1.0, W.p(), n);
    std::vector<double> h = dh.to_host(), h1 = dh1.to_host();
    for (size_t i = 0; i < h.size(); ++i) h[i] += h1[i];
    HostMat& H = st->H;
    H.grow(H.rows + bs, H.cols + bs);
    const int64_t R = H.rows, C = H.cols;
    HostMat Rf;
    if (!st->arnoldi) {
        const int64_t r0 = std::max<int64_t>(1, R - 3 * bs + 1) - 1;      // lanczos_krylov.m:88
        for (int64_t j = 0; j < bs; ++j)
            for (int64_t i = 0; i < c; ++i) H(r0 + i, C - bs + j) = h[(size_t)(i + j * c)];
        qr_thin(ctx, W.p(), n, bs, Rf);
        for (int64_t j = 0; j < bs; ++j)
            for (int64_t i = 0; i < bs; ++i) H(R - bs + i, C - bs + j) = Rf(i, j);
        st->lucky = fro_norm(Rf) < 1e-8;                                   // :91
        if (c == bs) {                                                     // :94-99
            CmMat Vn(ctx, n, 2 * bs);
            KR_CUDA(cudaMemcpyAsync(Vn.p(), st->V.p(), (size_t)n * bs * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            KR_CUDA(cudaMemcpyAsync(Vn.col(bs), W.p(), (size_t)n * bs * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            st->V = std::move(Vn);
            st->vcols = 2 * bs;
        } else {
            KR_CUDA(cudaMemcpyAsync(st->V.p(), st->V.col(bs), (size_t)n * bs * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            KR_CUDA(cudaMemcpyAsync(st->V.col(bs), W.p(), (size_t)n * bs * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    } else {
        HostMat& K = st->K;
        K.grow(K.rows + bs, K.cols + bs);
        for (int64_t j = 0; j < bs; ++j)
            for (int64_t i = 0; i < c; ++i) H(i, C - bs + j) = h[(size_t)(i + j * c)];     // arnoldi_krylov.m:96
        for (int64_t i = 0; i < bs; ++i) K(R - 2 * bs + i, C - bs + i) = 1.0;              // :97
        qr_thin(ctx, W.p(), n, bs, Rf);
        // ||r||_2 < 1e-12 (:100): spectral norm of a bs x bs triangular matrix = sqrt(max eig(R'R))
        {
            HostMat RtR(bs, bs);
            for (int64_t i = 0; i < bs; ++i)
                for (int64_t j = 0; j < bs; ++j) {
                    double s = 0;
                    for (int64_t k = 0; k < bs; ++k) s += Rf(k, i) * Rf(k, j);
                    RtR(i, j) = s;
                }
            // ||R||_F / sqrt(bs) <= ||R||_2 <= ||R||_F: away from breakdown the Frobenius norm already decides
            // and the bs x bs eigen-solve (one more device round trip per step) is skipped
            const double fro = fro_norm(Rf);
            if (bs == 1) st->lucky = std::abs(Rf(0, 0)) < 1e-12;
            else if (fro / std::sqrt((double)bs) >= 1e-12) st->lucky = false;
            else if (fro < 1e-12) st->lucky = true;
            else st->lucky = std::sqrt(sym_norm2_dev(ctx, RtR)) < 1e-12;
        }
        // third reorthogonalisation (:104-106)
        DevBuf<double> dhh(ctx, (size_t)c * bs);
        gram_tn(ctx, st->V.p(), n, c, W.p(), n, bs, n, dhh.p);
        gemm(ctx, false, false, n, bs, c, -1.0, st->V.p(), n, dhh.p, c, 1.0, W.p(), n);
        std::vector<double> hh = dhh.to_host();
        for (int64_t j = 0; j < bs; ++j)
            for (int64_t i = 0; i < c; ++i) {
                double s = 0;
                for (int64_t k = 0; k < bs; ++k) s += hh[(size_t)(i + k * c)] * Rf(k, j);
                H(i, C - bs + j) += s;
            }
        for (int64_t j = 0; j < bs; ++j)
            for (int64_t i = 0; i < bs; ++i) H(R - bs + i, C - bs + j) = Rf(i, j);
        // V = [V, w]
        if ((c + bs) * n > (int64_t)st->V.buf.count) {
            CmMat Vn(ctx, n, std::max<int64_t>(2 * (c + bs), 8 * bs));
            KR_CUDA(cudaMemcpyAsync(Vn.p(), st->V.p(), (size_t)n * c * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
            st->V = std::move(Vn);
        }
        KR_CUDA(cudaMemcpyAsync(st->V.col(c), W.p(), (size_t)n * bs * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        st->vcols = c + bs;
    }
    st->steps += 1;
    KR_CUDA(cudaStreamSynchronize(ctx->stream));
}

// start: V1 = qr(b, 0), one step    (lanczos_krylov.m:30-58, arnoldi_krylov.m:32-62); b on the device
inline std::unique_ptr<kr_krylov> krylov_start(kr_ctx* ctx, const kr_matrix* A, bool arnoldi, int64_t bs, CmMat&& b) {
    const int64_t n = A->dev.n;
    if (b.rows != n) fail(KR_ERR_ARG, "The block vector b has wrong number of rows");
    if (bs < 1) fail(KR_ERR_ARG, "empty starting block");
    std::unique_ptr<kr_krylov> st(new kr_krylov());
    st->ctx = ctx; st->A = A; st->arnoldi = arnoldi; st->n = n; st->bs = bs;
    HostMat R0;
    qr_thin(ctx, b.p(), n, bs, R0);
    st->V = std::move(b);
    st->vcols = bs;
    st->H = HostMat(bs, 0);
    st->K = HostMat(bs, 0);
    krylov_step(st.get());
    return st;
}

// Cm = (V1' U) B (V1' U)'   (trace_fun_update.m:65-66), V1 = first bs columns at start time
inline HostMat core_Cm(kr_ctx* ctx, const kr_krylov* st, const CmMat& U, const HostMat& B) {
    const int64_t n = st->n, bs = st->bs;
    DevBuf<double> dC(ctx, (size_t)bs * bs);
    gemm(ctx, true, false, bs, bs, n, 1.0, st->V.p(), n, U.p(), n, 0.0, dC.p, bs);
    std::vector<double> c = dC.to_host();
    HostMat T(bs, bs), Cm(bs, bs);
    for (int64_t i = 0; i < bs; ++i)
        for (int64_t j = 0; j < bs; ++j) {
            double s = 0;
            for (int64_t k = 0; k < bs; ++k) s += c[(size_t)(i + k * bs)] * B(k, j);
            T(i, j) = s;
        }
    for (int64_t i = 0; i < bs; ++i)
        for (int64_t j = 0; j < bs; ++j) {
            double s = 0;
            for (int64_t k = 0; k < bs; ++k) s += T(i, k) * c[(size_t)(j + k * bs)];
            Cm(i, j) = s;
        }
    return Cm;
}

inline bool is_symmetric(const HostMat& B) {
    for (int64_t i = 0; i < B.rows; ++i)
        for (int64_t j = 0; j < i; ++j)
            if (B(i, j) != B(j, i)) return false;
    return B.rows == B.cols;
}

// dense symmetric copy of A (+ U B U') for the small-n branches (host assembly, device eigen-solve)
inline HostMat dense_of(const kr_matrix* M) {
    const CsrHost& H = M->host;
    HostMat D(H.n, H.n);
    for (int64_t i = 0; i < H.n; ++i)
        for (int64_t p = H.row_ptr[i]; p < H.row_ptr[i + 1]; ++p) D(i, H.col[p]) += H.val[p];
    return D;
}

struct TfuResult { double Xm = 0; int64_t iter = 0; int lucky = 0; };

// trace_fun_update with one wide block (any rk)      (trace_fun_update.m:21-130)
inline TfuResult trace_fun_update_general(kr_ctx* ctx, const kr_matrix* M, int64_t rk, const double* U, int64_t ldu,
                                          const HostMat& B, double tol, int64_t it, int fun) {
    const int64_t n = M->dev.n;
    TfuResult out;
    if (!is_symmetric(B)) fail(KR_ERR_UNSUPPORTED, "trace_fun_update: the device path needs a symmetric (Hermitian) B");
    if (n <= 130) {                                           // :37-51 dense branch
        HostMat fA = dense_of(M), fAt = fA;
        // fAt = fA + U B U', symmetrised
        std::vector<double> UB((size_t)n * rk, 0.0);
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < rk; ++j) {
                double s = 0;
                for (int64_t k = 0; k < rk; ++k) s += U[i + k * ldu] * B(k, j);
                UB[(size_t)(i + j * n)] = s;
            }
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < n; ++j) {
                double s = 0;
                for (int64_t k = 0; k < rk; ++k) s += UB[(size_t)(i + k * n)] * U[j + k * ldu];
                fAt(i, j) += s;
            }
        HostMat S(n, n);
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = 0; j < n; ++j) S(i, j) = 0.5 * (fAt(i, j) + fAt(j, i));
        std::vector<double> d1, d2;
        sym_eig_dev(ctx, S, d1, fun, nullptr);
        sym_eig_dev(ctx, fA, d2, fun, nullptr);
        out.Xm = trace_formula_host(fun, d1, d2);
        return out;
    }
    CmMat Ud(ctx, n, rk), b0(ctx, n, rk);
    upload_host_cm(ctx, U, ldu, Ud);
    KR_CUDA(cudaMemcpyAsync(b0.p(), Ud.p(), (size_t)n * rk * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    std::unique_ptr<kr_krylov> st;
    HostMat Cm;
    double Xstop[2] = {0, 0};
    int64_t j = 0;
    for (j = 1; j <= it; ++j) {
        if (j == 1) {
            st = krylov_start(ctx, M, false, rk, std::move(b0));
            Cm = core_Cm(ctx, st.get(), Ud, B);
        } else {
            const double t0 = wide_prof().on ? WideProf::now() : 0.0;
            krylov_step(st.get());
            if (wide_prof().on) { KR_CUDA(cudaStreamSynchronize(ctx->stream)); wide_prof().step += WideProf::now() - t0; wide_prof().n_step++; }
        }
        const int64_t nn = st->H.rows - rk;
        HostMat G(nn, nn), tG(nn, nn);
        for (int64_t a = 0; a < nn; ++a)
            for (int64_t b = 0; b < nn; ++b) {
                double g = 0.5 * (st->H(a, b) + st->H(b, a));
                double cm = (a < rk && b < rk) ? 0.5 * (Cm(a, b) + Cm(b, a)) : 0.0;
                G(a, b) = g;
                tG(a, b) = g + cm;
            }
        std::vector<double> d1, d2;
        sym_eig_dev(ctx, tG, d1, fun, nullptr);
        sym_eig_dev(ctx, G, d2, fun, nullptr);
        out.Xm = trace_formula_host(fun, d1, d2);
        out.lucky = st->lucky;
        bool done = false;
        if (j <= 2) Xstop[j - 1] = out.Xm;
        else {
            if (std::abs(out.Xm - Xstop[0]) < tol) done = true;
            else { Xstop[0] = Xstop[1]; Xstop[1] = out.Xm; }
        }
        if (done || st->lucky) break;
    }
    out.iter = std::min(j, it);
    wide_prof().report("trace_fun_update");
    return out;
}

}  // namespace kr

// result of the last fun_update kept until fetched
struct FunUpdateResult {
    kr::HostMat Xm;
    kr::CmMat Um;           // n x dim (device)
    int64_t n = 0, dim = 0;
    bool identity_basis = false;
    bool has_basis = false;
};

namespace kr {

struct FuInfo { int64_t dim = 0, iter = 0; int lucky = 0; int dense_fallback = 0; };

// fun_update     (fun_update.m:20-137)
inline FuInfo fun_update_general(kr_ctx* ctx, const kr_matrix* M, int64_t rk, const double* U, int64_t ldu,
                                 const HostMat& B, int fun, double tol, int64_t it, bool want_basis,
                                 FunUpdateResult& res) {
    const int64_t n = M->dev.n;
    FuInfo info;
    const bool herm = is_symmetric(B);
    if (!herm) fail(KR_ERR_UNSUPPORTED, "fun_update: the device path needs a symmetric (Hermitian) B");
    CmMat Ud(ctx, n, rk), b0(ctx, n, rk);
    upload_host_cm(ctx, U, ldu, Ud);
    KR_CUDA(cudaMemcpyAsync(b0.p(), Ud.p(), (size_t)n * rk * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    std::unique_ptr<kr_krylov> st;
    HostMat Cm, Xm;
    std::vector<HostMat> Xstop;
    int64_t j = 0;
    for (j = 1; j <= it; ++j) {
        if (j == 1) {
            st = krylov_start(ctx, M, want_basis, rk, std::move(b0));
            Cm = core_Cm(ctx, st.get(), Ud, B);
        } else {
            const double t0 = wide_prof().on ? WideProf::now() : 0.0;
            krylov_step(st.get());
            if (wide_prof().on) { KR_CUDA(cudaStreamSynchronize(ctx->stream)); wide_prof().step += WideProf::now() - t0; wide_prof().n_step++; }
        }
        if (want_basis && 2 * st->vcols >= n) {     // :85-90 dense fallback
            HostMat fA = dense_of(M), fAt = fA;
            std::vector<double> UB((size_t)n * rk, 0.0);
            for (int64_t a = 0; a < n; ++a)
                for (int64_t q = 0; q < rk; ++q) {
                    double s = 0;
                    for (int64_t p = 0; p < rk; ++p) s += U[a + p * ldu] * B(p, q);
                    UB[(size_t)(a + q * n)] = s;
                }
            for (int64_t b = 0; b < n; ++b)
                for (int64_t q = 0; q < rk; ++q) {
                    const double u = U[b + q * ldu];
                    if (u == 0.0) continue;
                    for (int64_t a = 0; a < n; ++a) fAt(a, b) += UB[(size_t)(a + q * n)] * u;
                }
            std::vector<double> ev;
            HostMat F1, F0;
            sym_eig_dev(ctx, fAt, ev, fun, &F1);
            sym_eig_dev(ctx, fA, ev, fun, &F0);
            Xm = HostMat(n, n);
            for (size_t e = 0; e < Xm.a.size(); ++e) Xm.a[e] = F1.a[e] - F0.a[e];
            res.Xm = Xm;
            res.n = n; res.dim = n; res.identity_basis = true; res.has_basis = true;
            info.dim = n; info.iter = j; info.lucky = st->lucky; info.dense_fallback = 1;
            return info;
        }
        const int64_t nn = st->H.rows - rk;
        HostMat G(nn, nn), tG(nn, nn);
        for (int64_t a = 0; a < nn; ++a)
            for (int64_t b = 0; b < nn; ++b) {
                double g = 0.5 * (st->H(a, b) + st->H(b, a));                  // :94
                double cm = (a < rk && b < rk) ? 0.5 * (Cm(a, b) + Cm(b, a)) : 0.0;
                G(a, b) = g;
                tG(a, b) = g + cm;
            }
        std::vector<double> ev;
        HostMat F1, F0;
        sym_eig_dev(ctx, tG, ev, fun, &F1);
        sym_eig_dev(ctx, G, ev, fun, &F0);
        Xm = HostMat(nn, nn);
        for (size_t e = 0; e < Xm.a.size(); ++e) Xm.a[e] = F1.a[e] - F0.a[e];   // :106
        info.lucky = st->lucky;
        bool done = false;
        if (j <= 2) Xstop.push_back(Xm);
        else {
            HostMat D = Xm;                                                      // Xm - pad(Xstop{1})
            for (int64_t a = 0; a < Xstop[0].rows; ++a)
                for (int64_t b = 0; b < Xstop[0].cols; ++b) D(a, b) -= Xstop[0](a, b);
            // symmetric difference: ||.||_2 = max |eig|
            HostMat Ds(nn, nn);
            for (int64_t a = 0; a < nn; ++a)
                for (int64_t b = 0; b < nn; ++b) Ds(a, b) = 0.5 * (D(a, b) + D(b, a));
            if (sym_norm2_dev(ctx, Ds) < tol) done = true;
            else { Xstop[0] = Xstop[1]; Xstop[1] = Xm; }
        }
        if (done || st->lucky) break;
    }
    info.iter = std::min(j, it);
    info.dim = Xm.rows;
    res.Xm = Xm;
    res.n = n;
    res.dim = Xm.rows;
    res.identity_basis = false;
    res.has_basis = true;
    // Um(:, 1:size(Xm,1))  (:137) - for Lanczos this is the 2-block window, as in the reference
    const int64_t keep = std::min<int64_t>(st->vcols, Xm.rows);
    res.Um = CmMat(ctx, n, Xm.rows);
    res.Um.buf.zero();
    KR_CUDA(cudaMemcpyAsync(res.Um.p(), st->V.p(), (size_t)n * keep * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    KR_CUDA(cudaStreamSynchronize(ctx->stream));
    wide_prof().report("fun_update");
    return info;
}

// MATLAB normest(S, tol): power iteration on S'S from the column abs-sums
inline void normest_dev(kr_ctx* ctx, const kr_matrix* M, double tol, double* est, int64_t* count) {
    const CsrHost& H = M->host;
    const int64_t n = H.n;
    std::vector<double> x(n, 0.0);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t p = H.row_ptr[i]; p < H.row_ptr[i + 1]; ++p) x[H.col[p]] += std::abs(H.val[p]);
    double e = 0;
    for (double v : x) e += v * v;
    e = std::sqrt(e);
    int64_t cnt = 0;
    if (e == 0) { *est = 0; *count = 0; return; }
    for (auto& v : x) v /= e;
    CmMat dx(ctx, n, 1), dSx(ctx, n, 1), dy(ctx, n, 1);
    upload_host_cm(ctx, x.data(), n, dx);
    double e0 = 0;
    while (std::abs(e - e0) > tol * e) {
        e0 = e;
        spmm_cm(ctx, M, dx.p(), 1, dSx.p());                         // Sx = S*x
        double nSx = 0, nx = 0;
        KR_CUBLAS(cublasDnrm2(ctx->cublas, (int)n, dSx.p(), 1, &nSx));
        if (nSx == 0) fail(KR_ERR_UNSUPPORTED, "normest: S*x vanished");
        // x = S'*Sx
        {
            PanelBuf xb(ctx, n, 1), yb(ctx, n, 1);
            cm_to_panel(ctx, dSx.p(), n, xb);
            EpiPlain epi{yb.p(), xb.p(), 1.0, 0.0};
            launch_spmm(ctx, M->T(), xb.p(), 1, epi, nullptr, 1);
            panel_to_cm(ctx, yb, dy.p(), n);
        }
        KR_CUBLAS(cublasDnrm2(ctx->cublas, (int)n, dy.p(), 1, &nx));
        e = nx / nSx;
        const double inv = 1.0 / nx;
        KR_CUDA(cudaMemcpyAsync(dx.p(), dy.p(), (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        KR_CUBLAS(cublasDscal(ctx->cublas, (int)n, &inv, dx.p(), 1));
        cnt += 1;
        if (cnt > 100) break;
    }
    *est = e;
    *count = cnt;
}

}  // namespace kr
