// entries.cuh - function_multiple_entries (functions/function_multiple_entries.m): f(A)(i,j) for a
// list of index pairs.  The reference builds one bs = 1 Arnoldi space per DISTINCT row index and
// advances them one after the other (:86-110); here every distinct row index is one column of a
// panel-major block, so a step of ALL spaces is one n x R SpMM plus fused per-column
// Gram/update passes (CGS2 against the full basis + the third pass of arnoldi_krylov.m:104-106).
// The projected problem f(H) e1 of every column is solved by one Jacobi CTA (smalldense.cuh).
//
// Deviation noted in DESIGN.md: the reference applies expm/funm to the unsymmetrised Hessenberg
// matrix (:106,:118); A is symmetric on this path so H is tridiagonal up to rounding and the device
// uses (H + H')/2.  The difference is O(eps * |f(H)|), far below the 1e-10 parity bar.
#pragma once
#include "dense.cuh"
#include "smalldense.cuh"

namespace kr {

constexpr int MD_CHUNK = 8;
struct PtrChunk { const double* v[MD_CHUNK]; };

// partial[(rb*nl + l)*tc + col] = sum_rows V_l(r,col) * W(r,col),  l < nl <= 8
__global__ void __launch_bounds__(COL_THREADS)
multi_dot_kernel(PtrChunk V, int nl, const double* __restrict__ W, int64_t n, int tc, double* __restrict__ partial) {
    __shared__ double smem[COL_WARPS * LPT * 2];
    const int q = blockIdx.y, sub = threadIdx.x % LPT;
    const int64_t po = (int64_t)q * n * PW;
    const int64_t r0 = (int64_t)blockIdx.x * COL_ROWS_PER_CTA;
    const int64_t r1 = min(n, r0 + COL_ROWS_PER_CTA);
    double acc[MD_CHUNK][2];
#pragma unroll
    for (int l = 0; l < MD_CHUNK; ++l) acc[l][0] = acc[l][1] = 0.0;
    for (int64_t r = r0 + (threadIdx.x / LPT); r < r1; r += COL_THREADS / LPT) {
        const int64_t o = po + r * PW + sub * 2;
        double2 w = *reinterpret_cast<const double2*>(W + o);
#pragma unroll
        for (int l = 0; l < MD_CHUNK; ++l)
            if (l < nl) {
                double2 v = *reinterpret_cast<const double2*>(V.v[l] + o);
                acc[l][0] += v.x * w.x;
                acc[l][1] += v.y * w.y;
            }
    }
#pragma unroll
    for (int l = 0; l < MD_CHUNK; ++l) {
        if (l < nl) {                        // uniform across the CTA
            double a[2] = {acc[l][0], acc[l][1]};
            cta_reduce_by_sub<2>(a, smem);
            if (threadIdx.x < LPT) {
                double* o = partial + ((int64_t)blockIdx.x * nl + l) * tc + q * PW + threadIdx.x * 2;
                o[0] = a[0];
                o[1] = a[1];
            }
            __syncthreads();
        }
    }
}

// W(r,col) -= sum_l h[l][col] * V_l(r,col)
__global__ void __launch_bounds__(COL_THREADS)
multi_axpy_kernel(PtrChunk V, int nl, double* __restrict__ W, int64_t n, int tc, const double* __restrict__ h) {
    const int q = blockIdx.y, sub = threadIdx.x % LPT;
    const int64_t po = (int64_t)q * n * PW;
    const int col = q * PW + sub * 2;
    double hx[MD_CHUNK], hy[MD_CHUNK];
#pragma unroll
    for (int l = 0; l < MD_CHUNK; ++l) {
        hx[l] = l < nl ? h[(int64_t)l * tc + col] : 0.0;
        hy[l] = l < nl ? h[(int64_t)l * tc + col + 1] : 0.0;
    }
    const int64_t r0 = (int64_t)blockIdx.x * COL_ROWS_PER_CTA;
    const int64_t r1 = min(n, r0 + COL_ROWS_PER_CTA);
    for (int64_t r = r0 + (threadIdx.x / LPT); r < r1; r += COL_THREADS / LPT) {
        const int64_t o = po + r * PW + sub * 2;
        double2 w = *reinterpret_cast<const double2*>(W + o);
#pragma unroll
        for (int l = 0; l < MD_CHUNK; ++l)
            if (l < nl) {
                double2 v = *reinterpret_cast<const double2*>(V.v[l] + o);
                w.x -= hx[l] * v.x;
                w.y -= hy[l] * v.y;
            }
        *reinterpret_cast<double2*>(W + o) = w;
    }
}

// W(r,col) *= s[col]
__global__ void __launch_bounds__(COL_THREADS)
colscale_kernel(double* __restrict__ W, int64_t n, int tc, const double* __restrict__ s) {
    const int q = blockIdx.y, sub = threadIdx.x % LPT;
    const int64_t po = (int64_t)q * n * PW;
    const int col = q * PW + sub * 2;
    const double sx = s[col], sy = s[col + 1];
    const int64_t r0 = (int64_t)blockIdx.x * COL_ROWS_PER_CTA;
    const int64_t r1 = min(n, r0 + COL_ROWS_PER_CTA);
    for (int64_t r = r0 + (threadIdx.x / LPT); r < r1; r += COL_THREADS / LPT) {
        const int64_t o = po + r * PW + sub * 2;
        double2 w = *reinterpret_cast<const double2*>(W + o);
        w.x *= sx;
        w.y *= sy;
        *reinterpret_cast<double2*>(W + o) = w;
    }
}

// Hessenberg bookkeeping, one thread per column.  Hc[col][step][0..it] holds column `step` of H
// (entries 0..step: inner products, entry step+1: the sub-diagonal r).
// mode 0: Hc[..][l0+l] = h[l]            (first CGS pass, chunk starting at basis index l0)
// mode 1: Hc[..][l0+l] += h[l]           (second pass)
// mode 2: r = sqrt(nrm2); Hc[..][j+1] = r; inv[col] = 1/r
// mode 3: Hc[..][l0+l] += h[l] * r       (third pass, arnoldi_krylov.m:106)
__global__ void entries_accum_kernel(double* __restrict__ Hc, int it1, int j, int l0, int nl, int tc, int R,
                                     const double* __restrict__ h, double* __restrict__ inv, int mode) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= R) return;
    double* col = Hc + ((int64_t)c * it1 + j) * (it1 + 1);
    if (mode == 2) {
        double r = sqrt(h[c]);
        col[j + 1] = r;
        inv[c] = r > 0.0 ? 1.0 / r : 0.0;
        return;
    }
    const double r = mode == 3 ? col[j + 1] : 1.0;
    for (int l = 0; l < nl; ++l) {
        double v = h[(int64_t)l * tc + c] * r;
        if (mode == 0) col[l0 + l] = v;
        else col[l0 + l] += v;
    }
}

// e_h into column c of the start block
__global__ void entries_init_kernel(double* __restrict__ V0, int64_t n, const int64_t* __restrict__ rows, int R) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= R) return;
    V0[(int64_t)(c / PW) * n * PW + (rows[c] - 1) * PW + (c % PW)] = 1.0;
}

// lag-3 stopping test on ||x - pad(x_{jj-3})||_2 (function_multiple_entries.m:121-151) for the first column x of step jj;
// hc: ring of the last four first-columns [4][it1], x is stored in it.  Called by all JAC_THREADS threads of a CTA.
__device__ __forceinline__ bool entries_stop_test(const double* x, int it1, int jj, double tol, double* hc, JacobiShared* sh) {
    double err2 = 0.0;
    if (jj > 3) {
        const double* old = hc + (int64_t)((jj - 3) & 3) * it1;      // first column of step jj-3 (size jj-3)
        for (int i = threadIdx.x; i < jj; i += JAC_THREADS) {
            double d = x[i] - (i < jj - 3 ? old[i] : 0.0);
            err2 += d * d;
        }
    }
    err2 = block_sum(err2, sh->red);
    for (int i = threadIdx.x; i < jj; i += JAC_THREADS) hc[(int64_t)(jj & 3) * it1 + i] = x[i];
    return jj > 3 && !(sqrt(err2) > tol);
}

// x = f(Hsym) e1 for the jj x jj projection (jj = j+1 steps done, 1-based step number) and the lag-3 stopping test on
// ||x - pad(x_{jj-3})||_2 (function_multiple_entries.m:121-151).  Called by all JAC_THREADS threads of a CTA.
// H: columns of the Hessenberg matrix, column cc at H + cc*(it1+1); hc: ring of the last four first-columns [4][it1];
// dyn: scratch of 2*jj*(jj|1) + jj doubles.  Leaves x in dyn + 2*jj*(jj|1), stores it in the ring, returns "converged".
__device__ __forceinline__ bool entries_project_step(const double* H, int it1, int jj, int fun, double tol, double* hc,
                                                     double* dyn, JacobiShared* sh) {
    const int lda = jj | 1;
    double* A = dyn;
    double* V = A + jj * lda;
    double* x = V + jj * lda;
    for (int e = threadIdx.x; e < jj * jj; e += JAC_THREADS) {
        int r = e % jj, cc = e / jj;
        // H(r, cc) is stored in column cc at index r (valid for r <= cc + 1)
        double hrc = r <= cc + 1 ? H[(int64_t)cc * (it1 + 1) + r] : 0.0;
        double hcr = cc <= r + 1 ? H[(int64_t)r * (it1 + 1) + cc] : 0.0;
        A[r + cc * lda] = 0.5 * (hrc + hcr);
        V[r + cc * lda] = r == cc ? 1.0 : 0.0;
    }
    __syncthreads();
    block_jacobi(A, jj, lda, V, lda, sh);
    for (int i = threadIdx.x; i < jj; i += JAC_THREADS) {
        double s = 0.0;
        for (int k = 0; k < jj; ++k) s += V[i + k * lda] * fun_eval(fun, A[k + k * lda]) * V[0 + k * lda];
        x[i] = s;
    }
    __syncthreads();
    return entries_stop_test(x, it1, jj, tol, hc, sh);
}

// One CTA per distinct row index.  hist[col][4][it1]: ring of the last first-columns; xfin[col][it1] + nfin[col]:
// frozen at convergence.
__global__ void __launch_bounds__(JAC_THREADS)
entries_step_kernel(const double* __restrict__ Hc, int it1, int jj, int fun, double tol, double* __restrict__ hist,
                    double* __restrict__ xfin, int* __restrict__ nfin, int* __restrict__ conv, int* __restrict__ nactive) {
    extern __shared__ double dyn[];
    __shared__ JacobiShared sh;
    const int c = blockIdx.x;
    if (conv[c]) return;
    const double* x = dyn + 2 * jj * (jj | 1);
    const bool done = entries_project_step(Hc + (int64_t)c * it1 * (it1 + 1), it1, jj, fun, tol, hist + (int64_t)c * 4 * it1, dyn, &sh);
    for (int i = threadIdx.x; i < jj; i += JAC_THREADS) xfin[(int64_t)c * it1 + i] = x[i];
    if (threadIdx.x == 0) {
        nfin[c] = jj;
        if (done) {
            conv[c] = 1;
            atomicSub(nactive, 1);
        }
    }
}

// X[pair] = sum_{l < nfin} V_l(j2, col) * xfin[col][l]        (function_multiple_entries.m:162-164)
__global__ void entries_gather_kernel(const double* const* __restrict__ Vptr, int64_t n, const int64_t* __restrict__ j2,
                                      const int* __restrict__ colof, const double* __restrict__ xfin,
                                      const int* __restrict__ nfin, int it1, int k, double* __restrict__ X) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= k) return;
    const int c = colof[p];
    const int64_t off = (int64_t)(c / PW) * n * PW + (j2[p] - 1) * PW + (c % PW);
    double s = 0.0;
    for (int l = 0; l < nfin[c]; ++l) s += Vptr[l][off] * xfin[(int64_t)c * it1 + l];
    X[p] = s;
}

// A family of independent single-vector Arnoldi spaces K_j(A, e_h), one per column of panel-major
// blocks, advanced together (arnoldi_krylov.m with bs = 1, full CGS2 + third pass).  Shared by
// function_multiple_entries and multiple_frechet_eval.
struct ArnoldiBatch {
    kr_ctx* ctx;
    const CsrDev* A;
    int64_t n;
    int R, it1, panels, tc, rb;
    std::vector<std::unique_ptr<PanelBuf>> V;      // V[0..j]: basis blocks
    DevBuf<double> partial, hbuf, inv, Hc;          // Hc[col][step][0..it1]: columns of the Hessenberg matrices

    ArnoldiBatch(kr_ctx* c, const CsrDev& A_, const std::vector<int64_t>& rows, int it) : ctx(c), A(&A_) {
        n = A_.n;
        R = (int)rows.size();
        it1 = it + 1;
        V.emplace_back(new PanelBuf(ctx, n, R));
        V[0]->buf.zero();
        panels = V[0]->panels;
        tc = panels * PW;
        rb = col_row_blocks(n);
        DevBuf<int64_t> drows(ctx, R);
        drows.upload(rows.data(), R);
        KR_LAUNCH(ctx, entries_init_kernel, (int)ceil_div(R, 128), 128, 0, V[0]->p(), n, drows.p, R);
        KR_CUDA(cudaStreamSynchronize(ctx->stream));
        partial.reset(ctx, (size_t)rb * MD_CHUNK * tc);
        hbuf.reset(ctx, (size_t)MD_CHUNK * tc);
        inv.reset(ctx, tc);
        Hc.reset(ctx, (size_t)R * it1 * (it1 + 1));
        Hc.zero();
        inv.zero();
    }

    // step j (0-based): basis V_0..V_j exists; appends V_{j+1} and column j of every H
    void step(int j) {
        V.emplace_back(new PanelBuf(ctx, n, R));
        PanelBuf& W = *V.back();
        dim3 cgrid((unsigned)rb, (unsigned)panels);
        const int rbk = (int)ceil_div(R, 128);
        EpiPlain epi{W.p(), V[j]->p(), 1.0, 0.0};
        launch_spmm(ctx, *A, V[j]->p(), panels, epi, nullptr, R);
        auto gs_pass = [&](int mode) {
            for (int l0 = 0; l0 <= j; l0 += MD_CHUNK) {
                const int nl = std::min(MD_CHUNK, j + 1 - l0);
                PtrChunk pc;
                for (int l = 0; l < MD_CHUNK; ++l) pc.v[l] = l < nl ? V[l0 + l]->p() : nullptr;
                KR_LAUNCH(ctx, multi_dot_kernel, cgrid, COL_THREADS, 0, pc, nl, W.p(), n, tc, partial.p);
                sum_partials(ctx, partial.p, rb, nl * tc, hbuf.p);
                KR_LAUNCH(ctx, entries_accum_kernel, rbk, 128, 0, Hc.p, it1, j, l0, nl, tc, R, hbuf.p, inv.p, mode);
                KR_LAUNCH(ctx, multi_axpy_kernel, cgrid, COL_THREADS, 0, pc, nl, W.p(), n, tc, hbuf.p);
            }
        };
        // NOTE: classical Gram-Schmidt computes all inner products of a pass before updating; with the
        // basis processed in chunks of 8 the coefficients of later chunks see the update of earlier ones
        // (a block-modified variant).  Both are followed by the second and third pass, so the computed
        // H agrees to rounding; the chunked form is used for j + 1 > 8 only.
        gs_pass(0);
        gs_pass(1);
        KR_LAUNCH(ctx, colnorm2_kernel, cgrid, COL_THREADS, 0, W.p(), n, partial.p, tc);
        sum_partials(ctx, partial.p, rb, tc, hbuf.p);
        KR_LAUNCH(ctx, entries_accum_kernel, rbk, 128, 0, Hc.p, it1, j, 0, 0, tc, R, hbuf.p, inv.p, 2);
        KR_LAUNCH(ctx, colscale_kernel, cgrid, COL_THREADS, 0, W.p(), n, tc, inv.p);
        gs_pass(3);
    }

    // device array of the basis block base pointers (for gather kernels)
    DevBuf<const double*> block_pointers() {
        std::vector<const double*> ptrs(V.size());
        for (size_t l = 0; l < V.size(); ++l) ptrs[l] = V[l]->p();
        DevBuf<const double*> d(ctx, ptrs.size());
        d.upload(ptrs.data(), ptrs.size());
        KR_CUDA(cudaStreamSynchronize(ctx->stream));
        return d;
    }
};

struct EntriesResult { std::vector<double> X; int64_t iter = 0; };

// rows: distinct first indices (1-based, stable order); colof[p]: column of pair p; j2[p]: second index
inline EntriesResult entries_run(kr_ctx* ctx, const kr_matrix* M, const std::vector<int64_t>& rows,
                                 const std::vector<int>& colof, const std::vector<int64_t>& j2, int fun,
                                 double tol, int it) {
    const int64_t n = M->dev.n;
    const int R = (int)rows.size(), k = (int)colof.size();
    const int it1 = it + 1;
    if (it > 110) fail(KR_ERR_UNSUPPORTED, "function_multiple_entries: it > 110 exceeds the shared-memory solver");
    ArnoldiBatch B(ctx, M->dev, rows, it);
    DevBuf<int64_t> dj2(ctx, k);
    DevBuf<int> dcolof(ctx, k);
    dj2.upload(j2.data(), k);
    dcolof.upload(colof.data(), k);
    DevBuf<double> hist(ctx, (size_t)R * 4 * it1), xfin(ctx, (size_t)R * it1);
    DevBuf<int> istate(ctx, (size_t)2 * R + 1);
    hist.zero(); xfin.zero(); istate.zero();
    int* nfin = istate.p;
    int* conv = istate.p + R;
    int* nactive = istate.p + 2 * R;
    int nact = R;
    KR_CUDA(cudaMemcpyAsync(nactive, &nact, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    static bool attr_set[64] = {};              // per device: the attribute lives in the device's primary context
    if (first_use_on_device(attr_set, ctx->device))
        KR_CUDA(cudaFuncSetAttribute(entries_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JAC_SMEM_LIMIT));
    EntriesResult out;
    int j = 0;
    for (j = 0; j < it; ++j) {                 // step j+1 of the reference
        B.step(j);
        const int jj = j + 1;
        const size_t smem = (size_t)(2 * jj * (jj | 1) + jj) * sizeof(double);
        KR_LAUNCH(ctx, entries_step_kernel, R, JAC_THREADS, smem, B.Hc.p, it1, jj, fun, tol, hist.p, xfin.p, nfin, conv, nactive);
        KR_CUDA(cudaMemcpyAsync(&nact, nactive, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        KR_CUDA(cudaStreamSynchronize(ctx->stream));
        if (nact <= 0) { ++j; break; }
    }
    out.iter = std::min(j, it);
    DevBuf<const double*> dptr = B.block_pointers();
    DevBuf<double> dX(ctx, k);
    KR_LAUNCH(ctx, entries_gather_kernel, (int)ceil_div(k, 128), 128, 0, dptr.p, n, dj2.p, dcolof.p, xfin.p, nfin, it1, k, dX.p);
    out.X = dX.to_host();
    return out;
}

}  // namespace kr
