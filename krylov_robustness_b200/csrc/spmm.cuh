// spmm.cuh - CSR x dense-block product over k-wide fp64 blocks, with fused epilogues.
//
// Dense blocks on the device are PANEL-MAJOR: the k columns are cut into panels of PW = 8 columns;
// panel p is a contiguous row-major n x 8 array (64 B per row).  Element (i, c) lives at
// base[(c / 8) * n * 8 + i * 8 + (c % 8)].  A nonzero A(r, c) therefore gathers ONE 64 B row-tile
// of X per panel: four lanes x one 128-bit load.  A panel of X (64 MB at n = 1M) is what must stay
// L2-resident while the CSR arrays stream past it; blockIdx.y (slow index) walks the panels.
//
// Scheduling: one CTA (8 warps) per RowTile (csr.cuh).  The CTA first streams the tile's row
// pointers / row numbers / column indices (and values) into shared memory with coalesced loads, so
// the only dependent global loads left are the X gathers; inside a tile every row gets L = 4/8/16/32
// lanes (L/4 nonzeros in flight per step, 4-way unrolled).  Rows with >= 1024 nonzeros get a whole
// CTA.  Partial sums are folded with warp shuffles / shared memory in a fixed order, so results are
// bit-reproducible run to run (no float atomics).
//
// HBM roofline: algorithmic bytes per launch = 12*nnz + 4*(n+1) + 16*n*k.
#pragma once
#include "csr.cuh"

namespace kr {

constexpr int PW = 8;              // panel width (columns)
constexpr int SPMM_THREADS = 256;  // 8 warps
constexpr int SPMM_WARPS = SPMM_THREADS / 32;

struct PanelBlock {                // non-owning view of a panel-major block
    double* p;
    int64_t n;
    int cols;                      // logical columns
    int panels;                    // ceil(cols / 8)
    __host__ __device__ double* panel(int q) const { return p + (int64_t)q * n * PW; }
};

__device__ __forceinline__ double2 ld_x(const double* p) {      // gathered operand: keep in L1/L2
    return __ldg(reinterpret_cast<const double2*>(p));
}
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double* p, double2 v) {
    __stcs(reinterpret_cast<double2*>(p), v);
}
__device__ __forceinline__ double2 shfl_xor2(double2 v, int off) {
    v.x = __shfl_xor_sync(0xffffffffu, v.x, off);
    v.y = __shfl_xor_sync(0xffffffffu, v.y, off);
    return v;
}

// Sum NV per-thread values over all threads of the CTA that share (lane & 3), deterministically,
// and hand the totals to lanes 0..3 of warp 0.  smem: SPMM_WARPS * 4 * NV doubles.
template <int NV>
__device__ __forceinline__ void cta_reduce_by_sub(double (&v)[NV], double* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
        x += __shfl_xor_sync(0xffffffffu, x, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 8);
        x += __shfl_xor_sync(0xffffffffu, x, 16);
        v[i] = x;
    }
    if (lane < 4) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[(warp * 4 + lane) * NV + i] = v[i];
    }
    __syncthreads();
    if (warp == 0 && lane < 4) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < SPMM_WARPS; ++w) s += smem[(w * 4 + lane) * NV + i];
            v[i] = s;
        }
    }
}

// ------------------------------------------------------------------------------- epilogues
// An epilogue sees, for one (row, 2 columns) pair, the finished dot products y and may read other
// row-local operands.  finish() runs once per CTA.

// Y = alpha * (A*X - mu*X)
struct EpiPlain {
    double* __restrict__ Y;        // panel base
    const double* __restrict__ X;  // panel base (for the shift term)
    double alpha, mu;
    __device__ __forceinline__ void init(int q, int64_t stride) {
        Y += q * stride;
        X += q * stride;
    }
    double2 xr;
    __device__ __forceinline__ void pre(int r, int sub) {
        if (mu != 0.0) xr = *reinterpret_cast<const double2*>(X + (int64_t)r * PW + sub * 2);
    }
    __device__ __forceinline__ void row(int r, int sub, double2 y, unsigned) {
        if (mu != 0.0) {
            y.x -= mu * xr.x;
            y.y -= mu * xr.y;
        }
        y.x *= alpha;
        y.y *= alpha;
        st_stream(Y + (int64_t)r * PW + sub * 2, y);
    }
    __device__ __forceinline__ void finish(int, int, double*) {}
};

// Y = A*X and partial[tile][col] = sum_rows X(r,col) * Y(r,col)      (Lanczos alpha, SLQ path)
struct EpiDot {
    double* __restrict__ Y;
    const double* __restrict__ X;
    double* __restrict__ partial;  // [ntiles][panels*8]
    int total_cols;                // panels * 8
    double acc[2];
    __device__ __forceinline__ void init(int q, int64_t stride) {
        Y += q * stride;
        X += q * stride;
        acc[0] = acc[1] = 0.0;
    }
    double2 xr;
    __device__ __forceinline__ void pre(int r, int sub) {
        xr = *reinterpret_cast<const double2*>(X + (int64_t)r * PW + sub * 2);
    }
    __device__ __forceinline__ void row(int r, int sub, double2 y, unsigned) {
        acc[0] += xr.x * y.x;
        acc[1] += xr.y * y.y;
        st_stream(Y + (int64_t)r * PW + sub * 2, y);
    }
    __device__ __forceinline__ void finish(int tile, int panel, double* smem) {
        cta_reduce_by_sub<2>(acc, smem);
        if (threadIdx.x < 4) {
            double* o = partial + (int64_t)tile * total_cols + panel * PW + threadIdx.x * 2;
            o[0] = acc[0];
            o[1] = acc[1];
        }
    }
};

// Candidate pairs (columns 2c, 2c+1): Y = A*C and raw Gram blocks P'Y, C'Y (2x2 each)
// partial[tile][cand][8] = {P'Y (row-major 2x2), C'Y (row-major 2x2)}
struct EpiGram2 {
    double* __restrict__ Y;
    const double* __restrict__ P;  // previous block panel base (may be nullptr at step 1)
    const double* __restrict__ C;  // current block panel base (== X)
    double* __restrict__ partial;  // [ntiles][ncand_padded][8]
    int ncand;                     // panels * 4
    double acc[8];
    __device__ __forceinline__ void init(int q, int64_t stride) {
        Y += q * stride;
        C += q * stride;
        if (P) P += q * stride;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0;
    }
    double2 cr, pr;
    __device__ __forceinline__ void pre(int r, int sub) {
        const int64_t o = (int64_t)r * PW + sub * 2;
        cr = *reinterpret_cast<const double2*>(C + o);
        if (P) pr = *reinterpret_cast<const double2*>(P + o);
    }
    __device__ __forceinline__ void row(int r, int sub, double2 y, unsigned) {
        const int64_t o = (int64_t)r * PW + sub * 2;
        const double2 c = cr;
        if (P) {
            const double2 p = pr;
            acc[0] += p.x * y.x; acc[1] += p.x * y.y;
            acc[2] += p.y * y.x; acc[3] += p.y * y.y;
        }
        acc[4] += c.x * y.x; acc[5] += c.x * y.y;
        acc[6] += c.y * y.x; acc[7] += c.y * y.y;
        st_stream(Y + o, y);
    }
    __device__ __forceinline__ void finish(int tile, int panel, double* smem) {
        cta_reduce_by_sub<8>(acc, smem);
        if (threadIdx.x < 4) {
            double* o = partial + ((int64_t)tile * ncand + panel * 4 + threadIdx.x) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = acc[i];
        }
    }
};

// expmv Taylor term (functions/expmv.m:77-80): b' = coef*(A*b - mu*b); f += b';
// rowabs_b[panel][r] = sum over the panel's 8 columns of |b'|, same for f (matrix inf-norms are the
// max over rows of the sum over panels; reduced by expmv_norms_kernel).
// Skips all work once *done != 0 (early termination decided on the device).
struct EpiTaylor {
    double* __restrict__ Bn;       // output b' panel
    const double* __restrict__ Bo; // input b panel (== X)
    double* __restrict__ F;        // f panel (read-modify-write)
    double* __restrict__ rab;      // rowabs of b' for this panel [n]
    double* __restrict__ raf;      // rowabs of f  for this panel [n]
    double coef, mu;
    __device__ __forceinline__ void init(int q, int64_t stride) {
        Bn += q * stride;
        Bo += q * stride;
        F += q * stride;
        rab += q * (stride / PW);
        raf += q * (stride / PW);
    }
    double2 xr, fr;
    __device__ __forceinline__ void pre(int r, int sub) {
        const int64_t o = (int64_t)r * PW + sub * 2;
        if (mu != 0.0) xr = *reinterpret_cast<const double2*>(Bo + o);
        fr = *reinterpret_cast<const double2*>(F + o);
    }
    __device__ __forceinline__ void row(int r, int sub, double2 y, unsigned m) {
        const int64_t o = (int64_t)r * PW + sub * 2;
        if (mu != 0.0) {
            y.x -= mu * xr.x;
            y.y -= mu * xr.y;
        }
        y.x *= coef;
        y.y *= coef;
        double2 f = fr;
        f.x += y.x;
        f.y += y.y;
        *reinterpret_cast<double2*>(F + o) = f;
        st_stream(Bn + o, y);
        double ab = fabs(y.x) + fabs(y.y), af = fabs(f.x) + fabs(f.y);
        // the four slot-0 lanes of a row hold its 8 columns: fold over sub.  m = ballot of the
        // lanes that entered the epilogue (the four subs of a row always enter together).
        ab += __shfl_xor_sync(m, ab, 1);
        af += __shfl_xor_sync(m, af, 1);
        ab += __shfl_xor_sync(m, ab, 2);
        af += __shfl_xor_sync(m, af, 2);
        if (sub == 0) {
            rab[r] = ab;
            raf[r] = af;
        }
    }
    __device__ __forceinline__ void finish(int, int, double*) {}
};

// ------------------------------------------------------------------------------- kernel
struct SpmmSmem {
    int rp[SPMM_MAX_ROWS + 1];     // tile-relative row pointers
    int rid[SPMM_MAX_ROWS];        // original row numbers
    int col[SPMM_CAP];
};

// Multi-row tile: indices are already staged in shared memory; every row gets L lanes.
template <int L, bool HAS_VAL, class Epi>
__device__ __forceinline__ void spmm_tile(const RowTile t, const SpmmSmem& sm, const double* __restrict__ sval,
                                          const double* __restrict__ Xp, double uval, Epi& epi) {
    constexpr int S = L / 4;            // nonzero slots per row per step
    constexpr int RPW = 32 / L;         // rows per warp per pass
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 3;
    const int slot = (lane % L) >> 2;
    const int rlocal = lane / L;
    const double* xs = Xp + sub * 2;
    for (int base = warp * RPW; base < t.count; base += SPMM_WARPS * RPW) {
        const int rr = base + rlocal;
        const bool valid = rr < t.count;
        int row = 0, p0 = 0, p1 = 0;
        if (valid) {
            row = sm.rid[rr];
            p0 = sm.rp[rr];
            p1 = sm.rp[rr + 1];
        }
        const bool emit = valid && slot == 0;
        if (emit) epi.pre(row, sub);    // row-local operands: issue their loads before the gathers
        double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
        for (int p = p0 + slot; p < p1; p += 4 * S) {
            const bool b1 = p + S < p1, b2 = p + 2 * S < p1, b3 = p + 3 * S < p1;
            const int c0 = sm.col[p];
            const int c1 = b1 ? sm.col[p + S] : c0;
            const int c2 = b2 ? sm.col[p + 2 * S] : c0;
            const int c3 = b3 ? sm.col[p + 3 * S] : c0;
            double v0 = 1.0, v1 = b1 ? 1.0 : 0.0, v2 = b2 ? 1.0 : 0.0, v3 = b3 ? 1.0 : 0.0;
            if (HAS_VAL) {
                v0 = sval[p];
                if (b1) v1 = sval[p + S];
                if (b2) v2 = sval[p + 2 * S];
                if (b3) v3 = sval[p + 3 * S];
            }
            const double2 x0 = ld_x(xs + (int64_t)c0 * PW);
            const double2 x1 = ld_x(xs + (int64_t)c1 * PW);
            const double2 x2 = ld_x(xs + (int64_t)c2 * PW);
            const double2 x3 = ld_x(xs + (int64_t)c3 * PW);
            a0.x = fma(v0, x0.x, a0.x); a0.y = fma(v0, x0.y, a0.y);
            a1.x = fma(v1, x1.x, a1.x); a1.y = fma(v1, x1.y, a1.y);
            a2.x = fma(v2, x2.x, a2.x); a2.y = fma(v2, x2.y, a2.y);
            a3.x = fma(v3, x3.x, a3.x); a3.y = fma(v3, x3.y, a3.y);
        }
        double2 acc = make_double2((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
#pragma unroll
        for (int off = 4; off < L; off <<= 1) {
            double2 o = shfl_xor2(acc, off);
            acc.x += o.x;
            acc.y += o.y;
        }
        if (!HAS_VAL) {
            acc.x *= uval;
            acc.y *= uval;
        }
        const unsigned m = __ballot_sync(0xffffffffu, emit);
        if (emit) epi.row(row, sub, acc, m);
    }
}

// One long row processed by the whole CTA: 64 nonzero slots stride the row straight from global
// memory (coalesced index loads), partial sums meet in shared memory in a fixed order.
template <bool HAS_VAL, class Epi>
__device__ __forceinline__ void spmm_long_row(const CsrDevView& A, const RowTile t, const double* __restrict__ Xp,
                                              double* red, Epi& epi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 3;
    const int slot = threadIdx.x >> 2;                 // 0..63
    constexpr int S = SPMM_THREADS / 4;
    const int row = __ldg(A.row_order + t.start);
    const int p0 = __ldg(A.row_ptr + t.start), p1 = __ldg(A.row_ptr + t.start + 1);
    const bool emit = threadIdx.x < 4;
    if (emit) epi.pre(row, sub);
    const double* xs = Xp + sub * 2;
    double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
    for (int p = p0 + slot; p < p1; p += 4 * S) {
        const bool b1 = p + S < p1, b2 = p + 2 * S < p1, b3 = p + 3 * S < p1;
        const int c0 = ld_stream(A.col + p);
        const int c1 = b1 ? ld_stream(A.col + p + S) : c0;
        const int c2 = b2 ? ld_stream(A.col + p + 2 * S) : c0;
        const int c3 = b3 ? ld_stream(A.col + p + 3 * S) : c0;
        double v0 = 1.0, v1 = b1 ? 1.0 : 0.0, v2 = b2 ? 1.0 : 0.0, v3 = b3 ? 1.0 : 0.0;
        if (HAS_VAL) {
            v0 = ld_stream(A.val + p);
            if (b1) v1 = ld_stream(A.val + p + S);
            if (b2) v2 = ld_stream(A.val + p + 2 * S);
            if (b3) v3 = ld_stream(A.val + p + 3 * S);
        }
        const double2 x0 = ld_x(xs + (int64_t)c0 * PW);
        const double2 x1 = ld_x(xs + (int64_t)c1 * PW);
        const double2 x2 = ld_x(xs + (int64_t)c2 * PW);
        const double2 x3 = ld_x(xs + (int64_t)c3 * PW);
        a0.x = fma(v0, x0.x, a0.x); a0.y = fma(v0, x0.y, a0.y);
        a1.x = fma(v1, x1.x, a1.x); a1.y = fma(v1, x1.y, a1.y);
        a2.x = fma(v2, x2.x, a2.x); a2.y = fma(v2, x2.y, a2.y);
        a3.x = fma(v3, x3.x, a3.x); a3.y = fma(v3, x3.y, a3.y);
    }
    double v[2] = {(a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y)};
    cta_reduce_by_sub<2>(v, red);                      // totals in threads 0..3
    double2 acc = make_double2(v[0], v[1]);
    if (!HAS_VAL) {
        acc.x *= A.uval;
        acc.y *= A.uval;
    }
    (void)warp;
    const unsigned m = __ballot_sync(0xffffffffu, emit);
    if (emit) epi.row(row, sub, acc, m);
    __syncthreads();                                   // red[] is reused by epi.finish
}

// grid = (ntiles, panels); X panel q at X + q*n*8.  `done` (may be null): skip everything if set.
// Dynamic shared memory: SPMM_CAP doubles for the staged values when the matrix carries values.
template <class Epi, bool HAS_VAL>
__global__ void __launch_bounds__(SPMM_THREADS)
spmm_kernel(CsrDevView A, const double* __restrict__ X, Epi epi_proto, int64_t panel_stride,
            const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double red[SPMM_WARPS * 4 * 8];
    __shared__ SpmmSmem sm;
    extern __shared__ double sval[];
    const int tile = blockIdx.x, panel = blockIdx.y;
    const RowTile t = A.tiles[tile];
    Epi epi = epi_proto;
    epi.init(panel, panel_stride);
    const double* Xp = X + (int64_t)panel * panel_stride;
    if (t.lanes_log2 == 6) {
        spmm_long_row<HAS_VAL>(A, t, Xp, red, epi);
    } else {
        // stage the tile's row pointers, row numbers and column indices (coalesced streams)
        const int pb = __ldg(A.row_ptr + t.start);
        for (int i = threadIdx.x; i <= t.count; i += SPMM_THREADS) sm.rp[i] = __ldg(A.row_ptr + t.start + i) - pb;
        for (int i = threadIdx.x; i < t.count; i += SPMM_THREADS) sm.rid[i] = __ldg(A.row_order + t.start + i);
        const int nz = __ldg(A.row_ptr + t.start + t.count) - pb;
        for (int p = threadIdx.x; p < nz; p += SPMM_THREADS) sm.col[p] = ld_stream(A.col + pb + p);
        if (HAS_VAL)
            for (int p = threadIdx.x; p < nz; p += SPMM_THREADS) sval[p] = ld_stream(A.val + pb + p);
        __syncthreads();
        switch (t.lanes_log2) {
            case 2: spmm_tile<4, HAS_VAL>(t, sm, sval, Xp, A.uval, epi); break;
            case 3: spmm_tile<8, HAS_VAL>(t, sm, sval, Xp, A.uval, epi); break;
            case 4: spmm_tile<16, HAS_VAL>(t, sm, sval, Xp, A.uval, epi); break;
            default: spmm_tile<32, HAS_VAL>(t, sm, sval, Xp, A.uval, epi); break;
        }
    }
    epi.finish(tile, panel, red);
}

// Host-side launcher.  Epilogues carry panel-0 pointers; init() advances them to the CTA's panel.
template <class Epi>
inline void launch_spmm(kr_ctx* ctx, const CsrDev& A, const double* X, int panels, const Epi& epi,
                        const int* done = nullptr, int logical_cols = -1) {
    if (A.ntiles == 0 || panels == 0) return;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->timing) {
        KR_CUDA(cudaEventCreate(&e0));
        KR_CUDA(cudaEventCreate(&e1));
        KR_CUDA(cudaEventRecord(e0, ctx->stream));
    }
    dim3 grid((unsigned)A.ntiles, (unsigned)panels);
    if (A.pattern_only) {
        spmm_kernel<Epi, false><<<grid, SPMM_THREADS, 0, ctx->stream>>>(A.view(), X, epi, (int64_t)A.n * PW, done);
    } else {
        static bool attr_set = false;           // one flag per Epi instantiation
        const size_t dyn = SPMM_CAP * sizeof(double);
        if (!attr_set) {
            KR_CUDA(cudaFuncSetAttribute(spmm_kernel<Epi, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            attr_set = true;
        }
        spmm_kernel<Epi, true><<<grid, SPMM_THREADS, dyn, ctx->stream>>>(A.view(), X, epi, (int64_t)A.n * PW, done);
    }
    check_launch(ctx, "spmm_kernel");
    if (ctx->timing) {
        KR_CUDA(cudaEventRecord(e1, ctx->stream));
        ctx->spmm_events.emplace_back(e0, e1);
    }
    ctx->counters[1] += 1;
    ctx->counters[2] += logical_cols >= 0 ? logical_cols : (int64_t)panels * PW;
}

}  // namespace kr
