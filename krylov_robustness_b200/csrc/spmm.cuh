// spmm.cuh - CSR x dense-block product over k-wide fp64 blocks, with fused epilogues.
//
// Dense blocks on the device are PANEL-MAJOR: the k columns are cut into panels of PW columns
// (PW = 16 by default: 128 B = one cache line per row; PW = 8 also builds and passes the tests);
// panel p is a contiguous row-major n x PW array.  Element (i, c) lives at
// base[(c / PW) * n * PW + i * PW + (c % PW)].  A nonzero A(r, c) therefore gathers ONE row-tile of X
// per panel: LPT = PW/2 lanes x one 128-bit load, one L1 wavefront.  blockIdx.y (slow index) walks the
// panels so that one panel of X at a time competes for L2 with the streaming CSR arrays.
//
// Scheduling: one CTA (8 warps) per RowTile (csr.cuh).  The CTA first streams the tile's row
// pointers / row numbers / column indices (and values) into shared memory with coalesced loads, so
// the only dependent global loads left are the X gathers; inside a tile every row gets L = 4/8/16/32
// lanes and every lane issues U (4 or 8) independent 128-bit gathers before its first FMA.  Rows with
// >= 1024 nonzeros get a whole CTA.  Partial sums are folded with warp shuffles / shared memory in a fixed order, so results are
// bit-reproducible run to run (no float atomics).
//
// HBM roofline: algorithmic bytes per launch = 12*nnz + 4*(n+1) + 16*n*k.
#pragma once
#include "csr.cuh"

namespace kr {

#ifndef KR_PW
#define KR_PW 16
#endif
constexpr int PW = KR_PW;          // panel width (columns): 8 (64 B row-tiles) or 16 (128 B row-tiles)
constexpr int LPT = PW / 2;        // lanes per row-tile (one double2 per lane)
static_assert(PW == 8 || PW == 16, "panel width must be 8 or 16");
#ifndef KR_SPMM_THREADS
#define KR_SPMM_THREADS 256
#endif
constexpr int SPMM_THREADS = KR_SPMM_THREADS;  // warps per SpMM CTA = SPMM_THREADS / 32
constexpr int COL_WARPS = 8;                   // the streaming per-column kernels always use 256 threads
constexpr int SPMM_WARPS = SPMM_THREADS / 32;
#ifndef KR_SPMM_MIN_CTAS
#define KR_SPMM_MIN_CTAS 4
#endif
constexpr int SPMM_MIN_CTAS = KR_SPMM_MIN_CTAS;   // resident CTAs per SM the register allocation is bounded for
constexpr int SPMM_DEFAULT_UNROLL = 8;
constexpr int SPMM_DEFAULT_PANELS_PER_CTA = 1;  // consecutive panels walked with one staging of the tile's indices  // independent gathers per lane before the first FMA

struct PanelBlock {                // non-owning view of a panel-major block
    double* p;
    int64_t n;
    int cols;                      // logical columns
    int panels;                    // ceil(cols / 8)
    __host__ __device__ double* panel(int q) const { return p + (int64_t)q * n * PW; }
};

__device__ __forceinline__ double2 ld_x(const double* p) {      // gathered operand: keep in L1/L2
#if defined(KR_X_EVICT_LAST)
    // tuning build: mark gathered X lines evict_last in L2 so the streaming CSR / Y traffic cannot displace them
    double2 r;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
#elif defined(KR_X_LOAD_CG)
    return __ldcg(reinterpret_cast<const double2*>(p));          // tuning build: L2 only, no L1 allocation
#elif defined(KR_X_LOAD_NOALLOC)
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
#else
    return __ldg(reinterpret_cast<const double2*>(p));
#endif
}
// L2 prefetch of a row-tile of X that a later round will gather: the demand gathers are latency-bound
// (every lane already holds U row-tiles in flight in registers) and 39 % of them miss L2 because one
// panel of X is as large as L2; a prefetch holds no register, so the DRAM part of the latency is paid
// KR_SPMM_PF rounds ahead of the demand load (measured: 7.6 -> 6.6 ms per k=512 SpMM).
#ifndef KR_SPMM_PF
#define KR_SPMM_PF 1
#endif
#ifndef KR_SPMM_PF_EPI
#define KR_SPMM_PF_EPI 0
#endif
__device__ __forceinline__ void prefetch_x(const double* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
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

// Sum NV per-thread values over all threads of the CTA that share (lane % LPT), deterministically,
// and hand the totals to lanes 0..LPT-1 of warp 0.  smem: WARPS * LPT * NV doubles.
template <int NV, int WARPS = COL_WARPS>
__device__ __forceinline__ void cta_reduce_by_sub(double (&v)[NV], double* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int off = LPT; off < 32; off <<= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        v[i] = x;
    }
    if (lane < LPT) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[(warp * LPT + lane) * NV + i] = v[i];
    }
    __syncthreads();
    if (warp == 0 && lane < LPT) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < WARPS; ++w) s += smem[(w * LPT + lane) * NV + i];
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
    __device__ __forceinline__ void prefetch(int r) const {      // row-local operands of a later row -> L2
        if (mu != 0.0) prefetch_x(X + (int64_t)r * PW);
    }
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
    __device__ __forceinline__ void prefetch(int r) const { prefetch_x(X + (int64_t)r * PW); }
    __device__ __forceinline__ void pre(int r, int sub) {
        xr = *reinterpret_cast<const double2*>(X + (int64_t)r * PW + sub * 2);
    }
    __device__ __forceinline__ void row(int r, int sub, double2 y, unsigned) {
        acc[0] += xr.x * y.x;
        acc[1] += xr.y * y.y;
        st_stream(Y + (int64_t)r * PW + sub * 2, y);
    }
    __device__ __forceinline__ void finish(int tile, int panel, double* smem) {
        cta_reduce_by_sub<2, SPMM_WARPS>(acc, smem);
        if (threadIdx.x < LPT) {
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
    int ncand;                     // panels * LPT
    double acc[8];
    __device__ __forceinline__ void init(int q, int64_t stride) {
        Y += q * stride;
        C += q * stride;
        if (P) P += q * stride;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0;
    }
    double2 cr, pr;
    __device__ __forceinline__ void prefetch(int r) const {
        prefetch_x(C + (int64_t)r * PW);
        if (P) prefetch_x(P + (int64_t)r * PW);
    }
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
        cta_reduce_by_sub<8, SPMM_WARPS>(acc, smem);
        if (threadIdx.x < LPT) {
            double* o = partial + ((int64_t)tile * ncand + panel * LPT + threadIdx.x) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = acc[i];
        }
    }
};

// Device-side control block of one expmv call (expmv.cuh): the early-termination test of
// functions/expmv.m:83 is evaluated on the device so that a whole call is enqueued without host syncs.
struct TaylorCtl {
    double c1;                         // ||b||_inf of the previous term
    double c2, nf;                     // last ||b'||_inf, ||f||_inf (diagnostics)
    int done;                          // early-termination flag of the current stage
    int mv;                            // products actually computed
    unsigned long long mb_bits, mf_bits;   // running maxima (bit patterns of non-negative doubles)
    unsigned int ticket;               // CTAs that have finished the current term
};

__device__ __forceinline__ void taylor_decide(TaylorCtl* ctl, double c2, double nf, int full_term, double tol) {
    ctl->mv += 1;
    ctl->c2 = c2;
    ctl->nf = nf;
    if (!full_term) {
        if (ctl->c1 + c2 <= tol * nf) ctl->done = 1;
        else ctl->c1 = c2;
    }
}

// expmv Taylor term (functions/expmv.m:77-80): b' = coef*(A*b - mu*b); f += b'.  The matrix inf-norms
// (max over rows of the abs-sum over ALL columns) are needed for the termination test:
//  * one panel (q <= PW, the mc_trace case): row sums are complete inside the CTA; maxima go through
//    atomicMax on the bit patterns (max of non-negative doubles is order independent, hence still
//    deterministic) and the LAST CTA to finish evaluates the test - one launch per Taylor term;
//  * several panels: rowabs_b[panel][r] / rowabs_f[panel][r] are written and reduced by
//    taylor_norms_ctl_kernel (expmv.cuh) - two launches per term.
// Every launch returns immediately once *done != 0.
struct EpiTaylor {
    double* __restrict__ Bn;       // output b' panel
    const double* __restrict__ Bo; // input b panel (== X)
    double* __restrict__ F;        // f panel (read-modify-write)
    double* __restrict__ rab;      // rowabs of b' for this panel [n]   (multi-panel mode)
    double* __restrict__ raf;      // rowabs of f  for this panel [n]
    double coef, mu;
    TaylorCtl* ctl;                // non-null: single-panel fused control
    int total_ctas, full_term;
    double tol;
    double mb, mf;
    __device__ __forceinline__ void init(int q, int64_t stride) {
        Bn += q * stride;
        Bo += q * stride;
        F += q * stride;
        if (!ctl) {
            rab += q * (stride / PW);
            raf += q * (stride / PW);
        }
        mb = mf = 0.0;
    }
    double2 xr, fr;
    __device__ __forceinline__ void prefetch(int r) const {
        if (mu != 0.0) prefetch_x(Bo + (int64_t)r * PW);
        prefetch_x(F + (int64_t)r * PW);
    }
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
        // the LPT slot-0 lanes of a row hold its PW columns: fold over sub.  m = ballot of the lanes that
        // entered the epilogue (the subs of a row always enter together).
#pragma unroll
        for (int off = 1; off < LPT; off <<= 1) {
            ab += __shfl_xor_sync(m, ab, off);
            af += __shfl_xor_sync(m, af, off);
        }
        if (ctl) {
            mb = fmax(mb, ab);
            mf = fmax(mf, af);
        } else if (sub == 0) {
            rab[r] = ab;
            raf[r] = af;
        }
    }
    __device__ __forceinline__ void finish(int, int, double* smem) {
        if (!ctl) return;
        // CTA maxima -> global maxima -> last CTA decides
        double a = mb, b = mf;
        for (int off = 16; off; off >>= 1) {
            a = fmax(a, __shfl_xor_sync(0xffffffffu, a, off));
            b = fmax(b, __shfl_xor_sync(0xffffffffu, b, off));
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { smem[2 * warp] = a; smem[2 * warp + 1] = b; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < SPMM_WARPS; ++w) { a = fmax(a, smem[2 * w]); b = fmax(b, smem[2 * w + 1]); }
            atomicMax(&ctl->mb_bits, (unsigned long long)__double_as_longlong(a));
            atomicMax(&ctl->mf_bits, (unsigned long long)__double_as_longlong(b));
            __threadfence();
            const unsigned t = atomicAdd(&ctl->ticket, 1u);
            if (t == (unsigned)total_ctas - 1u) {
                __threadfence();
                const double c2 = __longlong_as_double((long long)atomicAdd(&ctl->mb_bits, 0ull));
                const double nf = __longlong_as_double((long long)atomicAdd(&ctl->mf_bits, 0ull));
                taylor_decide(ctl, c2, nf, full_term, tol);
                ctl->mb_bits = 0ull;
                ctl->mf_bits = 0ull;
                ctl->ticket = 0u;
                __threadfence();
            }
        }
    }
};

// ------------------------------------------------------------------------------- kernel
// dynamic shared memory carve-up (bytes): col | rp | rid | [val]
constexpr size_t SPMM_SMEM_PATTERN = SPMM_CAP * sizeof(int) + (SPMM_MAX_ROWS + 1 + SPMM_MAX_ROWS + 3) * sizeof(int);
constexpr size_t SPMM_SMEM_VALUED = SPMM_SMEM_PATTERN + SPMM_CAP * sizeof(double);

// One lane's share of a row: nonzeros p_first, p_first+stride, ... < p_end (tile-relative, indices in
// shared memory).  U independent 16-byte gathers are issued back to back before the first FMA, so a
// lane keeps U row-tiles of X in flight (measured: a cp.async ring that parks the in-flight data in
// shared memory instead of registers is SLOWER here - it triples the L1TEX data-pipe wavefronts, the
// unit that limits this kernel; profiles/r01_d_*).
template <bool HAS_VAL, int U>
__device__ __forceinline__ double2 gather_accumulate(int p_first, int p_end, int stride, const int* __restrict__ scol,
                                                     const double* __restrict__ sval, const double* __restrict__ xs,
                                                     const double* __restrict__ xpanel, int pf_sub) {
    double2 acc0 = make_double2(0.0, 0.0), acc1 = acc0;
    for (int p = p_first; p < p_end; p += U * stride) {
#if KR_SPMM_PF > 0
        {   // lane `pf_sub` of the group prefetches slot u = pf_sub of the round KR_SPMM_PF ahead
            const int qpf = p + (KR_SPMM_PF * U + pf_sub) * stride;
            if (pf_sub < U && qpf < p_end) prefetch_x(xpanel + (int64_t)scol[qpf] * PW);
        }
#endif
        double2 x[U];
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int q = p + u * stride;
            x[u] = make_double2(0.0, 0.0);
            v[u] = 1.0;
            if (q < p_end) {                                     // predicated: no index / X traffic past the row end
                x[u] = ld_x(xs + (int64_t)scol[q] * PW);
                if (HAS_VAL) v[u] = sval[q];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u & 1) {
                acc1.x = fma(v[u], x[u].x, acc1.x);
                acc1.y = fma(v[u], x[u].y, acc1.y);
            } else {
                acc0.x = fma(v[u], x[u].x, acc0.x);
                acc0.y = fma(v[u], x[u].y, acc0.y);
            }
        }
    }
    acc0.x += acc1.x;
    acc0.y += acc1.y;
    return acc0;
}

// Multi-row tile: indices are already staged in shared memory; every row gets L lanes.
template <int L, bool HAS_VAL, int U, class Epi>
__device__ __forceinline__ void spmm_tile(const RowTile t, const int* __restrict__ srp, const int* __restrict__ srid,
                                          const int* __restrict__ scol, const double* __restrict__ sval,
                                          const double* __restrict__ Xp, double uval, Epi& epi) {
    constexpr int S = L / LPT;          // nonzero slots per row
    constexpr int RPW = 32 / L;         // rows per warp per pass
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPT;
    const int slot = (lane % L) / LPT;
    const int rlocal = lane / L;
    const double* xs = Xp + sub * 2;
    for (int base = warp * RPW; base < t.count; base += SPMM_WARPS * RPW) {
        const int rr = base + rlocal;
        const bool valid = rr < t.count;
        int row = 0, p0 = 0, p1 = 0;
        if (valid) {
            row = srid[rr];
            p0 = srp[rr];
            p1 = srp[rr + 1];
        }
        const bool emit = valid && slot == 0;
        if (emit) epi.pre(row, sub);    // row-local operands: issue their loads before the gathers
#if KR_SPMM_PF > 0
        {   // the row this lane group serves KR_SPMM_PF passes from now: prefetch its first KR_SPMM_PF rounds
            const int rn = rr + KR_SPMM_PF * SPMM_WARPS * RPW;
            if (rn < t.count) {
                const int q0 = srp[rn], q1 = srp[rn + 1];
#if KR_SPMM_PF_EPI
                if (slot == 0 && sub == 0) epi.prefetch(srid[rn]);
#endif
#pragma unroll
                for (int k = 0; k < KR_SPMM_PF; ++k) {
                    const int q = q0 + slot + (k * U + sub) * S;
                    if (sub < U && q < q1) prefetch_x(Xp + (int64_t)scol[q] * PW);
                }
            }
        }
#endif
        double2 acc = gather_accumulate<HAS_VAL, U>(p0 + slot, p1, S, scol, sval, xs, Xp, sub);
#pragma unroll
        for (int off = LPT; off < L; off <<= 1) {
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

// grid = (ntiles, ceil(panels/ppc)); X panel q at X + q*n*8.  `done` (may be null): skip everything if set.
template <class Epi, bool HAS_VAL, int U>
__global__ void __launch_bounds__(SPMM_THREADS, SPMM_MIN_CTAS)
spmm_kernel(CsrDevView A, const double* __restrict__ X, Epi epi_proto, int64_t panel_stride,
            const int* __restrict__ done, int panels, int ppc, int panel0, const int* __restrict__ panel_active) {
    if (done && *done) return;
    __shared__ double red[SPMM_WARPS * LPT * 8];
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    int* scol = reinterpret_cast<int*>(dyn_smem);
    int* srp = scol + SPMM_CAP;
    int* srid = srp + SPMM_MAX_ROWS + 1;
    double* sval = reinterpret_cast<double*>(dyn_smem + SPMM_SMEM_PATTERN);
    const int tile = blockIdx.x;
    const RowTile t = A.tiles[tile];
    const int pb = __ldg(A.row_ptr + t.start);
    const int nz = __ldg(A.row_ptr + t.start + t.count) - pb;
    const bool long_row = t.lanes_log2 == 6;   // class 6: one row, whole CTA
    if (!long_row) {
        // stage the tile's row pointers, row numbers and column indices (coalesced streams) ONCE; the
        // CTA then walks `ppc` consecutive panels with them
        for (int i = threadIdx.x; i <= t.count; i += SPMM_THREADS) srp[i] = __ldg(A.row_ptr + t.start + i) - pb;
        for (int i = threadIdx.x; i < t.count; i += SPMM_THREADS) srid[i] = __ldg(A.row_order + t.start + i);
        for (int p = threadIdx.x; p < nz; p += SPMM_THREADS) scol[p] = ld_stream(A.col + pb + p);
        if (HAS_VAL)
            for (int p = threadIdx.x; p < nz; p += SPMM_THREADS) sval[p] = ld_stream(A.val + pb + p);
        __syncthreads();
    }
    const int panel_first = panel0 + blockIdx.y * ppc;
    const int panel_end = min(panels, panel_first + ppc);
    for (int panel = panel_first; panel < panel_end; ++panel) {
        if (panel_active && !panel_active[panel]) continue;   // every column of this panel has converged
        Epi epi = epi_proto;
        epi.init(panel, panel_stride);
        const double* Xp = X + (int64_t)panel * panel_stride;
        if (long_row) {
            // one long row, whole CTA: stage the indices chunk by chunk, 64 nonzero slots stride each chunk
            const int sub = threadIdx.x % LPT, slot = threadIdx.x / LPT;
            const int row = __ldg(A.row_order + t.start);
            const bool emit = threadIdx.x < LPT;
            if (emit) epi.pre(row, sub);
            double v[2] = {0.0, 0.0};
            for (int c0 = 0; c0 < nz; c0 += SPMM_CAP) {
                const int cn = min(SPMM_CAP, nz - c0);
                if (nz > SPMM_CAP || panel == panel_first) {     // a row that fits stays staged across panels
                    __syncthreads();
                    for (int p = threadIdx.x; p < cn; p += SPMM_THREADS) scol[p] = ld_stream(A.col + pb + c0 + p);
                    if (HAS_VAL)
                        for (int p = threadIdx.x; p < cn; p += SPMM_THREADS) sval[p] = ld_stream(A.val + pb + c0 + p);
                    __syncthreads();
                }
                double2 a = gather_accumulate<HAS_VAL, U>(slot, cn, SPMM_THREADS / LPT, scol, sval, Xp + sub * 2, Xp, sub);
                v[0] += a.x;
                v[1] += a.y;
            }
            cta_reduce_by_sub<2, SPMM_WARPS>(v, red);      // totals in threads 0..LPT-1
            double2 acc = make_double2(v[0], v[1]);
            if (!HAS_VAL) {
                acc.x *= A.uval;
                acc.y *= A.uval;
            }
            const unsigned m = __ballot_sync(0xffffffffu, emit);
            if (emit) epi.row(row, sub, acc, m);
            __syncthreads();                               // red[] is reused by epi.finish
        } else {
            switch (t.lanes_log2) {          // slots per row = 1 << lanes_log2, lanes per row = LPT << lanes_log2
                case 0: spmm_tile<LPT, HAS_VAL, U>(t, srp, srid, scol, sval, Xp, A.uval, epi); break;
                case 1: spmm_tile<2 * LPT, HAS_VAL, U>(t, srp, srid, scol, sval, Xp, A.uval, epi); break;
                case 2: spmm_tile<(4 * LPT > 32 ? 32 : 4 * LPT), HAS_VAL, U>(t, srp, srid, scol, sval, Xp, A.uval, epi); break;
                default: spmm_tile<32, HAS_VAL, U>(t, srp, srid, scol, sval, Xp, A.uval, epi); break;
            }
        }
        epi.finish(tile, panel, red);
        __syncthreads();                                   // red[] and the staged indices are reused by the next panel
    }
}

// Host-side launcher.  Epilogues carry panel-0 pointers; init() advances them to the CTA's panel.
template <class Epi>
inline void launch_spmm(kr_ctx* ctx, const CsrDev& A, const double* X, int panels, const Epi& epi,
                        const int* done = nullptr, int logical_cols = -1, const int* panel_active = nullptr) {
    if (A.ntiles == 0 || panels == 0) return;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->timing) {
        KR_CUDA(cudaEventCreate(&e0));
        KR_CUDA(cudaEventCreate(&e1));
        KR_CUDA(cudaEventRecord(e0, ctx->stream));
    }
    static bool attr_set = false;               // one flag per Epi instantiation
    static int variant = 4;
    static int ppc = SPMM_DEFAULT_PANELS_PER_CTA;
    if (!attr_set) {
        const char* pe = getenv("KR_SPMM_PPC");      // tuning knob: consecutive panels per CTA
        if (pe && atoi(pe) >= 1 && atoi(pe) <= 64) ppc = atoi(pe);
        KR_CUDA(cudaFuncSetAttribute(spmm_kernel<Epi, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPMM_SMEM_PATTERN));
        KR_CUDA(cudaFuncSetAttribute(spmm_kernel<Epi, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPMM_SMEM_VALUED));
        KR_CUDA(cudaFuncSetAttribute(spmm_kernel<Epi, false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPMM_SMEM_PATTERN));
        KR_CUDA(cudaFuncSetAttribute(spmm_kernel<Epi, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPMM_SMEM_VALUED));
        const char* e = getenv("KR_SPMM_UNROLL");   // tuning knob: gathers in flight per lane (4 or 8)
        variant = (e && atoi(e) == 8) ? 8 : (e && atoi(e) == 4) ? 4 : SPMM_DEFAULT_UNROLL;
        attr_set = true;
    }
    const int64_t ps = (int64_t)A.n * PW;
    auto launch = [&](dim3 grid, int npanels, int panel0) {
        if (A.pattern_only) {
            if (variant == 8) spmm_kernel<Epi, false, 8><<<grid, SPMM_THREADS, SPMM_SMEM_PATTERN, ctx->stream>>>(A.view(), X, epi, ps, done, npanels, ppc, panel0, panel_active);
            else spmm_kernel<Epi, false, 4><<<grid, SPMM_THREADS, SPMM_SMEM_PATTERN, ctx->stream>>>(A.view(), X, epi, ps, done, npanels, ppc, panel0, panel_active);
        } else {
            if (variant == 8) spmm_kernel<Epi, true, 8><<<grid, SPMM_THREADS, SPMM_SMEM_VALUED, ctx->stream>>>(A.view(), X, epi, ps, done, npanels, ppc, panel0, panel_active);
            else spmm_kernel<Epi, true, 4><<<grid, SPMM_THREADS, SPMM_SMEM_VALUED, ctx->stream>>>(A.view(), X, epi, ps, done, npanels, ppc, panel0, panel_active);
        }
    };
    launch(dim3((unsigned)A.ntiles, (unsigned)((panels + ppc - 1) / ppc)), panels, 0);
    check_launch(ctx, "spmm_kernel");
    if (ctx->timing) {
        KR_CUDA(cudaEventRecord(e1, ctx->stream));
        ctx->spmm_events.emplace_back(e0, e1);
    }
    ctx->counters[1] += 1;
    ctx->counters[2] += logical_cols >= 0 ? logical_cols : (int64_t)panels * PW;
}

}  // namespace kr
