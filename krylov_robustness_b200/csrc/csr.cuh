// csr.cuh - device-resident CSR adjacency + host-side preprocessing (row binning / tiling).
//
// Replaces MATLAB's sparse A on the path `w = A*w` (functions/lanczos_krylov.m:81,
// functions/arnoldi_krylov.m:86, functions/expmv.m:77, functions/normAm.m:20).
//
// Layout in HBM: the ROWS are stored sorted by decreasing length (`row_order[s]` = original index of
// stored row s): row_ptr int32[n+1] over the stored order, col_idx int32[nnz] (ORIGINAL column
// numbers - X and Y keep the caller's row numbering), val fp64[nnz] (omitted when every stored value
// is the same number -> 4 B per nonzero instead of 12).  `tiles` cuts the stored order into CTA work
// items: a tile is a run of rows of one lanes-per-row class whose nonzeros are CONTIGUOUS and fit the
// CTA's shared-memory staging buffer (SPMM_CAP), so the kernel streams a tile's indices with coalesced
// loads once and only the X gathers remain as dependent global loads.  Rows longer than
// SPMM_LONG_ROW get a tile (and a whole CTA) of their own.
#pragma once
#include <algorithm>
#include <numeric>

#include "common.cuh"

namespace kr {

struct RowTile {
    int start;      // offset into row_order
    int count;      // rows in this tile
    int lanes_log2; // log2 of the nonzero slots per row (0..3); 6 = one long row for the whole CTA
    int pad;
};

struct CsrHost {
    int64_t n = 0;
    std::vector<int64_t> row_ptr;
    std::vector<int32_t> col;
    std::vector<double> val;
};

struct CsrDevView {
    int n;
    int ntiles;
    const int* __restrict__ row_ptr;  // over the stored (sorted) row order
    const int* __restrict__ col;
    const double* __restrict__ val;   // nullptr => all values == uval
    double uval;
    const int* __restrict__ row_order;
    const RowTile* __restrict__ tiles;
};

#ifndef KR_SPMM_CAP
#define KR_SPMM_CAP 4096
#endif
#ifndef KR_SPMM_MAX_ROWS
#define KR_SPMM_MAX_ROWS 1024
#endif
constexpr int SPMM_CAP = KR_SPMM_CAP;             // nonzeros staged in shared memory per multi-row tile
constexpr int SPMM_MAX_ROWS = KR_SPMM_MAX_ROWS;   // rows per tile
constexpr int SPMM_LONG_ROW = 1024;   // rows at least this long are processed by a whole CTA

struct CsrDev {
    int64_t n = 0, nnz = 0;
    DevBuf<int> row_ptr, col, row_order;
    DevBuf<double> val;
    DevBuf<RowTile> tiles;
    int ntiles = 0;
    bool pattern_only = false;
    double uval = 1.0;
    CsrDevView view() const {
        CsrDevView v;
        v.n = (int)n;
        v.ntiles = ntiles;
        v.row_ptr = row_ptr.p;
        v.col = col.p;
        v.val = pattern_only ? nullptr : val.p;
        v.uval = uval;
        v.row_order = row_order.p;
        v.tiles = tiles.p;
        return v;
    }
};

inline CsrHost transpose(const CsrHost& A) {
    CsrHost T;
    T.n = A.n;
    const int64_t nnz = A.row_ptr[A.n];
    T.row_ptr.assign(A.n + 1, 0);
    T.col.resize(nnz);
    T.val.resize(nnz);
    for (int64_t p = 0; p < nnz; ++p) T.row_ptr[A.col[p] + 1]++;
    for (int64_t i = 0; i < A.n; ++i) T.row_ptr[i + 1] += T.row_ptr[i];
    std::vector<int64_t> next(T.row_ptr.begin(), T.row_ptr.end() - 1);
    for (int64_t i = 0; i < A.n; ++i)
        for (int64_t p = A.row_ptr[i]; p < A.row_ptr[i + 1]; ++p) {
            int64_t q = next[A.col[p]]++;
            T.col[q] = (int32_t)i;
            T.val[q] = A.val[p];
        }
    return T;
}

// Rows must have sorted, duplicate-free columns for the symmetry comparison to be exact.
inline void canonicalize(CsrHost& A) {
    const int64_t n = A.n;
    bool sorted = true;
    for (int64_t i = 0; i < n && sorted; ++i)
        for (int64_t p = A.row_ptr[i] + 1; p < A.row_ptr[i + 1]; ++p)
            if (A.col[p - 1] >= A.col[p]) { sorted = false; break; }
    if (sorted) return;
    CsrHost T = transpose(A);      // two transposes = stable counting sort by column
    A = transpose(T);
    // merge duplicates (sum, as MATLAB's sparse() does)
    std::vector<int64_t> rp(n + 1, 0);
    int64_t w = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t begin = w;
        for (int64_t p = A.row_ptr[i]; p < A.row_ptr[i + 1]; ++p) {
            if (w > begin && A.col[w - 1] == A.col[p]) A.val[w - 1] += A.val[p];
            else { A.col[w] = A.col[p]; A.val[w] = A.val[p]; ++w; }
        }
        rp[i + 1] = w;
    }
    A.row_ptr = rp;
    A.col.resize(w);
    A.val.resize(w);
}

inline void upload_csr(kr_ctx* ctx, const CsrHost& H, CsrDev& D) {
    const int64_t n = H.n, nnz = H.row_ptr[n];
    if (n >= (int64_t(1) << 31) - 64 || nnz >= (int64_t(1) << 31) - 64)
        fail(KR_ERR_UNSUPPORTED, "matrix too large for 32-bit device indices (n=%lld nnz=%lld)",
             (long long)n, (long long)nnz);
    D.n = n;
    D.nnz = nnz;
    std::vector<int> rp(n + 1);
    for (int64_t i = 0; i <= n; ++i) rp[i] = (int)H.row_ptr[i];
    bool uniform = nnz > 0;
    for (int64_t p = 1; p < nnz && uniform; ++p) uniform = (H.val[p] == H.val[0]);
    D.pattern_only = uniform;
    D.uval = uniform ? H.val[0] : 1.0;
    // ---- stored order: decreasing length (stable)
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return (rp[a + 1] - rp[a]) > (rp[b + 1] - rp[b]);
    });
    std::vector<int> srp(n + 1, 0);
    std::vector<int32_t> scol((size_t)std::max<int64_t>(nnz, 1));
    std::vector<double> sval(uniform ? 1 : (size_t)std::max<int64_t>(nnz, 1));
    for (int64_t s = 0; s < n; ++s) {
        const int r = order[s];
        const int len = rp[r + 1] - rp[r];
        srp[s + 1] = srp[s] + len;
        std::copy(H.col.begin() + rp[r], H.col.begin() + rp[r + 1], scol.begin() + srp[s]);
        if (!uniform) std::copy(H.val.begin() + rp[r], H.val.begin() + rp[r + 1], sval.begin() + srp[s]);
    }
    D.row_ptr.reset(ctx, n + 1);
    D.row_ptr.upload(srp.data(), n + 1);
    D.col.reset(ctx, std::max<int64_t>(nnz, 1));
    if (nnz) D.col.upload(scol.data(), nnz);
    if (!uniform) {
        D.val.reset(ctx, std::max<int64_t>(nnz, 1));
        if (nnz) D.val.upload(sval.data(), nnz);
    } else {
        D.val.free();
    }
    // ---- tiles: runs of one lane class, nonzeros <= SPMM_CAP, rows <= SPMM_MAX_ROWS; long rows alone
    auto lane_class = [](int len) {
        return len >= SPMM_LONG_ROW ? 6 : len >= 256 ? 3 : len >= 64 ? 2 : len >= 16 ? 1 : 0;
    };
    std::vector<RowTile> tiles;
    int64_t i = 0;
    while (i < n) {
        const int cls = lane_class(srp[i + 1] - srp[i]);
        int64_t j = i + 1;
        if (cls != 6) {
            int64_t acc = srp[i + 1] - srp[i];
            while (j < n && j - i < SPMM_MAX_ROWS) {
                const int len = srp[j + 1] - srp[j];
                if (lane_class(len) != cls || acc + len > SPMM_CAP) break;
                acc += len;
                ++j;
            }
        }
        tiles.push_back(RowTile{(int)i, (int)(j - i), cls, 0});
        i = j;
    }
    D.ntiles = (int)tiles.size();
    D.row_order.reset(ctx, std::max<int64_t>(n, 1));
    if (n) D.row_order.upload(order.data(), n);
    D.tiles.reset(ctx, std::max<size_t>(tiles.size(), 1));
    if (!tiles.empty()) D.tiles.upload(tiles.data(), tiles.size());
    KR_CUDA(cudaStreamSynchronize(ctx->stream));   // host staging vectors die here
}

}  // namespace kr

struct kr_matrix {
    kr_ctx* ctx = nullptr;
    kr::CsrHost host;          // kept for kr_matrix_set_edges and host-side glue
    kr::CsrDev dev;            // A
    kr::CsrDev devT;           // A' (only when !symmetric)
    bool symmetric = true;
    bool nonnegative = true;
    double trace = 0.0;
    double norm1 = 0.0;        // max column abs sum
    const kr::CsrDev& T() const { return symmetric ? dev : devT; }
};

namespace kr {

// keep_symmetric: the caller guarantees the edit preserved symmetry (kr_matrix_set_edges writes both
// triangles), so the O(nnz) transpose + comparison is skipped and ||A||_1 is the max row abs-sum.
inline void analyse_and_upload(kr_matrix* M, bool keep_symmetric = false) {
    CsrHost& H = M->host;
    canonicalize(H);
    const int64_t n = H.n;
    CsrHost T;
    if (keep_symmetric && M->symmetric) {
        M->symmetric = true;
    } else {
        T = transpose(H);
        M->symmetric = (T.row_ptr == H.row_ptr) && (T.col == H.col) && (T.val == H.val);
    }
    M->nonnegative = true;
    M->trace = 0.0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t p = H.row_ptr[i]; p < H.row_ptr[i + 1]; ++p) {
            if (H.val[p] < 0) M->nonnegative = false;
            if (H.col[p] == i) M->trace += H.val[p];
        }
    const CsrHost& C = M->symmetric ? H : T;       // column abs-sums of A = row abs-sums of A'
    M->norm1 = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double s = 0;
        for (int64_t p = C.row_ptr[i]; p < C.row_ptr[i + 1]; ++p) s += std::abs(C.val[p]);
        M->norm1 = std::max(M->norm1, s);
    }
    upload_csr(M->ctx, H, M->dev);
    if (!M->symmetric) upload_csr(M->ctx, T, M->devT);
    else { M->devT = CsrDev(); }
}

}  // namespace kr
