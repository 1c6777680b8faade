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
    const int* __restrict__ row_pos;  // inverse of row_order: stored position of original row i
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
    DevBuf<int> row_ptr, col, row_order, row_pos;
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
        v.row_pos = row_pos.p;
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
    int sorted = 1;
#pragma omp parallel for reduction(&& : sorted) schedule(static)
    for (int64_t i = 0; i < n; ++i)
        for (int64_t p = A.row_ptr[i] + 1; p < A.row_ptr[i + 1]; ++p)
            sorted = sorted && (A.col[p - 1] < A.col[p]);
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

// Host-side layout of the device CSR (everything upload_csr ships): rows in stored (length-sorted) order.
struct PreparedCsr {
    int64_t n = 0, nnz = 0;
    bool uniform = false;
    double uval = 1.0;
    std::vector<int> srp, order;
    std::vector<int32_t> scol;
    std::vector<double> sval;
    std::vector<RowTile> tiles;
};

inline int lane_class_of(int len) {
    return len >= SPMM_LONG_ROW ? 6 : len >= 256 ? 3 : len >= 64 ? 2 : len >= 16 ? 1 : 0;
}

// Pure host work (no CUDA calls): value uniformity, the stable decreasing-length row order (a counting sort,
// O(n + max length)), the re-packed index / value arrays (OpenMP over rows) and the tile list.
inline PreparedCsr prepare_csr(const CsrHost& H) {
    PreparedCsr P;
    const int64_t n = H.n, nnz = H.row_ptr[n];
    if (n >= (int64_t(1) << 31) - 64 || nnz >= (int64_t(1) << 31) - 64)
        fail(KR_ERR_UNSUPPORTED, "matrix too large for 32-bit device indices (n=%lld nnz=%lld)",
             (long long)n, (long long)nnz);
    P.n = n;
    P.nnz = nnz;
    int uniform = nnz > 0;
    const double v0 = nnz > 0 ? H.val[0] : 1.0;
#pragma omp parallel for reduction(&& : uniform) schedule(static)
    for (int64_t p = 0; p < nnz; ++p) uniform = uniform && (H.val[p] == v0);
    P.uniform = uniform != 0;
    P.uval = P.uniform ? v0 : 1.0;
    // ---- stored order: decreasing length, stable in the row index
    int maxlen = 0;
    for (int64_t i = 0; i < n; ++i) maxlen = std::max(maxlen, (int)(H.row_ptr[i + 1] - H.row_ptr[i]));
    std::vector<int64_t> start((size_t)maxlen + 2, 0);
    for (int64_t i = 0; i < n; ++i) start[(size_t)(maxlen - (int)(H.row_ptr[i + 1] - H.row_ptr[i])) + 1]++;
    for (size_t b = 1; b < start.size(); ++b) start[b] += start[b - 1];
    P.order.resize(n);
    for (int64_t i = 0; i < n; ++i) P.order[(size_t)start[(size_t)(maxlen - (int)(H.row_ptr[i + 1] - H.row_ptr[i]))]++] = (int)i;
    P.srp.assign(n + 1, 0);
    for (int64_t s = 0; s < n; ++s) {
        const int r = P.order[s];
        P.srp[s + 1] = P.srp[s] + (int)(H.row_ptr[r + 1] - H.row_ptr[r]);
    }
    P.scol.resize((size_t)std::max<int64_t>(nnz, 1));
    P.sval.resize(P.uniform ? 1 : (size_t)std::max<int64_t>(nnz, 1));
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t s = 0; s < n; ++s) {
        const int r = P.order[s];
        std::copy(H.col.begin() + H.row_ptr[r], H.col.begin() + H.row_ptr[r + 1], P.scol.begin() + P.srp[s]);
        if (!P.uniform) std::copy(H.val.begin() + H.row_ptr[r], H.val.begin() + H.row_ptr[r + 1], P.sval.begin() + P.srp[s]);
    }
    // ---- tiles: runs of one lane class, nonzeros <= SPMM_CAP, rows <= SPMM_MAX_ROWS; long rows alone
    int64_t i = 0;
    while (i < n) {
        const int cls = lane_class_of(P.srp[i + 1] - P.srp[i]);
        int64_t j = i + 1;
        if (cls != 6) {
            int64_t acc = P.srp[i + 1] - P.srp[i];
            while (j < n && j - i < SPMM_MAX_ROWS) {
                const int len = P.srp[j + 1] - P.srp[j];
                if (lane_class_of(len) != cls || acc + len > SPMM_CAP) break;
                acc += len;
                ++j;
            }
        }
        P.tiles.push_back(RowTile{(int)i, (int)(j - i), cls, 0});
        i = j;
    }
    return P;
}

inline void upload_csr(kr_ctx* ctx, const CsrHost& H, CsrDev& D) {
    const PreparedCsr P = prepare_csr(H);
    const int64_t n = P.n, nnz = P.nnz;
    D.n = n;
    D.nnz = nnz;
    D.pattern_only = P.uniform;
    D.uval = P.uval;
    D.row_ptr.reset(ctx, n + 1);
    D.row_ptr.upload(P.srp.data(), n + 1);
    D.col.reset(ctx, std::max<int64_t>(nnz, 1));
    if (nnz) D.col.upload(P.scol.data(), nnz);
    if (!P.uniform) {
        D.val.reset(ctx, std::max<int64_t>(nnz, 1));
        if (nnz) D.val.upload(P.sval.data(), nnz);
    } else {
        D.val.free();
    }
    D.ntiles = (int)P.tiles.size();
    D.row_order.reset(ctx, std::max<int64_t>(n, 1));
    if (n) D.row_order.upload(P.order.data(), n);
    std::vector<int> pos((size_t)std::max<int64_t>(n, 1), 0);
    for (int64_t s = 0; s < n; ++s) pos[(size_t)P.order[(size_t)s]] = (int)s;
    D.row_pos.reset(ctx, std::max<int64_t>(n, 1));
    if (n) D.row_pos.upload(pos.data(), n);
    D.tiles.reset(ctx, std::max<size_t>(P.tiles.size(), 1));
    if (!P.tiles.empty()) D.tiles.upload(P.tiles.data(), P.tiles.size());
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
// A == A' for a canonical (sorted, duplicate-free) CSR without building the transpose: walking the rows in
// ascending order, the entries (j, i) of row j are met in ascending i, so a per-row cursor finds the partner of
// every stored (i, j, v) in O(1) - the counting-sort transpose without its output arrays.
inline bool is_symmetric_csr(const CsrHost& H) {
    const int64_t n = H.n;
    std::vector<int64_t> cur(H.row_ptr.begin(), H.row_ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t p = H.row_ptr[i]; p < H.row_ptr[i + 1]; ++p) {
            const int32_t j = H.col[p];
            const int64_t q = cur[j];
            if (q >= H.row_ptr[j + 1] || H.col[q] != (int32_t)i || H.val[q] != H.val[p]) return false;
            cur[j] = q + 1;
        }
    return true;    // every row's cursor consumed exactly its entries (counts match because all nnz partners were found)
}

inline void analyse_and_upload(kr_matrix* M, bool keep_symmetric = false) {
    CsrHost& H = M->host;
    canonicalize(H);
    const int64_t n = H.n;
    CsrHost T;
    if (!(keep_symmetric && M->symmetric)) {
        M->symmetric = is_symmetric_csr(H);
        if (!M->symmetric) T = transpose(H);
    }
    int nonneg = 1;
    double tr = 0.0, norm1 = 0.0;
    const CsrHost& C = M->symmetric ? H : T;       // column abs-sums of A = row abs-sums of A'
    // the trace feeds expmv's shift mu = trace / n: summed sequentially in row order (diagonal entry by binary
    // search) so that it does not depend on the thread count
    for (int64_t i = 0; i < n; ++i) {
        const int32_t* b = H.col.data() + H.row_ptr[i];
        const int32_t* e = H.col.data() + H.row_ptr[i + 1];
        const int32_t* q = std::lower_bound(b, e, (int32_t)i);
        if (q != e && *q == (int32_t)i) tr += H.val[(size_t)(q - H.col.data())];
    }
#pragma omp parallel for reduction(&& : nonneg) reduction(max : norm1) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        for (int64_t p = H.row_ptr[i]; p < H.row_ptr[i + 1]; ++p)
            if (H.val[p] < 0) nonneg = 0;
        double sabs = 0;
        for (int64_t p = C.row_ptr[i]; p < C.row_ptr[i + 1]; ++p) sabs += std::abs(C.val[p]);
        norm1 = std::max(norm1, sabs);
    }
    M->nonnegative = nonneg != 0;
    M->trace = tr;
    M->norm1 = norm1;
    upload_csr(M->ctx, H, M->dev);
    if (!M->symmetric) upload_csr(M->ctx, T, M->devT);
    else { M->devT = CsrDev(); }
}

}  // namespace kr
