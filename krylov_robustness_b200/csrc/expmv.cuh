// expmv.cuh - Al-Mohy & Higham's expmv on the device: normAm, select_taylor_degree, Taylor stages.
// Reference: functions/expmv.m, functions/select_taylor_degree.m, functions/normAm.m.
//
// The Taylor inner loop (expmv.m:75-88) is ONE fused SpMM launch per term (b' = coef*(A-mu I)b,
// f += b', row abs-sums of b' and f); the early-termination test c1 + c2 <= tol*||f||_inf is evaluated
// ON THE DEVICE by the last CTA to finish (inside the SpMM itself when q <= 16, else in one small
// reduction kernel) and raises a flag; later launches of the stage see the flag and return
// immediately, so a whole expmv call is enqueued without a single host round-trip.  The 1-norm power sequence of normAm (A >= 0: e <- A'e, exact) is computed
// once for all p = 1..p_max (9 SpMVs instead of 44; the reported mv stays the reference's count).
#pragma once
#include <cmath>

#include "dense.cuh"
#include "theta_table.h"

namespace kr {

// ---------------------------------------------------------------------------------- norms
// ra[panel][r] = sum of |X(r, c)| over the panel's PW columns
__global__ void rowabs_kernel(const double* __restrict__ X, int64_t n, int panels, double* __restrict__ ra) {
    const int64_t total = n * panels;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const double* x = X + e * PW;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < PW; ++c) s += fabs(x[c]);
        ra[e] = s;
    }
}

// Several panels: out of the per-panel row abs-sums, ||b'||_inf and ||f||_inf (max over rows of the sum
// over panels); the last CTA to finish evaluates the termination test (expmv.m:78-86).
// mode 0 (stage start, expmv.m:74): c1 = ||b||_inf from rab only, done = 0.
__global__ void __launch_bounds__(256)
taylor_norms_ctl_kernel(const double* __restrict__ rab, const double* __restrict__ raf, int64_t n, int panels,
                        double* __restrict__ pb, double* __restrict__ pf, TaylorCtl* ctl, int mode, int full_term,
                        double tol) {
    if (mode == 1 && ctl->done) return;
    __shared__ double red[16];
    __shared__ int last;
    double mb = 0.0, mf = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        double sb = 0.0, sf = 0.0;
        for (int q = 0; q < panels; ++q) {
            sb += rab[(int64_t)q * n + r];
            if (mode == 1) sf += raf[(int64_t)q * n + r];
        }
        mb = fmax(mb, sb);
        mf = fmax(mf, sf);
    }
    for (int off = 16; off; off >>= 1) {
        mb = fmax(mb, __shfl_xor_sync(0xffffffffu, mb, off));
        mf = fmax(mf, __shfl_xor_sync(0xffffffffu, mf, off));
    }
    if ((threadIdx.x & 31) == 0) { red[2 * (threadIdx.x >> 5)] = mb; red[2 * (threadIdx.x >> 5) + 1] = mf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mb = fmax(mb, red[2 * w]); mf = fmax(mf, red[2 * w + 1]); }
        pb[blockIdx.x] = mb;
        pf[blockIdx.x] = mf;
        __threadfence();
        last = atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double c2 = 0.0, nf = 0.0;
        for (unsigned i = 0; i < gridDim.x; ++i) {          // fixed order (max is order independent anyway)
            c2 = fmax(c2, ((volatile double*)pb)[i]);
            nf = fmax(nf, ((volatile double*)pf)[i]);
        }
        ctl->ticket = 0u;
        if (mode == 0) {
            ctl->c1 = c2;
            ctl->done = 0;
        } else {
            taylor_decide(ctl, c2, nf, full_term, tol);
        }
        __threadfence();
    }
}

// f = eta * f ; b = f        (expmv.m:91)
__global__ void stage_end_kernel(double* __restrict__ F, double* __restrict__ B, int64_t total, double eta) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        double v = eta * F[e];
        F[e] = v;
        B[e] = v;
    }
}

__global__ void fill_kernel(double* __restrict__ X, int64_t total, double v) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) X[e] = v;
}

// max |X(r, 0)| of a one-column panel block -> out[0] (single CTA; n is modest work: one pass)
__global__ void col0_absmax_kernel(const double* __restrict__ X, int64_t n, double* __restrict__ out) {
    __shared__ double red[32];
    double m = 0.0;
    for (int64_t r = threadIdx.x; r < n; r += blockDim.x) m = fmax(m, fabs(X[r * PW]));
    for (int off = 16; off; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        out[0] = m;
    }
}

// ---------------------------------------------------------------------------------- shifted-matrix facts
struct ShiftFacts {
    double norm1;      // ||A - mu I||_1
    bool nonneg;       // every entry of (A - mu I) >= 0
};

inline ShiftFacts shifted_facts(const kr_matrix* M, double mu) {
    ShiftFacts f;
    if (mu == 0.0) {
        f.norm1 = M->norm1;
        f.nonneg = M->nonnegative;
        return f;
    }
    const CsrHost& H = M->host;
    const int64_t n = H.n;
    std::vector<double> colsum(n, 0.0);
    f.nonneg = true;
    for (int64_t i = 0; i < n; ++i) {
        bool diag = false;
        for (int64_t p = H.row_ptr[i]; p < H.row_ptr[i + 1]; ++p) {
            double v = H.val[p];
            if (H.col[p] == i) { v -= mu; diag = true; }
            if (v < 0) f.nonneg = false;
            colsum[H.col[p]] += std::abs(v);
        }
        if (!diag) {
            colsum[i] += std::abs(mu);
            if (-mu < 0) f.nonneg = false;
        }
    }
    f.norm1 = 0.0;
    for (int64_t i = 0; i < n; ++i) f.norm1 = std::max(f.norm1, colsum[i]);
    return f;
}

// y = alpha * (Aop - mu I) x on one-column panel blocks
inline void spmv_panel(kr_ctx* ctx, const CsrDev& Aop, const PanelBuf& x, PanelBuf& y, double alpha, double mu) {
    EpiPlain epi{y.p(), x.p(), alpha, mu};
    launch_spmm(ctx, Aop, x.p(), 1, epi, nullptr, 1);
}

// ||((scale*(A - mu I))')^j 1||_inf for j = 1..jmax, valid when scale*(A - mu I) >= 0 (normAm.m:17-23).
inline std::vector<double> power_norms_nonneg(kr_ctx* ctx, const kr_matrix* M, double scale, double mu, int jmax) {
    const int64_t n = M->dev.n;
    PanelBuf e0(ctx, n, 1), e1(ctx, n, 1);
    e0.buf.zero();
    e1.buf.zero();
    // ones in column 0 only
    {
        std::vector<double> ones((size_t)n * PW, 0.0);
        for (int64_t i = 0; i < n; ++i) ones[i * PW] = 1.0;
        e0.buf.upload(ones.data(), ones.size());
        KR_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    DevBuf<double> out(ctx, jmax);
    PanelBuf* a = &e0;
    PanelBuf* b = &e1;
    for (int j = 0; j < jmax; ++j) {
        spmv_panel(ctx, M->T(), *a, *b, scale, mu);
        KR_LAUNCH(ctx, col0_absmax_kernel, 1, 1024, 0, b->p(), n, out.p + j);
        std::swap(a, b);
    }
    return out.to_host();
}

struct OneNormEst { double est; int nprod; };

// Hager/Higham estimator with one column on (scale*(A-mu I))^m  - MATLAB normest1(@afun_power, 1)
// as called from normAm.m:25-26.  Vectors live on the device (column 0 of one-column panels); the
// handful of scalars per iteration come back through cuBLAS reductions.
inline OneNormEst onenormest_power(kr_ctx* ctx, const kr_matrix* M, double scale, double mu, int m) {
    const int64_t n = M->dev.n;
    PanelBuf x(ctx, n, 1), y(ctx, n, 1), z(ctx, n, 1), t(ctx, n, 1);
    x.buf.zero(); y.buf.zero(); z.buf.zero(); t.buf.zero();
    std::vector<double> hx((size_t)n * PW, 0.0);
    for (int64_t i = 0; i < n; ++i) hx[i * PW] = 1.0 / (double)n;
    x.buf.upload(hx.data(), hx.size());
    KR_CUDA(cudaStreamSynchronize(ctx->stream));
    auto power = [&](const CsrDev& Aop, PanelBuf& in, PanelBuf& outv) {
        // outv = Aop^m in (in is clobbered only through t)
        PanelBuf* a = &in;
        PanelBuf* b = &t;
        PanelBuf* fin = &outv;
        for (int i = 0; i < m; ++i) {
            PanelBuf* dst = (i == m - 1) ? fin : b;
            spmv_panel(ctx, Aop, *a, *dst, scale, mu);
            a = dst;
            b = (dst == &t) ? &z : &t;
            if (b == fin) b = &t;
        }
    };
    // Higham & Tisseur (2000), Algorithm 2.4 in the order of MathWorks' normest1.m for t = 1: estimate, test on the
    // estimate, the PARALLEL-SIGN test (before the transposed product, which it saves), the test of max|z| against the
    // entry of the unit vector in use, ties of max|z| to the LAST index (normest1 reverses an ascending sort).
    // Pinned against the reference's own normAm.m run on a matrix with a negative diagonal
    // (tests/golden/reference_golden.json: A0_loops_expmv_info, A0_loops_normAm5).
    int nprod = 0;
    double est_old = 0.0;
    std::vector<int64_t> hist;
    std::vector<double> hy((size_t)n * PW), hz((size_t)n * PW), hs_old((size_t)n, 0.0);
    int64_t est_j = -1, cur = -1;
    for (int itn = 1;; ++itn) {
        power(M->dev, x, y);
        nprod += 1;
        y.buf.download(hy.data(), hy.size());
        double est = 0.0;
        for (int64_t i = 0; i < n; ++i) est += std::abs(hy[i * PW]);
        if (est > est_old || itn == 2) est_j = cur;
        if (itn >= 2 && est <= est_old) break;
        est_old = est;
        if (itn > 5) break;
        double ss = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            const double sg = hy[i * PW] < 0 ? -1.0 : 1.0;
            ss += sg * hs_old[i];
            hs_old[i] = sg;
            hy[i * PW] = sg;
        }
        if (std::abs(ss) == (double)n) break;
        y.buf.upload(hy.data(), hy.size());
        PanelBuf s2(ctx, n, 1);
        s2.buf.zero();
        power(M->T(), y, s2);
        nprod += 1;
        s2.buf.download(hz.data(), hz.size());
        int64_t jmax = 0;
        double zmax = -1.0;
        for (int64_t i = 0; i < n; ++i) {
            const double a = std::abs(hz[i * PW]);
            if (a >= zmax) { zmax = a; jmax = i; }
        }
        if (itn >= 2 && est_j >= 0 && zmax == std::abs(hz[est_j * PW])) break;
        bool seen = false;
        for (int64_t v : hist) seen = seen || (v == jmax);
        if (itn >= 2 && seen) break;
        hist.push_back(jmax);
        cur = jmax;
        std::fill(hx.begin(), hx.end(), 0.0);
        hx[jmax * PW] = 1.0;
        x.buf.upload(hx.data(), hx.size());
        KR_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return OneNormEst{est_old, nprod};
}

struct NormAm { double c; int64_t mv; };

// [c, mv] = normAm(scale*(A - mu I), m)
inline NormAm normAm_dev(kr_ctx* ctx, const kr_matrix* M, double scale, double mu, int m) {
    ShiftFacts sf = shifted_facts(M, mu);
    if (sf.nonneg && scale >= 0.0) {
        std::vector<double> pn = power_norms_nonneg(ctx, M, scale, mu, m);
        return NormAm{pn[m - 1], m};
    }
    OneNormEst e = onenormest_power(ctx, M, scale, mu, m);
    return NormAm{e.est, (int64_t)e.nprod * m};
}

struct TaylorDegree {
    std::vector<double> Mtab;   // m_max x (p_max-1) column-major
    std::vector<double> alpha;
    int64_t mv = 0;
    int unA = 0;
};

// select_taylor_degree on scale*(A - mu I)        (select_taylor_degree.m:15-68)
inline TaylorDegree select_taylor_degree_dev(kr_ctx* ctx, const kr_matrix* M, double scale, double mu,
                                             int64_t ncols, int m_max, int p_max, bool force_estm) {
    if (p_max < 2 || m_max > 60 || m_max + 1 < p_max * (p_max - 1)) fail(KR_ERR_ARG, ">>> Invalid p_max or m_max.");
    TaylorDegree R;
    R.alpha.assign(p_max - 1, 0.0);
    ShiftFacts sf = shifted_facts(M, mu);
    const double normA = std::abs(scale) * sf.norm1;
    if (!force_estm && normA <= 4.0 * KR_THETA[m_max - 1] * p_max * (p_max + 3) / ((double)m_max * (double)ncols)) {
        R.unA = 1;
        for (auto& a : R.alpha) a = normA;
    } else {
        R.unA = 0;
        std::vector<double> eta(p_max);
        if (sf.nonneg && scale >= 0.0) {
            std::vector<double> pn = power_norms_nonneg(ctx, M, scale, mu, p_max + 1);
            for (int p = 1; p <= p_max; ++p) {
                eta[p - 1] = std::pow(pn[p], 1.0 / (p + 1));
                R.mv += p + 1;                 // the reference restarts from ones for every p (normAm.m:18-21)
            }
        } else {
            for (int p = 1; p <= p_max; ++p) {
                OneNormEst e = onenormest_power(ctx, M, scale, mu, p + 1);
                eta[p - 1] = std::pow(e.est, 1.0 / (p + 1));
                R.mv += (int64_t)e.nprod * (p + 1);
            }
        }
        for (int p = 1; p < p_max; ++p) R.alpha[p - 1] = std::max(eta[p - 1], eta[p]);
    }
    R.Mtab.assign((size_t)m_max * (p_max - 1), 0.0);
    for (int p = 2; p <= p_max; ++p)
        for (int m = p * (p - 1) - 1; m <= m_max; ++m)
            R.Mtab[(size_t)(p - 2) * m_max + (m - 1)] = R.alpha[p - 2] / KR_THETA[m - 1];
    return R;
}

// (m, s) from the cost matrix     (expmv.m:57-67)
inline void degree_from_M(const double* Mtab, int m_max, int pcols, double tt, int64_t* m_out, int64_t* s_out) {
    double best = INFINITY;
    int bestm = 1;
    for (int m = 1; m <= m_max; ++m) {
        double colmin = INFINITY;
        for (int p = 0; p < pcols; ++p) {
            double c = std::ceil(std::abs(tt) * Mtab[(size_t)p * m_max + (m - 1)]) * m;
            if (c == 0.0) c = INFINITY;
            colmin = std::min(colmin, c);
        }
        if (colmin < best) { best = colmin; bestm = m; }
    }
    double cost = std::isinf(best) ? 0.0 : best;
    double s = std::max(cost / bestm, 1.0);
    *m_out = bestm;
    *s_out = (int64_t)s;
}

struct ExpmvInfo { int64_t s = 1, m = 0, mv = 0, mvd = 0; int unA = 0; };

// F = exp(t A) B for panel-major device blocks.  B is clobbered.  Mtab may be null.
inline ExpmvInfo expmv_dev(kr_ctx* ctx, const kr_matrix* M, double t, PanelBuf& B, PanelBuf& F,
                           const double* Mtab, int m_max, int pcols, bool shift, bool full_term) {
    const int64_t n = M->dev.n;
    const int panels = B.panels;
    ExpmvInfo info;
    double mu = 0.0;
    if (shift) mu = M->trace / (double)n;
    double tt;
    TaylorDegree td;
    if (!Mtab) {
        tt = 1.0;
        td = select_taylor_degree_dev(ctx, M, t, mu, B.cols, 55, 8, false);
        Mtab = td.Mtab.data();
        m_max = 55;
        pcols = 7;
        info.mvd = td.mv;
        info.mv = td.mv;
        info.unA = td.unA;
    } else {
        tt = t;
    }
    const double tol = std::ldexp(1.0, -53);
    if (t == 0.0) { info.m = 0; info.s = 1; }
    else degree_from_M(Mtab, m_max, pcols, tt, &info.m, &info.s);
    const double eta = shift ? std::exp(t * mu / (double)info.s) : 1.0;
    const int64_t total = B.elems();
    KR_CUDA(cudaMemcpyAsync(F.p(), B.p(), (size_t)total * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    PanelBuf Bn(ctx, n, B.cols);
    DevBuf<double> rab(ctx, (size_t)n * panels), raf(ctx, (size_t)n * panels);
    const int nparts = std::min<int64_t>(ctx->num_sms * 4, std::max<int64_t>(1, ceil_div(n, 256)));
    DevBuf<double> pb(ctx, nparts), pf(ctx, nparts);
    DevBuf<TaylorCtl> ctl(ctx, 1);
    ctl.zero();
    const int egrid = (int)std::min<int64_t>(ctx->num_sms * 8, std::max<int64_t>(1, ceil_div(total, 256)));
    double* b = B.p();
    double* bn = Bn.p();
    const bool fused = panels == 1;                    // q <= PW: one launch per Taylor term
    const int spmm_ctas = M->dev.ntiles;
    for (int64_t i = 0; i < info.s; ++i) {
        KR_LAUNCH(ctx, rowabs_kernel, egrid, 256, 0, b, n, panels, rab.p);
        KR_LAUNCH(ctx, taylor_norms_ctl_kernel, nparts, 256, 0, rab.p, raf.p, n, panels, pb.p, pf.p, ctl.p, 0, (int)full_term, tol);
        const int* done = &ctl.p->done;
        for (int64_t k = 1; k <= info.m; ++k) {
            EpiTaylor epi;
            epi.Bn = bn; epi.Bo = b; epi.F = F.p(); epi.rab = rab.p; epi.raf = raf.p;
            epi.coef = t / ((double)info.s * (double)k);
            epi.mu = mu;
            epi.ctl = fused ? ctl.p : nullptr;
            epi.total_ctas = spmm_ctas;
            epi.full_term = (int)full_term;
            epi.tol = tol;
            launch_spmm(ctx, M->dev, b, panels, epi, done, 0);
            if (!fused)
                KR_LAUNCH(ctx, taylor_norms_ctl_kernel, nparts, 256, 0, rab.p, raf.p, n, panels, pb.p, pf.p, ctl.p, 1, (int)full_term, tol);
            std::swap(b, bn);
        }
        KR_LAUNCH(ctx, stage_end_kernel, egrid, 256, 0, F.p(), b, total, eta);
    }
    TaylorCtl h;
    KR_CUDA(cudaMemcpyAsync(&h, ctl.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    KR_CUDA(cudaStreamSynchronize(ctx->stream));
    info.mv += h.mv;
    ctx->counters[2] += (int64_t)h.mv * B.cols;
    return info;
}

}  // namespace kr
