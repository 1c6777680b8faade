// mctrace.cuh - mc_trace in PARITY mode (functions/mc_trace.m:20-62): 10-column blocks, sequential
// low-rank deflation, probes passed in by the caller.  Afun is @(x) A*x or
// @(x) expmv(1,A,x,[],'double') (functions/trace_exp.m:5).  Sequential by construction
// ("replicas only"); the throughput mode is slq.cuh.
#pragma once
#include "blockkrylov.cuh"
#include "expmv.cuh"

namespace kr {

struct McTraceResult { double tr = 0, res = 0; int64_t it = 0; };

inline McTraceResult mc_trace_dev(kr_ctx* ctx, const kr_matrix* M, int op, double tol, int64_t maxit,
                                  const double* probes) {
    const int64_t n = M->dev.n;
    const int m = 10;                                          // mc_trace.m:36
    const int64_t K = (maxit + 3 * m - 1) / (3 * m);           // :41
    std::vector<std::unique_ptr<CmMat>> Qs;                    // deflation bases Q_1..Q_i
    DevBuf<double> small(ctx, (size_t)m * m);

    auto project = [&](const CmMat& Q, CmMat& X) {             // X -= Q (Q' X)      (:47)
        gemm(ctx, true, false, m, m, n, 1.0, Q.p(), n, X.p(), n, 0.0, small.p, m);
        gemm(ctx, false, false, n, m, m, -1.0, Q.p(), n, small.p, m, 1.0, X.p(), n);
    };
    auto base = [&](CmMat& X, CmMat& Y) {                      // Y = Afun_0(X)
        if (op == 0) {
            spmm_cm(ctx, M, X.p(), m, Y.p());
        } else {
            PanelBuf b(ctx, n, m), f(ctx, n, m);
            cm_to_panel(ctx, X.p(), n, b);
            expmv_dev(ctx, M, 1.0, b, f, nullptr, 0, 0, true, false);
            panel_to_cm(ctx, f, Y.p(), n);
        }
    };
    // Afun_level(X): nested projectors (:48): innermost call sees aux_level(x) first
    auto apply = [&](int level, const CmMat& Xin, CmMat& Y) {
        CmMat X(ctx, n, m);
        KR_CUDA(cudaMemcpyAsync(X.p(), Xin.p(), (size_t)n * m * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        for (int l = level; l >= 1; --l) project(*Qs[l - 1], X);
        base(X, Y);
        for (int l = 1; l <= level; ++l) project(*Qs[l - 1], Y);
    };
    auto trace_of = [&](const CmMat& L, const CmMat& Rm) {     // trace(L' R)
        gemm(ctx, true, false, m, m, n, 1.0, L.p(), n, Rm.p(), n, 0.0, small.p, m);
        std::vector<double> h = small.to_host();
        double t = 0;
        for (int i = 0; i < m; ++i) t += h[(size_t)i * m + i];
        return t;
    };

    McTraceResult out;
    double tr = 0, tr_old = 0, tr_new = 0;
    for (int64_t it = 1; it <= K; ++it) {
        CmMat S(ctx, n, m), G(ctx, n, m), Y(ctx, n, m);
        upload_host_cm(ctx, probes + (size_t)(2 * (it - 1)) * n * m, n, S);
        upload_host_cm(ctx, probes + (size_t)(2 * (it - 1) + 1) * n * m, n, G);
        const int level = (int)Qs.size();
        apply(level, S, Y);                                    // Afun(S)
        HostMat Rq;
        qr_thin(ctx, Y.p(), n, m, Rq);                         // [Q,~] = qr(Afun(S),0)   (:45)
        std::unique_ptr<CmMat> Q(new CmMat(ctx, n, m));
        KR_CUDA(cudaMemcpyAsync(Q->p(), Y.p(), (size_t)n * m * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        CmMat AQ(ctx, n, m);
        apply(level, *Q, AQ);
        tr += trace_of(*Q, AQ);                                // :46
        Qs.push_back(std::move(Q));
        CmMat AG(ctx, n, m);
        apply(level + 1, G, AG);
        tr_new = tr + trace_of(G, AG) / m;                     // :49
        out.res = std::abs(tr_new - tr_old) / std::max(std::abs(tr_new), std::abs(tr_old));
        out.it = it;
        if (out.res < tol) break;
        tr_old = tr_new;
    }
    out.tr = tr_new;
    return out;
}

}  // namespace kr
