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
    std::vector<std::unique_ptr<PanelBuf>> Qs;                 // deflation bases Q_1..Q_i (panel-major, one panel each)
    const int mp = PW;                                         // padded width of a 10-column block
    DevBuf<double> small(ctx, (size_t)mp * mp), scratch;
    ThinQrWork qr;

    auto project = [&](const PanelBuf& Q, PanelBuf& X) {       // X -= Q (Q' X)      (:47)
        const PanelList Ql = list_of(Q), Xl = list_of(X);
        ts_gram(ctx, Ql, Xl, n, small.p, scratch);
        ts_update(ctx, Ql, Xl, n, small.p);
    };
    auto base = [&](PanelBuf& X, PanelBuf& Y) {                // Y = Afun_0(X)
        if (op == 0) {
            EpiPlain epi{Y.p(), X.p(), 1.0, 0.0};
            launch_spmm(ctx, M->dev, X.p(), X.panels, epi, nullptr, m);
        } else {
            expmv_dev(ctx, M, 1.0, X, Y, nullptr, 0, 0, true, false);
        }
    };
    // Afun_level(X): nested projectors (:48): innermost call sees aux_level(x) first
    auto apply = [&](int level, const PanelBuf& Xin, PanelBuf& Y) {
        PanelBuf X(ctx, n, m);
        KR_CUDA(cudaMemcpyAsync(X.p(), Xin.p(), (size_t)X.elems() * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        for (int l = level; l >= 1; --l) project(*Qs[l - 1], X);
        base(X, Y);
        for (int l = 1; l <= level; ++l) project(*Qs[l - 1], Y);
    };
    auto trace_of = [&](const PanelBuf& L, const PanelBuf& Rm) {   // trace(L' R)
        ts_gram(ctx, list_of(L), list_of(Rm), n, small.p, scratch);
        std::vector<double> h = small.to_host();
        double t = 0;
        for (int i = 0; i < m; ++i) t += h[(size_t)i * mp + i];
        return t;
    };

    McTraceResult out;
    double tr = 0, tr_old = 0, tr_new = 0;
    for (int64_t it = 1; it <= K; ++it) {
        PanelBuf S(ctx, n, m), G(ctx, n, m);
        std::unique_ptr<PanelBuf> Q(new PanelBuf(ctx, n, m));
        upload_cm_block(ctx, probes + (size_t)(2 * (it - 1)) * n * m, n, S);
        upload_cm_block(ctx, probes + (size_t)(2 * (it - 1) + 1) * n * m, n, G);
        const int level = (int)Qs.size();
        apply(level, S, *Q);                                   // Afun(S)
        thin_qr(ctx, *Q, m, qr);                  // [Q,~] = qr(Afun(S),0)   (:45)
        PanelBuf AQ(ctx, n, m);
        apply(level, *Q, AQ);
        tr += trace_of(*Q, AQ);                                // :46
        Qs.push_back(std::move(Q));
        PanelBuf AG(ctx, n, m);
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
