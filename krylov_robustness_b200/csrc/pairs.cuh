// pairs.cuh - trace_fun_update for MANY candidate edges at once (the candidate loop of
// functions/krylov_miobi.m:76-99 -> functions/trace_fun_update.m:60-125 -> functions/lanczos_krylov.m:73-101).
//
// Candidate h owns columns (2h, 2h+1) of three panel-major blocks P (previous basis block),
// C (current), Y (work).  Basis blocks are never normalised in memory: the orthonormal block is
// V = raw * T with a per-candidate 2x2 matrix T, so the thin QR of every step costs no pass over
// the data.  One block-Lanczos step of the reference = 1 SpMM + 2 fused update passes:
//   pass A  spmm<EpiGram2>: Y = A*C, raw Grams P'Y, C'Y               (CGS pass 1 coefficients)
//   pass B  W = Y*Ma - P*Mp - C*Mc, raw Grams P'W, C'W                (CGS pass 2 coefficients)
//   pass C  W = W - P*Mp - C*Mc, Gram W'W                             (R of the thin QR)
// then one small CTA per candidate: QR factor from the Gram (Householder conventions of LAPACK for
// exactly-zero columns, see DESIGN.md "rank-deficient blocks"), projected matrices Gm / tGm, two
// Jacobi eigen-solves, the trace formula and the reference's lag-2 stopping rule.
#pragma once
#include "dense.cuh"
#include "smalldense.cuh"

namespace kr {

// U block: column 2h = e_{i_h}, column 2h+1 = e_{j_h}    (functions/krylov_miobi.m:85-87)
__global__ void pair_init_kernel(double* __restrict__ C, int64_t n, const int64_t* __restrict__ E,
                                 int64_t nE, int64_t e0, int ncand) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= ncand) return;
    int64_t i = E[e0 + h] - 1, j = E[nE + e0 + h] - 1;
    int c0 = 2 * h, c1 = 2 * h + 1;
    C[(int64_t)(c0 / PW) * n * PW + i * PW + (c0 % PW)] = 1.0;
    C[(int64_t)(c1 / PW) * n * PW + j * PW + (c1 % PW)] = 1.0;
}

// W = Y*Ma - P*Mp - C*Mc per candidate (coef[cand][12] = Ma, Mp, Mc row-major 2x2), in place in Y.
// MODE 0: partial[rb][cand][8] = {P'W, C'W};  MODE 1: partial[rb][cand][4] = {w1'w1, w1'w2, w2'w2, 0}
template <int MODE>
__global__ void __launch_bounds__(COL_THREADS)
pair_update_kernel(double* __restrict__ Y, const double* __restrict__ P, const double* __restrict__ C,
                   int64_t n, const double* __restrict__ coef, int ncand, double* __restrict__ partial,
                   const int* __restrict__ panel_active) {
    constexpr int NV = MODE == 0 ? 8 : 4;
    if (!panel_active[blockIdx.y]) return;            // all candidates of this panel have converged
    __shared__ double smem[COL_WARPS * LPT * NV];
    __shared__ double cf[LPT][12];
    const int q = blockIdx.y, sub = threadIdx.x % LPT;
    if (threadIdx.x < LPT * 12) {
        int cand = q * LPT + threadIdx.x / 12;
        cf[threadIdx.x / 12][threadIdx.x % 12] = cand < ncand ? coef[(int64_t)cand * 12 + threadIdx.x % 12] : 0.0;
    }
    __syncthreads();
    const double* m = cf[sub];
    const int64_t po = (int64_t)q * n * PW;
    const int64_t r0 = (int64_t)blockIdx.x * COL_ROWS_PER_CTA;
    const int64_t r1 = min(n, r0 + COL_ROWS_PER_CTA);
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    for (int64_t r = r0 + (threadIdx.x / LPT); r < r1; r += COL_THREADS / LPT) {
        const int64_t o = po + r * PW + sub * 2;
        double2 y = *reinterpret_cast<const double2*>(Y + o);
        double2 c = *reinterpret_cast<const double2*>(C + o);
        double2 p = make_double2(0.0, 0.0);
        if (P) p = *reinterpret_cast<const double2*>(P + o);
        double2 w;
        w.x = y.x * m[0] + y.y * m[2] - (p.x * m[4] + p.y * m[6]) - (c.x * m[8] + c.y * m[10]);
        w.y = y.x * m[1] + y.y * m[3] - (p.x * m[5] + p.y * m[7]) - (c.x * m[9] + c.y * m[11]);
        *reinterpret_cast<double2*>(Y + o) = w;
        if (MODE == 0) {
            acc[0] += p.x * w.x; acc[1] += p.x * w.y; acc[2] += p.y * w.x; acc[3] += p.y * w.y;
            acc[4] += c.x * w.x; acc[5] += c.x * w.y; acc[6] += c.y * w.x; acc[7] += c.y * w.y;
        } else {
            acc[0] += w.x * w.x; acc[1] += w.x * w.y; acc[2] += w.y * w.y;
        }
    }
    cta_reduce_by_sub<NV>(acc, smem);
    if (threadIdx.x < LPT) {
        int cand = q * LPT + threadIdx.x;
        double* o = partial + ((int64_t)blockIdx.x * (gridDim.y * LPT) + cand) * NV;
#pragma unroll
        for (int i = 0; i < NV; ++i) o[i] = acc[i];
    }
}

// 2x2 helpers, row-major [a b; c d]
__device__ __forceinline__ void mm2(const double* A, const double* B, double* C) {
    double c0 = A[0] * B[0] + A[1] * B[2], c1 = A[0] * B[1] + A[1] * B[3];
    double c2 = A[2] * B[0] + A[3] * B[2], c3 = A[2] * B[1] + A[3] * B[3];
    C[0] = c0; C[1] = c1; C[2] = c2; C[3] = c3;
}
__device__ __forceinline__ void mtm2(const double* A, const double* B, double* C) {   // A' * B
    double c0 = A[0] * B[0] + A[2] * B[2], c1 = A[0] * B[1] + A[2] * B[3];
    double c2 = A[1] * B[0] + A[3] * B[2], c3 = A[1] * B[1] + A[3] * B[3];
    C[0] = c0; C[1] = c1; C[2] = c2; C[3] = c3;
}

struct PairState {
    int ncand, it, fun;
    double tol, b_off;
    double* Tp;       // [ncand][4]
    double* Tc;       // [ncand][4]
    double* hp;       // [ncand][4] accumulated CGS2 coefficients vs previous block (this step)
    double* hc;       // [ncand][4] vs current block
    double* coef;     // [ncand][12]
    double* Hd;       // [ncand][it][4] diagonal blocks
    double* Hs;       // [ncand][it][4] super-diagonal blocks (block (l-1, l))
    double* Hr;       // [ncand][it][4] sub-diagonal blocks (block (l+1, l))
    double* Xstop;    // [ncand][2]
    double* Xm;       // [ncand]
    int* iter;        // [ncand]
    int* lucky;       // [ncand]
    int* active;      // [ncand]
    int* nactive;     // [1]
};

// after pass A: h = T' G T (first CGS pass), coefficients of pass B
__global__ void pair_coef1_kernel(PairState st, const double* __restrict__ G, int first) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= st.ncand) return;
    const double* g = G + (int64_t)h * 8;
    const double* Tc = st.Tc + h * 4;
    const double* Tp = st.Tp + h * 4;
    double tmp[4], hp[4] = {0, 0, 0, 0}, hc[4], Mp[4] = {0, 0, 0, 0}, Mc[4];
    mm2(g + 4, Tc, tmp);       // (C'Y) Tc
    mtm2(Tc, tmp, hc);         // Tc' (C'Y) Tc
    mm2(Tc, hc, Mc);
    if (!first) {
        mm2(g, Tc, tmp);
        mtm2(Tp, tmp, hp);
        mm2(Tp, hp, Mp);
    }
    double* cf = st.coef + (int64_t)h * 12;
    for (int i = 0; i < 4; ++i) {
        cf[i] = Tc[i];
        cf[4 + i] = Mp[i];
        cf[8 + i] = Mc[i];
        st.hp[h * 4 + i] = hp[i];
        st.hc[h * 4 + i] = hc[i];
    }
}

// after pass B: h1 = T' G (W already carries T), h += h1, coefficients of pass C
__global__ void pair_coef2_kernel(PairState st, const double* __restrict__ G, int first) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= st.ncand) return;
    const double* g = G + (int64_t)h * 8;
    const double* Tc = st.Tc + h * 4;
    const double* Tp = st.Tp + h * 4;
    double hp[4] = {0, 0, 0, 0}, hc[4], Mp[4] = {0, 0, 0, 0}, Mc[4];
    mtm2(Tc, g + 4, hc);
    mm2(Tc, hc, Mc);
    if (!first) {
        mtm2(Tp, g, hp);
        mm2(Tp, hp, Mp);
    }
    double* cf = st.coef + (int64_t)h * 12;
    cf[0] = 1.0; cf[1] = 0.0; cf[2] = 0.0; cf[3] = 1.0;
    for (int i = 0; i < 4; ++i) {
        cf[4 + i] = Mp[i];
        cf[8 + i] = Mc[i];
        st.hp[h * 4 + i] += hp[i];
        st.hc[h * 4 + i] += hc[i];
    }
}

// One CTA per candidate, after pass C of step j (1-based).  G3[cand][4] = {g11, g12, g22, -}.
// W is the panel-major block holding the orthogonalised (un-normalised) new block.
__global__ void __launch_bounds__(JAC_THREADS)
pair_step_kernel(PairState st, const double* __restrict__ G3, double* __restrict__ W, int64_t n, int j) {
    extern __shared__ double dyn[];
    __shared__ JacobiShared sh;
    __shared__ double Tn[4];
    __shared__ int s_lucky;
    const int h = blockIdx.x;
    if (!st.active[h]) return;
    const int it = st.it;
    double* Hd = st.Hd + ((int64_t)h * it) * 4;
    double* Hs = st.Hs + ((int64_t)h * it) * 4;
    double* Hr = st.Hr + ((int64_t)h * it) * 4;
    if (threadIdx.x == 0) {
        const double g11 = G3[h * 4 + 0], g12 = G3[h * 4 + 1], g22 = G3[h * 4 + 2];
        const int c0 = 2 * h, c1 = 2 * h + 1;
        double* w1 = W + (int64_t)(c0 / PW) * n * PW + (c0 % PW);    // element (r, c0) at w1[r*PW]
        double* w2 = W + (int64_t)(c1 / PW) * n * PW + (c1 % PW);
        double R[4] = {0, 0, 0, 0}, T[4] = {1, 0, 0, 1};
        // thin QR of the n x 2 block from its Gram matrix; exactly-zero columns follow LAPACK's
        // dgeqr2/dorg2r (tau = 0 -> the completion is a coordinate vector), functions/lanczos_krylov.m:90
        if (g11 == 0.0) {
            // H1 = I, R11 = 0, q1 = e_0
            const double b0 = w2[0], b1 = n > 1 ? w2[PW] : 0.0;
            w1[0] = 1.0;
            R[1] = b0;
            const double tail = g22 - b0 * b0 - b1 * b1;
            if (!(tail > 0.0)) {
                R[3] = b1;                       // tau2 = 0: q2 = e_1
                w2[0] = 0.0;
                if (n > 1) w2[PW] = 1.0;
            } else {
                const double beta2 = -copysign(sqrt(g22 - b0 * b0), b1);
                R[3] = beta2;
                w2[0] = 0.0;                     // q2 = [0; w2(2:n)] / beta2
                T[3] = 1.0 / beta2;
            }
        } else if (g22 == 0.0 && g12 == 0.0) {
            const double a0 = w1[0], a1 = n > 1 ? w1[PW] : 0.0;
            const double tail = g11 - a0 * a0;
            if (!(tail > 0.0)) {
                R[0] = a0;                       // w1 = a0 e_0: tau1 = 0, q1 = e_0, q2 = e_1
                w1[0] = 1.0;
                if (n > 1) w2[PW] = 1.0;
            } else {
                const double beta = -copysign(sqrt(g11), a0);
                const double kappa = a1 / (beta * (a0 - beta));
                R[0] = beta;
                T[0] = 1.0 / beta;
                T[1] = kappa;                    // q2 = kappa*w1 + (e_1 - kappa*beta*e_0)
                w2[0] = -kappa * beta;
                if (n > 1) w2[PW] = 1.0;
            }
        } else {
            const double r11 = sqrt(g11);
            const double r12 = g12 / r11;
            const double d = g22 - r12 * r12;
            R[0] = r11;
            R[1] = r12;
            T[0] = 1.0 / r11;
            if (d > 1e-28 * g22) {
                const double r22 = sqrt(d);
                R[3] = r22;
                T[1] = -r12 / (r11 * r22);
                T[3] = 1.0 / r22;
            } else {                             // numerically dependent column: deflate it
                R[3] = 0.0;
                T[1] = 0.0;
                T[3] = 0.0;
            }
        }
        for (int i = 0; i < 4; ++i) {
            Tn[i] = T[i];
            Hd[(j - 1) * 4 + i] = st.hc[h * 4 + i];
            Hs[(j - 1) * 4 + i] = st.hp[h * 4 + i];
            Hr[(j - 1) * 4 + i] = R[i];
        }
        s_lucky = sqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2] + R[3] * R[3]) < 1e-8;   // lanczos_krylov.m:91
    }
    __syncthreads();
    // ---- projected matrices: Gm = H(1:2j, 1:2j) symmetrised, tGm = Gm + Cm    (trace_fun_update.m:72-81)
    const int nn = 2 * j, lda = nn | 1;
    double* G = dyn;
    double* tG = dyn + nn * lda;
    double* d1 = tG + nn * lda;
    double* d2 = d1 + nn;
    for (int e = threadIdx.x; e < nn * nn; e += JAC_THREADS) {
        int r = e % nn, c = e / nn;
        int br = r >> 1, bc = c >> 1, ir = r & 1, ic = c & 1;
        // H(r, c): block (br, bc); column block bc is step bc+1
        auto Hval = [&](int rr, int cc, int irr, int icc) -> double {
            if (rr == cc) return Hd[cc * 4 + irr * 2 + icc];
            if (rr == cc - 1) return Hs[cc * 4 + irr * 2 + icc];
            if (rr == cc + 1) return Hr[cc * 4 + irr * 2 + icc];
            return 0.0;
        };
        double v = 0.5 * (Hval(br, bc, ir, ic) + Hval(bc, br, ic, ir));
        G[r + c * lda] = v;
        // Cm = B in the leading 2x2 block (V1 = U, so V1'U = I): B = b_off * [0 1; 1 0]
        double cm = (r < 2 && c < 2 && r != c) ? st.b_off : 0.0;
        tG[r + c * lda] = v + cm;
    }
    __syncthreads();
    block_jacobi(tG, nn, lda, nullptr, 0, &sh);
    block_sorted_diag(tG, nn, lda, d1);
    block_jacobi(G, nn, lda, nullptr, 0, &sh);
    block_sorted_diag(G, nn, lda, d2);
    const double Xm = block_trace_formula(st.fun, d1, d2, nn, sh.red);
    if (threadIdx.x == 0) {
        bool done = false;
        double* Xs = st.Xstop + h * 2;
        if (j <= 2) {
            Xs[j - 1] = Xm;
        } else {
            const double err = fabs(Xm - Xs[0]);
            if (err < st.tol) done = true;
            else { Xs[0] = Xs[1]; Xs[1] = Xm; }
        }
        if (!done && s_lucky) done = true;
        if (j == it) done = true;
        st.Xm[h] = Xm;
        st.iter[h] = j;
        st.lucky[h] = s_lucky;
        if (done) {
            st.active[h] = 0;
            atomicSub(st.nactive, 1);
        }
        // rotate transforms: previous <- current, current <- new
        for (int i = 0; i < 4; ++i) {
            st.Tp[h * 4 + i] = st.Tc[h * 4 + i];
            st.Tc[h * 4 + i] = Tn[i];
        }
    }
}

// panel_active[q] = any candidate of panel q still iterating (lets the SpMM and the update passes skip
// whole panels: candidates converge after 3..11 steps)
__global__ void pair_panel_active_kernel(const int* __restrict__ active, int ncp, int* __restrict__ panel_active) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q * LPT >= ncp) return;
    int any = 0;
    for (int i = 0; i < LPT; ++i) any |= active[q * LPT + i];
    panel_active[q] = any;
}

struct PairResult {
    std::vector<double> Xm;
    std::vector<int> iter, lucky;
};

// Scores candidates E[e0 .. e0+ncand) (E is nE x 2 column-major on the DEVICE, 1-based, i != j).
inline void pairs_run_chunk(kr_ctx* ctx, const kr_matrix* M, const int64_t* E_dev, int64_t nE, int64_t e0,
                            int ncand, double b_off, double tol, int it, int fun, double* Xm_out,
                            int64_t* iter_out, int* lucky_out) {
    const CsrDev& A = M->dev;
    const int64_t n = A.n;
    const int cols = 2 * ncand;
    PanelBuf B0(ctx, n, cols), B1(ctx, n, cols), B2(ctx, n, cols);
    const int panels = B0.panels;
    const int ncp = panels * LPT;                     // candidates padded to whole panels
    B0.buf.zero();
    B1.buf.zero();
    B2.buf.zero();
    const int rb = col_row_blocks(n);
    const int nparts = std::max(rb, A.ntiles);
    DevBuf<double> partial(ctx, (size_t)nparts * ncp * 8), Gsum(ctx, (size_t)ncp * 8);
    DevBuf<double> dstate(ctx, (size_t)ncp * (4 + 4 + 4 + 4 + 12 + 2 + 1) + (size_t)ncp * it * 12);
    DevBuf<int> istate(ctx, (size_t)ncp * 3 + 1), pact(ctx, panels);
    dstate.zero();
    istate.zero();
    PairState st;
    st.ncand = ncand; st.it = it; st.fun = fun; st.tol = tol; st.b_off = b_off;
    double* d = dstate.p;
    st.Tp = d; d += ncp * 4;
    st.Tc = d; d += ncp * 4;
    st.hp = d; d += ncp * 4;
    st.hc = d; d += ncp * 4;
    st.coef = d; d += ncp * 12;
    st.Xstop = d; d += ncp * 2;
    st.Xm = d; d += ncp;
    st.Hd = d; d += (size_t)ncp * it * 4;
    st.Hs = d; d += (size_t)ncp * it * 4;
    st.Hr = d;
    st.iter = istate.p;
    st.lucky = istate.p + ncp;
    st.active = istate.p + 2 * ncp;
    st.nactive = istate.p + 3 * ncp;
    // identity transforms, everything active
    std::vector<double> Tinit((size_t)ncp * 4, 0.0);
    for (int h = 0; h < ncp; ++h) { Tinit[h * 4] = 1.0; Tinit[h * 4 + 3] = 1.0; }
    KR_CUDA(cudaMemcpyAsync(st.Tc, Tinit.data(), Tinit.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    KR_CUDA(cudaMemcpyAsync(st.Tp, Tinit.data(), Tinit.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int> act(ncp + 1, 0);
    for (int h = 0; h < ncand; ++h) act[h] = 1;
    KR_CUDA(cudaMemcpyAsync(st.active, act.data(), (size_t)ncp * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    int nact = ncand;
    KR_CUDA(cudaMemcpyAsync(st.nactive, &nact, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    KR_LAUNCH(ctx, pair_panel_active_kernel, (int)ceil_div(panels, 128), 128, 0, st.active, ncp, pact.p);
    KR_CUDA(cudaStreamSynchronize(ctx->stream));

    double* C = B0.p();
    double* P = nullptr;
    double* Y = B1.p();
    double* spare = B2.p();
    KR_LAUNCH(ctx, pair_init_kernel, (int)ceil_div(ncand, 128), 128, 0, C, n, E_dev, nE, e0, ncand);
    const int cb = (int)ceil_div(ncp, 128);
    const int sb8 = (int)ceil_div((int64_t)ncp * 8, 128), sb4 = (int)ceil_div((int64_t)ncp * 4, 128);
    dim3 ugrid((unsigned)rb, (unsigned)panels);
    static bool attr_set = false;
    if (!attr_set) {
        KR_CUDA(cudaFuncSetAttribute(pair_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JAC_SMEM_LIMIT));
        attr_set = true;
    }
    for (int j = 1; j <= it; ++j) {
        const int first = (j == 1);
        EpiGram2 epi;
        epi.Y = Y; epi.P = P; epi.C = C; epi.partial = partial.p; epi.ncand = ncp;
        launch_spmm(ctx, A, C, panels, epi, nullptr, cols, pact.p);
        sum_partials(ctx, partial.p, A.ntiles, ncp * 8, Gsum.p);
        KR_LAUNCH(ctx, pair_coef1_kernel, cb, 128, 0, st, Gsum.p, first);
        KR_LAUNCH(ctx, pair_update_kernel<0>, ugrid, COL_THREADS, 0, Y, P, C, n, st.coef, ncp, partial.p, pact.p);
        sum_partials(ctx, partial.p, rb, ncp * 8, Gsum.p);
        KR_LAUNCH(ctx, pair_coef2_kernel, cb, 128, 0, st, Gsum.p, first);
        KR_LAUNCH(ctx, pair_update_kernel<1>, ugrid, COL_THREADS, 0, Y, P, C, n, st.coef, ncp, partial.p, pact.p);
        sum_partials(ctx, partial.p, rb, ncp * 4, Gsum.p);
        const int nn = 2 * j;
        const size_t smem = (size_t)(2 * nn * (nn | 1) + 2 * nn) * sizeof(double);
        if (smem > JAC_SMEM_LIMIT)
            fail(KR_ERR_UNSUPPORTED, "trace_fun_update_edges: projected size %d exceeds the shared-memory solver", nn);
        KR_LAUNCH(ctx, pair_step_kernel, ncand, JAC_THREADS, smem, st, Gsum.p, Y, n, j);
        KR_LAUNCH(ctx, pair_panel_active_kernel, (int)ceil_div(panels, 128), 128, 0, st.active, ncp, pact.p);
        // rotate blocks: previous <- current, current <- W
        double* oldP = P ? P : spare;
        P = C;
        C = Y;
        Y = oldP;
        KR_CUDA(cudaMemcpyAsync(&nact, st.nactive, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        KR_CUDA(cudaStreamSynchronize(ctx->stream));
        if (nact <= 0) break;
    }
    std::vector<double> xm(ncp);
    std::vector<int> itv(ncp), lk(ncp);
    KR_CUDA(cudaMemcpyAsync(xm.data(), st.Xm, ncp * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    KR_CUDA(cudaMemcpyAsync(itv.data(), st.iter, ncp * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KR_CUDA(cudaMemcpyAsync(lk.data(), st.lucky, ncp * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KR_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->counters[4] += ncp * 16;
    for (int h = 0; h < ncand; ++h) {
        Xm_out[h] = xm[h];
        iter_out[h] = itv[h];
        lucky_out[h] = lk[h];
    }
}

}  // namespace kr
