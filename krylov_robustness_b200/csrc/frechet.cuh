// frechet.cuh - Hessian callbacks of the weighted experiments:
// functions/multiple_frechet_eval.m + functions/hessianfcn_exp.m / hessianfcn_fun.m.
//
// For every pair (i, j) in omega the reference builds K(A, e_i) and K(A', e_j) with bs = 1 Arnoldi and
// takes the (1,2) block of f([G, C; 0, H']) with C = (V'e_i)(W'e_j)' = +-e1 e1'
// (multiple_frechet_eval.m:150-159).  Here the two families of spaces are two ArnoldiBatch objects (one
// wide SpMM per family per step) and the block-triangular function is evaluated per pair by one CTA in
// shared memory through the Daleckii-Krein formula: with G = P diag(lambda) P', H = Q diag(mu) Q'
// (both symmetric up to rounding on this path; symmetrised as in entries.cuh),
//     X = P [ f[lambda_a, mu_b] * P(1,a) Q(1,b) ]_{ab} Q'
// where f[.,.] is the first divided difference, written in the cancellation-free product form
// exp[l,m] = e^{(l+m)/2} sinh(d)/d, sinh[l,m] = cosh((l+m)/2) sinh(d)/d, cosh[l,m] = sinh((l+m)/2) sinh(d)/d,
// d = (l-m)/2.  The lag-3 stopping rule uses ||X_j - pad(X_{j-3})||_2 (:172) = sqrt(lambda_max(D'D)).
#pragma once
#include "entries.cuh"

namespace kr {

constexpr int FRECHET_CAP = 64;     // largest projected size handled in shared memory (5 matrices)

__device__ __forceinline__ double divided_difference(int fun, double l, double m) {
    const double mean = 0.5 * (l + m), d = 0.5 * (l - m);
    const double shc = d == 0.0 ? 1.0 : sinh(d) / d;
    const double w = fun == KR_FUN_EXP ? exp(mean) : fun == KR_FUN_SINH ? cosh(mean) : sinh(mean);
    return w * shc;
}

// One CTA per pair.  HcI/HcJ: Hessenberg columns of the row / column families (entries.cuh layout);
// sa[p], sb[p]: space (column) of pair p in each family.  hist[pair][4][CAP*CAP]; xfin[pair][CAP*CAP].
__global__ void __launch_bounds__(JAC_THREADS)
frechet_step_kernel(const double* __restrict__ HcI, const double* __restrict__ HcJ, int it1, int jj, int fun, double tol,
                    const int* __restrict__ sa, const int* __restrict__ sb, double* __restrict__ hist,
                    double* __restrict__ xfin, int* __restrict__ nfin, int* __restrict__ conv, int* __restrict__ nactive) {
    extern __shared__ double dyn[];
    __shared__ JacobiShared sh;
    __shared__ double lam[FRECHET_CAP], mu[FRECHET_CAP];
    const int p = blockIdx.x;
    if (conv[p]) return;
    const int lda = jj | 1;
    double* G = dyn;
    double* P = G + jj * lda;
    double* H = P + jj * lda;
    double* Q = H + jj * lda;
    double* T = Q + jj * lda;
    const double* HI = HcI + (int64_t)sa[p] * it1 * (it1 + 1);
    const double* HJ = HcJ + (int64_t)sb[p] * it1 * (it1 + 1);
    for (int e = threadIdx.x; e < jj * jj; e += JAC_THREADS) {
        const int r = e % jj, c = e / jj;
        const double g1 = r <= c + 1 ? HI[(int64_t)c * (it1 + 1) + r] : 0.0;
        const double g2 = c <= r + 1 ? HI[(int64_t)r * (it1 + 1) + c] : 0.0;
        const double h1 = r <= c + 1 ? HJ[(int64_t)c * (it1 + 1) + r] : 0.0;
        const double h2 = c <= r + 1 ? HJ[(int64_t)r * (it1 + 1) + c] : 0.0;
        G[r + c * lda] = 0.5 * (g1 + g2);
        H[r + c * lda] = 0.5 * (h1 + h2);
        P[r + c * lda] = r == c ? 1.0 : 0.0;
        Q[r + c * lda] = r == c ? 1.0 : 0.0;
    }
    __syncthreads();
    block_jacobi(G, jj, lda, P, lda, &sh);
    block_jacobi(H, jj, lda, Q, lda, &sh);
    for (int i = threadIdx.x; i < jj; i += JAC_THREADS) {
        lam[i] = G[i + i * lda];
        mu[i] = H[i + i * lda];
    }
    __syncthreads();
    // T = [f[lam_a, mu_b] P(0,a) Q(0,b)]
    for (int e = threadIdx.x; e < jj * jj; e += JAC_THREADS) {
        const int a = e % jj, b = e / jj;
        T[a + b * lda] = divided_difference(fun, lam[a], mu[b]) * P[0 + a * lda] * Q[0 + b * lda];
    }
    __syncthreads();
    // G <- P * T
    for (int e = threadIdx.x; e < jj * jj; e += JAC_THREADS) {
        const int r = e % jj, c = e / jj;
        double s = 0.0;
        for (int k = 0; k < jj; ++k) s += P[r + k * lda] * T[k + c * lda];
        G[r + c * lda] = s;
    }
    __syncthreads();
    // H <- X = G * Q'
    for (int e = threadIdx.x; e < jj * jj; e += JAC_THREADS) {
        const int r = e % jj, c = e / jj;
        double s = 0.0;
        for (int k = 0; k < jj; ++k) s += G[r + k * lda] * Q[c + k * lda];
        H[r + c * lda] = s;
    }
    __syncthreads();
    double* hp = hist + (int64_t)p * 4 * FRECHET_CAP * FRECHET_CAP;
    double err = 0.0;
    if (jj > 3) {
        const double* old = hp + (int64_t)((jj - 3) & 3) * FRECHET_CAP * FRECHET_CAP;   // (jj-3) x (jj-3), ld jj-3
        const int on = jj - 3;
        for (int e = threadIdx.x; e < jj * jj; e += JAC_THREADS) {
            const int r = e % jj, c = e / jj;
            T[r + c * lda] = H[r + c * lda] - ((r < on && c < on) ? old[r + c * on] : 0.0);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < jj * jj; e += JAC_THREADS) {       // G <- T' T
            const int r = e % jj, c = e / jj;
            double s = 0.0;
            for (int k = 0; k < jj; ++k) s += T[k + r * lda] * T[k + c * lda];
            G[r + c * lda] = s;
        }
        __syncthreads();
        block_jacobi(G, jj, lda, nullptr, 0, &sh);
        double m = 0.0;
        for (int i = threadIdx.x; i < jj; i += JAC_THREADS) m = fmax(m, G[i + i * lda]);
        for (int off = 16; off; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = m;
        __syncthreads();
        for (int w = 0; w < JAC_THREADS / 32; ++w) m = fmax(m, sh.red[w]);
        err = sqrt(fmax(m, 0.0));
    }
    double* cur = hp + (int64_t)(jj & 3) * FRECHET_CAP * FRECHET_CAP;
    double* fin = xfin + (int64_t)p * FRECHET_CAP * FRECHET_CAP;
    for (int e = threadIdx.x; e < jj * jj; e += JAC_THREADS) {
        const int r = e % jj, c = e / jj;
        const double v = H[r + c * lda];
        cur[r + c * jj] = v;
        fin[r + c * jj] = v;
    }
    if (threadIdx.x == 0) {
        nfin[p] = jj;
        if (jj > 3 && !(err > tol)) {
            conv[p] = 1;
            atomicSub(nactive, 1);
        }
    }
}

// Hes(j, l) = -2 * U_{a_j}(O(l,1), 1:sz) * X_j * V_{b_j}(O(l,2), 1:sz)'  for l >= j, mirrored
// (hessianfcn_exp.m:9-15).  One CTA per j, one thread per l.
__global__ void hessian_assemble_kernel(const double* const* __restrict__ VI, const double* const* __restrict__ VJ,
                                        int64_t n, const int64_t* __restrict__ omega, int m, const int* __restrict__ sa,
                                        const int* __restrict__ sb, const double* __restrict__ xfin,
                                        const int* __restrict__ nfin, double* __restrict__ Hes) {
    const int j = blockIdx.x;
    const int sz = nfin[j];
    const double* X = xfin + (int64_t)j * FRECHET_CAP * FRECHET_CAP;      // sz x sz, ld sz
    const int ca = sa[j], cb = sb[j];
    for (int l = j + threadIdx.x; l < m; l += blockDim.x) {
        const int64_t offa = (int64_t)(ca / PW) * n * PW + (omega[l] - 1) * PW + (ca % PW);
        const int64_t offb = (int64_t)(cb / PW) * n * PW + (omega[m + l] - 1) * PW + (cb % PW);
        double u[FRECHET_CAP];
        for (int r = 0; r < sz; ++r) u[r] = VI[r][offa];
        double acc = 0.0;
        for (int c = 0; c < sz; ++c) {
            double t = 0.0;
            for (int r = 0; r < sz; ++r) t += u[r] * X[r + c * sz];
            acc += t * VJ[c][offb];
        }
        Hes[j + (int64_t)l * m] = -2.0 * acc;
        Hes[l + (int64_t)j * m] = -2.0 * acc;
    }
}

struct HessianResult { std::vector<double> Hes; int64_t iter = 0; };

// omega: m x 2 column-major on the host (1-based)
inline HessianResult frechet_hessian_run(kr_ctx* ctx, const kr_matrix* M, int64_t m, const int64_t* omega, int fun,
                                         double tol, int it) {
    const int64_t n = M->dev.n;
    it = std::min(it, FRECHET_CAP);
    const int it1 = it + 1;
    std::vector<int64_t> rowsI, rowsJ;
    std::vector<int> sa(m), sb(m);
    std::map<int64_t, int> pi, pj;
    for (int64_t p = 0; p < m; ++p) {
        const int64_t i = omega[p], j = omega[m + p];
        if (i < 1 || i > n || j < 1 || j > n) fail(KR_ERR_ARG, "Omega index out of range");
        auto fi = pi.find(i);
        if (fi == pi.end()) { pi[i] = (int)rowsI.size(); sa[p] = (int)rowsI.size(); rowsI.push_back(i); } else sa[p] = fi->second;
        auto fj = pj.find(j);
        if (fj == pj.end()) { pj[j] = (int)rowsJ.size(); sb[p] = (int)rowsJ.size(); rowsJ.push_back(j); } else sb[p] = fj->second;
    }
    ArnoldiBatch BI(ctx, M->dev, rowsI, it), BJ(ctx, M->T(), rowsJ, it);
    DevBuf<int> dsa(ctx, m), dsb(ctx, m), istate(ctx, (size_t)2 * m + 1);
    DevBuf<int64_t> dom(ctx, 2 * m);
    dsa.upload(sa.data(), m);
    dsb.upload(sb.data(), m);
    dom.upload(omega, 2 * m);
    DevBuf<double> hist(ctx, (size_t)m * 4 * FRECHET_CAP * FRECHET_CAP), xfin(ctx, (size_t)m * FRECHET_CAP * FRECHET_CAP);
    hist.zero(); xfin.zero(); istate.zero();
    int* nfin = istate.p;
    int* conv = istate.p + m;
    int* nactive = istate.p + 2 * m;
    int nact = (int)m;
    KR_CUDA(cudaMemcpyAsync(nactive, &nact, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    static bool attr_set[64] = {};              // per device: the attribute lives in the device's primary context
    if (first_use_on_device(attr_set, ctx->device))
        KR_CUDA(cudaFuncSetAttribute(frechet_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JAC_SMEM_LIMIT));
    HessianResult out;
    int j = 0;
    for (j = 0; j < it; ++j) {
        BI.step(j);
        BJ.step(j);
        const int jj = j + 1;
        const size_t smem = (size_t)(5 * jj * (jj | 1)) * sizeof(double);
        KR_LAUNCH(ctx, frechet_step_kernel, (int)m, JAC_THREADS, smem, BI.Hc.p, BJ.Hc.p, it1, jj, fun, tol, dsa.p, dsb.p,
                  hist.p, xfin.p, nfin, conv, nactive);
        KR_CUDA(cudaMemcpyAsync(&nact, nactive, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        KR_CUDA(cudaStreamSynchronize(ctx->stream));
        if (nact <= 0) { ++j; break; }
    }
    out.iter = std::min(j, it);
    DevBuf<const double*> pI = BI.block_pointers(), pJ = BJ.block_pointers();
    DevBuf<double> dH(ctx, (size_t)m * m);
    dH.zero();
    KR_LAUNCH(ctx, hessian_assemble_kernel, (int)m, 64, 0, pI.p, pJ.p, n, dom.p, (int)m, dsa.p, dsb.p, xfin.p, nfin, dH.p);
    out.Hes = dH.to_host();
    return out;
}

}  // namespace kr
