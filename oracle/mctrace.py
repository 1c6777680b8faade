"""Oracle: stochastic trace estimators.  TEST INFRASTRUCTURE ONLY.

``mc_trace`` / ``trace_exp`` restate functions/mc_trace.m and functions/trace_exp.m.  The reference
draws its probes from MATLAB's unseeded global stream (mc_trace.m:43-44); here the probe blocks are
ARGUMENTS (``probes = [(S_1, G_1), (S_2, G_2), ...]``, each n x 10, entries +-1) so that the device
path and the oracle consume identical inputs.  The draw order (S then G, both before any operator
application) is the reference's, so pre-drawn blocks are equivalent.

``slq_trace`` is NOT a reference function: it is the throughput-mode estimator SURVEY.md section 8(d)
defines for config C3 (all probes at once, single-vector Lanczos per probe, Gauss quadrature
``||z||^2 e1' f(T) e1``).  It is restated here so the device kernel has a CPU checker.
"""
import math

import numpy as np

from .expmv import expmv
from .updates import fun_name, _SCALAR


def mc_trace(Afun, n, tol=1e-3, maxit=10, isAreal=0, debug=0, probes=None, rng=None):
    """[tr_new, res, it] = mc_trace(Afun,n,tol,maxit,isAreal,debug)  (mc_trace.m:1-62)."""
    if not callable(Afun):                                  # :32-34
        Amat = Afun
        Afun = lambda x: np.asarray(Amat @ x)
    tr = 0.0
    tr_old = 0.0
    m = 10                                                  # :36
    K = int(math.ceil(maxit / (3 * m)))                     # :41
    tr_new, res, it = 0.0, float("nan"), 0
    for it in range(1, K + 1):
        if probes is not None:
            S, G = probes[it - 1]
        else:
            rng = rng or np.random.default_rng(0)
            S = np.sign(rng.standard_normal((n, m)))
            G = np.sign(rng.standard_normal((n, m)))
        Q, _ = np.linalg.qr(Afun(S), mode="reduced")        # :45
        tr = tr + np.trace(Q.T @ Afun(Q))                   # :46
        prev = Afun

        def Afun(x, Q=Q, prev=prev):                        # :47-48 nested deflation
            x = x - Q @ (Q.T @ x)
            y = prev(x)
            return y - Q @ (Q.T @ y)

        tr_new = tr + np.trace(G.T @ Afun(G)) / m           # :49
        res = abs(tr_new - tr_old) / max(abs(tr_new), abs(tr_old))
        if res < tol:
            break
        tr_old = tr_new
    if isAreal == 1:
        tr_new = float(np.real(tr_new))
    return tr_new, res, it


def trace_exp(A, probes=None, rng=None):
    """tr = trace_exp(A)  (trace_exp.m:1-7)."""
    Afun = lambda x: expmv(1, A, x, None, "double")[0]
    return mc_trace(Afun, A.shape[0], 1e-4, 1000, 1, probes=probes, rng=rng)[0]


def slq_trace(A, Z, m, fun="exp"):
    """Throughput-mode estimator: trace(f(A)) ~ mean_z ||z||^2 e1' f(T_z) e1
    with T_z the m-step single-vector Lanczos tridiagonal started at z/||z|| (no reorthogonalisation).
    Returns (estimate, per-probe values, alpha (m x k), beta (m x k))."""
    name = fun_name(fun)
    f = _SCALAR[name]
    Z = np.asarray(Z, dtype=np.float64)
    n, k = Z.shape
    nrm2 = np.einsum("ij,ij->j", Z, Z)
    v = Z / np.sqrt(nrm2)
    vprev = np.zeros_like(v)
    beta_prev = np.zeros(k)
    alpha = np.zeros((m, k))
    beta = np.zeros((m, k))
    steps = np.full(k, m)
    alive = np.ones(k, dtype=bool)
    for j in range(m):
        w = np.asarray(A @ v)
        a = np.einsum("ij,ij->j", v, w)
        w = w - v * a - vprev * beta_prev
        b = np.sqrt(np.einsum("ij,ij->j", w, w))
        alpha[j] = a
        beta[j] = b
        dead = alive & ~(b > 0)
        steps[dead] = j + 1
        alive &= b > 0
        inv = np.where(b > 0, 1.0 / np.where(b > 0, b, 1.0), 0.0)
        vprev = v
        v = w * inv
        beta_prev = b
    vals = np.zeros(k)
    for c in range(k):
        s = int(steps[c])
        T = np.diag(alpha[:s, c]) + np.diag(beta[:s - 1, c], 1) + np.diag(beta[:s - 1, c], -1)
        lam, Q = np.linalg.eigh(T)
        vals[c] = nrm2[c] * float(np.sum(Q[0, :] ** 2 * f(lam)))
    return float(vals.mean()), vals, alpha, beta
