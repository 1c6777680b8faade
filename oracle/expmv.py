"""Oracle: expmv / select_taylor_degree / normAm (Al-Mohy & Higham 2011, Alg. 3.2).

TEST INFRASTRUCTURE ONLY.  Restates functions/expmv.m, functions/select_taylor_degree.m and
functions/normAm.m (vendored there from github.com/higham/expmv, README.md:43-47).
Only the double-precision, unbalanced path is restated (``bal=false``; ``prec='double'``), which
is the only one the reference's callers take (functions/trace_exp.m:5).
"""
import math

import numpy as np
import scipy.sparse as sp

from .theta import THETA


def _norm1(A):
    if sp.issparse(A):
        return float(abs(A).sum(axis=0).max()) if A.nnz else 0.0
    return float(np.abs(A).sum(axis=0).max())


def _norminf_mat(b):
    # MATLAB norm(b, inf) of an n x q matrix: max row abs-sum (expmv.m:74,80,83)
    b = np.asarray(b)
    if b.ndim == 1:
        return float(np.abs(b).max())
    return float(np.abs(b).sum(axis=1).max())


def normAm(A, m):
    """[c, mv] = normAm(A, m)  (normAm.m:1-52)."""
    n = A.shape[0]
    nonneg = (A.data >= 0).all() if sp.issparse(A) else bool((A >= 0).all())
    if nonneg:                                              # :17-23
        e = np.ones(n)
        AT = A.T
        for _ in range(m):
            e = AT @ e
        e = np.asarray(e).ravel()
        return float(np.abs(e).max()), m
    # :25-26 normest1(@afun_power, 1): Hager/Higham estimator with t = 1 column on A^m;
    # mv = it(2)*t*m where it(2) counts block products.
    c, nprod = _onenormest_power(A, m)
    return c, nprod * m


def _onenormest_power(A, m):
    # Hager/Higham 1-norm estimator with one column (normest1 with t = 1), products with A^m / (A^m)'.
    n = A.shape[0]
    AT = A.T

    def fwd(x):
        for _ in range(m):
            x = A @ x
        return np.asarray(x).ravel()

    def bwd(x):
        for _ in range(m):
            x = AT @ x
        return np.asarray(x).ravel()

    # Higham & Tisseur (2000), Algorithm 2.4, as MathWorks' normest1.m states it for t = 1 column: the estimate,
    # then the test on the estimate, then the PARALLEL-SIGN test (before the transposed product, which it saves),
    # then the test on max|z| against the entry of the unit vector in use, ties of max|z| to the LAST index
    # (normest1 reverses an ascending sort).  Pinned by tests/golden/reference_golden.json (A0_loops_*).
    x = np.ones(n) / n
    nprod = 0
    est_old = 0.0
    ind_hist = []
    est_j = -1
    cur = -1
    s = np.zeros(n)
    it = 0
    while True:
        it += 1
        y = fwd(x)
        nprod += 1
        est = float(np.abs(y).sum())
        if est > est_old or it == 2:
            est_j = cur
        if it >= 2 and est <= est_old:
            est = est_old
            break
        est_old = est
        if it > 5:
            break
        s_old = s
        s = np.sign(y)
        s[s == 0] = 1.0
        if abs(float(s_old @ s)) == n:
            break
        z = bwd(s)
        nprod += 1
        h = np.abs(z)
        if it >= 2 and h.max() == h[est_j]:
            break
        jmax = int(np.argsort(h, kind="stable")[-1])
        if it >= 2 and jmax in ind_hist:
            break
        ind_hist.append(jmax)
        cur = jmax
        x = np.zeros(n)
        x[jmax] = 1.0
    return est_old, nprod


def select_taylor_degree(A, b, m_max=55, p_max=8, prec="double", shift=False, bal=False,
                         force_estm=False):
    """[M, mv, alpha, unA] = select_taylor_degree(A,b,m_max,p_max,prec,shift,bal,force_estm)
    (select_taylor_degree.m:1-68).  ``b`` is only consulted for its column count (:44)."""
    if m_max is None:
        m_max = 55
    if p_max is None:
        p_max = 8
    if p_max < 2 or m_max > 60 or m_max + 1 < p_max * (p_max - 1):
        raise ValueError(">>> Invalid p_max or m_max.")
    if prec != "double" or bal:
        raise NotImplementedError("oracle restates the double / unbalanced path only")
    n = A.shape[0]
    theta = THETA
    if shift:                                               # :37-40
        mu = A.diagonal().sum() / n
        A = A - mu * sp.identity(n, format="csr")
    ncols = 1 if np.ndim(b) == 1 else np.shape(b)[1]
    mv = 0
    normA = None if force_estm else _norm1(A)
    if (not force_estm) and normA <= 4 * theta[m_max - 1] * p_max * (p_max + 3) / (m_max * ncols):
        unA = 1                                             # :44-49
        alpha = normA * np.ones(p_max - 1)
    else:
        unA = 0                                             # :50-62
        eta = np.zeros(p_max)
        alpha = np.zeros(p_max - 1)
        for p in range(1, p_max + 1):
            c, k = normAm(A, p + 1)
            eta[p - 1] = c ** (1.0 / (p + 1))
            mv += k
        for p in range(1, p_max):
            alpha[p - 1] = max(eta[p - 1], eta[p])
    M = np.zeros((m_max, p_max - 1))                        # :63-68
    for p in range(2, p_max + 1):
        for m in range(p * (p - 1) - 1, m_max + 1):
            M[m - 1, p - 2] = alpha[p - 2] / theta[m - 1]
    return M, mv, alpha, unA


def degree_from_M(M, tt):
    """expmv.m:57-67: (m, s) from the cost matrix."""
    m_max, p = M.shape
    U = np.diag(np.arange(1, m_max + 1, dtype=np.float64))
    C = np.ceil(abs(tt) * M).T @ U
    C[C == 0] = np.inf
    if p > 1:
        colmin = C.min(axis=0)
        m = int(np.argmin(colmin)) + 1
        cost = colmin[m - 1]
    else:
        m = int(np.argmin(C.ravel())) + 1
        cost = C.ravel()[m - 1]
    if cost == np.inf:
        cost = 0
    s = max(cost / m, 1)
    return m, int(s)


def expmv(t, A, b, M=None, prec="double", shift=True, bal=False, full_term=False, prnt=False):
    """[f, s, m, mv, mvd, unA] = expmv(t,A,b,M,prec,shift,bal,full_term,prnt)  (expmv.m:1-94)."""
    if bal or prec != "double":
        raise NotImplementedError("oracle restates the double / unbalanced path only")
    b = np.array(b, dtype=np.float64)
    n = A.shape[0]
    mu = 0.0
    if shift:                                               # :32-36
        mu = float(A.diagonal().sum() / n)
        A = A - mu * sp.identity(n, format="csr")
    unA = None
    if M is None:                                           # :39-45
        tt = 1
        M, mvd, _, unA = select_taylor_degree(t * A, b, None, None, prec, False, False)
        mv = mvd
    else:
        tt = t
        mv = 0
        mvd = 0
    tol = 2.0 ** -53
    s = 1
    if t == 0:
        m = 0
    else:
        m, s = degree_from_M(np.asarray(M), tt)
    eta = 1.0
    if shift:
        eta = math.exp(t * mu / s)
    f = b.copy()
    for _ in range(int(s)):                                 # :73-92
        c1 = _norminf_mat(b)
        for k in range(1, m + 1):
            b = (t / (s * k)) * (A @ b)
            mv += 1
            f = f + b
            c2 = _norminf_mat(b)
            if not full_term:
                if c1 + c2 <= tol * _norminf_mat(f):
                    break
                c1 = c2
        f = eta * f
        b = f
    return f, s, m, mv, mvd, unA
