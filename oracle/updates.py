"""Oracle: low-rank-update evaluators, entries of f(A), objective/gradient callbacks.

TEST INFRASTRUCTURE ONLY.  Restates functions/trace_fun_update.m, fun_update.m,
function_multiple_entries.m, fun_and_grad_krylov_exp.m, fun_and_grad_krylov_fun.m, plus
MATLAB's ``normest`` (power iteration on A'A) which those callbacks use.

Function selectors: the reference compares function handles with ``isequal(f,@exp)``; here a
selector is one of the strings 'exp' / 'sinh' / 'cosh' (NumPy ufuncs np.exp / np.sinh / np.cosh
are accepted and mapped to the strings).
"""
import warnings

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from .krylov import lanczos_krylov, arnoldi_krylov

_SCALAR = {"exp": np.exp, "sinh": np.sinh, "cosh": np.cosh}


def fun_name(fun):
    if isinstance(fun, str):
        if fun not in _SCALAR:
            raise ValueError("unsupported function selector %r" % (fun,))
        return fun
    for k, v in _SCALAR.items():
        if fun is v:
            return k
    raise ValueError("unsupported function handle %r" % (fun,))


def _full(A):
    return A.toarray() if sp.issparse(A) else np.array(A, dtype=np.float64)


def _ishermitian(B):
    B = _full(B) if sp.issparse(B) else np.atleast_2d(np.asarray(B))
    return B.shape[0] == B.shape[1] and np.array_equal(B, B.T)


def _eig_sorted(M):
    # MATLAB eig() of an exactly symmetric matrix takes the symmetric path; otherwise general.
    if np.array_equal(M, M.T):
        return np.linalg.eigvalsh(M)
    # a non-Hermitian B (never on the device path, which rejects it) makes the projection unsymmetric: MATLAB's eig
    # returns complex eigenvalues and sort() orders them by modulus, then phase; the trace formula is evaluated in
    # complex arithmetic and its imaginary parts cancel over the conjugate pairs (found by the differential test
    # against the reference's sources, tests/test_mlab_differential.py: taking real parts first changes the value)
    w = np.linalg.eigvals(M)
    return w[np.lexsort((np.angle(w), np.abs(w)))]


def _trace_formula(name, d1, d2):
    # trace_fun_update.m:43-47 / :85-89
    if name == "exp":
        return float(np.real(np.sum(np.exp(d1) * (1.0 - np.exp(d2 - d1)))))
    f = _SCALAR[name]
    return float(np.real(np.sum(f(d1) - f(d2))))


def _matfun(name):
    # fun_update.m:43-59
    if name == "exp":
        return lambda M: sla.expm(M)
    if name == "sinh":
        return lambda M: (sla.expm(M) - sla.expm(-M)) / 2
    if name == "cosh":
        return lambda M: (sla.expm(M) + sla.expm(-M)) / 2
    raise ValueError(name)


def _matfun_entries(name):
    # function_multiple_entries.m:48-60 : exp -> expm ; everything else -> funm(M, f)
    if name == "exp":
        return lambda M: sla.expm(M)
    if name == "sinh":
        return lambda M: sla.sinhm(M)
    if name == "cosh":
        return lambda M: sla.coshm(M)
    raise ValueError(name)


def _pad(M, nn):
    if M.shape[0] >= nn:
        return M
    out = np.zeros((nn, nn))
    out[:M.shape[0], :M.shape[1]] = M
    return out


# --------------------------------------------------------------------------- normest
def normest(S, tol=1e-6):
    """MATLAB ``normest(S,tol)``: 2-norm estimate by power iteration on S'S started from the
    column abs-sums; stops when the relative change is <= tol (at most 100 iterations).
    Used at fun_and_grad_krylov_exp.m:26, fun_and_grad_krylov_fun.m:27, Tests/test_unweighted_break.m:56."""
    S = sp.csr_matrix(S) if not sp.issparse(S) else S.tocsr()
    x = np.asarray(abs(S).sum(axis=0)).ravel().astype(np.float64)
    cnt = 0
    e = float(np.linalg.norm(x))
    if e == 0:
        return e, cnt
    x = x / e
    e0 = 0.0
    while abs(e - e0) > tol * e:
        e0 = e
        Sx = S @ x
        if not np.any(Sx):
            Sx = np.random.default_rng(0).random(Sx.shape)
        x = S.T @ Sx
        normx = float(np.linalg.norm(x))
        e = normx / float(np.linalg.norm(Sx))
        x = x / normx
        cnt += 1
        if cnt > 100:
            warnings.warn("normest: not converged")
            break
    return e, cnt


# --------------------------------------------------------------------------- trace_fun_update
def trace_fun_update(A, U, B, tol=1e-12, it=None, debug=0, fun="exp"):
    """[Xm, iter, lucky] = trace_fun_update(A,U,B,tol,it,debug,fun)  (trace_fun_update.m:1-130)."""
    name = fun_name(fun)
    U = np.asarray(U.toarray() if sp.issparse(U) else U, dtype=np.float64)
    if U.ndim == 1:
        U = U[:, None]
    B = np.atleast_2d(np.asarray(B, dtype=np.float64))
    n = A.shape[0]
    if it is None:
        it = min(100, n)
    if U.shape[0] <= 130:                                   # :37-51 dense branch
        fA = _full(A)
        fAt = fA + U @ B @ U.T
        fAt = (fAt + fAt.T) / 2
        d1 = _eig_sorted(fAt)
        d2 = _eig_sorted(fA)
        return _trace_formula(name, d1, d2), 0, 0
    rk = U.shape[1]
    herm = _ishermitian(B)
    d = 2
    Xstop = [0.0] * d
    lucky = False
    Xm = 0.0
    j = 0
    for j in range(1, it + 1):
        if j == 1:
            Um, HA, param_A, lucky = lanczos_krylov(A, U)   # :64
            Cm = Um[:, :Um.shape[1] - rk].T @ U             # :65
            Cm = Cm @ B @ Cm.T                              # :66
        else:
            Um, HA, param_A, lucky = lanczos_krylov(Um, HA, param_A)
        Gm = HA[:HA.shape[0] - rk, :]                       # :72
        nn = Gm.shape[0]
        Cm = _pad(Cm, nn)
        tGm = Gm + Cm
        if herm:                                            # :78-81
            Gm = (Gm + Gm.T) / 2
            tGm = (tGm + tGm.T) / 2
        d1 = _eig_sorted(tGm)
        d2 = _eig_sorted(Gm)
        Xm = _trace_formula(name, d1, d2)
        if j <= d:                                          # :104-118
            Xstop[j - 1] = Xm
        else:
            err = abs(Xm - Xstop[0])
            if err < tol:
                break
            Xstop = Xstop[1:] + [Xm]
        if lucky:                                           # :119-124
            break
    it_used = j
    if it_used == it:
        warnings.warn("TRACE_FUN_UPDATE:: Reached maximum number of iterations")
    return Xm, it_used, bool(lucky)


# --------------------------------------------------------------------------- fun_update
def fun_update(A, U, B, fun, tol=1e-12, it=None, debug=0, want_basis=False):
    """[Xm, iter, lucky, Um] = fun_update(A,U,B,fun,tol,it,debug)  (fun_update.m:1-137).

    ``want_basis`` stands in for MATLAB's ``nargout == 4`` (fun_update.m:69,77): False -> block
    Lanczos (returned Um is the 2-block window, as in the reference), True -> block Arnoldi."""
    name = fun_name(fun)
    f = _matfun(name)
    U = np.asarray(U.toarray() if sp.issparse(U) else U, dtype=np.float64)
    if U.ndim == 1:
        U = U[:, None]
    B = np.atleast_2d(np.asarray(B, dtype=np.float64))
    n = A.shape[0]
    if it is None:
        it = min(100, n)
    rk = U.shape[1]
    herm = _ishermitian(B)
    d = 2
    Xstop = []
    lucky = False
    j = 0
    Xm = None
    for j in range(1, it + 1):
        if not want_basis:                                  # :69-76
            if j == 1:
                Um, HA, param_A, lucky = lanczos_krylov(A, U)
                Cm = Um[:, :Um.shape[1] - rk].T @ U
                Cm = Cm @ B @ Cm.T
            else:
                Um, HA, param_A, lucky = lanczos_krylov(Um, HA, param_A)
        else:                                               # :77-91
            if j == 1:
                Um, KA, HA, param_A, lucky = arnoldi_krylov(A, U)
                Cm = Um[:, :Um.shape[1] - rk].T @ U
                Cm = Cm @ B @ Cm.T
            else:
                Um, KA, HA, param_A, lucky = arnoldi_krylov(Um, KA, HA, param_A)
            if Um.shape[1] >= Um.shape[0] / 2:              # :85-90 dense fallback
                Um = np.eye(U.shape[0])
                fA = _full(A)
                Xm = f(fA + U @ B @ U.T) - f(fA)
                return Xm, j, bool(lucky), Um
        Gm = HA[:HA.shape[0] - rk, :]
        Gm = (Gm + Gm.T) / 2                                # :94
        nn = Gm.shape[0]
        Cm = _pad(Cm, nn)
        if herm:
            tGm = Gm + (Cm + Cm.T) / 2
        else:
            tGm = Gm + Cm
        Xm = f(tGm) - f(Gm)                                 # :106
        if j <= d:
            Xstop.append(Xm)
        else:
            nn = Xm.shape[0]
            Xstop[0] = _pad(Xstop[0], nn)
            err = np.linalg.norm(Xm - Xstop[0], 2)          # :114
            if err < tol:
                break
            Xstop = Xstop[1:] + [Xm]
        if lucky:
            warnings.warn("FUN_UPDATE:: Detected lucky breakdown")
            break
    it_used = j
    if it_used == it:
        warnings.warn("FUN_UPDATE:: Reached maximum number of iterations")
    Um = Um[:, :Xm.shape[0]]                                # :137
    return Xm, it_used, bool(lucky), Um


# --------------------------------------------------------------------------- function_multiple_entries
def function_multiple_entries(A, omega, f, tol=1e-12, it=None, poles=np.inf, debug=0):
    """[X, iter] = function_multiple_entries(A,omega,f,tol,it,poles,debug)
    (function_multiple_entries.m:1-172).  ``omega`` is k x 2 with 1-based indices."""
    name = fun_name(f)
    fM = _matfun_entries(name)
    omega = np.atleast_2d(np.asarray(omega)).astype(np.int64)
    n = A.shape[0]
    if it is None:
        it = min(100, n)
    if not (np.isscalar(poles) and poles == np.inf):
        raise ValueError("FUNCTION_MULTIPLE_ENTRIES::Unsupported rational Krylov yet")
    k = omega.shape[0]
    notconverged = list(range(k))
    rk = 1

    def unique_stable(v):
        seen, out = set(), []
        for x in v:
            if x not in seen:
                seen.add(x)
                out.append(int(x))
        return out

    I0 = unique_stable(omega[:, 0])                         # :42 (captured by ``row``)
    row = {t: i for i, t in enumerate(I0)}
    I = list(I0)
    d = 3                                                   # :63
    Xstop = [[] for _ in range(k)]
    nI = len(I0)
    Um = [None] * nI
    KA = [None] * nI
    HA = [None] * nI
    PA = [None] * nI
    Gm = [None] * nI
    Uaux = np.zeros(nI)
    Xm = [None] * k
    j = 0
    for j in range(1, it + 1):
        for h in list(I):                                   # :86-110
            r = row[h]
            if j == 1:
                Uv = np.zeros((n, 1))
                Uv[h - 1, 0] = 1.0
                Um[r], KA[r], HA[r], PA[r], _ = arnoldi_krylov(A, Uv)
                temp = Um[r].T @ Uv
                Uaux[r] = temp[0, 0]                        # :94-95
            else:
                Um[r], KA[r], HA[r], PA[r], _ = arnoldi_krylov(Um[r], KA[r], HA[r], PA[r])
            Gm[r] = HA[r][:HA[r].shape[0] - rk, :]          # :106
        stop = 1
        for h in list(notconverged):                        # :113-152
            r = row[int(omega[h, 0])]
            Xm[h] = fM(Gm[r])                               # :118
            if j <= d:
                Xstop[h].append(Xm[h])
                stop = 0
            else:
                nn = Xm[h].shape[0]
                Xstop[h][0] = _pad(Xstop[h][0], nn)
                err = np.linalg.norm((Xm[h] - Xstop[h][0])[:, 0])   # :129
                if err > tol:
                    stop = 0
                else:
                    notconverged = [x for x in notconverged if x != h]
                    I = unique_stable(omega[notconverged, 0]) if notconverged else []
                Xstop[h] = Xstop[h][1:] + [Xm[h]]
        if stop == 1:
            break
    it_used = j
    if it_used == it:
        warnings.warn("FUNCTION_MULTIPLE_ENTRIES:: Reached maximum number of iterations")
    X = np.zeros(k)
    for q in range(k):                                      # :162-164
        r = row[int(omega[q, 0])]
        sz = Xm[q].shape[0]
        X[q] = Um[r][int(omega[q, 1]) - 1, :sz] @ Xm[q][:, 0] * Uaux[r]
    return X, it_used


# --------------------------------------------------------------------------- fmincon callbacks
def _low_rank_from_omega(X, Omega, n):
    # fun_and_grad_krylov_fun.m:38-54 (identical in _exp.m:57-73)
    aux = np.unique(Omega.ravel())
    k = aux.size
    pos = {int(a): i for i, a in enumerate(aux)}
    U = np.zeros((n, k))
    B = np.zeros((k, k))
    for jj in range(k):
        U[aux[jj] - 1, jj] = 1.0
    for jj in range(Omega.shape[0]):
        i1 = pos[int(Omega[jj, 0])]
        i2 = pos[int(Omega[jj, 1])]
        B[i1, i2] = X[jj]
        B[i2, i1] = X[jj]
    return U, B


def _check_hermitian(A, msg):
    As = sp.csr_matrix(A)
    if (As != As.T).nnz != 0:
        raise ValueError(msg)


def fun_and_grad_krylov_exp(X, A, Omega, eA, tol, it, debug=False):
    """[f, gr] = fun_and_grad_krylov_exp(X,A,Omega,eA,tol,it,debug)  (fun_and_grad_krylov_exp.m:1-113)."""
    _check_hermitian(A, "FUN_AND_GRAD_KRYLOV:: matrix A is not Hermitian")
    X = np.asarray(X, dtype=np.float64).ravel()
    Omega = np.atleast_2d(np.asarray(Omega)).astype(np.int64)
    eA = np.asarray(eA, dtype=np.float64).ravel()
    n = A.shape[0]
    nrmA, _ = normest(A, 1e-2)                              # :26
    if np.sum(np.abs(X)) == 0:                              # :30-54
        return 0.0, -2 * eA
    U, B = _low_rank_from_omega(X, Omega, n)
    eXm, _, _, Um = fun_update(A, U, B, "exp", tol * np.exp(nrmA), it, False, want_basis=True)  # :83
    f = -np.trace(eXm)                                      # :84
    L = Um[Omega[:, 0] - 1, :] @ eXm                        # :85-86 (diagonal of the product only)
    DeA = np.einsum("ij,ij->i", L, Um[Omega[:, 1] - 1, :])
    gr = -2 * (eA + DeA)                                    # :88
    return float(f), gr


def fun_and_grad_krylov_fun(X, A, Omega, fun, dfun, dfA, tol, it, debug=False, fun_M=None):
    """[f, gr] = fun_and_grad_krylov_fun(X,A,Omega,fun,dfun,dfA,tol,it,debug,fun_M)
    (fun_and_grad_krylov_fun.m:1-71)."""
    _check_hermitian(A, "FUN_AND_GRAD_KRYLOV_FCONNECTIVITY:: matrix A is not Hermitian")
    fn, dfn = fun_name(fun), fun_name(dfun)
    X = np.asarray(X, dtype=np.float64).ravel()
    Omega = np.atleast_2d(np.asarray(Omega)).astype(np.int64)
    dfA = np.asarray(dfA, dtype=np.float64).ravel()
    n = A.shape[0]
    nrmA, _ = normest(A, 1e-2)                              # :27
    if np.sum(np.abs(X)) == 0:                              # :31-35
        return 0.0, -2 * dfA
    U, B = _low_rank_from_omega(X, Omega, n)
    dfXm, _, _, Um = fun_update(A, U, B, dfn, tol * _SCALAR[dfn](nrmA), it, False, want_basis=True)  # :64
    tr, _, _ = trace_fun_update(A, U, B, tol * _SCALAR[fn](nrmA), it, False, fn)                     # :65
    f = -tr
    L = Um[Omega[:, 0] - 1, :] @ dfXm                       # :67-68
    DdfA = np.einsum("ij,ij->i", L, Um[Omega[:, 1] - 1, :])
    gr = -2 * (dfA + DdfA)                                  # :70
    return float(f), gr
