"""Oracle: Frechet-derivative factors and the Hessian callbacks.  TEST INFRASTRUCTURE ONLY.

Restates functions/multiple_frechet_eval.m, functions/hessianfcn_exp.m and functions/hessianfcn_fun.m
(row f3 of SURVEY.md section 8).
"""
import warnings

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from .krylov import arnoldi_krylov
from .updates import fun_name, _pad


def _matfun_frechet(name):
    # multiple_frechet_eval.m:60-76
    if name == "exp":
        return lambda M: sla.expm(M)
    if name == "sinh":
        return lambda M: (sla.expm(M) - sla.expm(-M)) / 2
    if name == "cosh":
        return lambda M: (sla.expm(M) + sla.expm(-M)) / 2
    raise ValueError(name)


def _unique_stable(v):
    seen, out = set(), []
    for x in v:
        if x not in seen:
            seen.add(x)
            out.append(int(x))
    return out


def multiple_frechet_eval(A, omega, f, tol=1e-12, it=None, poles=np.inf, debug=0):
    """[Um, Xm, Vm, row, col, iter] = multiple_frechet_eval(A,omega,f,tol,it,poles,debug)
    (multiple_frechet_eval.m:1-211).  Returns dict-based ``row`` / ``col`` lookups."""
    name = fun_name(f)
    fM = _matfun_frechet(name)
    omega = np.atleast_2d(np.asarray(omega)).astype(np.int64)
    n = A.shape[0]
    if it is None:
        it = min(100, n)
    if not (np.isscalar(poles) and poles == np.inf):
        raise ValueError("MULTIPLE_FRECHET_EVAL::Unsupported rational Krylov yet")
    k = omega.shape[0]
    notconverged = list(range(k))
    rk = 1
    AT = A.T.tocsr() if sp.issparse(A) else A.T
    I0 = _unique_stable(omega[:, 0])
    J0 = _unique_stable(omega[:, 1])
    row = {t: i for i, t in enumerate(I0)}
    col = {t: i for i, t in enumerate(J0)}
    I, J = list(I0), list(J0)
    d = 3
    Xstop = [[] for _ in range(k)]
    Um, KA, HA, PA, Gm = ([None] * len(I0) for _ in range(5))
    Vm, KB, HB, PB, Hm = ([None] * len(J0) for _ in range(5))
    Uaux = np.zeros(len(I0))
    Vaux = np.zeros(len(J0))
    Xm = [None] * k
    j = 0
    for j in range(1, it + 1):
        for h in list(I):                                   # :98-121 rows
            r = row[h]
            if j == 1:
                U = np.zeros((n, 1))
                U[h - 1, 0] = 1.0
                Um[r], KA[r], HA[r], PA[r], _ = arnoldi_krylov(A, U)
                Uaux[r] = (Um[r].T @ U)[0, 0]
            else:
                Um[r], KA[r], HA[r], PA[r], _ = arnoldi_krylov(Um[r], KA[r], HA[r], PA[r])
            Gm[r] = HA[r][:HA[r].shape[0] - rk, :]
        for h in list(J):                                   # :124-145 columns
            c = col[h]
            if j == 1:
                V = np.zeros((n, 1))
                V[h - 1, 0] = 1.0
                Vm[c], KB[c], HB[c], PB[c], _ = arnoldi_krylov(AT, V)
                Vaux[c] = (Vm[c].T @ V)[0, 0]
            else:
                Vm[c], KB[c], HB[c], PB[c], _ = arnoldi_krylov(Vm[c], KB[c], HB[c], PB[c])
            Hm[c] = HB[c][:HB[c].shape[0] - rk, :] @ np.linalg.inv(KB[c][:KB[c].shape[0] - rk, :])   # :144 (K = I)
        stop = 1
        for h in list(notconverged):                        # :148-195
            G = Gm[row[int(omega[h, 0])]]
            H = Hm[col[int(omega[h, 1])]]
            Cm = np.zeros((G.shape[0], H.shape[0]))
            Cm[0, 0] = Uaux[row[int(omega[h, 0])]] * Vaux[col[int(omega[h, 1])]]
            Fm = np.block([[G, Cm], [np.zeros((H.shape[1], G.shape[1])), H.T]])
            Fm = fM(Fm)
            Xm[h] = Fm[:G.shape[0], G.shape[1]:]
            if j <= d:
                Xstop[h].append(Xm[h])
                stop = 0
            else:
                nn = Xm[h].shape[0]
                Xstop[h][0] = _pad(Xstop[h][0], nn)
                err = np.linalg.norm(Xm[h] - Xstop[h][0], 2)
                if err > tol:
                    stop = 0
                else:
                    notconverged = [x for x in notconverged if x != h]
                    I = _unique_stable(omega[notconverged, 0]) if notconverged else []
                    J = _unique_stable(omega[notconverged, 1]) if notconverged else []
                Xstop[h] = Xstop[h][1:] + [Xm[h]]
        if stop == 1:
            break
    it_used = j
    if it_used == it:
        warnings.warn("MULTIPLE_FRECHET_EVAL:: Reached maximum number of iterations")
    Um = [u[:, :u.shape[1] - rk] for u in Um]               # :203-209
    Vm = [v[:, :v.shape[1] - rk] for v in Vm]
    return Um, Xm, Vm, row, col, it_used


def _hessian(X, A, Omega, f, tol, it):
    # hessianfcn_exp.m:3-16 / hessianfcn_fun.m:3-16
    n = A.shape[0]
    Omega = np.atleast_2d(np.asarray(Omega)).astype(np.int64)
    X = np.asarray(X, dtype=np.float64).ravel()
    XX = sp.csr_matrix((X, (Omega[:, 0] - 1, Omega[:, 1] - 1)), shape=(n, n))
    Atilde = (sp.csr_matrix(A) + XX + XX.T).tocsr()
    Um, Xm, Vm, row, col, _ = multiple_frechet_eval(Atilde, Omega, f, tol, it, np.inf, False)
    m = Omega.shape[0]
    Hes = np.zeros((m, m))
    for j in range(m):
        h, k = int(Omega[j, 0]), int(Omega[j, 1])
        s0, s1 = Xm[j].shape
        for l in range(j, m):
            Hes[j, l] = Um[row[h]][Omega[l, 0] - 1, :s0] @ Xm[j] @ Vm[col[k]][Omega[l, 1] - 1, :s1]
        Hes[j + 1:, j] = Hes[j, j + 1:]
    return -2 * Hes


def hessianfcn_exp(X, A, Omega, tol, it):
    """Hes = hessianfcn_exp(X,A,Omega,tol,it)  (hessianfcn_exp.m:1-17)."""
    return _hessian(X, A, Omega, "exp", tol, it)


def hessianfcn_fun(X, A, Omega, f, tol, it):
    """Hes = hessianfcn_fun(X,A,Omega,f,tol,it)  (hessianfcn_fun.m:1-17)."""
    return _hessian(X, A, Omega, f, tol, it)
