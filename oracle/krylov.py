"""Oracle: block Lanczos / block Arnoldi basis builders.  TEST INFRASTRUCTURE ONLY.

Restates functions/lanczos_krylov.m and functions/arnoldi_krylov.m of the reference.
Dense blocks are NumPy arrays (n x bs), ``A`` is a SciPy sparse matrix, a dense array, or any
object with a ``multiply(alpha, beta, w)`` method (the reference's operator-struct plug-in point,
lanczos_krylov.m:78-79).
"""
import numpy as np


class KrylovParams:
    """The reference's ``params`` struct (lanczos_krylov.m:55-57): continuation block + operator."""

    def __init__(self, last, A):
        self.last = last
        self.A = A


def _is_operator(A):
    return hasattr(A, "multiply") and not hasattr(A, "shape")


def _apply(A, w):
    # lanczos_krylov.m:78-82 / arnoldi_krylov.m:83-87
    if _is_operator(A):
        return A.multiply(1.0, 0.0, w)
    return np.asarray(A @ w)


def _qr0(w):
    # MATLAB qr(w, 0): Householder thin QR (LAPACK dgeqrf + dorgqr); numpy calls the same routines.
    return np.linalg.qr(w, mode="reduced")


def _cgs2(V, w):
    # lanczos_krylov.m:109-115 / arnoldi_krylov.m:119-125 ("mgs_orthogonalize": CGS applied twice)
    h = V.T @ w
    w = w - V @ h
    h1 = V.T @ w
    h = h + h1
    w = w - V @ h1
    return w, h


def _grow(H, bs):
    out = np.zeros((H.shape[0] + bs, H.shape[1] + bs))
    out[:H.shape[0], :H.shape[1]] = H
    return out


def _check(A, b):
    # lanczos_krylov.m:32-43
    if not _is_operator(A):
        m, n = A.shape
        if m != n:
            raise ValueError("The matrix A should be square")
        if n != b.shape[0]:
            raise ValueError("The block vector b has wrong number of rows")


# --------------------------------------------------------------------------- Lanczos
def _lanczos_add_inf_pole(V, H, A, w):
    """lanczos_krylov.m:73-101."""
    lucky_tol = 1e-8
    lucky = False
    bs = w.shape[1]
    w = _apply(A, w)
    H = _grow(H, bs)                                        # :85
    R, C = H.shape
    w, h = _cgs2(V, w)                                      # :88
    r0 = max(1, R - 3 * bs + 1) - 1                         # 0-based start row
    H[r0:R - bs, C - bs:C] = h
    w, rfac = _qr0(w)                                       # :90
    H[R - bs:R, C - bs:C] = rfac
    if np.linalg.norm(rfac, "fro") < lucky_tol:             # :91-93
        lucky = True
    if V.shape[1] == bs:                                    # :94-99
        V = np.hstack([V, w])
    else:
        V = np.hstack([V[:, bs:2 * bs], w])
    return V, H, w, lucky


def lanczos_krylov(*args):
    """[V,H,params,lucky] = lanczos_krylov(A,b) | lanczos_krylov(V,H,params)  (lanczos_krylov.m:1-67)."""
    if len(args) not in (2, 3):
        raise ValueError("Called with the wrong number of arguments")
    if len(args) == 2:
        A, b = args
        b = np.asarray(b, dtype=np.float64)
        if b.ndim == 1:
            b = b[:, None]
        _check(A, b)
        bs = b.shape[1]
        V, _ = _qr0(b)                                      # :48
        H = np.zeros((bs, 0))
        V, H, w, lucky = _lanczos_add_inf_pole(V, H, A, V)  # :52
        return V, H, KrylovParams(w, A), lucky
    V, H, params = args
    V, H, w, lucky = _lanczos_add_inf_pole(V, H, params.A, params.last)
    params.last = w
    return V, H, params, lucky


# --------------------------------------------------------------------------- Arnoldi
def _arnoldi_add_inf_pole(V, K, H, A, w):
    """arnoldi_krylov.m:78-111."""
    lucky_tol = 1e-12
    lucky = False
    bs = w.shape[1]
    w = _apply(A, w)
    w, h = _cgs2(V, w)                                      # :90
    H = _grow(H, bs)
    K = _grow(K, bs)
    R, C = H.shape
    H[:R - bs, C - bs:C] = h                                # :96
    K[R - 2 * bs:R - bs, C - bs:C] = np.eye(bs)             # :97
    w, r = _qr0(w)                                          # :99
    if np.linalg.norm(r, 2) < lucky_tol:                    # :100-102
        lucky = True
    hh = V.T @ w                                            # :104-106 third reorthogonalisation
    w = w - V @ hh
    H[:R - bs, C - bs:C] += hh @ r
    H[R - bs:R, C - bs:C] = r                               # :108
    V = np.hstack([V, w])                                   # :110
    return V, K, H, w, lucky


def arnoldi_krylov(*args):
    """[V,K,H,params,lucky] = arnoldi_krylov(A,b) | arnoldi_krylov(V,K,H,params)  (arnoldi_krylov.m:1-72)."""
    if len(args) not in (2, 4):
        raise ValueError("Called with the wrong number of arguments")
    if len(args) == 2:
        A, b = args
        b = np.asarray(b, dtype=np.float64)
        if b.ndim == 1:
            b = b[:, None]
        _check(A, b)
        bs = b.shape[1]
        V, _ = _qr0(b)
        H = np.zeros((bs, 0))
        K = np.zeros((bs, 0))
        V, K, H, w, lucky = _arnoldi_add_inf_pole(V, K, H, A, V)
        return V, K, H, KrylovParams(w, A), lucky
    V, K, H, params = args
    V, K, H, w, lucky = _arnoldi_add_inf_pole(V, K, H, params.A, params.last)
    params.last = w
    return V, K, H, params, lucky
