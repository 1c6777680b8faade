"""Host-side mirror of the reference's function interface (functions/*.m), same names, argument
order, defaults and error strings - each a thin marshaller into the C ABI of libkrylov_b200.so.

``A`` may be a SciPy sparse matrix / dense array (uploaded on each call) or an ``engine.Matrix``
(device-resident, the analogue of passing the operator struct of lanczos_krylov.m:78-79).
Index lists are 1-based as in the reference.  Function selectors are 'exp' / 'sinh' / 'cosh'
(or np.exp / np.sinh / np.cosh).  No function here computes on the CPU.
"""
import ctypes as C
import math
import warnings

import numpy as np
import scipy.sparse as sp

from ._lib import check
from .engine import Context, Matrix, Dense, fun_id, _ptr, _f64_cm


def _mat(A, ctx=None):
    return Matrix.wrap(A, ctx)


# ----------------------------------------------------------------------------- throughput mode
def slq_trace(A, Z, m=30, fun="exp", return_details=False):
    """Throughput-mode estimator: mean_z ||z||^2 e1' f(T_z) e1 over all probe
    columns of Z at once.  Z: n x k array (host) or engine.Dense (device-resident)."""
    M = _mat(A)
    lib, ctx = M.ctx.lib, M.ctx
    tr = C.c_double()
    if isinstance(Z, Dense):
        k = Z.k
    else:
        Z, ldz = _f64_cm(Z)
        k = Z.shape[1]
    vals = np.empty(k) if return_details else None
    al = np.empty((m, k), order="F") if return_details else None
    be = np.empty((m, k), order="F") if return_details else None
    if isinstance(Z, Dense):
        check(lib.kr_slq_trace_dev(ctx.h, M.h, Z.h, m, fun_id(fun), C.byref(tr), _ptr(vals), _ptr(al), _ptr(be)))
    else:
        check(lib.kr_slq_trace(ctx.h, M.h, k, _ptr(Z), ldz, m, fun_id(fun), C.byref(tr), _ptr(vals), _ptr(al),
                               _ptr(be)))
    if return_details:
        return tr.value, vals, al, be
    return tr.value
