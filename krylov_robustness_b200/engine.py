"""Thin object layer over the C ABI: Context (one per GPU), Matrix (device CSR), Dense (device block).

Everything numeric happens inside libkrylov_b200.so; this module only marshals NumPy / SciPy
buffers into the plain pointers the C ABI takes.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import check

FUN = {"exp": 0, "sinh": 1, "cosh": 2}


def fun_id(fun):
    """Map a function selector ('exp'/'sinh'/'cosh' or np.exp/np.sinh/np.cosh) to enum kr_fun.
    Mirrors the reference's isequal(f,@exp) dispatch (functions/fun_update.m:43-59)."""
    if isinstance(fun, str):
        if fun in FUN:
            return FUN[fun]
    else:
        for name, uf in (("exp", np.exp), ("sinh", np.sinh), ("cosh", np.cosh)):
            if fun is uf:
                return FUN[name]
    raise ValueError("unsupported function handle %r (device path supports exp, sinh, cosh)" % (fun,))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64_cm(a):
    """Column-major fp64 view/copy + leading dimension."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    if not a.flags.f_contiguous:
        a = np.asfortranarray(a)
    return a, max(a.shape[0], 1)


class Context:
    _default = {}

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.kr_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def sync(self):
        check(self.lib.kr_ctx_sync(self.h))

    def counters(self):
        out = (C.c_int64 * 5)()
        check(self.lib.kr_ctx_counters(self.h, out))
        return dict(zip(("launches", "spmm_launches", "matvecs", "h2d_bytes", "d2h_bytes"), list(out)))

    def set_timing(self, on):
        check(self.lib.kr_ctx_set_timing(self.h, int(bool(on))))

    def spmm_time(self, reset=True):
        ms = C.c_double()
        n = C.c_int64()
        check(self.lib.kr_ctx_spmm_time(self.h, int(reset), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def stream(self):
        return self.lib.kr_ctx_stream(self.h)

    def close(self):
        if self.h:
            self.lib.kr_ctx_destroy(self.h)
            self.h = None


class Matrix:
    """Device-resident CSR of A.  Accepts any SciPy sparse matrix / dense array."""

    def __init__(self, A, ctx=None):
        self.ctx = ctx or Context.default()
        A = sp.csr_matrix(A)
        if A.shape[0] != A.shape[1]:
            raise ValueError("The matrix A should be square")
        A = A.astype(np.float64)
        self.n = A.shape[0]
        self.shape = A.shape
        rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
        ci = np.ascontiguousarray(A.indices, dtype=np.int64)
        va = np.ascontiguousarray(A.data, dtype=np.float64)
        h = C.c_void_p()
        check(self.ctx.lib.kr_matrix_create(self.ctx.h, self.n, int(A.nnz), _ptr(rp), _ptr(ci), _ptr(va),
                                            C.byref(h)))
        self.h = h

    @classmethod
    def from_csr_arrays(cls, n, indptr, indices, data, ctx=None):
        """Device matrix straight from CSR arrays (0-based; `data` may be a scalar for a uniform-valued pattern):
        skips the SciPy conversions of the constructor, which copy the arrays several times - noticeable at
        2^28 stored entries (config C4)."""
        self = cls.__new__(cls)
        self.ctx = ctx or Context.default()
        self.n = int(n)
        self.shape = (self.n, self.n)
        rp = np.ascontiguousarray(indptr, dtype=np.int64)
        ci = np.ascontiguousarray(indices, dtype=np.int64)
        va = np.full(ci.size, float(data)) if np.isscalar(data) else np.ascontiguousarray(data, dtype=np.float64)
        h = C.c_void_p()
        check(self.ctx.lib.kr_matrix_create(self.ctx.h, self.n, int(ci.size), _ptr(rp), _ptr(ci), _ptr(va), C.byref(h)))
        self.h = h
        return self

    @staticmethod
    def wrap(A, ctx=None):
        return A if isinstance(A, Matrix) else Matrix(A, ctx)

    def info(self):
        n, nnz = C.c_int64(), C.c_int64()
        s, p, nn = C.c_int(), C.c_int(), C.c_int()
        check(self.ctx.lib.kr_matrix_info(self.h, C.byref(n), C.byref(nnz), C.byref(s), C.byref(p), C.byref(nn)))
        return dict(n=n.value, nnz=nnz.value, symmetric=bool(s.value), pattern_only=bool(p.value),
                    nonnegative=bool(nn.value))

    def set_edges(self, i, j, v):
        i = np.ascontiguousarray(i, dtype=np.int64).ravel()
        j = np.ascontiguousarray(j, dtype=np.int64).ravel()
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), i.shape))
        check(self.ctx.lib.kr_matrix_set_edges(self.h, i.size, _ptr(i), _ptr(j), _ptr(v)))

    def replicate(self, devices=None):
        """Multi-GPU in one process (kr_matrix_replicate): a replica of A on every listed device (None: all visible
        ones).  Afterwards trace_fun_update_edges splits its candidates and slq_trace (host probes) its columns across
        the replicas; set_edges edits all of them.  Returns the number of GPUs holding A."""
        if devices is None:
            check(self.ctx.lib.kr_matrix_replicate(self.h, 0, None))
        else:
            d = np.ascontiguousarray(devices, dtype=np.int32)
            check(self.ctx.lib.kr_matrix_replicate(self.h, int(d.size), _ptr(d)))
        return self.replicas()

    def replicas(self):
        return int(self.ctx.lib.kr_matrix_replicas(self.h))

    def multiply(self, alpha, beta, w):
        """The reference's operator-struct plug-in point: A.multiply(1.0, 0.0, w)
        (functions/lanczos_krylov.m:78-79)."""
        return alpha * self.matmul(w)

    def matmul(self, X):
        X, ldx = _f64_cm(X)
        if X.shape[0] != self.n:
            raise ValueError("The block vector b has wrong number of rows")
        Y = np.empty_like(X, order="F")
        check(self.ctx.lib.kr_spmm(self.ctx.h, self.h, X.shape[1], _ptr(X), ldx, _ptr(Y), ldx))
        return Y

    __matmul__ = matmul

    def close(self):
        if self.h:
            self.ctx.lib.kr_matrix_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Dense:
    """Device-resident n x k fp64 block (panel-major on the device)."""

    def __init__(self, n, k, ctx=None):
        self.ctx = ctx or Context.default()
        self.n, self.k = int(n), int(k)
        h = C.c_void_p()
        check(self.ctx.lib.kr_dense_create(self.ctx.h, self.n, self.k, C.byref(h)))
        self.h = h

    def upload(self, host, ld=None):
        """host: column-major fp64 array (n x k) or a raw pointer (int) with leading dimension ld."""
        if isinstance(host, int):
            check(self.ctx.lib.kr_dense_upload(self.h, C.c_void_p(host), ld or self.n))
            return self
        a, lda = _f64_cm(host)
        assert a.shape == (self.n, self.k)
        check(self.ctx.lib.kr_dense_upload(self.h, _ptr(a), lda))
        return self

    def download(self):
        out = np.empty((self.n, self.k), dtype=np.float64, order="F")
        check(self.ctx.lib.kr_dense_download(self.h, _ptr(out), max(self.n, 1)))
        return out

    def fill_rademacher(self, seed, col_offset=0):
        check(self.ctx.lib.kr_dense_fill_rademacher(self.h, C.c_uint64(seed), int(col_offset)))
        return self

    def close(self):
        if self.h:
            self.ctx.lib.kr_dense_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rademacher_host(n, k, seed, col_offset=0):
    """NumPy twin of the device's counter-based Rademacher stream (csrc/dense.cuh splitmix64)."""
    i = np.arange(n, dtype=np.uint64)[:, None]
    c = (np.arange(k, dtype=np.uint64) + np.uint64(col_offset))[None, :]
    x = np.uint64(seed) ^ ((c << np.uint64(32)) | i)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return np.where((x >> np.uint64(63)) == 1, -1.0, 1.0)
