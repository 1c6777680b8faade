"""MEX host for the MATLAB-subset interpreter (oracle/mlab): TEST INFRASTRUCTURE ONLY.

Builds krylov_robustness_b200/mex/kr_mex.c against the stub MEX runtime of this directory as a shared library and
exposes it to interpreted MATLAB code as the function ``kr_mex(op, ...)``, marshalling interpreter values to mxArrays
and back.  With it the drop-in wrappers krylov_robustness_b200/matlab/*.m run as MATLAB code - the language of the
reference - end to end: wrapper .m -> mexFunction -> C ABI -> CUDA (tests/test_gpu_dropin_matlab.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
STUB = os.path.join(ROOT, "tests", "mex_stub")
PKG = os.path.join(ROOT, "krylov_robustness_b200")

mxDOUBLE, mxLOGICAL, mxCHAR, mxUINT64 = 6, 3, 4, 13


class UInt64:
    """A uint64 scalar (device handles) travelling through the interpreter as an opaque value."""
    __slots__ = ("v",)
    shape = (1, 1)

    def __init__(self, v):
        self.v = int(v)

    def __repr__(self):
        return "uint64(%d)" % self.v


class OnCleanup:
    """onCleanup(@() ...): runs the handle when the last reference to the object goes away."""

    def __init__(self, interp, fh):
        self.interp, self.fh = interp, fh

    def __del__(self):
        try:
            self.interp.call_handle(self.fh, [], 0)
        except Exception:
            pass


def build_host(outdir):
    so = os.path.join(str(outdir), "libkr_mex_host.so")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-shared", "-fPIC", "-I" + STUB, "-I" + os.path.join(ROOT, "include"),
           os.path.join(PKG, "mex", "kr_mex.c"), os.path.join(STUB, "mex_stub.c"), "-o", so,
           "-L" + PKG, "-l:libkrylov_b200.so", "-Wl,-rpath," + PKG, "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the MEX host failed:\n" + r.stderr)
    return so


class MexHost:
    def __init__(self, so):
        L = self.L = C.CDLL(so)
        vp, sz = C.c_void_p, C.c_size_t
        L.mxCreateDoubleMatrix.restype = vp
        L.mxCreateDoubleMatrix.argtypes = [sz, sz, C.c_int]
        L.mxCreateLogicalScalar.restype = vp
        L.mxCreateLogicalScalar.argtypes = [C.c_int]
        L.mxCreateNumericMatrix.restype = vp
        L.mxCreateNumericMatrix.argtypes = [sz, sz, C.c_int, C.c_int]
        L.mxCreateSparse.restype = vp
        L.mxCreateSparse.argtypes = [sz, sz, sz, C.c_int]
        L.mxCreateString.restype = vp
        L.mxCreateString.argtypes = [C.c_char_p]
        L.mxDestroyArray.argtypes = [vp]
        for f in ("mxGetPr", "mxGetData", "mxGetJc", "mxGetIr"):
            getattr(L, f).restype = vp
            getattr(L, f).argtypes = [vp]
        for f in ("mxGetM", "mxGetN"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp]
        for f in ("mxIsSparse", "mxGetClassID"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [vp]
        L.kr_stub_call.restype = C.c_int
        L.kr_stub_call.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp), C.c_char_p, sz]
        self.calls = {}

    # ---- python value -> mxArray*
    def to_mx(self, v):
        L = self.L
        if isinstance(v, UInt64):
            a = L.mxCreateNumericMatrix(1, 1, mxUINT64, 0)
            C.cast(L.mxGetData(a), C.POINTER(C.c_uint64))[0] = v.v
            return a
        if isinstance(v, str):
            return L.mxCreateString(v.encode())
        if sp.issparse(v):
            S = sp.csc_matrix(v).astype(np.float64)
            S.sort_indices()
            a = L.mxCreateSparse(S.shape[0], S.shape[1], max(S.nnz, 1), 0)
            jc = np.ascontiguousarray(S.indptr, dtype=np.uint64)      # keep the temporaries alive across memmove
            ir = np.ascontiguousarray(S.indices, dtype=np.uint64)
            pr = np.ascontiguousarray(S.data, dtype=np.float64)
            C.memmove(L.mxGetJc(a), jc.ctypes.data, (S.shape[1] + 1) * 8)
            if S.nnz:
                C.memmove(L.mxGetIr(a), ir.ctypes.data, S.nnz * 8)
                C.memmove(L.mxGetPr(a), pr.ctypes.data, S.nnz * 8)
            return a
        v = np.asarray(v)
        if v.dtype == np.bool_ and v.size == 1:
            return L.mxCreateLogicalScalar(int(bool(v.reshape(-1)[0])))
        v = np.atleast_2d(v.astype(np.float64))
        a = L.mxCreateDoubleMatrix(v.shape[0], v.shape[1], 0)
        if v.size:
            f = np.asfortranarray(v)
            C.memmove(L.mxGetPr(a), f.ctypes.data, v.size * 8)
        return a

    # ---- mxArray* -> python value
    def from_mx(self, a):
        L = self.L
        m, n = L.mxGetM(a), L.mxGetN(a)
        cls = L.mxGetClassID(a)
        if L.mxIsSparse(a):
            jc = np.ctypeslib.as_array(C.cast(L.mxGetJc(a), C.POINTER(C.c_uint64)), (n + 1,)).astype(np.int64)
            nnz = int(jc[-1])
            ir = np.ctypeslib.as_array(C.cast(L.mxGetIr(a), C.POINTER(C.c_uint64)), (max(nnz, 1),))[:nnz].astype(np.int64)
            pr = np.ctypeslib.as_array(C.cast(L.mxGetPr(a), C.POINTER(C.c_double)), (max(nnz, 1),))[:nnz].copy()
            return sp.csc_matrix((pr, ir, jc), shape=(m, n))
        if cls == mxUINT64:
            return UInt64(C.cast(L.mxGetData(a), C.POINTER(C.c_uint64))[0])
        if cls == mxLOGICAL:
            return np.array([[bool(C.cast(L.mxGetData(a), C.POINTER(C.c_ubyte))[0])]])
        if cls == mxCHAR:
            return C.string_at(L.mxGetData(a), m * n).decode()
        if m * n == 0:
            return np.zeros((m, n))
        flat = np.ctypeslib.as_array(C.cast(L.mxGetPr(a), C.POINTER(C.c_double)), (m * n,)).copy()
        return flat.reshape((m, n), order="F")

    def call(self, args, nlhs):
        L = self.L
        self.calls[args[0]] = self.calls.get(args[0], 0) + 1
        nout = max(nlhs, 1)
        prhs = (C.c_void_p * len(args))(*[self.to_mx(v) for v in args])
        plhs = (C.c_void_p * nout)()
        err = C.create_string_buffer(1024)
        rc = L.kr_stub_call(nlhs, plhs, len(args), prhs, err, 1024)
        for p in prhs:
            L.mxDestroyArray(p)
        if rc:
            from oracle.mlab import MatlabError
            raise MatlabError(err.value.decode(errors="replace"))
        out = []
        for p in plhs:
            if p:
                out.append(self.from_mx(p))
                L.mxDestroyArray(p)
        return out

    def install(self, interp):
        """Make kr_mex(...) and onCleanup(...) available to interpreted code."""
        interp.builtins["kr_mex"] = lambda I, a, n: self.call(list(a), n)
        interp.builtins["onCleanup"] = lambda I, a, n: OnCleanup(I, a[0])
        return interp
