"""Built-in functions of the MATLAB-subset interpreter (see oracle/mlab/__init__.py).  TEST INFRASTRUCTURE ONLY.

MATLAB's built-ins are closed source; these stand-ins call the same LAPACK routines through NumPy / SciPy (`qr` ->
dgeqrf + dorgqr, `eig` of a Hermitian matrix -> dsyevd, `expm` -> SciPy's Al-Mohy-Higham scaling and squaring) and
restate the two MathWorks m-files the path uses (`normest`: power iteration on A'A; `normest1`: Hager / Higham-Tisseur
block 1-norm estimator, written here for the t = 1 column the reference asks for) from their published algorithms."""
import os
import tempfile

import numpy as np
import scipy.io as sio
import scipy.linalg as sla
import scipy.sparse as sp

from .values import MatlabError, Cell, Struct, FH, EMPTY, scalar, is_num, dense, as_float, to_float, to_int
from . import ops

TABLE = {}


def builtin(*names):
    def deco(f):
        for n in names:
            TABLE[n] = f
        return f
    return deco


def _isstr(v, s=None):
    return isinstance(v, str) and (s is None or v == s)


def _dims(args):
    """zeros(n) / zeros(m,n) / zeros([m n]) -> (m, n)"""
    if not args:
        return 1, 1
    if len(args) == 1:
        d = dense(args[0]).reshape(-1)
        if d.size == 1:
            return max(int(d[0]), 0), max(int(d[0]), 0)
        return max(int(d[0]), 0), max(int(d[1]), 0)
    return max(to_int(args[0]), 0), max(to_int(args[1]), 0)


# ------------------------------------------------------------------------------------------------- constants
@builtin("pi")
def _pi(I, a, n):
    return scalar(np.pi)


@builtin("eps")
def _eps(I, a, n):
    if a:
        return np.spacing(np.abs(as_float(a[0])))
    return scalar(np.finfo(float).eps)


@builtin("inf", "Inf")
def _inf(I, a, n):
    m, k = _dims(a)
    return np.full((m, k), np.inf)


@builtin("nan", "NaN")
def _nan(I, a, n):
    m, k = _dims(a)
    return np.full((m, k), np.nan)


@builtin("true")
def _true(I, a, n):
    m, k = _dims(a)
    return np.ones((m, k), dtype=bool)


@builtin("false")
def _false(I, a, n):
    m, k = _dims(a)
    return np.zeros((m, k), dtype=bool)


# ------------------------------------------------------------------------------------------------- construction
@builtin("zeros")
def _zeros(I, a, n):
    return np.zeros(_dims(a))


@builtin("ones")
def _ones(I, a, n):
    return np.ones(_dims(a))


@builtin("eye")
def _eye(I, a, n):
    m, k = _dims(a)
    return np.eye(m, k)


@builtin("speye")
def _speye(I, a, n):
    m, k = _dims(a)
    return sp.eye(m, k, format="csc")


@builtin("sparse")
def _sparse(I, a, n):
    if len(a) == 1:
        return sp.csc_matrix(dense(a[0]).astype(np.float64) if not sp.issparse(a[0]) else a[0])
    if len(a) == 2:
        return sp.csc_matrix((to_int(a[0]), to_int(a[1])))
    i = dense(a[0]).reshape(-1, order="F").astype(np.int64) - 1
    j = dense(a[1]).reshape(-1, order="F").astype(np.int64) - 1
    v = as_float(a[2]).reshape(-1, order="F")
    k = max(i.size, j.size, v.size)
    i = np.broadcast_to(i, (k,)) if i.size == 1 else i
    j = np.broadcast_to(j, (k,)) if j.size == 1 else j
    v = np.broadcast_to(v, (k,)) if v.size == 1 else v
    if len(a) >= 5:
        shape = (to_int(a[3]), to_int(a[4]))
    else:
        shape = (int(i.max()) + 1 if k else 0, int(j.max()) + 1 if k else 0)
    if k and (i.min() < 0 or j.min() < 0 or i.max() >= shape[0] or j.max() >= shape[1]):
        raise MatlabError("Index exceeds matrix dimensions (sparse).")
    m = sp.coo_matrix((v, (i, j)), shape=shape).tocsc()     # duplicates are summed, as in MATLAB
    m.eliminate_zeros()
    return m


@builtin("full")
def _full(I, a, n):
    return dense(a[0])


@builtin("double")
def _double(I, a, n):
    return as_float(a[0]) if not sp.issparse(a[0]) else a[0]


@builtin("logical")
def _logical(I, a, n):
    return dense(a[0]) != 0


@builtin("diag")
def _diag(I, a, n):
    v = a[0]
    k = to_int(a[1]) if len(a) > 1 else 0
    if sp.issparse(v):
        if min(v.shape) == 1:
            return sp.diags(v.toarray().reshape(-1), k, format="csc")
        return sp.csc_matrix(v.diagonal(k).reshape(-1, 1))
    v = dense(v)
    if min(v.shape) == 1 or v.size == 0:
        return np.diag(v.reshape(-1), k)
    return np.diag(v, k).reshape(-1, 1)


@builtin("tril")
def _tril(I, a, n):
    k = to_int(a[1]) if len(a) > 1 else 0
    if sp.issparse(a[0]):
        m = sp.tril(a[0], k, format="csc")
        m.eliminate_zeros()
        return m
    return np.tril(a[0], k)


@builtin("triu")
def _triu(I, a, n):
    k = to_int(a[1]) if len(a) > 1 else 0
    if sp.issparse(a[0]):
        m = sp.triu(a[0], k, format="csc")
        m.eliminate_zeros()
        return m
    return np.triu(a[0], k)


@builtin("struct")
def _struct(I, a, n):
    s = Struct()
    for k in range(0, len(a), 2):
        s.f[a[k]] = a[k + 1]
    return s


@builtin("cell")
def _cell(I, a, n):
    m, k = _dims(a)
    c = np.empty((m, k), dtype=object)
    for i in range(m):
        for j in range(k):
            c[i, j] = EMPTY
    return Cell(c)


# ------------------------------------------------------------------------------------------------- shape queries
def _shape(v):
    if isinstance(v, str):
        return (1, len(v)) if v else (0, 0)
    if isinstance(v, Cell) or is_num(v):
        return v.shape
    return (1, 1)


@builtin("size")
def _size(I, a, n):
    shp = _shape(a[0])
    if len(a) > 1:
        d = to_int(a[1])
        return scalar(float(shp[d - 1] if d <= 2 else 1))
    if n <= 1:
        return np.array([[float(shp[0]), float(shp[1])]])
    out = [scalar(float(shp[0])), scalar(float(shp[1]))]
    return tuple(out + [scalar(1.0)] * (n - 2))


@builtin("length")
def _length(I, a, n):
    shp = _shape(a[0])
    return scalar(float(0 if 0 in shp else max(shp)))


@builtin("numel")
def _numel(I, a, n):
    shp = _shape(a[0])
    return scalar(float(shp[0] * shp[1]))


@builtin("isempty")
def _isempty(I, a, n):
    shp = _shape(a[0])
    return scalar(shp[0] * shp[1] == 0)


@builtin("isscalar")
def _isscalar(I, a, n):
    return scalar(_shape(a[0]) == (1, 1))


@builtin("isvector")
def _isvector(I, a, n):
    s = _shape(a[0])
    return scalar((s[0] == 1 or s[1] == 1) and s[0] * s[1] >= 1)


@builtin("isstruct")
def _isstruct(I, a, n):
    return scalar(isinstance(a[0], Struct))


@builtin("iscell")
def _iscell(I, a, n):
    return scalar(isinstance(a[0], Cell))


@builtin("ischar")
def _ischar(I, a, n):
    return scalar(isinstance(a[0], str))


@builtin("isnumeric")
def _isnumeric(I, a, n):
    return scalar(is_num(a[0]) and a[0].dtype != np.bool_)


@builtin("isfloat")
def _isfloat(I, a, n):
    return scalar(is_num(a[0]) and a[0].dtype != np.bool_)


@builtin("islogical")
def _islogical(I, a, n):
    return scalar(is_num(a[0]) and a[0].dtype == np.bool_)


@builtin("issparse")
def _issparse(I, a, n):
    return scalar(sp.issparse(a[0]))


@builtin("isreal")
def _isreal(I, a, n):
    return scalar(is_num(a[0]) and not np.iscomplexobj(a[0]))


@builtin("isa")
def _isa(I, a, n):
    cls = a[1]
    v = a[0]
    if cls == "function_handle":
        return scalar(isinstance(v, FH))
    if cls in ("double", "float", "numeric"):
        return scalar(is_num(v) and v.dtype != np.bool_)
    return scalar(False)


@builtin("class")
def _class(I, a, n):
    v = a[0]
    if isinstance(v, str):
        return "char"
    if isinstance(v, Cell):
        return "cell"
    if isinstance(v, Struct):
        return "struct"
    if isinstance(v, FH):
        return "function_handle"
    return "logical" if v.dtype == np.bool_ else "double"


@builtin("nnz")
def _nnz(I, a, n):
    v = a[0]
    if sp.issparse(v):
        return scalar(float(v.count_nonzero()))
    return scalar(float(np.count_nonzero(dense(v))))


def _sym_equal(A, B):
    if sp.issparse(A):
        return (A != B).nnz == 0
    return np.array_equal(A, B)


@builtin("issymmetric")
def _issymmetric(I, a, n):
    A = a[0]
    if A.shape[0] != A.shape[1]:
        return scalar(False)
    return scalar(_sym_equal(A, ops.transpose(A, False)))


@builtin("ishermitian")
def _ishermitian(I, a, n):
    A = a[0]
    if A.shape[0] != A.shape[1]:
        return scalar(False)
    return scalar(_sym_equal(A, ops.transpose(A, True)))


def _isequal2(x, y):
    if isinstance(x, FH) or isinstance(y, FH):
        return isinstance(x, FH) and isinstance(y, FH) and x.name is not None and x.name == y.name
    if isinstance(x, str) or isinstance(y, str):
        if isinstance(x, str) and isinstance(y, str):
            return x == y
        xs, ys = dense(x), dense(y)
        return xs.shape == ys.shape and np.array_equal(xs, ys)
    if isinstance(x, Cell) or isinstance(y, Cell):
        if not (isinstance(x, Cell) and isinstance(y, Cell)) or x.shape != y.shape:
            return False
        return all(_isequal2(p, q) for p, q in zip(x.a.reshape(-1), y.a.reshape(-1)))
    if isinstance(x, Struct) or isinstance(y, Struct):
        if not (isinstance(x, Struct) and isinstance(y, Struct)) or set(x.f) != set(y.f):
            return False
        return all(_isequal2(x.f[k], y.f[k]) for k in x.f)
    if x.shape != y.shape:
        return False
    if sp.issparse(x) and sp.issparse(y):
        return (x != y).nnz == 0
    return np.array_equal(dense(x), dense(y))


@builtin("isequal")
def _isequal(I, a, n):
    return scalar(all(_isequal2(a[0], b) for b in a[1:]))


@builtin("strcmp")
def _strcmp(I, a, n):
    return scalar(isinstance(a[0], str) and isinstance(a[1], str) and a[0] == a[1])


@builtin("strcmpi")
def _strcmpi(I, a, n):
    return scalar(isinstance(a[0], str) and isinstance(a[1], str) and a[0].lower() == a[1].lower())


# ------------------------------------------------------------------------------------------------- elementwise
def _elementwise(name, f, sparse_ok=False):
    def g(I, a, n):
        v = a[0]
        if sp.issparse(v) and sparse_ok:
            out = v.copy().astype(np.float64)
            out.data = f(out.data)
            return out
        with np.errstate(all="ignore"):
            return f(as_float(v))
    TABLE[name] = g


for _n, _f, _s in (("abs", np.abs, True), ("exp", np.exp, False), ("sqrt", None, True), ("log", None, False),
                   ("sin", np.sin, True), ("cos", np.cos, False), ("tan", np.tan, True), ("sinh", np.sinh, True),
                   ("cosh", np.cosh, False), ("tanh", np.tanh, True), ("ceil", np.ceil, True),
                   ("floor", np.floor, True), ("round", None, True), ("fix", np.trunc, True), ("sign", np.sign, True),
                   ("real", np.real, True), ("imag", np.imag, True), ("conj", np.conj, True), ("isnan", np.isnan, False),
                   ("isinf", np.isinf, False), ("isfinite", np.isfinite, False), ("log2", None, False),
                   ("log10", None, False), ("gamma", None, False)):
    if _f is None:
        if _n == "sqrt":
            _f = lambda x: np.sqrt(x.astype(np.complex128)) if np.any(np.real(x) < 0) and not np.iscomplexobj(x) else np.sqrt(x)
        elif _n == "log":
            _f = lambda x: np.log(x.astype(np.complex128)) if np.any(np.real(x) < 0) and not np.iscomplexobj(x) else np.log(x)
        elif _n == "round":
            _f = lambda x: np.sign(x) * np.floor(np.abs(x) + 0.5)      # MATLAB rounds halves away from zero
        elif _n == "log2":
            _f = np.log2
        elif _n == "log10":
            _f = np.log10
        elif _n == "gamma":
            import scipy.special as _ss
            _f = _ss.gamma
    _elementwise(_n, _f, _s)


@builtin("mod")
def _mod(I, a, n):
    x, y = as_float(a[0]), as_float(a[1])
    with np.errstate(all="ignore"):
        r = np.mod(x, y)
    return np.where(y == 0, x, r)


@builtin("rem")
def _rem(I, a, n):
    x, y = as_float(a[0]), as_float(a[1])
    with np.errstate(all="ignore"):
        return np.where(y == 0, np.nan, np.fmod(x, y))


@builtin("factorial")
def _factorial(I, a, n):
    import scipy.special as ss
    return ss.factorial(as_float(a[0]))


# ------------------------------------------------------------------------------------------------- reductions
def _red_dim(v, a, k=1):
    """Default reduction dimension: the first non-singleton one."""
    if len(a) > k and not isinstance(a[k], str) and dense(a[k]).size == 1:
        return to_int(a[k]) - 1
    return 0 if v.shape[0] != 1 else 1


@builtin("sum")
def _sum(I, a, n):
    v = a[0]
    d = _red_dim(v, a)
    if sp.issparse(v):
        if v.shape == (0, 0):
            return scalar(0.0)
        return np.asarray(v.astype(np.float64).sum(axis=d)).reshape((1, -1) if d == 0 else (-1, 1))
    v = as_float(v)
    if v.shape == (0, 0):
        return scalar(0.0)
    return v.sum(axis=d, keepdims=True)


@builtin("prod")
def _prod(I, a, n):
    v = as_float(a[0])
    if v.shape == (0, 0):
        return scalar(1.0)
    return v.prod(axis=_red_dim(v, a), keepdims=True)


@builtin("cumsum")
def _cumsum(I, a, n):
    v = as_float(a[0])
    return np.cumsum(v, axis=_red_dim(v, a))


@builtin("mean")
def _mean(I, a, n):
    v = as_float(a[0])
    return v.mean(axis=_red_dim(v, a), keepdims=True)


@builtin("any")
def _any(I, a, n):
    v = dense(a[0]) != 0
    if v.size == 0:
        return scalar(False)
    return v.any(axis=_red_dim(v, a), keepdims=True)


@builtin("all")
def _all(I, a, n):
    v = dense(a[0]) != 0
    if v.size == 0:
        return scalar(True)
    return v.all(axis=_red_dim(v, a), keepdims=True)


def _minmax(a, nout, is_max):
    v = a[0]
    if len(a) >= 2 and not (is_num(a[1]) and a[1].shape == (0, 0)):
        x, y = as_float(v), as_float(a[1])
        # NaNs are ignored by MATLAB's two-argument min / max
        return (np.fmax if is_max else np.fmin)(x, y)
    v = as_float(v)
    if v.size == 0:
        return (EMPTY, EMPTY) if nout > 1 else EMPTY
    d = to_int(a[2]) - 1 if len(a) >= 3 else (0 if v.shape[0] != 1 else 1)
    key = np.real(v) if not np.iscomplexobj(v) else np.abs(v)
    if np.any(np.isnan(key)):
        key = np.where(np.isnan(key), -np.inf if is_max else np.inf, key)
    idx = (np.argmax if is_max else np.argmin)(key, axis=d)      # first occurrence, as in MATLAB
    val = np.take_along_axis(v, np.expand_dims(idx, d), axis=d)
    if nout > 1:
        return val, np.expand_dims(idx, d).astype(np.float64) + 1.0
    return val


@builtin("max")
def _max(I, a, n):
    return _minmax(a, n, True)


@builtin("min")
def _min(I, a, n):
    return _minmax(a, n, False)


@builtin("trace")
def _trace(I, a, n):
    v = a[0]
    if sp.issparse(v):
        return sp.csc_matrix(np.array([[v.diagonal().sum()]]))     # MATLAB: trace of a sparse matrix is sparse
    return scalar(np.trace(v))


@builtin("norm")
def _norm(I, a, n):
    v = a[0]
    p = a[1] if len(a) > 1 else scalar(2.0)
    vec = min(v.shape) == 1 if v.size else True
    if _isstr(p):
        if p == "fro":
            if sp.issparse(v):
                return scalar(float(np.sqrt((abs(v).power(2)).sum())))
            return scalar(float(np.linalg.norm(v, "fro")))
        if p.lower() == "inf":
            p = scalar(np.inf)
        else:
            raise MatlabError("norm: unknown norm type %s" % p)
    p = to_float(p)
    if v.size == 0:
        return scalar(0.0)
    if sp.issparse(v) and not vec:
        if p == 1:
            return scalar(float(abs(v).sum(axis=0).max()))
        if np.isinf(p):
            return scalar(float(abs(v).sum(axis=1).max()))
        raise MatlabError("norm(S,2) of a sparse matrix is not available; use normest")
    v = dense(v)
    if vec:
        x = v.reshape(-1)
        if np.isinf(p):
            return scalar(float(np.abs(x).max()))
        if p == 2:
            return scalar(float(np.linalg.norm(x)))
        if p == 1:
            return scalar(float(np.abs(x).sum()))
        return scalar(float((np.abs(x) ** p).sum() ** (1.0 / p)))
    if p == 2:
        return scalar(float(np.linalg.norm(v, 2)))
    if p == 1:
        return scalar(float(np.abs(v).sum(axis=0).max()))
    if np.isinf(p):
        return scalar(float(np.abs(v).sum(axis=1).max()))
    raise MatlabError("norm: unsupported p")


# ------------------------------------------------------------------------------------------------- searching / sets
@builtin("find")
def _find(I, a, n):
    v = a[0]
    lim = to_int(a[1]) if len(a) > 1 else None
    if sp.issparse(v):
        c = sp.coo_matrix(v)
        keep = c.data != 0
        r, cc, dd = c.row[keep], c.col[keep], c.data[keep]
        order = np.lexsort((r, cc))                  # column-major order
        r, cc, dd = r[order], cc[order], dd[order]
    else:
        v = dense(v)
        cc, r = np.nonzero(v.T)                      # column-major order
        dd = v[r, cc]
    if lim is not None:
        r, cc, dd = r[:lim], cc[:lim], dd[:lim]
    if n <= 1:
        k = (cc.astype(np.float64) * v.shape[0] + r + 1.0)
        if v.shape[0] == 1 and v.shape[1] != 1:
            return k.reshape(1, -1)
        return k.reshape(-1, 1)
    shape = (1, -1) if (v.shape[0] == 1 and v.shape[1] != 1) else (-1, 1)
    out = [(r + 1.0).reshape(shape), (cc + 1.0).reshape(shape)]
    if n >= 3:
        out.append(np.asarray(dd, dtype=np.float64).reshape(shape))
    return tuple(out)


@builtin("sort")
def _sort(I, a, n):
    v = as_float(a[0])
    desc = False
    dim = None
    for x in a[1:]:
        if _isstr(x):
            desc = x.lower() == "descend"
        else:
            dim = to_int(x) - 1
    if v.size == 0:
        return (v, v) if n > 1 else v
    if dim is None:
        dim = 0 if v.shape[0] != 1 else 1
    key = v if not np.iscomplexobj(v) else np.abs(v)
    # MATLAB's sort is stable; NaNs go last ascending / first descending
    if desc:
        idx = np.argsort(-key, axis=dim, kind="stable")
    else:
        idx = np.argsort(key, axis=dim, kind="stable")
    out = np.take_along_axis(v, idx, axis=dim)
    if n > 1:
        return out, idx.astype(np.float64) + 1.0
    return out


@builtin("unique")
def _unique(I, a, n):
    v = as_float(a[0])
    stable = any(_isstr(x, "stable") for x in a[1:])
    flat = v.reshape(-1, order="F")
    u, first = np.unique(flat, return_index=True)
    if stable:
        first = np.sort(first)
        u = flat[first]
    row = v.shape[0] == 1 and v.shape[1] != 1
    out = u.reshape(1, -1) if row else u.reshape(-1, 1)
    if n > 1:
        if not stable:
            # MATLAB returns the FIRST occurrence by default (R2013a onwards)
            pass
        ia = first.astype(np.float64) + 1.0
        return out, ia.reshape(-1, 1)
    return out


@builtin("setdiff")
def _setdiff(I, a, n):
    x, y = as_float(a[0]), as_float(a[1])
    u = np.setdiff1d(x.reshape(-1, order="F"), y.reshape(-1, order="F"))
    col = x.shape[1] == 1 and x.shape[0] != 1 and not (y.shape[0] == 1 and y.shape[1] > 1)
    return u.reshape(-1, 1) if col else u.reshape(1, -1)


@builtin("union")
def _union(I, a, n):
    x, y = as_float(a[0]), as_float(a[1])
    u = np.union1d(x.reshape(-1), y.reshape(-1))
    col = x.shape[1] == 1 and x.shape[0] > 1 and y.shape[1] == 1
    return u.reshape(-1, 1) if col else u.reshape(1, -1)


@builtin("intersect")
def _intersect(I, a, n):
    x, y = as_float(a[0]), as_float(a[1])
    u = np.intersect1d(x.reshape(-1), y.reshape(-1))
    col = x.shape[1] == 1 and x.shape[0] > 1 and y.shape[1] == 1
    return u.reshape(-1, 1) if col else u.reshape(1, -1)


@builtin("ismember")
def _ismember(I, a, n):
    x, y = as_float(a[0]), as_float(a[1])
    return np.isin(x, y.reshape(-1))


@builtin("ind2sub")
def _ind2sub(I, a, n):
    shp = dense(a[0]).reshape(-1)
    m = int(shp[0])
    k = as_float(a[1]) - 1.0
    if n <= 1:
        return k + 1.0
    return np.mod(k, m) + 1.0, np.floor(k / m) + 1.0


@builtin("sub2ind")
def _sub2ind(I, a, n):
    shp = dense(a[0]).reshape(-1)
    m = int(shp[0])
    return (as_float(a[1]) - 1.0) + (as_float(a[2]) - 1.0) * m + 1.0


@builtin("fliplr")
def _fliplr(I, a, n):
    return dense(a[0])[:, ::-1]


@builtin("flipud")
def _flipud(I, a, n):
    return dense(a[0])[::-1, :]


@builtin("reshape")
def _reshape(I, a, n):
    v = dense(a[0])
    if len(a) == 2:
        d = dense(a[1]).reshape(-1)
        m, k = int(d[0]), int(d[1])
    else:
        total = v.size
        m = None if (is_num(a[1]) and a[1].shape == (0, 0)) else to_int(a[1])
        k = None if (is_num(a[2]) and a[2].shape == (0, 0)) else to_int(a[2])
        if m is None:
            m = total // k
        if k is None:
            k = total // m
    if m * k != v.size:
        raise MatlabError("To RESHAPE the number of elements must not change.")
    return v.reshape((m, k), order="F")


@builtin("repmat")
def _repmat(I, a, n):
    m, k = _dims(a[1:])
    return np.tile(dense(a[0]), (m, k))


@builtin("kron")
def _kron(I, a, n):
    return np.kron(dense(a[0]), dense(a[1]))


@builtin("linspace")
def _linspace(I, a, n):
    return np.linspace(to_float(a[0]), to_float(a[1]), to_int(a[2]) if len(a) > 2 else 100).reshape(1, -1)


# ------------------------------------------------------------------------------------------------- linear algebra
@builtin("qr")
def _qr(I, a, n):
    A = dense(a[0])
    econ = len(a) > 1
    A = as_float(A)
    if A.shape[1] == 0:
        q, r = np.zeros((A.shape[0], 0)), np.zeros((0, 0))
    elif econ or A.shape[0] <= A.shape[1]:
        # LAPACK dgeqrf + dorgqr, the routines behind MATLAB's dense qr; no sign normalisation on either side
        q, r = sla.qr(A, mode="economic")
    else:
        q, r = sla.qr(A, mode="full")
    if n <= 1:
        return np.triu(r) if not econ else r
    return q, r


@builtin("eig")
def _eig(I, a, n):
    A = as_float(dense(a[0]))
    vector = any(_isstr(x, "vector") for x in a[1:])
    if A.shape[0] != A.shape[1]:
        raise MatlabError("Input matrix must be square.")
    if A.size == 0:
        return (EMPTY, EMPTY) if n > 1 else EMPTY
    herm = np.array_equal(A, A.conj().T)
    if n <= 1:
        if herm:
            return np.linalg.eigvalsh(A).reshape(-1, 1)          # ascending, like MATLAB for Hermitian input
        return np.linalg.eigvals(A).reshape(-1, 1)
    if herm:
        w, V = np.linalg.eigh(A)
    else:
        w, V = np.linalg.eig(A)
    return V, (w.reshape(-1, 1) if vector else np.diag(w))


@builtin("expm")
def _expm(I, a, n):
    A = as_float(dense(a[0]))
    if A.size == 0:
        return A
    return sla.expm(A)


@builtin("logm")
def _logm(I, a, n):
    return sla.logm(as_float(dense(a[0])))


@builtin("sqrtm")
def _sqrtm(I, a, n):
    return sla.sqrtm(as_float(dense(a[0])))


@builtin("funm")
def _funm(I, a, n):
    A = as_float(dense(a[0]))
    f = a[1]
    name = f.name if isinstance(f, FH) else None
    if name in ("sin", "cos", "sinh", "cosh", "exp", "log"):
        g = {"sin": sla.sinm, "cos": sla.cosm, "sinh": sla.sinhm, "cosh": sla.coshm, "exp": sla.expm, "log": sla.logm}[name]
        return g(A)
    raise MatlabError("funm: only @sin/@cos/@sinh/@cosh/@exp/@log are available")


@builtin("inv")
def _inv(I, a, n):
    return np.linalg.inv(as_float(dense(a[0])))


@builtin("det")
def _det(I, a, n):
    return scalar(float(np.linalg.det(as_float(dense(a[0])))))


@builtin("rank")
def _rank(I, a, n):
    return scalar(float(np.linalg.matrix_rank(as_float(dense(a[0])))))


@builtin("svd")
def _svd(I, a, n):
    A = as_float(dense(a[0]))
    if n <= 1:
        return np.linalg.svd(A, compute_uv=False).reshape(-1, 1)
    U, s, Vh = np.linalg.svd(A, full_matrices=len(a) == 1)
    S = np.zeros((U.shape[1], Vh.shape[0]))
    S[:s.size, :s.size] = np.diag(s)
    return U, S, Vh.conj().T


@builtin("chol")
def _chol(I, a, n):
    return np.linalg.cholesky(as_float(dense(a[0]))).conj().T


@builtin("normest")
def _normest(I, a, n):
    """MathWorks normest.m: power iteration on S'*S started from the column abs-sums, relative change <= tol."""
    S = a[0]
    tol = to_float(a[1]) if len(a) > 1 else 1e-6
    S = sp.csc_matrix(S) if not sp.issparse(S) else S
    x = np.asarray(abs(S).sum(axis=0)).reshape(-1, 1).astype(np.float64)
    cnt = 0
    e = float(np.linalg.norm(x))
    if e == 0:
        return (scalar(e), scalar(float(cnt)))[:max(n, 1)] if n > 1 else scalar(e)
    x = x / e
    e0 = 0.0
    St = S.T.tocsc()
    while abs(e - e0) > tol * e:
        e0 = e
        Sx = S @ x
        if not np.any(Sx):
            Sx = np.random.default_rng(0).random(Sx.shape)
        x = St @ Sx
        normx = float(np.linalg.norm(x))
        e = normx / float(np.linalg.norm(Sx))
        x = x / normx
        cnt += 1
        if cnt > 100:
            I.warnings.append("normest: not converged")
            break
    if n > 1:
        return scalar(e), scalar(float(cnt))
    return scalar(e)


@builtin("normest1")
def _normest1(I, a, n):
    """MathWorks normest1.m for a function-handle operand (Higham & Tisseur 2000, Algorithm 2.4) restricted to the
    t = 1 column that functions/normAm.m:25 asks for, where it reduces to Hager's / Higham's estimator.  Returns
    [est, v, w, it] with it = [iterations, products]."""
    afun = a[0]
    t = to_int(a[1]) if len(a) > 1 else 2
    if not isinstance(afun, FH):
        A = a[0]
        nn = A.shape[0]
        mul = lambda X: ops.binop("*", A, X)
        mult = lambda X: ops.binop("*", ops.transpose(A, True), X)
    else:
        nn = to_int(I.call_handle(afun, ["dim", EMPTY], 1)[0])
        mul = lambda X: dense(I.call_handle(afun, ["notransp", X], 1)[0])
        mult = lambda X: dense(I.call_handle(afun, ["transp", X], 1)[0])
        I.call_handle(afun, ["real", EMPTY], 1)
    if t != 1:
        raise MatlabError("normest1: only t = 1 is implemented (the reference calls normest1(@afun_power, 1))")
    X = np.ones((nn, 1)) / nn
    itmax = 5
    it = 0
    nmv = 0
    est_old = 0.0
    est = 0.0
    ind_hist = []
    est_j = 0
    w = None
    S = np.zeros((nn, 1))
    while True:
        it += 1
        Y = mul(X)
        nmv += 1
        est = float(np.abs(Y).sum())
        if est > est_old or it == 2:
            if it >= 2:
                est_j = ind_hist[-1] if ind_hist else 0
            w = Y
        if it >= 2 and est <= est_old:
            est = est_old
            break
        est_old = est
        if it > itmax:
            it = itmax
            break
        S_old = S
        S = np.sign(Y)
        S[S == 0] = 1.0
        if it >= 2 and float(abs((S_old.T @ S)[0, 0])) == nn:
            break                                           # parallel sign vectors: converged
        Z = mult(S)
        nmv += 1
        h = np.abs(Z).reshape(-1)
        j = int(np.argsort(h, kind="stable")[-1])       # ties to the LAST index (reversed ascending sort)
        if it >= 2 and h.max() == h[est_j]:
            break
        if j in ind_hist:
            break                                           # every candidate column has been visited already
        ind_hist.append(j)
        X = np.zeros((nn, 1))
        X[j, 0] = 1.0
    v = np.zeros((nn, 1))
    v[est_j, 0] = 1.0
    return scalar(est), v, (w if w is not None else np.zeros((nn, 1))), np.array([[float(it), float(nmv)]])


# ------------------------------------------------------------------------------------------------- random
@builtin("randn")
def _randn(I, a, n):
    raise MatlabError("randn is not available: parity runs replay committed probes (put a randn.m shim on the path)")


@builtin("rand")
def _rand(I, a, n):
    raise MatlabError("rand is not available in parity runs")


# ------------------------------------------------------------------------------------------------- workspace / control
@builtin("exist")
def _exist(I, a, n):
    name = a[0]
    kind = a[1] if len(a) > 1 else None
    fr = I.frames[-1]
    if kind in (None, "var") and I.has_var(fr, name):
        return scalar(1.0)
    if kind == "var":
        return scalar(0.0)
    if kind in (None, "file", "builtin") and (I.find_function(name, fr) is not None):
        return scalar(2.0)
    if kind in (None, "builtin") and name in I.builtins:
        return scalar(5.0)
    if kind in (None, "file", "dir") and os.path.exists(name):
        return scalar(7.0 if os.path.isdir(name) else 2.0)
    return scalar(0.0)


@builtin("clear")
def _clear(I, a, n):
    fr = I.frames[-1]
    if not a:
        fr.ws.clear()
    for name in a:
        fr.ws.pop(name, None)
    return []


@builtin("error")
def _error(I, a, n):
    if not a:
        raise MatlabError("error")
    msg = a[0]
    if len(a) > 1:
        import re
        if re.fullmatch(r"\w+(:\w+)+", a[0]):          # error('pkg:id', 'format', ...)
            msg = _sprintf(a[1], a[2:])
        else:
            msg = _sprintf(a[0], a[1:])
    raise MatlabError(msg)


@builtin("warning")
def _warning(I, a, n):
    if a and a[0] in ("off", "on"):
        return []
    msg = _sprintf(a[0], a[1:]) if len(a) > 1 else (a[0] if a else "")
    I.warnings.append(msg)
    return []


@builtin("keyboard", "pause", "drawnow", "tic", "format", "hold", "close", "clc", "figure")
def _noop(I, a, n):
    return []


@builtin("toc")
def _toc(I, a, n):
    return scalar(0.0)


@builtin("disp", "display")
def _disp(I, a, n):
    I.stdout.write(_to_text(a[0]) + "\n")
    return []


def _to_text(v):
    if isinstance(v, str):
        return v
    if is_num(v):
        return np.array2string(dense(v))
    return repr(v)


def _sprintf(fmt, args):
    """C-style formatting with MATLAB's cycling of the format over the flattened arguments."""
    import re
    flat = []
    for v in args:
        if isinstance(v, str):
            flat.append(v)
        else:
            d = dense(v)
            if d.dtype == np.bool_:
                d = d.astype(np.float64)
            flat.extend(list(d.reshape(-1, order="F")))
    fmt = fmt.replace("\\n", "\n").replace("\\t", "\t").replace("\\\\", "\\")
    spec = re.compile(r"%(?:%|[-+ 0#]*\d*(?:\.\d+)?[diouxXeEfgGcs])")
    specs = [m for m in spec.finditer(fmt) if m.group() != "%%"]
    if not specs:
        return fmt.replace("%%", "%")
    out = []
    pos = 0
    first = True
    while first or pos < len(flat):
        first = False
        last = 0
        piece = []
        stop = False
        for m in spec.finditer(fmt):
            piece.append(fmt[last:m.start()])
            last = m.end()
            s = m.group()
            if s == "%%":
                piece.append("%")
                continue
            if pos >= len(flat):
                stop = True
                break
            v = flat[pos]
            pos += 1
            conv = s[-1]
            if conv in "di":
                if isinstance(v, str):
                    piece.append(v)
                elif float(np.real(v)) == int(np.real(v)):
                    piece.append((s[:-1] + "d") % int(np.real(v)))
                else:
                    piece.append((s[:-1].split(".")[0] + "e") % float(np.real(v)))
            elif conv in "eEfgG":
                piece.append(s % float(np.real(v)) if not isinstance(v, str) else v)
            elif conv == "s":
                piece.append(s % (v if isinstance(v, str) else ("%g" % float(np.real(v)))))
            elif conv == "c":
                piece.append(v if isinstance(v, str) else chr(int(np.real(v))))
            else:
                piece.append(s % int(np.real(v)))
        if not stop:
            piece.append(fmt[last:])
        out.append("".join(piece))
        if not specs:
            break
    return "".join(out)


@builtin("sprintf")
def _sprintf_b(I, a, n):
    return _sprintf(a[0], a[1:])


@builtin("num2str")
def _num2str(I, a, n):
    if isinstance(a[0], str):
        return a[0]
    if len(a) > 1 and isinstance(a[1], str):
        return _sprintf(a[1], [a[0]])
    x = to_float(a[0])
    return "%d" % int(x) if x == int(x) else "%.5g" % x


@builtin("fprintf")
def _fprintf(I, a, n):
    if a and not isinstance(a[0], str):
        fid = to_int(a[0])
        text = _sprintf(a[1], a[2:])
        if fid == 1:
            I.stdout.write(text)
        elif fid == 2:
            pass
        else:
            I.files[fid].write(text)
        return []
    I.stdout.write(_sprintf(a[0], a[1:]))
    return []


@builtin("fopen")
def _fopen(I, a, n):
    mode = a[1] if len(a) > 1 else "r"
    try:
        fh = open(a[0], mode.replace("t", ""))
    except OSError:
        return scalar(-1.0)
    fid = I.next_fid
    I.next_fid += 1
    I.files[fid] = fh
    return scalar(float(fid))


@builtin("fclose")
def _fclose(I, a, n):
    fid = to_int(a[0])
    fh = I.files.pop(fid, None)
    if fh is not None:
        fh.close()
    return scalar(0.0)


@builtin("fullfile")
def _fullfile(I, a, n):
    return os.path.join(*[x for x in a if x != ""])


@builtin("fileparts")
def _fileparts(I, a, n):
    d, base = os.path.split(a[0])
    name, ext = os.path.splitext(base)
    return d, name, ext


@builtin("mfilename")
def _mfilename(I, a, n):
    fr = I.frames[-1]
    f = fr.fn.fname if fr.fn is not None else I.script_file
    if f is None:
        return ""
    full = os.path.splitext(os.path.abspath(f))[0]
    if a and a[0] == "fullpath":
        return full
    return os.path.basename(full)


@builtin("pwd")
def _pwd(I, a, n):
    return os.getcwd()


@builtin("addpath")
def _addpath(I, a, n):
    for d in a:
        if isinstance(d, str) and not d.startswith("-"):
            for part in d.split(os.pathsep):
                I.addpath(part)
    return []


@builtin("rmpath")
def _rmpath(I, a, n):
    for d in a:
        I.rmpath(d)
    return []


@builtin("tempname")
def _tempname(I, a, n):
    return os.path.join(tempfile.gettempdir(), "mlab_" + next(tempfile._get_candidate_names()))


@builtin("mkdir")
def _mkdir(I, a, n):
    os.makedirs(os.path.join(*a), exist_ok=True)
    return scalar(1.0)


@builtin("fieldnames")
def _fieldnames(I, a, n):
    names = list(a[0].f)
    c = np.empty((len(names), 1), dtype=object)
    for k, nm in enumerate(names):
        c[k, 0] = nm
    return Cell(c)


@builtin("isfield")
def _isfield(I, a, n):
    return scalar(isinstance(a[0], Struct) and a[1] in a[0].f)


@builtin("load")
def _load(I, a, n):
    path = I.which_data_file(a[0])
    raw = sio.loadmat(path)
    vals = {}
    for k, v in raw.items():
        if k.startswith("__"):
            continue
        if sp.issparse(v):
            vals[k] = sp.csc_matrix(v).astype(np.float64)
        elif isinstance(v, np.ndarray) and v.dtype.kind in "fiub":
            vals[k] = np.atleast_2d(v).astype(np.float64 if v.dtype.kind != "b" else np.bool_)
        elif isinstance(v, np.ndarray) and v.dtype.kind == "c":
            vals[k] = np.atleast_2d(v).astype(np.complex128)
        elif isinstance(v, np.ndarray) and v.dtype.kind == "U":
            vals[k] = str(v.reshape(-1)[0]) if v.size else ""
        else:
            continue                                   # struct / cell variables are not needed on this path
    want = [x for x in a[1:] if isinstance(x, str)]
    if want:
        vals = {k: v for k, v in vals.items() if k in want}
    if n >= 1:
        return Struct(vals)
    fr = I.frames[-1]
    for k, v in vals.items():
        I.setvar(fr, k, v)
    return []


@builtin("func2str")
def _func2str(I, a, n):
    return a[0].name if a[0].name else "@(...)"


@builtin("str2func")
def _str2func(I, a, n):
    return FH(a[0].lstrip("@"), frame=I.frames[-1] if I.frames else None)


@builtin("feval")
def _feval(I, a, n):
    f = a[0]
    if isinstance(f, str):
        return tuple(I.call_named(f, list(a[1:]), max(n, 1), I.frames[-1]))
    return tuple(I.call_handle(f, list(a[1:]), max(n, 1)))


@builtin("cellfun")
def _cellfun(I, a, n):
    f, c = a[0], a[1]
    vals = [I.call_handle(f, [x], 1)[0] if isinstance(f, FH) else I.call_named(f, [x], 1, I.frames[-1])[0]
            for x in c.a.reshape(-1, order="F")]
    return np.array([to_float(v) for v in vals]).reshape(c.a.shape, order="F")


@builtin("balance")
def _balance(I, a, n):
    raise MatlabError("balance is not available (bal = true is never used on this path)")


@builtin("eigs")
def _eigs(I, a, n):
    """eigs(A, k) / eigs(Afun, n, k): the k = 1 eigenpair of largest magnitude (ARPACK, as in MATLAB).  The sign of
    the vector is arbitrary on both sides; compute_centrality.m takes abs()."""
    import scipy.sparse.linalg as spla
    if isinstance(a[0], FH):
        nn = to_int(a[1])
        k = to_int(a[2]) if len(a) > 2 else 6
        op = spla.LinearOperator((nn, nn), matvec=lambda x: dense(I.call_handle(a[0], [np.asarray(x, dtype=np.float64).reshape(-1, 1)], 1)[0]).reshape(-1),
                                 dtype=np.float64)
        w, V = spla.eigs(op, k=k, which="LM", v0=np.ones(nn), tol=0)
        if np.all(np.abs(w.imag) <= 1e-14 * np.abs(w)) and np.all(np.abs(V.imag) <= 1e-12):
            w, V = w.real, V.real
    else:
        A = a[0]
        k = to_int(a[1]) if len(a) > 1 else 6
        A = sp.csc_matrix(A).astype(np.float64)
        sym = (A != A.T).nnz == 0
        if sym:
            w, V = spla.eigsh(A, k=k, which="LM", v0=np.ones(A.shape[0]), tol=0)
        else:
            w, V = spla.eigs(A, k=k, which="LM", v0=np.ones(A.shape[0]), tol=0)
    order = np.argsort(-np.abs(w), kind="stable")
    w, V = w[order], V[:, order]
    if n <= 1:
        return w.reshape(-1, 1)
    return V, np.diag(w)
