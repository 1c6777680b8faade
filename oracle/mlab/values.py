"""Value model of the MATLAB-subset interpreter (see oracle/mlab/__init__.py).  TEST INFRASTRUCTURE ONLY.

Everything numeric is a 2-D float64 / complex128 / bool ``numpy.ndarray`` (scalars are 1 x 1) or a
``scipy.sparse.csc_matrix``; char arrays are Python ``str``; cell arrays, scalar structs and function handles have the
small classes below.  Values are never mutated in place once they are bound to a name: every indexed assignment copies
(MATLAB's value semantics), which is what makes the shallow workspace snapshot of an anonymous function correct."""
import numpy as np
import scipy.sparse as sp


class MatlabError(Exception):
    """error('...') raised by interpreted code or by a built-in, with MATLAB's message."""


class Colon:
    def __repr__(self):
        return ":"


COLON = Colon()


class Cell:
    __slots__ = ("a",)

    def __init__(self, a=None):
        if a is None:
            a = np.empty((0, 0), dtype=object)
        self.a = a

    @staticmethod
    def row(items):
        a = np.empty((1, len(items)), dtype=object)
        for k, v in enumerate(items):
            a[0, k] = v
        return Cell(a) if items else Cell()

    @property
    def shape(self):
        return self.a.shape

    def __repr__(self):
        return "Cell%s" % (self.a.shape,)


class Struct:
    __slots__ = ("f",)

    def __init__(self, f=None):
        self.f = dict(f) if f else {}

    def __repr__(self):
        return "Struct(%s)" % ", ".join(self.f)


class FH:
    """Function handle: named (@exp) or anonymous (@(x) ...) with its captured workspace."""
    __slots__ = ("name", "params", "body", "env", "frame")

    def __init__(self, name=None, params=None, body=None, env=None, frame=None):
        self.name, self.params, self.body, self.env, self.frame = name, params, body, env, frame

    def __repr__(self):
        return "@%s" % self.name if self.name else "@(%s)<anon>" % ",".join(self.params)


EMPTY = np.zeros((0, 0))


def scalar(x):
    return np.array([[x]], dtype=np.complex128 if isinstance(x, complex) else
                    (np.bool_ if isinstance(x, (bool, np.bool_)) else np.float64))


def norm_val(v):
    """Bring whatever NumPy / SciPy returned into the value model."""
    if isinstance(v, np.matrix):
        return np.asarray(v)
    if isinstance(v, np.ndarray):
        if v.ndim == 2:
            return v
        if v.ndim == 0:
            return v.reshape(1, 1)
        if v.ndim == 1:
            return v.reshape(1, -1)
        raise MatlabError("N-d arrays are not supported")
    if sp.issparse(v):
        return v if isinstance(v, sp.csc_matrix) else sp.csc_matrix(v)
    if isinstance(v, (bool, np.bool_)):
        return scalar(bool(v))
    if isinstance(v, (int, float, np.integer, np.floating)):
        return scalar(float(v))
    if isinstance(v, (complex, np.complexfloating)):
        return scalar(complex(v))
    return v


def is_num(v):
    return isinstance(v, np.ndarray) or sp.issparse(v)


def is_scalar(v):
    return is_num(v) and v.shape == (1, 1)


def dense(v):
    if sp.issparse(v):
        return v.toarray()
    if isinstance(v, str):
        return np.array([[float(ord(c)) for c in v]]) if v else np.zeros((0, 0))
    return v


def as_float(v):
    v = dense(v)
    if v.dtype == np.bool_:
        return v.astype(np.float64)
    return v


def to_float(v):
    """Scalar value -> Python float."""
    v = dense(norm_val(v))
    if v.size != 1:
        raise MatlabError("scalar expected, got %dx%d" % v.shape)
    x = v.reshape(-1)[0]
    if np.iscomplexobj(x):
        if x.imag != 0:
            raise MatlabError("real scalar expected")
        x = x.real
    return float(x)


def to_int(v):
    x = to_float(v)
    if x != int(x):
        raise MatlabError("integer expected, got %r" % x)
    return int(x)


def truth(v):
    """`if v`: true when v is non-empty and all its elements are non-zero."""
    if isinstance(v, str):
        return len(v) > 0
    if sp.issparse(v):
        return v.shape[0] * v.shape[1] > 0 and v.nnz == v.shape[0] * v.shape[1] and bool(np.all(v.data != 0))
    if isinstance(v, np.ndarray):
        return v.size > 0 and bool(np.all(v != 0))
    raise MatlabError("condition must be numeric or logical")


# --------------------------------------------------------------------------------------------- indexing
def _index_vector(ix, extent, what="Index"):
    """One subscript -> 0-based integer array (and the shape of the subscript)."""
    if ix is COLON:
        return np.arange(extent), None
    ix = dense(ix) if not isinstance(ix, str) else dense(ix)
    if ix.dtype == np.bool_:
        flat = ix.reshape(-1, order="F")
        if flat.size > extent and np.any(flat[extent:]):
            raise MatlabError("%s exceeds matrix dimensions (logical mask)" % what)
        k = np.flatnonzero(flat)
        shape = (1, k.size) if ix.shape[0] == 1 else (k.size, 1)
        return k, shape
    if np.iscomplexobj(ix):
        raise MatlabError("Subscript indices must be real positive integers")
    k = ix.reshape(-1, order="F")
    ki = k.astype(np.int64)
    if np.any(ki != k) or np.any(ki < 1):
        raise MatlabError("Subscript indices must either be real positive integers or logicals")
    return ki - 1, ix.shape


def index_get(obj, args):
    if isinstance(obj, Cell):
        sub = index_get_array(obj.a, args)
        return Cell(sub)
    if isinstance(obj, str):
        arr = np.array([list(obj)], dtype=object) if obj else np.empty((0, 0), dtype=object)
        sub = index_get_array(arr, args)
        return "".join(sub.reshape(-1, order="F"))
    if isinstance(obj, Struct):
        if all(np.all(dense(a) == 1) for a in args if a is not COLON):
            return obj
        raise MatlabError("Index exceeds the number of array elements (1)")
    if not is_num(obj):
        raise MatlabError("value of this type cannot be indexed with ()")
    return index_get_array(obj, args)


def index_get_array(obj, args):
    m, n = obj.shape
    sparse = sp.issparse(obj)
    if len(args) == 1:
        ix = args[0]
        if ix is COLON:
            if sparse:
                return sp.csc_matrix(obj.toarray().reshape(-1, 1, order="F"))
            return obj.reshape(-1, 1, order="F")
        k, shape = _index_vector(ix, m * n)
        if k.size and k.max() >= m * n:
            raise MatlabError("Index exceeds the number of array elements (%d)" % (m * n))
        # a vector subscript into a vector follows the SOURCE's orientation, everything else the subscript's shape
        # (a logical mask behaves as find(mask): a row for a row mask, a column otherwise)
        vec_src = (m == 1) != (n == 1)
        vec_ix = shape[0] == 1 or shape[1] == 1
        if vec_src and vec_ix:
            out_shape = (1, k.size) if m == 1 else (k.size, 1)
        else:
            out_shape = shape
        if sparse:
            r, c = k % m, k // m
            vals = np.asarray(obj[r, c]).reshape(-1) if k.size else np.zeros(0)
            return sp.csc_matrix(vals.reshape(out_shape, order="F"))
        vals = obj.reshape(-1, order="F")[k]
        return vals.reshape(out_shape, order="F")
    if len(args) == 2:
        r, _ = _index_vector(args[0], m)
        c, _ = _index_vector(args[1], n)
        if (r.size and r.max() >= m) or (c.size and c.max() >= n):
            raise MatlabError("Index exceeds matrix dimensions")
        if sparse:
            return sp.csc_matrix(obj[r, :][:, c])
        return obj[np.ix_(r, c)]
    # trailing singleton subscripts a(i,j,1)
    if all((a is not COLON and np.all(dense(a) == 1)) or a is COLON for a in args[2:]):
        return index_get_array(obj, args[:2])
    raise MatlabError("N-d indexing is not supported")


def _result_dtype(obj, rhs):
    if obj.dtype == object:
        return object
    if np.iscomplexobj(obj) or np.iscomplexobj(rhs):
        return np.complex128
    if obj.dtype == np.bool_ and rhs.dtype == np.bool_:
        return np.bool_
    return np.float64


def index_set(obj, args, rhs):
    """obj(args) = rhs  ->  the new value of obj (obj itself is never modified)."""
    if isinstance(obj, Cell):
        if isinstance(rhs, Cell):
            return Cell(index_set_array(obj.a, args, rhs.a, fill=None))
        raise MatlabError("Conversion to cell from double is not possible")
    if isinstance(obj, Struct):
        raise MatlabError("struct arrays are not supported")
    if obj is None:
        obj = np.empty((0, 0), dtype=object) if isinstance(rhs, Cell) else EMPTY
        if isinstance(rhs, Cell):
            return Cell(index_set_array(obj, args, rhs.a, fill=None))
    if isinstance(rhs, str):
        rhs = dense(rhs)
    if not is_num(rhs):
        raise MatlabError("cannot assign a value of this type into a numeric array")
    if sp.issparse(obj):
        return _sparse_set(obj, args, rhs)
    rhs = dense(rhs)
    if rhs.shape == (0, 0) and obj.dtype != object:
        return _delete(obj, args)
    return index_set_array(obj, args, rhs, fill=0)


def _delete(obj, args):
    m, n = obj.shape
    if len(args) == 1:
        k, _ = _index_vector(args[0], m * n)
        keep = np.ones(m * n, dtype=bool)
        keep[k] = False
        flat = obj.reshape(-1, order="F")[keep]
        return flat.reshape(1, -1) if m == 1 or not (n == 1) else flat.reshape(-1, 1)
    if len(args) == 2:
        if args[0] is COLON:
            c, _ = _index_vector(args[1], n)
            keep = np.ones(n, dtype=bool)
            keep[c] = False
            return obj[:, keep]
        if args[1] is COLON:
            r, _ = _index_vector(args[0], m)
            keep = np.ones(m, dtype=bool)
            keep[r] = False
            return obj[keep, :]
    raise MatlabError("A null assignment can have only one non-colon index")


def index_set_array(obj, args, rhs, fill=0):
    m, n = obj.shape
    dt = _result_dtype(obj, rhs)

    def grown(mm, nn):
        new = np.empty((mm, nn), dtype=dt)
        if dt == object:
            for i in range(mm):
                for j in range(nn):
                    new[i, j] = EMPTY
        else:
            new[...] = fill
        new[:m, :n] = obj
        return new

    if len(args) == 1:
        ix = args[0]
        if ix is COLON:
            k = np.arange(m * n)
        else:
            k, _ = _index_vector(ix, m * n)
        need = int(k.max()) + 1 if k.size else 0
        if need > m * n:
            if m == 0 and n == 0:
                new = grown(1, need)
            elif m == 1:
                new = grown(1, need)
            elif n == 1:
                new = grown(need, 1)
            else:
                raise MatlabError("Attempt to grow array along ambiguous dimension")
        else:
            new = obj.astype(dt, copy=True)
        flat = new.reshape(-1, order="F")         # a copy unless the array is a vector
        vals = rhs.reshape(-1, order="F")
        if dt == object:
            if vals.size not in (1, k.size):
                raise MatlabError("cell assignment: the left and right sides have a different number of elements")
            for q, kk in enumerate(k):
                flat[kk] = vals[0] if vals.size == 1 else vals[q]
        elif vals.size == 1:
            flat[k] = vals[0]
        elif vals.size == k.size:
            flat[k] = vals
        else:
            raise MatlabError("Unable to perform assignment because the left and right sides have a different number of elements")
        return flat.reshape(new.shape, order="F")
    if len(args) == 2:
        a0, a1 = args
        if a0 is COLON:
            r = np.arange(m if (m or n) else rhs.shape[0])
        else:
            r, _ = _index_vector(a0, m)
        if a1 is COLON:
            c = np.arange(n if (m or n) else rhs.shape[1])
        else:
            c, _ = _index_vector(a1, n)
        if a0 is COLON and m == 0 and rhs.size > 1:
            r = np.arange(rhs.shape[0] if rhs.shape[1] == c.size else rhs.size // max(c.size, 1))
        if a1 is COLON and n == 0 and rhs.size > 1:
            c = np.arange(rhs.shape[1] if rhs.shape[0] == r.size else rhs.size // max(r.size, 1))
        mm = max(m, int(r.max()) + 1 if r.size else 0)
        nn = max(n, int(c.max()) + 1 if c.size else 0)
        new = grown(mm, nn) if (mm, nn) != (m, n) else obj.astype(dt, copy=True)
        if r.size * c.size == 0 and rhs.size == 0:
            pass                                   # empty = empty, whatever the orientations
        elif dt == object:
            if rhs.size != 1 and rhs.shape != (r.size, c.size):
                raise MatlabError("cell assignment: size mismatch")
            for p_, i_ in enumerate(r):
                for q_, j_ in enumerate(c):
                    new[i_, j_] = rhs.reshape(-1)[0] if rhs.size == 1 else rhs[p_, q_]
        elif rhs.size == 1:
            new[np.ix_(r, c)] = rhs.reshape(-1)[0]
        elif rhs.shape == (r.size, c.size):
            new[np.ix_(r, c)] = rhs
        elif rhs.size == r.size * c.size and (min(rhs.shape) == 1) and (r.size == 1 or c.size == 1):
            new[np.ix_(r, c)] = rhs.reshape(r.size, c.size, order="F")
        else:
            raise MatlabError("Unable to perform assignment because the size of the left side is %d-by-%d and the size "
                              "of the right side is %d-by-%d" % (r.size, c.size, rhs.shape[0], rhs.shape[1]))
        return new
    raise MatlabError("N-d indexed assignment is not supported")


def _sparse_set(obj, args, rhs):
    import warnings
    m, n = obj.shape
    rhs = dense(rhs)
    if len(args) == 1:
        k, _ = _index_vector(args[0], m * n)
        if k.size and k.max() >= m * n:
            if n == 1:
                obj = sp.vstack([obj, sp.csc_matrix((int(k.max()) + 1 - m, 1))]).tocsc()
                m = obj.shape[0]
            elif m == 1:
                obj = sp.hstack([obj, sp.csc_matrix((1, int(k.max()) + 1 - n))]).tocsc()
                n = obj.shape[1]
            else:
                raise MatlabError("Attempt to grow sparse matrix along ambiguous dimension")
        r, c = k % m, k // m
        vals = rhs.reshape(-1, order="F")
    elif len(args) == 2:
        r0, _ = _index_vector(args[0], m)
        c0, _ = _index_vector(args[1], n)
        mm = max(m, int(r0.max()) + 1 if r0.size else 0)
        nn = max(n, int(c0.max()) + 1 if c0.size else 0)
        if (mm, nn) != (m, n):
            obj = sp.csc_matrix((obj.data, obj.indices, np.concatenate([obj.indptr, np.full(nn - n, obj.indptr[-1])])),
                                shape=(mm, nn))
            m, n = mm, nn
        c, r = np.meshgrid(c0, r0)
        r, c = r.reshape(-1, order="F"), c.reshape(-1, order="F")
        if rhs.size not in (1, r.size):
            raise MatlabError("sparse indexed assignment: size mismatch")
        vals = rhs.reshape(-1, order="F")
    else:
        raise MatlabError("N-d indexed assignment is not supported")
    new = obj.tolil(copy=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if vals.size == 1:
            for i, j in zip(r, c):
                new[i, j] = vals[0]
        else:
            for i, j, v in zip(r, c, vals):
                new[i, j] = v
    new = new.tocsc()
    new.eliminate_zeros()          # MATLAB drops entries that are set to zero
    return new


# --------------------------------------------------------------------------------------------- concatenation
def _wrap_cell(v):
    a = np.empty((1, 1), dtype=object)
    a[0, 0] = v
    return a


def concat(rows):
    """rows: list of lists of values -> the value of [r1; r2; ...]."""
    flat = [v for r in rows for v in r]
    if not flat:
        return EMPTY
    if any(isinstance(v, Cell) for v in flat):
        rws = []
        for r in rows:
            parts = [(v.a if isinstance(v, Cell) else _wrap_cell(v)) for v in r]
            parts = [p for p in parts if p.shape != (0, 0)]
            if parts:
                rws.append(np.concatenate(parts, axis=1))
        if not rws:
            return Cell()
        return Cell(np.concatenate(rws, axis=0))
    if any(isinstance(v, (Struct, FH)) for v in flat):
        if len(flat) == 1:
            return flat[0]
        raise MatlabError("arrays of structs / function handles are not supported")
    if all(isinstance(v, str) for v in flat):
        if len(rows) == 1:
            return "".join(rows[0])
        raise MatlabError("multi-row char arrays are not supported")
    any_sparse = any(sp.issparse(v) for v in flat)
    out_rows = []
    for r in rows:
        parts = [(dense(v) if isinstance(v, str) else v) for v in r]
        parts = [p for p in parts if p.shape != (0, 0)]
        if not parts:
            continue
        if len({p.shape[0] for p in parts}) != 1:
            raise MatlabError("Dimensions of arrays being concatenated are not consistent.")
        if any_sparse:
            out_rows.append(sp.hstack([sp.csc_matrix(p) for p in parts]).tocsc())
        else:
            out_rows.append(np.concatenate([as_float(p) if not all(q.dtype == np.bool_ for q in parts) else p
                                            for p in parts], axis=1) if len(parts) > 1 else parts[0])
    if not out_rows:
        return EMPTY
    if len(out_rows) == 1:
        return out_rows[0]
    if len({p.shape[1] for p in out_rows}) != 1:
        raise MatlabError("Dimensions of arrays being concatenated are not consistent.")
    if any_sparse:
        return sp.vstack([sp.csc_matrix(p) for p in out_rows]).tocsc()
    if not all(p.dtype == np.bool_ for p in out_rows):
        out_rows = [as_float(p) for p in out_rows]
    return np.concatenate(out_rows, axis=0)
