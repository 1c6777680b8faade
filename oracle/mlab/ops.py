"""Operators of the MATLAB-subset interpreter (see oracle/mlab/__init__.py).  TEST INFRASTRUCTURE ONLY.

Sparse / full result classes follow MATLAB: sparse (+,-) scalar or full -> full; sparse * scalar, sparse .* full,
sparse / scalar -> sparse; sparse * full -> full."""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .values import MatlabError, norm_val, is_num, dense, as_float, Cell


def _num(v, op):
    if isinstance(v, str):
        return dense(v)
    if not is_num(v):
        raise MatlabError("Undefined operator '%s' for input arguments of this type." % op)
    return v


def _arith(v):
    """bool -> float64 for arithmetic; sparse stays sparse."""
    if sp.issparse(v):
        return v if v.dtype != np.bool_ else v.astype(np.float64)
    return as_float(v)


def _is_scalar(v):
    return v.shape == (1, 1)


def _sval(v):
    return (v.toarray() if sp.issparse(v) else v).reshape(-1)[0]


def _check_bcast(a, b, op):
    for x, y in zip(a.shape, b.shape):
        if x != y and x != 1 and y != 1:
            raise MatlabError("Matrix dimensions must agree (%s: %dx%d vs %dx%d)." % (op, a.shape[0], a.shape[1],
                                                                                    b.shape[0], b.shape[1]))


def binop(op, a, b):
    a, b = _num(a, op), _num(b, op)
    if op in ("==", "~=", "<", "<=", ">", ">="):
        a, b = dense(a), dense(b)
        _check_bcast(a, b, op)
        if np.iscomplexobj(a) or np.iscomplexobj(b):
            if op not in ("==", "~="):
                a, b = a.real, b.real
        f = {"==": np.equal, "~=": np.not_equal, "<": np.less, "<=": np.less_equal, ">": np.greater,
             ">=": np.greater_equal}[op]
        return f(a, b)
    if op in ("&", "|"):
        a, b = dense(a) != 0, dense(b) != 0
        _check_bcast(a, b, op)
        return (a & b) if op == "&" else (a | b)
    a, b = _arith(a), _arith(b)
    sa, sb = sp.issparse(a), sp.issparse(b)
    if op in ("+", "-"):
        if sa and sb:
            if a.shape != b.shape:
                if _is_scalar(a) or _is_scalar(b):
                    a, b = a.toarray(), b.toarray()
                    return a + b if op == "+" else a - b
                raise MatlabError("Matrix dimensions must agree.")
            return norm_val(a + b if op == "+" else a - b)
        if sa or sb:
            a, b = dense(a), dense(b)
        _check_bcast(a, b, op)
        return a + b if op == "+" else a - b
    if op == ".*":
        if sa or sb:
            if _is_scalar(a) or _is_scalar(b):
                sc, m = (a, b) if _is_scalar(a) else (b, a)
                if sp.issparse(m):
                    return norm_val(sp.csc_matrix(m * _sval(sc)))
                return dense(a) * dense(b)
            if a.shape == b.shape:
                A = a if sa else sp.csc_matrix(b)
                other = b if sa else a
                return norm_val(sp.csc_matrix(A.multiply(other)))
            a, b = dense(a), dense(b)
            _check_bcast(a, b, op)
            return sp.csc_matrix(a * b)
        _check_bcast(a, b, op)
        return a * b
    if op == "./":
        if sa and _is_scalar(b) and not sb:
            return norm_val(a / _sval(b))
        a, b = dense(a), dense(b)
        _check_bcast(a, b, op)
        with np.errstate(divide="ignore", invalid="ignore"):
            return a / b
    if op == ".\\":
        a, b = dense(a), dense(b)
        _check_bcast(a, b, op)
        with np.errstate(divide="ignore", invalid="ignore"):
            return b / a
    if op == ".^":
        a, b = dense(a), dense(b)
        _check_bcast(a, b, op)
        return _power(a, b)
    if op == "*":
        if _is_scalar(a) or _is_scalar(b):
            if sa and not _is_scalar(a):
                return norm_val(sp.csc_matrix(a * _sval(b)))
            if sb and not _is_scalar(b):
                return norm_val(sp.csc_matrix(b * _sval(a)))
            return dense(a) * dense(b)
        if a.shape[1] != b.shape[0]:
            raise MatlabError("Incorrect dimensions for matrix multiplication (%dx%d * %dx%d)."
                              % (a.shape[0], a.shape[1], b.shape[0], b.shape[1]))
        if sa and sb:
            return norm_val(a @ b)
        if sa:
            return norm_val(np.asarray(a @ b))
        if sb:
            return norm_val(np.asarray((b.T @ a.T).T))
        return a @ b
    if op == "/":
        if _is_scalar(b):
            if sa:
                return norm_val(a / _sval(b))
            with np.errstate(divide="ignore", invalid="ignore"):
                return a / dense(b)
        # X / K = (K' \ X')'
        r = binop("\\", transpose(b, True), transpose(a, True))
        return transpose(r, True)
    if op == "\\":
        if _is_scalar(a):
            return binop("/", b, a) if not _is_scalar(b) else dense(b) / dense(a)
        if a.shape[0] != b.shape[0]:
            raise MatlabError("Matrix dimensions must agree.")
        if sa:
            if a.shape[0] != a.shape[1]:
                raise MatlabError("sparse least squares is not supported")
            x = spla.spsolve(sp.csc_matrix(a), dense(b))
            return norm_val(np.asarray(x).reshape(b.shape[0] if False else a.shape[1], -1))
        b = dense(b)
        if a.shape[0] == a.shape[1]:
            # MATLAB's mldivide: triangular -> substitution, otherwise LU with partial pivoting
            if np.array_equal(a, np.triu(a)):
                return sla.solve_triangular(a, b, lower=False)
            if np.array_equal(a, np.tril(a)):
                return sla.solve_triangular(a, b, lower=True)
            return np.linalg.solve(a, b)
        return np.linalg.lstsq(a, b, rcond=None)[0]
    if op == "^":
        if _is_scalar(a) and _is_scalar(b):
            return _power(dense(a), dense(b))
        if _is_scalar(b):
            p = float(_sval(b).real)
            if p == int(p) and p >= 0:
                if sa:
                    out = sp.identity(a.shape[0], format="csc")
                    for _ in range(int(p)):
                        out = out @ a
                    return norm_val(out)
                return np.linalg.matrix_power(a, int(p))
            return sla.fractional_matrix_power(dense(a), p)
        raise MatlabError("matrix exponent of a matrix power is not supported")
    raise MatlabError("unknown operator %s" % op)


def _power(a, b):
    if (not np.iscomplexobj(a)) and (not np.iscomplexobj(b)) and np.any((a < 0) & (b != np.floor(b))):
        a = a.astype(np.complex128)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return np.power(a, b)


def unop(op, a):
    a = _num(a, op)
    if op == "-":
        return norm_val(-_arith(a))
    if op == "+":
        return _arith(a)
    if op == "~":
        return dense(a) == 0
    raise MatlabError("unknown unary operator %s" % op)


def transpose(a, conj):
    if isinstance(a, str):
        if len(a) <= 1:
            return a
        raise MatlabError("transpose of a char row is not supported")
    if isinstance(a, Cell):
        return Cell(a.a.T.copy())
    if not is_num(a):
        raise MatlabError("transpose of a value of this type")
    if sp.issparse(a):
        t = a.T
        if conj and np.iscomplexobj(t):
            t = t.conj()
        return sp.csc_matrix(t)
    t = a.T
    if conj and np.iscomplexobj(t):
        t = t.conj()
    return t


def make_range(a, step, b):
    a = float(dense(a).reshape(-1)[0].real) if dense(a).size else None
    b = float(dense(b).reshape(-1)[0].real) if dense(b).size else None
    s = 1.0 if step is None else (float(dense(step).reshape(-1)[0].real) if dense(step).size else None)
    if a is None or b is None or s is None or s == 0 or (s > 0 and a > b) or (s < 0 and a < b):
        return np.zeros((1, 0))
    n = int(np.floor((b - a) / s * (1 + 2 * np.finfo(float).eps) + 1e-10)) + 1
    if a == int(a) and s == int(s):
        n = int((b - a) // s) + 1
    return (a + s * np.arange(n, dtype=np.float64)).reshape(1, -1)
