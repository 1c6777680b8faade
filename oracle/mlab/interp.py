"""Evaluator of the MATLAB-subset interpreter (see oracle/mlab/__init__.py).  TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np
import scipy.sparse as sp

from .parser import parse_source, Function
from .values import (MatlabError, COLON, Cell, Struct, FH, EMPTY, scalar, norm_val, is_num, dense, truth, index_get,
                     index_set, concat)


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    pass


class Frame:
    def __init__(self, fn=None, nargin=0, nargout=0, parent=None):
        self.ws = {}
        self.fn = fn
        self.nargin = nargin
        self.nargout = nargout
        self.parent = parent          # workspace frame of the enclosing function (nested functions only)
        self.globals = set()

    def lookup(self, name):
        f = self
        while f is not None:
            if name in f.ws:
                return f
            f = f.parent
        return None


class Interpreter:
    """Runs MATLAB source files.  `path` is the list of directories searched for `<name>.m`."""

    def __init__(self, path=(), stdout=None):
        from . import builtins as B
        self.path = [os.path.abspath(p) for p in path]
        self.builtins = dict(B.TABLE)  # per instance: a host may add functions (tests/mex_stub/bridge.py: kr_mex)
        self.cache = {}               # file -> (script_body, {name: Function})
        self.lookup_cache = {}
        self.globals = {}
        self.warnings = []
        self.stdout = stdout if stdout is not None else sys.stdout
        self.files = {}               # fid -> python file
        self.next_fid = 3
        self.end_stack = []
        self.frames = []
        self.script_file = None
        self.calls = {}               # user-function call counts (evidence of what was executed)

    # ------------------------------------------------------------------ files and functions
    def load_file(self, fname):
        fname = os.path.abspath(fname)
        if fname not in self.cache:
            with open(fname, "r", encoding="utf-8", errors="replace") as fh:
                src = fh.read()
            body, funs = parse_source(src, fname)
            self.cache[fname] = (body, {f.name: f for f in funs}, funs[0] if funs else None)
        return self.cache[fname]

    def addpath(self, d, front=True):
        d = os.path.abspath(d)
        if d in self.path:
            self.path.remove(d)
        if front:
            self.path.insert(0, d)
        else:
            self.path.append(d)
        self.lookup_cache.clear()

    def rmpath(self, d):
        d = os.path.abspath(d)
        if d in self.path:
            self.path.remove(d)
        self.lookup_cache.clear()

    def find_function(self, name, frame=None):
        # nested functions, then the other functions of the calling file, then the path
        fn = frame.fn if frame is not None else None
        g = fn
        while g is not None:
            if name in g.nested:
                return g.nested[name]
            g = g.parent
        if fn is not None:
            top = fn
            while top.parent is not None:
                top = top.parent
            if name in top.locals:
                return top.locals[name]
        if name in self.lookup_cache:
            return self.lookup_cache[name]
        found = None
        for d in self.path:
            p = os.path.join(d, name + ".m")
            if os.path.isfile(p):
                body, funs, main = self.load_file(p)
                found = main if main is not None else ("script", p)
                break
        self.lookup_cache[name] = found
        return found

    def which_data_file(self, name):
        for cand in (name, name + ".mat"):
            if os.path.isfile(cand):
                return cand
        for d in self.path:
            for cand in (name, name + ".mat"):
                p = os.path.join(d, cand)
                if os.path.isfile(p):
                    return p
        raise MatlabError("Unable to read file '%s'. No such file or directory." % name)

    # ------------------------------------------------------------------ entry points
    def run_script(self, fname, ws=None):
        body, funs, main = self.load_file(fname)
        if body is None:
            raise MatlabError("%s is a function file" % fname)
        fr = Frame()
        if ws:
            fr.ws.update(ws)
        old = self.script_file
        self.script_file = os.path.abspath(fname)
        self.frames.append(fr)
        try:
            self.exec_block(body, fr)
        except _Return:
            pass
        finally:
            self.frames.pop()
            self.script_file = old
        return fr.ws

    def call(self, name, *args, nargout=1):
        """Call a function on the path (or a built-in) from Python; returns a list of nargout values."""
        args = [norm_val(a) for a in args]
        out = self.call_named(name, list(args), nargout, None)
        return out

    # ------------------------------------------------------------------ calls
    def call_named(self, name, args, nargout, frame):
        fn = self.find_function(name, frame)
        if isinstance(fn, Function):
            return self.call_function(fn, args, nargout, frame)
        if isinstance(fn, tuple):
            self.run_script_in(fn[1], frame)
            return []
        b = self.builtins.get(name)
        if b is None:
            raise MatlabError("Undefined function or variable '%s'." % name)
        out = b(self, args, nargout)
        if isinstance(out, tuple):
            out = list(out)
        elif not isinstance(out, list):
            out = [out]
        return [norm_val(v) for v in out]

    def run_script_in(self, fname, frame):
        body, funs, main = self.load_file(fname)
        self.exec_block(body, frame)

    def call_function(self, fn, args, nargout, caller):
        if len(args) > len(fn.params) and not (fn.params and fn.params[-1] == "varargin"):
            raise MatlabError("Too many input arguments (%s)." % fn.name)
        if nargout > max(len(fn.outs), 1) and not (fn.outs and fn.outs[-1] == "varargout"):
            raise MatlabError("Too many output arguments (%s)." % fn.name)
        self.calls[fn.name] = self.calls.get(fn.name, 0) + 1
        fr = Frame(fn, len(args), nargout)
        if fn.parent is not None:
            # nested function: shares the workspace of the function it is defined in
            p = caller
            while p is not None and p.fn is not fn.parent:
                p = p.parent
            fr.parent = p
        for k, pname in enumerate(fn.params):
            if pname == "varargin" and k == len(fn.params) - 1:
                fr.ws["varargin"] = Cell.row(args[k:])
                break
            if k < len(args) and pname != "~":
                fr.ws[pname] = args[k]
        self.frames.append(fr)
        try:
            self.exec_block(fn.body, fr)
        except _Return:
            pass
        finally:
            self.frames.pop()
        outs = []
        for k, oname in enumerate(fn.outs):
            if k >= max(nargout, 1):
                break
            if oname == "varargout":
                vo = fr.ws.get("varargout")
                if isinstance(vo, Cell):
                    outs.extend(list(vo.a.reshape(-1, order="F"))[:max(nargout, 1) - k])
                break
            if oname not in fr.ws:
                if k == 0 and nargout == 0:
                    break
                raise MatlabError("Output argument \"%s\" (and maybe others) not assigned during call to \"%s\"."
                                  % (oname, fn.name))
            outs.append(fr.ws[oname])
        return outs

    def call_handle(self, fh, args, nargout, frame=None):
        if fh.name is not None:
            if fh.frame is not None:
                fn = self.find_function(fh.name, fh.frame)
                if isinstance(fn, Function):
                    # a handle to a nested function runs in the workspace of the frame that created the handle
                    return self.call_function(fn, args, nargout, fh.frame)
            return self.call_named(fh.name, args, nargout, fh.frame)
        fr = Frame(fh.frame.fn if fh.frame is not None else None, len(args), nargout)
        fr.ws = dict(fh.env)
        for k, p in enumerate(fh.params):
            if p == "varargin" and k == len(fh.params) - 1:
                fr.ws["varargin"] = Cell.row(args[k:])
                break
            if k < len(args):
                if p != "~":
                    fr.ws[p] = args[k]
        if len(args) > len(fh.params) and not (fh.params and fh.params[-1] == "varargin"):
            raise MatlabError("Too many input arguments (anonymous function).")
        self.frames.append(fr)
        try:
            vals = self.eval_multi(fh.body, fr, nargout)
        finally:
            self.frames.pop()
        return vals

    # ------------------------------------------------------------------ statements
    def exec_block(self, body, fr):
        for st in body:
            self.exec_stmt(st, fr)

    def exec_stmt(self, st, fr):
        kind = st[0]
        try:
            if kind == "assign":
                val = self.eval(st[2], fr)
                self.assign(st[1], val, fr)
            elif kind == "expr":
                e = st[1]
                if e[0] == "id" and fr.lookup(e[1]) is None and e[1] not in ("nargin", "nargout"):
                    vals = self.call_named(e[1], [], 0, fr)
                else:
                    vals = self.eval_multi(e, fr, 0)
                if vals:
                    self.setvar(fr, "ans", vals[0])
            elif kind == "massign":
                lhs = st[1]
                vals = self.eval_multi(st[2], fr, len(lhs))
                # fewer outputs than targets is only fine when the missing ones are ~ placeholders at the end
                if len(vals) < len(lhs) and any(l is not None for l in lhs[len(vals):]):
                    raise MatlabError("Too many output arguments.")
                for l, v in zip(lhs, vals):
                    if l is not None:
                        self.assign(l, v, fr)
            elif kind == "if":
                for cond, body in st[1]:
                    if truth(self.eval(cond, fr)):
                        self.exec_block(body, fr)
                        break
                else:
                    if st[2] is not None:
                        self.exec_block(st[2], fr)
            elif kind == "for":
                it = self.eval(st[2], fr)
                var = st[1]
                if isinstance(it, Cell):
                    cols = [Cell(it.a[:, j:j + 1]) for j in range(it.a.shape[1])] if it.a.size else []
                elif is_num(it):
                    if sp.issparse(it):
                        it = it.toarray()
                    cols = [it[:, j:j + 1] for j in range(it.shape[1])] if it.size else []
                elif isinstance(it, str):
                    cols = list(it)
                else:
                    cols = [it]
                for c in cols:
                    self.setvar(fr, var, c)
                    try:
                        self.exec_block(st[3], fr)
                    except _Break:
                        break
                    except _Continue:
                        continue
                if not cols:
                    self.setvar(fr, var, np.zeros((1, 0)))
            elif kind == "while":
                while truth(self.eval(st[1], fr)):
                    try:
                        self.exec_block(st[2], fr)
                    except _Break:
                        break
                    except _Continue:
                        continue
            elif kind == "switch":
                v = self.eval(st[1], fr)
                done = False
                for ce, body in st[2]:
                    c = self.eval(ce, fr)
                    cands = list(c.a.reshape(-1, order="F")) if isinstance(c, Cell) else [c]
                    if any(self.switch_match(v, x) for x in cands):
                        self.exec_block(body, fr)
                        done = True
                        break
                if not done and st[3] is not None:
                    self.exec_block(st[3], fr)
            elif kind == "try":
                try:
                    self.exec_block(st[1], fr)
                except MatlabError as err:
                    if st[2] is not None:
                        self.setvar(fr, st[2], Struct({"message": err.args[0] if err.args else "", "identifier": ""}))
                    self.exec_block(st[3], fr)
            elif kind == "break":
                raise _Break()
            elif kind == "continue":
                raise _Continue()
            elif kind == "return":
                raise _Return()
            elif kind == "global":
                for name in st[1]:
                    fr.globals.add(name)
                    if name not in self.globals:
                        self.globals[name] = EMPTY
            elif kind == "command":
                self.call_named(st[1], list(st[2]), 0, fr)
            else:
                raise MatlabError("unknown statement %s" % kind)
        except MatlabError as e:
            if not getattr(e, "located", False):
                e.located = True
                where = fr.fn.fname if fr.fn is not None else (self.script_file or "<script>")
                e.args = ("%s  [%s:%d]" % (e.args[0] if e.args else "", os.path.basename(where), st[-1]),)
            raise

    @staticmethod
    def switch_match(v, c):
        if isinstance(v, str) or isinstance(c, str):
            return isinstance(v, str) and isinstance(c, str) and v == c
        return is_num(v) and is_num(c) and dense(v).size == 1 and dense(c).size == 1 and \
            dense(v).reshape(-1)[0] == dense(c).reshape(-1)[0]

    # ------------------------------------------------------------------ variables
    def getvar(self, fr, name):
        if name in fr.globals:
            return self.globals[name]
        f = fr.lookup(name)
        return None if f is None else f.ws[name]

    def setvar(self, fr, name, val):
        if name in fr.globals:
            self.globals[name] = val
            return
        f = fr.lookup(name) if fr.parent is not None else None
        # a nested function writes to the enclosing workspace only for names that live there and are
        # neither its own parameters nor its outputs
        if f is not None and f is not fr and fr.fn is not None and name not in fr.fn.params and name not in fr.fn.outs:
            f.ws[name] = val
        else:
            fr.ws[name] = val

    def has_var(self, fr, name):
        return name in fr.globals or fr.lookup(name) is not None

    # ------------------------------------------------------------------ assignment
    def assign(self, target, val, fr):
        if target[0] == "id":
            self.setvar(fr, target[1], val)
            return
        # unwind the accessor chain down to the variable
        chain = []
        node = target
        while node[0] in ("index", "field"):
            chain.append(node)
            node = node[1]
        if node[0] != "id":
            raise MatlabError("invalid assignment target")
        chain.reverse()
        name = node[1]
        cur = self.getvar(fr, name)
        new = self.assign_into(cur, chain, val, fr)
        self.setvar(fr, name, new)

    def eval_index_args(self, args, obj, fr):
        out = []
        n = len(args)
        for k, a in enumerate(args):
            if a[0] == "colon":
                out.append(COLON)
                continue
            self.end_stack.append((obj, k, n))
            try:
                if a[0] == "index" and a[2] == "{}" or a[0] == "id":
                    vals = self.eval_multi(a, fr, 1)
                    out.extend(vals)
                else:
                    out.append(self.eval(a, fr))
            finally:
                self.end_stack.pop()
        return out

    def assign_into(self, cur, chain, val, fr):
        if not chain:
            return val
        acc = chain[0]
        rest = chain[1:]
        if acc[0] == "field":
            fname = acc[2] if not acc[3] else self.eval(acc[2], fr)
            if not isinstance(fname, str):
                raise MatlabError("dynamic field name must be a string")
            if cur is None or (is_num(cur) and cur.shape == (0, 0)):
                cur = Struct()
            if not isinstance(cur, Struct):
                raise MatlabError("field assignment into a non-struct value")
            new = Struct(cur.f)
            new.f[fname] = self.assign_into(cur.f.get(fname), rest, val, fr)
            return new
        kind = acc[2]
        idx = self.eval_index_args(acc[3], cur if cur is not None else EMPTY, fr)
        if kind == "()":
            if not rest:
                return index_set(cur, idx, val)
            sub = index_get(cur, idx) if cur is not None else None
            newsub = self.assign_into(sub, rest, val, fr)
            return index_set(cur, idx, newsub)
        # brace index into a cell
        if cur is None or (is_num(cur) and cur.shape == (0, 0)):
            cur = Cell()
        if not isinstance(cur, Cell):
            raise MatlabError("brace indexing assignment into a non-cell value")
        sub = None
        if rest:
            try:
                got = index_get(cur, idx)
                if got.a.size == 1:
                    sub = got.a.reshape(-1)[0]
                    if is_num(sub) and sub.shape == (0, 0):
                        sub = sub        # keep []: deeper accessors decide what it becomes
            except MatlabError:
                sub = None
        newsub = self.assign_into(sub, rest, val, fr)
        wrapped = np.empty((1, 1), dtype=object)
        wrapped[0, 0] = newsub
        return index_set(cur, idx, Cell(wrapped))

    # ------------------------------------------------------------------ expressions
    def eval(self, e, fr):
        vals = self.eval_multi(e, fr, 1)
        if not vals:
            raise MatlabError("expression produced no value")
        return vals[0]

    def eval_args(self, args, fr):
        out = []
        for a in args:
            if a[0] == "colon":
                out.append(":")
            elif a[0] in ("index", "id"):
                out.extend(self.eval_multi(a, fr, 1))
            else:
                out.append(self.eval(a, fr))
        return out

    def eval_multi(self, e, fr, nargout):
        k = e[0]
        if k == "num":
            return [scalar(e[1])]
        if k == "str":
            return [e[1]]
        if k == "paren":
            return [self.eval(e[1], fr)]
        if k == "id":
            name = e[1]
            if name == "nargin" and not self.has_var(fr, name):
                return [scalar(float(fr.nargin))]
            if name == "nargout" and not self.has_var(fr, name):
                return [scalar(float(fr.nargout))]
            v = self.getvar(fr, name)
            if v is not None:
                return [v]
            return self.call_named(name, [], max(nargout, 1) if nargout else 1, fr)
        if k == "end":
            if not self.end_stack:
                raise MatlabError("`end` outside an index expression")
            obj, pos, n = self.end_stack[-1]
            shp = obj.shape if hasattr(obj, "shape") else ((1, len(obj)) if isinstance(obj, str) else (1, 1))
            if n == 1:
                return [scalar(float(shp[0] * shp[1]))]
            if pos < 2:
                return [scalar(float(shp[pos]))]
            return [scalar(1.0)]
        if k == "colon":
            return [":"]
        if k == "binop":
            from .ops import binop
            return [binop(e[1], self.eval(e[2], fr), self.eval(e[3], fr))]
        if k == "unop":
            from .ops import unop
            return [unop(e[1], self.eval(e[2], fr))]
        if k == "postfix":
            from .ops import transpose
            return [transpose(self.eval(e[2], fr), e[1] == "'")]
        if k == "andand":
            a = self.eval(e[1], fr)
            if not self.logical_scalar(a, "&&"):
                return [scalar(False)]
            return [scalar(self.logical_scalar(self.eval(e[2], fr), "&&"))]
        if k == "oror":
            a = self.eval(e[1], fr)
            if self.logical_scalar(a, "||"):
                return [scalar(True)]
            return [scalar(self.logical_scalar(self.eval(e[2], fr), "||"))]
        if k == "range":
            from .ops import make_range
            a = self.eval(e[1], fr)
            s = self.eval(e[2], fr) if e[2] is not None else None
            b = self.eval(e[3], fr)
            return [make_range(a, s, b)]
        if k == "matrix":
            rows = [self.eval_args(r, fr) for r in e[1]]
            return [norm_val(concat(rows))]
        if k == "cell":
            rows = [self.eval_args(r, fr) for r in e[1]]
            rows = [r for r in rows if r]
            if not rows:
                return [Cell()]
            if len({len(r) for r in rows}) != 1:
                raise MatlabError("Dimensions of arrays being concatenated are not consistent.")
            a = np.empty((len(rows), len(rows[0])), dtype=object)
            for i, r in enumerate(rows):
                for j, v in enumerate(r):
                    a[i, j] = v
            return [Cell(a)]
        if k == "anon":
            # snapshot of the defining workspace (values are immutable, so a shallow copy is a value capture)
            env = {}
            f = fr
            chain = []
            while f is not None:
                chain.append(f)
                f = f.parent
            for f in reversed(chain):
                env.update(f.ws)
            for g in fr.globals:
                env[g] = self.globals[g]
            return [FH(None, e[1], e[2], env, fr)]
        if k == "fhandle":
            return [FH(e[1], frame=fr)]
        if k == "field":
            base = self.eval(e[1], fr)
            fname = e[2] if not e[3] else self.eval(e[2], fr)
            if not isinstance(base, Struct):
                raise MatlabError("Dot indexing is not supported for variables of this type.")
            if fname not in base.f:
                raise MatlabError("Unrecognized field name \"%s\"." % fname)
            return [base.f[fname]]
        if k == "index":
            return self.eval_index(e, fr, nargout)
        raise MatlabError("cannot evaluate node %s" % k)

    @staticmethod
    def logical_scalar(v, op):
        if isinstance(v, str):
            raise MatlabError("Operands to the %s operator must be convertible to logical scalar values." % op)
        d = dense(v)
        if d.size != 1:
            raise MatlabError("Operands to the || and && operators must be convertible to logical scalar values.")
        return bool(d.reshape(-1)[0] != 0)

    def eval_index(self, e, fr, nargout):
        base, kind, args = e[1], e[2], e[3]
        # name(...) where name is not a variable: a function call
        if base[0] == "id" and kind == "()" and not self.has_var(fr, base[1]) and base[1] not in ("nargin", "nargout"):
            argv = self.eval_args(args, fr)
            return self.call_named(base[1], argv, nargout, fr)
        obj = self.eval(base, fr)
        if kind == "()":
            if isinstance(obj, FH):
                argv = self.eval_args(args, fr)
                return self.call_handle(obj, argv, max(nargout, 1), fr)
            idx = self.eval_index_args(args, obj, fr)
            return [norm_val(index_get(obj, idx))]
        # brace: comma-separated list of cell contents
        if not isinstance(obj, Cell):
            raise MatlabError("Brace indexing is not supported for variables of this type.")
        idx = self.eval_index_args(args, obj, fr)
        sub = index_get(obj, idx)
        return list(sub.a.reshape(-1, order="F"))
