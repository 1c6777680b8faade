"""Recursive-descent parser of the MATLAB-subset interpreter (see oracle/mlab/__init__.py).  TEST INFRASTRUCTURE ONLY.

Produces a small tuple AST.  Operator precedence follows MATLAB's table: `||` < `&&` < `|` < `&` < comparisons < `:`
< `+ -` < `* / \\ .* ./ .\\` < unary `+ - ~` < `^ .^` < postfix (transpose, indexing, field access)."""
from .lexer import tokenize


class ParseError(Exception):
    pass


class Function:
    def __init__(self, name, params, outs, body, fname):
        self.name, self.params, self.outs, self.body, self.fname = name, params, outs, body, fname
        self.nested = {}          # nested functions (share the parent's workspace)
        self.parent = None
        self.locals = {}          # the other functions of the same file

    def __repr__(self):
        return "<function %s (%s)>" % (self.name, self.fname)


BLOCK_OPENERS = {"if", "for", "while", "switch", "try", "parfor"}


class Parser:
    def __init__(self, src, fname="<string>"):
        self.fname = fname
        self.toks = tokenize(src, fname)
        self.i = 0
        kws = [t for t in self.toks if t.kind == "KW" and t.depth == 0]
        n_end = sum(1 for t in kws if t.val == "end")
        n_open = sum(1 for t in kws if t.val in BLOCK_OPENERS)
        n_fun = sum(1 for t in kws if t.val == "function")
        # functions closed by `end` (lanczos_krylov.m, normAm.m) or running to the next `function` / EOF (expmv.m)
        self.end_style = n_fun > 0 and n_end == n_open + n_fun

    # ---- token helpers
    @property
    def t(self):
        return self.toks[self.i]

    def peek(self, k=1):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def err(self, msg):
        raise ParseError("%s:%d: %s (at %r)" % (self.fname, self.t.line, msg, self.t.val))

    def is_op(self, *vals):
        return self.t.kind == "OP" and self.t.val in vals

    def is_kw(self, *vals):
        return self.t.kind == "KW" and self.t.val in vals

    def eat_op(self, val):
        if not self.is_op(val):
            self.err("expected %r" % val)
        self.i += 1

    def eat_kw(self, val):
        if not self.is_kw(val):
            self.err("expected %r" % val)
        self.i += 1

    def skip_newlines(self):
        while self.t.kind == "NL" or self.is_op(";", ","):
            self.i += 1

    # ---- file level
    def parse_file(self):
        """-> (script_body or None, [Function...])"""
        self.skip_newlines()
        if not self.is_kw("function"):
            body = self.parse_block(())
            if self.t.kind != "EOF":
                self.err("unexpected token in script")
            return body, []
        funs = []
        while self.is_kw("function"):
            funs.append(self.parse_function(None))
            self.skip_newlines()
        if self.t.kind != "EOF":
            self.err("unexpected token after functions")
        for f in funs:
            f.locals = {g.name: g for g in funs}
        return None, funs

    def parse_function(self, parent):
        self.eat_kw("function")
        outs = []
        # [o1, o2] = name(...)  |  o = name(...)  |  name(...)
        if self.is_op("["):
            self.i += 1
            while not self.is_op("]"):
                if self.is_op(","):
                    self.i += 1
                    continue
                if self.t.kind != "ID":
                    self.err("output name expected")
                outs.append(self.t.val)
                self.i += 1
            self.i += 1
            self.eat_op("=")
            name = self.t.val
            self.i += 1
        else:
            name = self.t.val
            if self.t.kind != "ID":
                self.err("function name expected")
            self.i += 1
            if self.is_op("="):
                self.i += 1
                outs = [name]
                name = self.t.val
                self.i += 1
        params = []
        if self.is_op("("):
            self.i += 1
            while not self.is_op(")"):
                if self.is_op(","):
                    self.i += 1
                    continue
                if self.is_op("~"):
                    params.append("~")
                elif self.t.kind == "ID":
                    params.append(self.t.val)
                else:
                    self.err("parameter name expected")
                self.i += 1
            self.i += 1
        f = Function(name, params, outs, None, self.fname)
        f.parent = parent
        body = []
        while True:
            self.skip_newlines()
            if self.t.kind == "EOF":
                if self.end_style:
                    self.err("function %s: missing end" % name)
                break
            if self.is_kw("function"):
                if self.end_style:
                    g = self.parse_function(f)
                    f.nested[g.name] = g
                    continue
                break
            if self.is_kw("end"):
                if not self.end_style:
                    self.err("unexpected end")
                self.i += 1
                break
            body.append(self.parse_statement())
        f.body = body
        return f

    def parse_block(self, terminators):
        body = []
        while True:
            self.skip_newlines()
            if self.t.kind == "EOF":
                if terminators:
                    self.err("unexpected end of file, expected one of %s" % (terminators,))
                return body
            if self.t.kind == "KW" and self.t.val in terminators:
                return body
            if self.is_kw("function"):
                self.err("function definition inside a block")
            body.append(self.parse_statement())

    # ---- statements
    def end_of_statement(self):
        """-> suppress flag"""
        if self.is_op(";"):
            self.i += 1
            if self.t.kind == "NL":
                self.i += 1
            return True
        if self.is_op(","):
            self.i += 1
            return False
        if self.t.kind == "NL":
            self.i += 1
            return False
        if self.t.kind == "EOF":
            return False
        if self.t.kind == "KW":           # `if c, x = 1; end` / `case 'a', tol = 1;`
            return False
        self.err("end of statement expected")

    def parse_statement(self):
        t = self.t
        line = t.line
        if t.kind == "KW":
            kw = t.val
            if kw == "if":
                return self.parse_if(line)
            if kw == "for":
                self.i += 1
                paren = self.is_op("(")
                if paren:
                    self.i += 1
                var = self.t.val
                self.i += 1
                self.eat_op("=")
                e = self.parse_expr()
                if paren:
                    self.eat_op(")")
                body = self.parse_block(("end",))
                self.eat_kw("end")
                return ("for", var, e, body, line)
            if kw == "while":
                self.i += 1
                c = self.parse_expr()
                body = self.parse_block(("end",))
                self.eat_kw("end")
                return ("while", c, body, line)
            if kw == "switch":
                self.i += 1
                e = self.parse_expr()
                self.skip_newlines()
                cases, other = [], None
                while not self.is_kw("end"):
                    if self.is_kw("case"):
                        self.i += 1
                        ce = self.parse_expr()
                        body = self.parse_block(("case", "otherwise", "end"))
                        cases.append((ce, body))
                    elif self.is_kw("otherwise"):
                        self.i += 1
                        other = self.parse_block(("case", "otherwise", "end"))
                    else:
                        self.err("case expected")
                self.eat_kw("end")
                return ("switch", e, cases, other, line)
            if kw == "try":
                self.i += 1
                body = self.parse_block(("catch", "end"))
                ident, cbody = None, []
                if self.is_kw("catch"):
                    self.i += 1
                    if self.t.kind == "ID" and self.t.line == self.toks[self.i - 1].line:
                        ident = self.t.val
                        self.i += 1
                    cbody = self.parse_block(("end",))
                self.eat_kw("end")
                return ("try", body, ident, cbody, line)
            if kw in ("break", "continue", "return"):
                self.i += 1
                self.end_of_statement()
                return (kw, line)
            if kw in ("global", "persistent"):
                self.i += 1
                names = []
                while self.t.kind == "ID":
                    names.append(self.t.val)
                    self.i += 1
                self.end_of_statement()
                return ("global", names, line)
            self.err("unexpected keyword")
        # command syntax:  load theta_taylor / clear a b / hold on
        if t.kind == "ID" and self.peek().kind == "ID" and self.peek().space and self.peek().line == t.line:
            name = t.val
            self.i += 1
            words = []
            while self.t.kind in ("ID", "NUM") and self.t.line == line:
                words.append(str(self.t.val))
                self.i += 1
            self.end_of_statement()
            return ("command", name, words, line)
        # multi-assignment  [a, b, ~] = f(...)
        if self.is_op("["):
            j, depth = self.i, 0
            while True:
                tk = self.toks[j]
                if tk.kind == "OP" and tk.val in "([{":
                    depth += 1
                elif tk.kind == "OP" and tk.val in ")]}":
                    depth -= 1
                    if depth == 0:
                        break
                elif tk.kind in ("NL", "EOF"):
                    break
                j += 1
            nxt = self.toks[j + 1] if j + 1 < len(self.toks) else None
            if nxt is not None and nxt.kind == "OP" and nxt.val == "=":
                self.i += 1
                lhs = []
                while not self.is_op("]"):
                    if self.is_op(","):
                        self.i += 1
                        continue
                    if self.is_op("~") and (self.peek().kind == "OP" and self.peek().val in (",", "]")):
                        lhs.append(None)
                        self.i += 1
                        continue
                    lhs.append(self.parse_postfix())
                self.i += 1
                self.eat_op("=")
                rhs = self.parse_expr()
                sup = self.end_of_statement()
                return ("massign", lhs, rhs, line)
        e = self.parse_expr()
        if self.is_op("="):
            if e[0] not in ("id", "index", "field"):
                self.err("invalid assignment target")
            self.i += 1
            rhs = self.parse_expr()
            self.end_of_statement()
            return ("assign", e, rhs, line)
        sup = self.end_of_statement()
        return ("expr", e, sup, line)

    def parse_if(self, line):
        self.eat_kw("if")
        clauses = []
        c = self.parse_expr()
        body = self.parse_block(("elseif", "else", "end"))
        clauses.append((c, body))
        other = None
        while True:
            if self.is_kw("elseif"):
                self.i += 1
                c = self.parse_expr()
                body = self.parse_block(("elseif", "else", "end"))
                clauses.append((c, body))
            elif self.is_kw("else"):
                self.i += 1
                other = self.parse_block(("end",))
            else:
                self.eat_kw("end")
                break
        return ("if", clauses, other, line)

    # ---- expressions
    def parse_expr(self):
        if self.is_op("@"):
            return self.parse_at()
        return self.parse_oror()

    def parse_at(self):
        self.eat_op("@")
        if self.is_op("("):
            self.i += 1
            params = []
            while not self.is_op(")"):
                if self.is_op(","):
                    self.i += 1
                    continue
                params.append("~" if self.is_op("~") else self.t.val)
                self.i += 1
            self.i += 1
            body = self.parse_expr()
            return ("anon", params, body)
        name = self.t.val
        if self.t.kind != "ID":
            self.err("function name expected after @")
        self.i += 1
        return ("fhandle", name)

    def parse_oror(self):
        a = self.parse_andand()
        while self.is_op("||"):
            self.i += 1
            a = ("oror", a, self.parse_andand())
        return a

    def parse_andand(self):
        a = self.parse_or()
        while self.is_op("&&"):
            self.i += 1
            a = ("andand", a, self.parse_or())
        return a

    def parse_or(self):
        a = self.parse_and()
        while self.is_op("|"):
            self.i += 1
            a = ("binop", "|", a, self.parse_and())
        return a

    def parse_and(self):
        a = self.parse_cmp()
        while self.is_op("&"):
            self.i += 1
            a = ("binop", "&", a, self.parse_cmp())
        return a

    def parse_cmp(self):
        a = self.parse_range()
        while self.is_op("==", "~=", "<", "<=", ">", ">="):
            op = self.t.val
            self.i += 1
            a = ("binop", op, a, self.parse_range())
        return a

    def in_index_colon(self):
        """A bare `:` used as an index (followed by `,` or `)` or `}`)."""
        return self.is_op(":") and self.peek().kind == "OP" and self.peek().val in (",", ")", "}")

    def parse_range(self):
        a = self.parse_add()
        if self.is_op(":") and not self.in_index_colon():
            self.i += 1
            b = self.parse_add()
            if self.is_op(":") and not self.in_index_colon():
                self.i += 1
                c = self.parse_add()
                return ("range", a, b, c)
            return ("range", a, None, b)
        return a

    def parse_add(self):
        a = self.parse_mul()
        while self.is_op("+", "-"):
            op = self.t.val
            self.i += 1
            a = ("binop", op, a, self.parse_mul())
        return a

    def parse_mul(self):
        a = self.parse_unary()
        while self.is_op("*", "/", "\\", ".*", "./", ".\\"):
            op = self.t.val
            self.i += 1
            a = ("binop", op, a, self.parse_unary())
        return a

    def parse_unary(self):
        if self.is_op("-", "+", "~"):
            op = self.t.val
            self.i += 1
            return ("unop", op, self.parse_unary())
        return self.parse_power()

    def parse_power(self):
        a = self.parse_postfix()
        while self.is_op("^", ".^"):
            op = self.t.val
            self.i += 1
            # the exponent may carry its own unary sign: 2^-1
            if self.is_op("-", "+", "~"):
                uop = self.t.val
                self.i += 1
                b = ("unop", uop, self.parse_postfix())
            else:
                b = self.parse_postfix()
            a = ("binop", op, a, b)
        return a

    def parse_args(self, close):
        args = []
        while not self.is_op(close):
            if self.is_op(","):
                self.i += 1
                continue
            if self.is_op(":") and self.peek().kind == "OP" and self.peek().val in (",", close):
                self.i += 1
                args.append(("colon",))
                continue
            args.append(self.parse_expr())
        self.i += 1
        return args

    def parse_postfix(self):
        a = self.parse_primary()
        while True:
            t = self.t
            if t.kind != "OP":
                break
            if t.val == "(":
                self.i += 1
                a = ("index", a, "()", self.parse_args(")"))
            elif t.val == "{":
                # inside a matrix / cell literal `a {b}` (with a blank) was already split by the lexer's comma
                self.i += 1
                a = ("index", a, "{}", self.parse_args("}"))
            elif t.val == "." and self.peek().kind == "ID" and not self.peek().space:
                self.i += 1
                a = ("field", a, self.t.val, False)
                self.i += 1
            elif t.val == "." and self.peek().kind == "OP" and self.peek().val == "(":
                self.i += 2
                e = self.parse_expr()
                self.eat_op(")")
                a = ("field", a, e, True)
            elif t.val in ("'", ".'"):
                self.i += 1
                a = ("postfix", t.val, a)
            else:
                break
        return a

    def parse_primary(self):
        t = self.t
        if t.kind == "NUM":
            self.i += 1
            return ("num", t.val)
        if t.kind == "STR":
            self.i += 1
            return ("str", t.val)
        if t.kind == "ID":
            self.i += 1
            return ("id", t.val)
        if t.kind == "KW" and t.val == "end" and t.depth > 0:
            self.i += 1
            return ("end",)
        if t.kind == "OP":
            if t.val == "(":
                self.i += 1
                e = self.parse_expr()
                self.eat_op(")")
                return ("paren", e)
            if t.val == "[":
                self.i += 1
                return ("matrix", self.parse_rows("]"))
            if t.val == "{":
                self.i += 1
                return ("cell", self.parse_rows("}"))
            if t.val == "@":
                return self.parse_at()
            if t.val == ":":
                self.i += 1
                return ("colon",)
        self.err("expression expected")

    def parse_rows(self, close):
        rows, row = [], []
        while not self.is_op(close):
            if self.is_op(","):
                self.i += 1
                continue
            if self.is_op(";"):
                self.i += 1
                if row:
                    rows.append(row)
                row = []
                continue
            if self.t.kind == "NL":
                self.i += 1
                continue
            row.append(self.parse_expr())
        self.i += 1
        if row:
            rows.append(row)
        return rows


def parse_source(src, fname="<string>"):
    return Parser(src, fname).parse_file()
