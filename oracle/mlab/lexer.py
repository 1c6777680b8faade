"""Lexer of the MATLAB-subset interpreter (see oracle/mlab/__init__.py).  TEST INFRASTRUCTURE ONLY.

Handles what the reference's functions/*.m use: `%` comments, `...` continuations, the quote that is a transpose
after a value and a string otherwise, numbers such as `1.` `.5` `1e-12` (with `1./x` lexed as `1` `./`), and the
white-space rules inside `[ ]` / `{ }` (a blank between two values separates elements: `[cost m]`, `[a -b]`,
`[V, ~]`) which are resolved HERE by emitting an explicit `,` token so that the parser never sees white space."""

KEYWORDS = {"function", "end", "if", "elseif", "else", "for", "while", "switch", "case", "otherwise",
            "break", "continue", "return", "global", "persistent", "try", "catch"}

OPS3 = ("...",)
OPS2 = ("==", "~=", "!=", "<=", ">=", "&&", "||", ".*", "./", ".\\", ".^", ".'")
OPS1 = "+-*/\\^'<>=&|~!,;:()[]{}.@"


class Tok:
    __slots__ = ("kind", "val", "line", "space", "depth")

    def __init__(self, kind, val, line, space, depth):
        self.kind = kind      # NUM STR ID KW OP NL EOF
        self.val = val
        self.line = line
        self.space = space    # white space immediately before the token
        self.depth = depth    # bracket nesting depth at the token (0 = statement level)

    def __repr__(self):
        return "Tok(%s,%r,l%d)" % (self.kind, self.val, self.line)


class LexError(Exception):
    pass


def _value_end(t):
    """Can the previous token end a value (so that a following quote is a transpose, a blank a separator)?"""
    if t is None:
        return False
    if t.kind in ("NUM", "STR", "ID"):
        return True
    if t.kind == "KW":
        return t.val == "end" and t.depth > 0
    return t.kind == "OP" and t.val in (")", "]", "}", "'", ".'")


def tokenize(src, fname="<string>"):
    toks = []
    stack = []            # open brackets
    i, n, line = 0, len(src), 1
    space = False

    def prev():
        return toks[-1] if toks else None

    def emit(kind, val):
        nonlocal space
        toks.append(Tok(kind, val, line, space, len(stack)))
        space = False

    while i < n:
        c = src[i]
        if c in " \t\r":
            i += 1
            space = True
            continue
        if c == "%" or c == "#":
            # block comment %{ ... %} on lines of their own
            if src.startswith("%{", i) and src[i:src.find("\n", i) if src.find("\n", i) >= 0 else n].strip() == "%{":
                j = src.find("\n%}", i)
                if j < 0:
                    raise LexError("%s:%d unterminated block comment" % (fname, line))
                line += src.count("\n", i, j + 3)
                i = j + 3
                continue
            while i < n and src[i] != "\n":
                i += 1
            continue
        if src.startswith("...", i):
            while i < n and src[i] != "\n":
                i += 1
            i += 1
            line += 1
            space = True
            continue
        if c == "\n":
            if stack and stack[-1] in "[{":
                # a line break inside a matrix / cell literal ends the row
                p = prev()
                if p is not None and not (p.kind == "OP" and p.val in (";", "[", "{")):
                    emit("OP", ";")
            elif not stack:
                emit("NL", "\n")
            i += 1
            line += 1
            space = True
            continue
        in_mat = bool(stack) and stack[-1] in "[{"
        p = prev()
        # ---- implicit element separator inside [ ] and { }
        if in_mat and space and _value_end(p):
            starts = False
            if c.isalnum() or c == "_" or c in "([{@'\"":
                starts = True
            elif c == "." and i + 1 < n and src[i + 1].isdigit():
                starts = True
            elif c in "+-" and i + 1 < n and src[i + 1] not in " \t=":
                starts = True          # [a -b] is two elements, [a - b] and [a-b] are one
            elif c == "~" and not src.startswith("~=", i):
                starts = True
            if starts:
                emit("OP", ",")
                space = True
                p = prev()
        # ---- numbers
        if c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            j = i
            while j < n and src[j].isdigit():
                j += 1
            if j < n and src[j] == "." and not (j + 1 < n and src[j + 1] in "*/\\^'"):
                j += 1
                while j < n and src[j].isdigit():
                    j += 1
            if j < n and src[j] in "eEdD":
                k = j + 1
                if k < n and src[k] in "+-":
                    k += 1
                if k < n and src[k].isdigit():
                    while k < n and src[k].isdigit():
                        k += 1
                    j = k
            text = src[i:j].replace("d", "e").replace("D", "e")
            val = float(text)
            if j < n and src[j] in "ij" and not (j + 1 < n and (src[j + 1].isalnum() or src[j + 1] == "_")):
                val = complex(0.0, val)
                j += 1
            emit("NUM", val)
            i = j
            continue
        # ---- identifiers / keywords
        if c.isalpha() or c == "_":
            j = i
            while j < n and (src[j].isalnum() or src[j] == "_"):
                j += 1
            w = src[i:j]
            is_field = p is not None and p.kind == "OP" and p.val == "." and not p.space
            if w in KEYWORDS and not is_field and not (w == "end" and stack):
                emit("KW", w)
            elif w == "end" and stack and not is_field:
                emit("KW", "end")      # the `end` of an index expression (depth > 0)
            else:
                emit("ID", w)
            i = j
            continue
        # ---- strings and transposes
        if c == "'":
            if _value_end(p) and not (space and in_mat):
                emit("OP", "'")
                i += 1
                continue
            j = i + 1
            buf = []
            while True:
                if j >= n or src[j] == "\n":
                    raise LexError("%s:%d unterminated string" % (fname, line))
                if src[j] == "'":
                    if j + 1 < n and src[j + 1] == "'":
                        buf.append("'")
                        j += 2
                        continue
                    break
                buf.append(src[j])
                j += 1
            emit("STR", "".join(buf))
            i = j + 1
            continue
        if c == '"':
            j = i + 1
            buf = []
            while True:
                if j >= n or src[j] == "\n":
                    raise LexError("%s:%d unterminated string" % (fname, line))
                if src[j] == '"':
                    if j + 1 < n and src[j + 1] == '"':
                        buf.append('"')
                        j += 2
                        continue
                    break
                buf.append(src[j])
                j += 1
            emit("STR", "".join(buf))
            i = j + 1
            continue
        # ---- operators
        two = src[i:i + 2]
        if two in OPS2:
            emit("OP", "~=" if two == "!=" else two)
            i += 2
            continue
        if c in OPS1:
            if c in "([{":
                emit("OP", c)
                stack.append(c)
            elif c in ")]}":
                if not stack:
                    raise LexError("%s:%d unbalanced %s" % (fname, line, c))
                stack.pop()
                emit("OP", c)
            else:
                emit("OP", "~" if c == "!" else c)
            i += 1
            continue
        raise LexError("%s:%d unexpected character %r" % (fname, line, c))
    if stack:
        raise LexError("%s: unbalanced %s at end of file" % (fname, stack[-1]))
    emit("NL", "\n")
    emit("EOF", None)
    return toks
