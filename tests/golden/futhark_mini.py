"""TEST INFRASTRUCTURE: a small interpreter for the subset of Futhark that flowamok is written in, so that the reference's OWN
source text can be executed here -- there is no Futhark compiler in the image -- and its outputs committed as golden vectors:
src/steppers.fut, helpers.fut, utils.fut, option.fut, reduce1.fut, flowamok.fut (make_stepper_fixture.py), README.fut /
animation.fut, explorer/explorer.fut + explorer_scenario_helper.fut + scenarios.fut, designer/designer.fut (make_lys_fixture.py,
make_explorer_runs.py).  It reads the reference files where they lie; nothing of them is copied.  Its own semantics are
held by known-answer programs in tests/test_futhark_mini.py.

Subset: declarations `def` / `local def` / `let` / `type` (skipped, except the size parameters of record abbreviations) /
`module X = { ... }`, module ascriptions, parametric modules and their application to module expressions, `import`,
`open`; expressions `let` (tuple / record patterns, local functions, `let a[i] = v`), `if`, `match` (tuple / constructor /
literal / record patterns), `loop ... while|for`, lambdas and sections, curried application, records with field access and
`with` updates (nested fields, array indices), tuples, arrays, slices, ranges, `#constructors`, strings and characters, `:>`,
the operators of the sources, and the prelude functions they call (tabulate_2d, map, map2, map3, zip, unzip, filter, all,
reduce, flatten, unflatten, replicate, length, copy, const, i64.sum, conversions).  Definitions are not recursive and later
ones shadow earlier ones (lexical scoping per declaration).  Types are parsed only far enough to skip them and to bind size
parameters (`[gh][gw]`, also through `state [] [gh] [gw]`) from the shapes of the arguments.

Values: integers and booleans are Python's (`/` and `%` round towards negative infinity, as in Futhark), every
floating-point value is numpy.float32, records are dicts, tuples are tuples, arrays are lists, strings are str, `#tag x` is
('#tag', x...).  Nothing is mutated in place.  A tuple component whose evaluation fails is poisoned instead of failing the
tuple (the reference relies on the compiler dropping unused components, explorer_scenario_helper.fut:26-31).  Modules the
sources import but that are not in the reference tree (diku-dk/cpprandom, athas/matte, diku-dk/lys) are supplied by the caller
as stubs: the steppers take `choose_direction` as a parameter, and where the frontends draw random numbers the stubs are the
repository's restated pcg32 (unpinned arithmetic).
"""
from __future__ import annotations

import re
import sys
from pathlib import Path

import numpy as np

sys.setrecursionlimit(20000)
F32 = np.float32      # every floating-point value of the sources is f32

KEYWORDS = {"let", "in", "if", "then", "else", "match", "case", "loop", "for", "while", "do", "with", "def", "local",
            "module", "import", "open", "type", "entry", "true", "false", "assert", "val"}
DECL_START = {"def", "local", "module", "import", "open", "type", "entry"}

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<comment>--[^\n]*)
  | (?P<char>'[^'\\]')
  | (?P<num>0b[01]+|0x[0-9a-fA-F]+|\d+(\.\d+)?(i8|i16|i32|i64|u8|u16|u32|u64|f32|f64)?)
  | (?P<str>"[^"]*")
  | (?P<tag>\#[A-Za-z_][A-Za-z_0-9']*)
  | (?P<id>[A-Za-z_][A-Za-z_0-9']*)
  | (?P<op>\.\.<|\.\.\.|->|==|!=|<=|>=|&&|\|\||\+\+|:>|[-+*/%<>=!\\(){}\[\],.:|'~^])
""", re.X)


class Tok:
    __slots__ = ("kind", "text", "space_before", "line")

    def __init__(self, kind, text, space_before, line):
        self.kind, self.text, self.space_before, self.line = kind, text, space_before, line

    def __repr__(self):
        return f"{self.kind}:{self.text!r}@{self.line}"


def lex(src):
    toks, pos, space, line = [], 0, True, 1
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot lex at line {line}: {src[pos:pos + 30]!r}")
        kind, text = m.lastgroup, m.group()
        pos = m.end()
        if kind in ("ws", "comment"):
            space = True
            line += text.count("\n")
            continue
        toks.append(Tok(kind, text, space, line))
        space = False
    toks.append(Tok("eof", "", True, line))
    return toks


# ---------------------------------------------------------------------------------------------------------------------
# parser: AST nodes are tuples ('kind', ...)
# ---------------------------------------------------------------------------------------------------------------------
BINOPS = [  # lowest precedence first
    ("..<",), ("||",), ("&&",), ("==", "!=", "<", "<=", ">", ">="), ("++",), ("+", "-"), ("*", "/", "%"),
]
LEVEL = {op: i for i, ops in enumerate(BINOPS) for op in ops}


class Parser:
    def __init__(self, toks, name="?"):
        self.t, self.i, self.name = toks, 0, name

    # -- token helpers
    def peek(self, k=0):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def at(self, text):
        return self.peek().text == text and self.peek().kind != "str"

    def eat(self, text=None):
        tok = self.peek()
        if text is not None and tok.text != text:
            raise SyntaxError(f"{self.name}:{tok.line}: expected {text!r}, found {tok.text!r}")
        self.i += 1
        return tok

    # -- types: skipped, except for the leading dimensions of a parameter's type
    def skip_type(self, stop):
        """skip a type up to (not including) one of the `stop` tokens at nesting depth 0; returns the leading dims"""
        dims, depth, leading = [], 0, True
        self.last_tyapp = None
        first = self.peek()
        if first.kind == "id" and first.text not in KEYWORDS and self.peek(1).text == "[":     # `state [] [gh] [gw]`
            j, args = 1, []
            while self.peek(j).text == "[":
                if self.peek(j + 1).text == "]":
                    args.append(None)
                    j += 2
                elif self.peek(j + 2).text == "]":
                    args.append(self.peek(j + 1).text)
                    j += 3
                else:
                    break
            self.last_tyapp = (first.text, args)
        while True:
            tok = self.peek()
            if tok.kind == "eof":
                break
            if depth == 0 and tok.text in stop:
                break
            if tok.text in "([{" and tok.kind == "op":
                if leading and tok.text == "[" and self.peek(1).kind in ("id", "num") and self.peek(2).text == "]":
                    dims.append(self.peek(1).text)
                    self.i += 3
                    continue
                depth += 1
                leading = False
            elif tok.text in ")]}" and tok.kind == "op":
                if depth == 0:
                    break
                depth -= 1
            elif tok.text == "*" and leading:
                pass
            else:
                leading = False
            self.i += 1
        return dims

    # -- patterns
    def pattern_atom(self):
        tok = self.peek()
        if tok.kind == "id" and tok.text not in KEYWORDS:
            self.eat()
            return ("pvar", tok.text) if tok.text != "_" else ("pwild",)
        if tok.text in ("true", "false"):
            self.eat()
            return ("plit", tok.text == "true")
        if tok.kind == "num":
            self.eat()
            return ("plit", number(tok.text))
        if tok.text == "-" and self.peek(1).kind == "num":
            self.eat()
            return ("plit", -number(self.eat().text))
        if tok.kind == "tag":
            self.eat()
            args = []
            while self.pattern_starts():
                args.append(self.pattern_atom())
            return ("pcon", tok.text, args)
        if tok.text == "(":
            self.eat("(")
            if self.at(")"):
                self.eat(")")
                return ("ptuple", [])
            items = [self.pattern()]
            while self.at(","):
                self.eat(",")
                items.append(self.pattern())
            self.eat(")")
            return items[0] if len(items) == 1 else ("ptuple", items)
        if tok.text == "{":
            self.eat("{")
            fields = []
            while not self.at("}"):
                name = self.eat().text
                sub = ("pvar", name)
                if self.at("="):
                    self.eat("=")
                    sub = self.pattern()
                fields.append((name, sub))
                if self.at(","):
                    self.eat(",")
            self.eat("}")
            return ("prec", fields)
        raise SyntaxError(f"{self.name}:{tok.line}: pattern expected, found {tok.text!r}")

    def pattern_starts(self):
        tok = self.peek()
        return (tok.kind == "id" and tok.text not in KEYWORDS) or tok.kind in ("num", "tag") or tok.text in ("(", "{", "true", "false")

    def pattern(self):
        p = self.pattern_atom()
        if self.at(":"):
            self.eat(":")
            dims = self.skip_type({",", ")", "="})
            p = ("ptyped", p, dims, self.last_tyapp)
        return p

    # -- expressions
    def expr(self):
        tok = self.peek()
        if tok.text == "let":
            return self.let()
        if tok.text == "if":
            self.eat("if")
            c = self.expr()
            self.eat("then")
            a = self.expr()
            self.eat("else")
            return ("if", c, a, self.expr())
        if tok.text == "match":
            self.eat("match")
            scrut = self.expr()
            cases = []
            while self.at("case"):
                self.eat("case")
                p = self.pattern()
                self.eat("->")
                cases.append((p, self.expr()))
            return ("match", scrut, cases)
        if tok.text == "loop":
            return self.loop()
        if tok.text == "\\":
            self.eat("\\")
            params = []
            while not self.at("->"):
                params.append(self.pattern_atom_typed())
            self.eat("->")
            return ("lam", params, self.expr())
        return self.binop(0)

    def pattern_atom_typed(self):
        """a lambda / function parameter: an atom pattern, possibly `(p: type)`"""
        return self.pattern_atom()

    def let(self):
        self.eat("let")
        tok = self.peek()
        # let name[idx] = value  (in-place update sugar)
        if tok.kind == "id" and self.peek(1).text == "[" and not self.peek(1).space_before:
            name = self.eat().text
            self.eat("[")
            idx = self.index_list()
            self.eat("]")
            self.eat("=")
            val = self.expr()
            body = self.let_body()
            return ("let", ("pvar", name), ("with_idx", ("var", name), idx, val), body)
        # let f p1 p2 ... [: type] = body   (local function) -- a name followed by parameters
        if tok.kind == "id" and tok.text not in KEYWORDS and (self.peek(1).text not in ("=", ":", ",")):
            name = self.eat().text
            sizes, params = self.params()
            if self.at(":"):
                self.eat(":")
                self.skip_type({"="})
            self.eat("=")
            fbody = self.expr()
            body = self.let_body()
            return ("letfun", name, sizes, params, fbody, body)
        p = self.pattern()
        self.eat("=")
        val = self.expr()
        return ("let", p, val, self.let_body())

    def let_body(self):
        if self.at("in"):
            self.eat("in")
            return self.expr()
        if self.at("let"):
            return self.let()
        tok = self.peek()
        raise SyntaxError(f"{self.name}:{tok.line}: `in` or `let` expected after a let binding, found {tok.text!r}")

    def params(self):
        sizes, params = [], []
        while True:
            tok = self.peek()
            if tok.text == "'":                      # type parameter 'aux
                self.eat()
                self.eat()
            elif tok.text == "[" and self.peek(2).text == "]":   # size parameter [gh]
                self.eat()
                sizes.append(self.eat().text)
                self.eat("]")
            elif self.pattern_starts():
                params.append(self.pattern_atom())
            else:
                return sizes, params

    def loop(self):
        self.eat("loop")
        p = self.pattern()
        init = None
        if self.at("="):
            self.eat("=")
            init = self.expr()
        else:                                          # `loop x for ...`: x starts as the variable of the same name
            init = pattern_as_expr(p)
        if self.at("while"):
            self.eat("while")
            cond = self.expr()
            self.eat("do")
            return ("loop_while", p, init, cond, self.expr())
        self.eat("for")
        var = self.eat().text
        self.eat("<")
        bound = self.expr()
        self.eat("do")
        return ("loop_for", p, init, var, bound, self.expr())

    def binop(self, level):
        if level == len(BINOPS):
            return self.with_expr()
        left = self.binop(level + 1)
        while self.peek().kind == "op" and self.peek().text in BINOPS[level]:
            op = self.eat().text
            nxt = self.peek().text
            right = self.expr() if nxt in ("if", "let", "match", "loop", "\\") else self.binop(level + 1)
            left = ("bin", op, left, right)
        return left

    COERCION_STOP = {"let", "in", "then", "else", "do", "case", ")", ",", "]", "}", "with", "def", "local", "module", "import", "open",
                     "type", "entry"}

    def with_expr(self):
        e = self.unary()
        if self.at(":>"):                              # size coercion: a no-op here
            self.eat()
            self.skip_type(self.COERCION_STOP)
        while self.at("with"):
            self.eat("with")
            if self.at("["):
                self.eat("[")
                idx = self.index_list()
                self.eat("]")
                self.eat("=")
                e = ("with_idx", e, idx, self.rhs_of_with())
            else:
                path = [self.eat().text]
                while self.at("."):
                    self.eat(".")
                    path.append(self.eat().text)
                self.eat("=")
                e = ("with_field", e, path, self.rhs_of_with())
        return e

    def rhs_of_with(self):
        nxt = self.peek().text
        if nxt in ("if", "let", "match", "loop", "\\"):
            return self.expr()
        # an operator expression that does not itself swallow the next `with` of the chain
        return self.binop_nowith(0)

    def binop_nowith(self, level):
        if level == len(BINOPS):
            return self.unary()
        left = self.binop_nowith(level + 1)
        while self.peek().kind == "op" and self.peek().text in BINOPS[level]:
            op = self.eat().text
            left = ("bin", op, left, self.binop_nowith(level + 1))
        return left

    def unary(self):
        if self.at("!"):
            self.eat("!")
            return ("not", self.unary())
        if self.at("-") :
            self.eat("-")
            return ("neg", self.unary())
        return self.application()

    def atom_starts(self):
        tok = self.peek()
        if tok.kind in ("num", "tag", "str", "char"):
            return True
        if tok.kind == "id":
            return tok.text not in KEYWORDS or tok.text in ("true", "false")
        return tok.text in ("(", "{", "[")

    def application(self):
        if self.at("assert"):
            self.eat("assert")
            c = self.atom()
            return ("assert", c, self.atom())
        f = self.atom()
        if f[0] == "con":
            args = []
            while self.atom_starts():
                args.append(self.atom())
            return ("con", f[1], args)
        while self.atom_starts():
            f = ("app", f, self.atom())
        return f

    def index_list(self):
        idx = []
        while True:
            if self.at(":"):                            # [:n]
                self.eat(":")
                idx.append(("slice", None, self.expr()))
            else:
                e = self.expr()
                if self.at(":"):
                    self.eat(":")
                    hi = None if self.at("]") or self.at(",") else self.expr()
                    idx.append(("slice", e, hi))
                else:
                    idx.append(e)
            if self.at(","):
                self.eat(",")
                continue
            return idx

    def atom(self):
        tok = self.peek()
        if tok.kind == "num":
            self.eat()
            e = ("lit", number(tok.text))
        elif tok.kind == "str":
            self.eat()
            e = ("lit", tok.text[1:-1].replace("\\n", "\n"))
        elif tok.kind == "char":
            self.eat()
            e = ("lit", ord(tok.text[1]))
        elif tok.text in ("true", "false"):
            self.eat()
            e = ("lit", tok.text == "true")
        elif tok.kind == "tag":
            self.eat()
            return ("con", tok.text)
        elif tok.kind == "id" and tok.text not in KEYWORDS:
            self.eat()
            e = ("var", tok.text)
        elif tok.text == "(":
            self.eat("(")
            if self.at(")"):
                self.eat(")")
                e = ("tuple", [])
            elif self.at(".") and self.peek(1).kind in ("num", "id") and self.peek(2).text == ")":   # (.0) projection section
                self.eat(".")
                fld = self.eat().text
                self.eat(")")
                e = ("lam", [("pvar", "$p")], ("field", ("var", "$p"), fld))
            elif self.peek().kind == "op" and self.peek().text in LEVEL and self.peek(1).text == ")":   # (+) operator section
                op = self.eat().text
                self.eat(")")
                e = ("lam", [("pvar", "$a"), ("pvar", "$b")], ("bin", op, ("var", "$a"), ("var", "$b")))
            else:
                items = [self.expr()]
                while self.at(","):
                    self.eat(",")
                    items.append(self.expr())
                if self.at(":>") or self.at(":"):       # coercion / ascription inside parentheses
                    self.eat()
                    self.skip_type({")"})
                self.eat(")")
                e = items[0] if len(items) == 1 else ("tuple", items)
        elif tok.text == "{":
            self.eat("{")
            fields = []
            while not self.at("}"):
                name = self.eat().text
                if self.at("="):
                    self.eat("=")
                    fields.append((name, self.expr()))
                else:
                    fields.append((name, ("var", name)))
                if self.at(","):
                    self.eat(",")
            self.eat("}")
            e = ("record", fields)
        elif tok.text == "[":
            self.eat("[")
            items = []
            while not self.at("]"):
                items.append(self.expr())
                if self.at(","):
                    self.eat(",")
            self.eat("]")
            e = ("array", items)
        else:
            raise SyntaxError(f"{self.name}:{tok.line}: expression expected, found {tok.text!r}")
        # postfix: .field  and  [index] (only when it follows without a space)
        while True:
            nxt = self.peek()
            if nxt.text == "." and nxt.kind == "op" and not nxt.space_before and self.peek(1).kind in ("id", "num"):
                self.eat(".")
                e = ("field", e, self.eat().text)
            elif nxt.text == "[" and not nxt.space_before:
                self.eat("[")
                idx = self.index_list()
                self.eat("]")
                e = ("index", e, idx)
            else:
                return e

    # -- declarations
    def decls(self, until=None):
        out = []
        while True:
            tok = self.peek()
            if tok.kind == "eof" or (until and tok.text == until):
                return out
            local = False
            if tok.text == "local":
                self.eat()
                local = True
                tok = self.peek()
            if tok.text in ("def", "entry", "let"):
                self.eat()
                name = self.eat().text
                sizes, params = self.params()
                if self.at(":"):
                    self.eat(":")
                    self.skip_type({"="})
                self.eat("=")
                out.append(("def", name, sizes, params, self.expr(), local))
            elif tok.text == "open":
                self.eat()
                if self.at("import"):
                    self.eat()
                    out.append(("import", self.eat().text[1:-1], True))
                else:
                    out.append(("open", self.eat().text))
            elif tok.text == "import":
                self.eat()
                out.append(("import", self.eat().text[1:-1], False))
            elif tok.text == "module":
                self.eat()
                if self.at("type"):                      # module type: skipped
                    self.eat()
                    self.eat()
                    self.eat("=")
                    self.skip_braces()
                    continue
                name = self.eat().text
                fparams = []
                while self.at("("):                      # parametric module: (p: signature)
                    self.eat("(")
                    fparams.append(self.eat().text)
                    self.eat(":")
                    self.skip_type({")"})
                    self.eat(")")
                if self.at(":"):                         # signature ascription, possibly `sig with t = u`
                    self.eat(":")
                    self.skip_module_sig()
                self.eat("=")
                mexp = self.module_expr()
                out.append(("module", name, fparams, mexp, local))
            elif tok.text == "type":
                self.eat()
                td = self.typedef()
                if td:
                    out.append(td)
                while self.peek().kind != "eof" and self.peek().text not in DECL_START and not (until and self.peek().text == until
                                                                                              and self.depth0()):
                    if self.at("{"):
                        self.skip_braces()
                    else:
                        self.eat()
            else:
                raise SyntaxError(f"{self.name}:{tok.line}: declaration expected, found {tok.text!r}")

    def depth0(self):
        return True

    def typedef(self):
        """`type name [p] [q] = {field: [p][q]elem, ...}`: which field and axis carries each size parameter (look-ahead only)"""
        j = 0
        if self.peek(j).text in ("~", "^"):
            j += 1
        name = self.peek(j).text
        j += 1
        params = []
        while self.peek(j).text == "[" and self.peek(j + 2).text == "]":
            params.append(self.peek(j + 1).text)
            j += 3
        if not params or self.peek(j).text != "=" or self.peek(j + 1).text != "{":
            return None
        j += 2
        where, depth = {}, 1
        while depth and self.peek(j).kind != "eof":
            tok = self.peek(j)
            if depth == 1 and tok.kind == "id" and self.peek(j + 1).text == ":":
                field, k, axis = tok.text, j + 2, 0
                while self.peek(k).text == "[" and self.peek(k + 2).text == "]":
                    where.setdefault(self.peek(k + 1).text, (field, axis))
                    axis += 1
                    k += 3
            depth += (tok.text == "{") - (tok.text == "}")
            j += 1
        return ("typedef", name, params, where)

    def skip_module_sig(self):
        """skip `sig` or `sig with t = u ...` up to the `=` that introduces the module's body"""
        depth = 0
        while True:
            tok = self.peek()
            if tok.kind == "eof":
                return
            if tok.text in ("(", "{", "["):
                depth += 1
            elif tok.text in (")", "}", "]"):
                depth -= 1
            elif tok.text == "=" and depth == 0:
                nxt = self.peek(1)
                if nxt.text in ("{", "import") or (nxt.kind == "id" and self.peek(2).text != "="):
                    return
            self.i += 1

    def module_expr(self):
        """{ decls } | import "f" | name | name.name | application of a parametric module to module expressions"""
        def atom():
            if self.at("{"):
                self.eat("{")
                body = self.decls(until="}")
                self.eat("}")
                return ("mbody", body)
            if self.at("import"):
                self.eat()
                return ("mimport", self.eat().text[1:-1])
            if self.at("("):
                self.eat("(")
                m = self.module_expr()
                self.eat(")")
                return m
            path = [self.eat().text]
            while self.at(".") and not self.peek().space_before:
                self.eat(".")
                path.append(self.eat().text)
            return ("mname", path)
        m = atom()
        while self.peek().kind != "eof" and self.peek().text not in DECL_START and not self.at("}") and not self.at(")") and \
                (self.at("{") or self.at("(") or (self.peek().kind == "id" and self.peek().text not in KEYWORDS)):
            m = ("mapp", m, atom())
        return m

    def skip_braces(self):
        self.eat("{")
        depth = 1
        while depth:
            t = self.eat().text
            depth += (t == "{") - (t == "}")


def number(text):
    if text.startswith("0b"):
        return int(text[2:], 2)
    if text.startswith("0x"):
        return int(text[2:], 16)
    m = re.match(r"(\d+(\.\d+)?)", text)
    if "." in m.group(1) or text.endswith(("f32", "f64")):
        return F32(m.group(1))
    return int(m.group(1))


def pattern_as_expr(p):
    if p[0] == "pvar":
        return ("var", p[1])
    if p[0] == "ptyped":
        return pattern_as_expr(p[1])
    if p[0] == "ptuple":
        return ("tuple", [pattern_as_expr(q) for q in p[1]])
    raise SyntaxError("loop without an initial value needs a variable pattern")


# ---------------------------------------------------------------------------------------------------------------------
# evaluator
# ---------------------------------------------------------------------------------------------------------------------
class Env:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None, vars=None):
        self.vars, self.parent = vars if vars is not None else {}, parent

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars:
                return e.vars[name]
            e = e.parent
        raise NameError(f"unbound Futhark name {name!r}")


class Module:
    def __init__(self, names):
        self.names = names


class Fn:
    """a curried function value: user-defined (params, body, env) or builtin (arity, python function)"""
    __slots__ = ("name", "sizes", "params", "body", "env", "args", "py")

    def __init__(self, name, sizes, params, body, env, args=(), py=None):
        self.name, self.sizes, self.params, self.body, self.env, self.args, self.py = name, sizes, params, body, env, args, py


def builtin(name, arity, py):
    return Fn(name, (), [None] * arity, None, None, (), py)


TYPEDEFS = {}      # record types with size parameters: name -> (parameters, {parameter: (field, axis)})


class FutharkError(Exception):
    pass


class Poison:
    """the value of a tuple component whose evaluation failed; fails again when it is looked at"""

    def __init__(self, exc):
        self.exc = exc

    def _fail(self, *_a, **_k):
        raise FutharkError(f"use of a failed value: {self.exc}")
    __getitem__ = __len__ = __iter__ = __add__ = __radd__ = __eq__ = __ne__ = __lt__ = __bool__ = __call__ = _fail


def apply(f, arg):
    if not isinstance(f, Fn):
        raise TypeError(f"applying a non-function {f!r}")
    args = f.args + (arg,)
    if len(args) < len(f.params):
        return Fn(f.name, f.sizes, f.params, f.body, f.env, args, f.py)
    if f.py is not None:
        return f.py(*args)
    env = Env(f.env)
    for p, a in zip(f.params, args):
        if not bind(p, a, env, f.sizes):
            raise FutharkError(f"{f.name}: argument does not match parameter pattern")
    return ev(f.body, env)


def call(f, *args):
    for a in args:
        f = apply(f, a)
    return f


def bind(p, v, env, sizes=()):
    k = p[0]
    if k == "pvar":
        env.vars[p[1]] = v
        return True
    if k == "pwild":
        return True
    if k == "ptyped":
        if sizes and p[2]:
            x = v
            for d in p[2]:
                if d in sizes and isinstance(x, list):
                    if d in env.vars and env.vars[d] != len(x) and len(x) > 0:
                        raise FutharkError(f"size parameter {d}: {env.vars[d]} vs {len(x)}")
                    env.vars.setdefault(d, len(x))
                x = x[0] if isinstance(x, list) and x else None
        if sizes and len(p) > 3 and p[3] and p[3][0] in TYPEDEFS and isinstance(v, dict):
            params, where = TYPEDEFS[p[3][0]]
            for formal, actual in zip(params, p[3][1]):
                if actual in sizes and formal in where:
                    field, axis = where[formal]
                    x = v[field]
                    for _ in range(axis):
                        x = x[0] if x else []
                    env.vars.setdefault(actual, len(x))
        return bind(p[1], v, env, sizes)
    if k == "ptuple":
        if not isinstance(v, tuple) or len(v) != len(p[1]):
            return False
        return all(bind(q, w, env, sizes) for q, w in zip(p[1], v))
    if k == "plit":
        return v == p[1] and type(v) is type(p[1])
    if k == "pcon":
        if not (isinstance(v, tuple) and v and v[0] == p[1] and len(v) - 1 == len(p[2])):
            return False
        return all(bind(q, w, env, sizes) for q, w in zip(p[2], v[1:]))
    if k == "prec":
        return all(bind(q, v[name], env, sizes) for name, q in p[1])
    raise FutharkError(f"pattern {k}")


def set_path(rec, path, val):
    if not path:
        return val
    new = dict(rec)
    new[path[0]] = set_path(rec[path[0]], path[1:], val)
    return new


def set_index(arr, idx, val):
    i = idx[0]
    if not 0 <= i < len(arr):
        raise FutharkError(f"index {i} out of bounds for array of {len(arr)}")
    new = list(arr)
    new[i] = val if len(idx) == 1 else set_index(arr[i], idx[1:], val)
    return new


def ev(e, env):
    k = e[0]
    if k == "var":
        name = e[1]
        if name == "unflatten":      # result shape comes from the enclosing function's sizes (type inference in Futhark)
            gh, gw = env.lookup("gh"), env.lookup("gw")
            return builtin("unflatten", 1, lambda xs: unflatten(xs, gh, gw))
        return env.lookup(name)
    if k == "lit":
        return e[1]
    if k == "app":
        return apply(ev(e[1], env), ev(e[2], env))
    if k == "field":
        base = ev(e[1], env)
        if isinstance(base, Module):
            return base.names[e[2]]
        if isinstance(base, tuple):
            return base[int(e[2])]
        return base[e[2]]
    if k == "index":
        v = ev(e[1], env)
        for ix in e[2]:
            if isinstance(ix, tuple) and ix[0] == "slice":
                lo = ev(ix[1], env) if ix[1] is not None else 0
                hi = ev(ix[2], env) if ix[2] is not None else len(v)
                v = v[lo:hi]
            else:
                i = ev(ix, env)
                if not 0 <= i < len(v):
                    raise FutharkError(f"index {i} out of bounds for array of {len(v)}")
                v = v[i]
        return v
    if k == "bin":
        op = e[1]
        if op == "&&":
            return ev(e[2], env) and ev(e[3], env)
        if op == "||":
            return ev(e[2], env) or ev(e[3], env)
        a, b = ev(e[2], env), ev(e[3], env)
        if op == "==":
            return a == b
        if op == "!=":
            return a != b
        if op == "<":
            return a < b
        if op == "<=":
            return a <= b
        if op == ">":
            return a > b
        if op == ">=":
            return a >= b
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*":
            return a * b
        if op == "/":
            if isinstance(a, (float, np.floating)) or isinstance(b, (float, np.floating)):
                return F32(a) / F32(b)
            return a // b      # Futhark's integer division rounds towards negative infinity, like Python's
        if op == "%":
            return a % b
        if op == "++":
            if isinstance(a, str) or isinstance(b, str):      # strings are arrays of u8
                def as_str(v):
                    return v if isinstance(v, str) else "".join(chr(c) for c in v)
                return as_str(a) + as_str(b)
            return a + b
        if op == "..<":
            return list(range(a, b))
        raise FutharkError(f"operator {op}")
    if k == "if":
        return ev(e[2], env) if ev(e[1], env) else ev(e[3], env)
    if k == "let":
        v = ev(e[2], env)
        new = Env(env)
        if not bind(e[1], v, new):
            raise FutharkError("let pattern does not match")
        return ev(e[3], new)
    if k == "letfun":
        new = Env(env)
        new.vars[e[1]] = Fn(e[1], e[2], e[3], e[4], new)
        return ev(e[5], new)
    if k == "lam":
        return Fn("<lambda>", (), e[1], e[2], env)
    if k == "tuple":
        # A component that fails is poisoned instead of failing the tuple: the reference computes e.g. a scenario's NAME from
        # (scenario.init dummy_1x1_grid, ..., scenario.name ()) (explorer_scenario_helper.fut:26-31, :43-44) and relies on the
        # compiler to drop the unused components, out-of-bounds edits included.
        out = []
        for x in e[1]:
            try:
                out.append(ev(x, env))
            except FutharkError as exc:
                out.append(Poison(exc))
        return tuple(out)
    if k == "record":
        return {name: ev(x, env) for name, x in e[1]}
    if k == "array":
        return [ev(x, env) for x in e[1]]
    if k == "con":
        return (e[1],) + tuple(ev(x, env) for x in (e[2] if len(e) > 2 else []))
    if k == "not":
        return not ev(e[1], env)
    if k == "neg":
        return -ev(e[1], env)
    if k == "with_field":
        return set_path(ev(e[1], env), e[2], ev(e[3], env))
    if k == "with_idx":
        return set_index(ev(e[1], env), [ev(i, env) for i in e[2]], ev(e[3], env))
    if k == "match":
        v = ev(e[1], env)
        for p, body in e[2]:
            new = Env(env)
            if bind(p, v, new):
                return ev(body, new)
        raise FutharkError(f"no case matches {v!r}")
    if k == "loop_while":
        state = ev(e[2], env)
        while True:
            new = Env(env)
            bind(e[1], state, new)
            if not ev(e[3], new):
                return state
            state = ev(e[4], new)
    if k == "loop_for":
        state = ev(e[2], env)
        for i in range(ev(e[4], env)):
            new = Env(env, {e[3]: i})
            bind(e[1], state, new)
            state = ev(e[5], new)
        return state
    if k == "assert":
        if not ev(e[1], env):
            raise FutharkError("assertion failed")
        return ev(e[2], env)
    raise FutharkError(f"cannot evaluate {k}")


def unflatten(xs, gh, gw):
    if len(xs) != gh * gw:
        raise FutharkError(f"unflatten: {len(xs)} elements into [{gh}][{gw}]")
    return [xs[r * gw:(r + 1) * gw] for r in range(gh)]


def reduce_(f, ne, xs):
    acc = ne
    for x in xs:
        acc = call(f, acc, x)
    return acc


def prelude():
    ident = builtin("id", 1, lambda x: x)
    ints = Module({"i8": ident, "i16": ident, "i32": ident, "i64": ident, "u8": ident, "u32": ident, "u64": ident,
                   "f32": builtin("int.f32", 1, lambda v: int(v)), "f64": builtin("int.f64", 1, lambda v: int(v)),
                   "bool": builtin("bool", 1, lambda b: int(bool(b))),
                   "sum": builtin("sum", 1, lambda xs: sum(xs)),
                   "abs": builtin("abs", 1, abs), "sgn": builtin("sgn", 1, lambda v: (v > 0) - (v < 0)),
                   "max": builtin("max", 2, max), "min": builtin("min", 2, min)})
    return {
        "i8": ints, "i16": ints, "i32": ints, "i64": ints, "u8": ints, "u32": ints, "u64": ints,
        "tabulate_2d": builtin("tabulate_2d", 3, lambda h, w, f: [[call(f, y, x) for x in range(w)] for y in range(h)]),
        "tabulate": builtin("tabulate", 2, lambda n, f: [apply(f, i) for i in range(n)]),
        "map": builtin("map", 2, lambda f, xs: [apply(f, x) for x in xs]),
        "map2": builtin("map2", 3, lambda f, xs, ys: [call(f, x, y) for x, y in zip(xs, ys)]),
        "map3": builtin("map3", 4, lambda f, xs, ys, zs: [call(f, x, y, z) for x, y, z in zip(xs, ys, zs)]),
        "zip": builtin("zip", 2, lambda xs, ys: [(x, y) for x, y in zip(xs, ys)]),
        "unzip": builtin("unzip", 1, lambda ps: ([p[0] for p in ps], [p[1] for p in ps])),
        "filter": builtin("filter", 2, lambda f, xs: [x for x in xs if apply(f, x)]),
        "all": builtin("all", 2, lambda f, xs: all(apply(f, x) for x in xs)),
        "any": builtin("any", 2, lambda f, xs: any(apply(f, x) for x in xs)),
        "reduce": builtin("reduce", 3, reduce_),
        "flatten": builtin("flatten", 1, lambda xss: [x for xs in xss for x in xs]),
        "replicate": builtin("replicate", 2, lambda n, x: [x] * n),
        "length": builtin("length", 1, len),
        "copy": ident,
        "iota": builtin("iota", 1, lambda n: list(range(n))),
        "const": builtin("const", 2, lambda x, _y: x),
        "t32": builtin("t32", 1, lambda v: int(v)),
        "f32": Module({"i32": builtin("f32.i32", 1, F32), "i64": builtin("f32.i64", 1, F32), "f64": builtin("f32.f64", 1, F32)}),
    }


class Program:
    """loads .fut files (imports resolved relative to the importing file); `stubs` maps import paths that are not in the
    reference tree to ready-made name tables"""

    def __init__(self, stubs=None):
        self.stubs, self.cache = stubs or {}, {}

    def load(self, path: Path) -> dict:
        path = Path(path).resolve()
        if path in self.cache:
            return self.cache[path]
        toks = lex(path.read_text())
        decls = Parser(toks, path.name).decls()
        env = Env(None, dict(prelude()))
        exported = {}
        self.cache[path] = exported
        self.run_decls(decls, env, exported, path)
        return exported

    def resolve_import(self, name, here: Path):
        for key, table in self.stubs.items():
            if name.endswith(key):
                return table
        target = (here.parent / (name + ".fut")).resolve()
        return self.load(target)

    def run_decls(self, decls, env, exported, here):
        """Definitions are not recursive and may shadow earlier ones (designer.fut defines `step` twice): every declaration
        opens a new scope on top of the previous ones, and a function closes over the scope BEFORE its own name."""
        for d in decls:
            k = d[0]
            if k == "def":
                _k, name, sizes, params, body, local = d
                val = Fn(name, sizes, params, body, env) if params else ev(body, env)
                env = Env(env, {name: val})
                if not local:
                    exported[name] = val
            elif k == "typedef":
                TYPEDEFS[d[1]] = (d[2], d[3])
            elif k == "import":
                table = self.resolve_import(d[1], here)
                env = Env(env, dict(table))
                if d[2]:                                 # `open import`: re-exported
                    exported.update(table)
            elif k == "open":
                mod = env.lookup(d[1])
                env = Env(env, dict(mod.names))
                exported.update(mod.names)
            elif k == "module":
                _k, name, fparams, mexp, local = d
                if fparams:
                    val = Functor(fparams, mexp, env, here, self)
                else:
                    val = self.eval_module(mexp, env, here)
                env = Env(env, {name: val})
                if not local:
                    exported[name] = val
        return env

    def eval_module(self, mexp, env, here):
        k = mexp[0]
        if k == "mbody":
            inner_env, inner_exp = Env(env), {}
            self.run_decls(mexp[1], inner_env, inner_exp, here)
            return Module(inner_exp)
        if k == "mimport":
            return Module(self.resolve_import(mexp[1], here))
        if k == "mname":
            try:
                m = env.lookup(mexp[1][0])
                for part in mexp[1][1:]:
                    m = m.names[part]
                return m
            except (NameError, KeyError):
                return Module({})                      # a module of a library that is not in the reference tree
        if k == "mapp":
            f = self.eval_module(mexp[1], env, here)
            arg = self.eval_module(mexp[2], env, here)
            if not isinstance(f, Functor):
                return Module({})                      # e.g. uniform_int_distribution i8 rnge (cpprandom): stubbed elsewhere
            return f.apply(arg)
        raise FutharkError(f"module expression {k}")


class Functor:
    def __init__(self, params, body, env, here, program):
        self.params, self.body, self.env, self.here, self.program, self.args = params, body, env, here, program, []

    def apply(self, arg):
        args = self.args + [arg]
        if len(args) < len(self.params):
            f = Functor(self.params, self.body, self.env, self.here, self.program)
            f.args = args
            return f
        env = Env(self.env, dict(zip(self.params, args)))
        return self.program.eval_module(self.body, env, self.here)
