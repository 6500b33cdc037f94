"""jl_interp.py -- executes the reference's Julia source TEXT with numpy.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ and by
tests/golden/make_jl_fixtures.py, never by the product package.

Why it exists.  Julia is not in the image, so the reference cannot run here, and its only golden
vector is stale.  The C oracle (ns3d_oracle.c) is a *re-typed* restatement of the reference's
formulas: a transcription slip shared with the CUDA kernels would pass every test.  This module
removes the re-typing: it tokenises and parses the reference scripts themselves
(scripts/NavierStokes3D_multi_gpu.jl, scripts/NavierStokes3D_gpu.jl) -- the `macro ∇V`, the
`@parallel function` kernels, the `@parallel_indices` kernels, `backtrack!`, `lerp`, `set_bc_*!`,
the parameter block and the time loop of the run functions -- and evaluates that text.  What is
still restated (the single remaining point of trust, ~60 lines below) is
the meaning of names the scripts import from packages that are not in the reference tree:

* ParallelStencil.FiniteDifferences3D: `@all @inn @d_xa @d_ya @d_za @d_xi @d_yi @d_zi @d2_xi
  @d2_yi @d2_zi`, the per-statement `@within` guards of `@parallel function`, the launch ranges
  of `@parallel f!(args)` (1:max over the array arguments' sizes) and `@parallel ranges f!(args)`,
  `@zeros`;
* ImplicitGlobalGrid (single rank): `init_global_grid`, `nx_g/ny_g/nz_g`, `x_g/y_g/z_g`,
  `update_halo!` (no-op on one rank);
* Base Julia: operator precedence and associativity, `2μ` juxtaposition, `x^2 -> x*x`
  (`Base.literal_pow`), `%` = `rem` (C `fmod`), `floor(Int,x)`, `clamp`, `LinRange` (`lerpi`),
  NaN-propagating `maximum`, short-circuit `&&`/`||`, IEEE double arithmetic without contraction
  (numpy element-wise `+ - * /` are exactly that).

Execution model.  `@parallel function` kernels run one statement at a time over the statement's
guarded index box with numpy slices; `@parallel_indices` kernels (and the `@inline` functions they
call) run SIMT-style: every thread variable is a vector over the launch box ("lanes"), an `if`
executes its body on the compressed set of lanes whose condition holds.  Both are equivalent to
the reference's per-thread sequential execution because no kernel of the scripts reads a location
that another thread of the same launch writes.  Every array access is bounds-checked, so the
fixtures also certify that the kernels never index out of bounds for the tested shapes.
"""
from __future__ import annotations

import math
import os
import re
import threading

import numpy as np

# ------------------------------------------------------------------------------------------------
# tokenizer
# ------------------------------------------------------------------------------------------------
_OPS3 = ["...", "===", "!=="]
_OPS2 = ["<:", ".=", ".*", "./", ".+", ".-", "==", "!=", "<=", ">=", "&&", "||", "+=", "-=", "*=", "/=", "=>", "::", "->"]
_BLOCK_OPEN = {"function", "macro", "if", "for", "while", "begin", "let", "struct", "try", "quote", "module", "do"}


class Tok:
    __slots__ = ("kind", "val", "start", "end", "line")

    def __init__(self, kind, val, start, end, line):
        self.kind, self.val, self.start, self.end, self.line = kind, val, start, end, line

    def __repr__(self):
        return f"{self.kind}:{self.val!r}@{self.line}"


def _is_id_start(c: str) -> bool:
    return c.isalpha() or c == "_" or (ord(c) > 127 and not c.isspace() and c not in "≈≠≤≥÷∈")


def _is_id_char(c: str) -> bool:
    return c.isalnum() or c == "_" or (ord(c) > 127 and not c.isspace() and c not in "≈≠≤≥÷∈")


def _skip_string(s: str, i: int) -> int:
    """s[i] == '"'; returns the index just behind the closing quote (handles \"\"\", escapes, $(...))."""
    if s.startswith('"""', i):
        j = s.index('"""', i + 3)
        return j + 3
    j = i + 1
    while True:
        c = s[j]
        if c == "\\":
            j += 2
        elif c == '"':
            return j + 1
        elif c == "$" and s[j + 1] == "(":
            depth, j = 1, j + 2
            while depth:
                if s[j] == '"':
                    j = _skip_string(s, j)
                    continue
                depth += (s[j] == "(") - (s[j] == ")")
                j += 1
        else:
            j += 1


def tokenize(s: str, cont_ops=()) -> list:
    """`cont_ops`: operators after which a line break does not end the statement (`a =\n b`, `f(x),\n g(y)`)."""
    toks, i, line, depth, n = [], 0, 1, 0, len(s)
    while i < n:
        c = s[i]
        if c == "\n":
            if depth == 0 and not (toks and toks[-1].kind == "op" and toks[-1].val in cont_ops):
                toks.append(Tok("nl", "\n", i, i + 1, line))
            line += 1
            i += 1
        elif c.isspace():
            i += 1
        elif c == "#":
            if s.startswith("#=", i):
                j = s.index("=#", i) + 2
                line += s.count("\n", i, j)
                i = j
            else:
                while i < n and s[i] != "\n":
                    i += 1
        elif c == '"':
            j = _skip_string(s, i)
            toks.append(Tok("str", s[i:j], i, j, line))
            line += s.count("\n", i, j)
            i = j
        elif c.isdigit() or (c == "." and i + 1 < n and s[i + 1].isdigit()):
            m = re.compile(r"\d*\.?\d*(?:[eE][+-]?\d+)?").match(s, i)
            j = m.end()
            text = s[i:j]
            # "1:nt" / "2:end": a trailing '.' belongs to the number only if no identifier follows ("1.0")
            val = float(text) if any(ch in text for ch in ".eE") else int(text)
            toks.append(Tok("num", val, i, j, line))
            i = j
        elif c == "@":
            j = i + 1
            while j < n and (_is_id_char(s[j]) or (s[j] == "!" and s[j + 1:j + 2] != "=")):
                j += 1
            toks.append(Tok("macro", s[i + 1:j], i, j, line))
            i = j
        elif _is_id_start(c):
            j = i + 1
            while j < n and (_is_id_char(s[j]) or (s[j] == "!" and s[j + 1:j + 2] != "=")):
                j += 1
            toks.append(Tok("id", s[i:j], i, j, line))
            i = j
        else:
            op = next((o for o in _OPS3 + _OPS2 if s.startswith(o, i)), c)
            if op in "([{":
                depth += 1
            elif op in ")]}":
                depth -= 1
            toks.append(Tok("op", op, i, i + len(op), line))
            i += len(op)
    toks.append(Tok("nl", "\n", n, n, line))
    toks.append(Tok("eof", None, n, n, line))
    return toks


# ------------------------------------------------------------------------------------------------
# parser: token list -> nested tuples
# ------------------------------------------------------------------------------------------------
class ParseError(Exception):
    pass


class Parser:
    def __init__(self, toks):
        self.t, self.i, self.in_index = toks, 0, 0

    # -- helpers
    @property
    def cur(self):
        return self.t[self.i]

    def peek(self, k=1):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def is_op(self, v):
        return self.cur.kind == "op" and self.cur.val == v

    def is_kw(self, v):
        return self.cur.kind == "id" and self.cur.val == v

    def eat(self, v):
        if self.cur.val != v or self.cur.kind not in ("op", "id"):
            raise ParseError(f"expected {v!r}, got {self.cur!r}")
        self.i += 1

    def skip_seps(self):
        while self.cur.kind == "nl" or self.is_op(";"):
            self.i += 1

    def at_stmt_end(self):
        return self.cur.kind in ("nl", "eof") or self.is_op(";") or self.cur.val in ("end", "else", "elseif")

    # -- statements
    def block(self):
        out = []
        while True:
            self.skip_seps()
            if self.cur.kind == "eof" or (self.cur.kind == "id" and self.cur.val in ("end", "else", "elseif")):
                return out
            out.append(self.statement())

    def statement(self):
        c = self.cur
        if c.kind == "id":
            if c.val == "if":
                return self.if_stmt()
            if c.val == "for":
                self.i += 1
                var = self.cur.val
                self.i += 1
                self.eat("=")
                rng = self.expr()
                body = self.block()
                self.eat("end")
                return ("for", var, rng, body, c.line)
            if c.val == "return":
                self.i += 1
                if self.at_stmt_end():
                    return ("return", None, c.line)
                vals = self.expr_list()
                return ("return", vals[0] if len(vals) == 1 else ("tuple", vals), c.line)
            if c.val == "break":
                self.i += 1
                return ("break", c.line)
        if c.kind == "macro" and c.val == "parallel" and self.peek().start > c.end:
            self.i += 1
            first = self.expr()
            if self.at_stmt_end():
                return ("parallel", None, first, c.line)
            return ("parallel", first, self.expr(), c.line)
        lhs = self.expr_list()
        if self.cur.kind == "op" and self.cur.val in ("=", ".=", "+=", "-=", "*=", "/="):
            op = self.cur.val
            self.i += 1
            rhs = self.expr_list()
            return ("assign", lhs, op, rhs, c.line)
        if len(lhs) != 1:
            raise ParseError(f"tuple expression statement at line {c.line}")
        return ("expr", lhs[0], c.line)

    def if_stmt(self):
        line = self.cur.line
        self.i += 1   # if / elseif
        cond = self.expr()
        body = self.block()
        orelse = []
        if self.is_kw("elseif"):
            orelse = [self.if_stmt()]
            return ("if", cond, body, orelse, line)   # the nested if_stmt consumed the shared 'end'
        if self.is_kw("else"):
            self.i += 1
            orelse = self.block()
        self.eat("end")
        return ("if", cond, body, orelse, line)

    def expr_list(self):
        out = [self.expr()]
        while self.is_op(","):
            self.i += 1
            out.append(self.expr())
        return out

    # -- expressions (Julia precedence, lowest first)
    def expr(self):
        a = self.p_or()
        if self.is_op("=>"):                 # a pair, as in Dict("Pr"=>Array(Pr))
            self.i += 1
            return ("tuple", [a, self.expr()])
        return a

    def p_or(self):
        a = self.p_and()
        while self.is_op("||"):
            self.i += 1
            a = ("or", a, self.p_and())
        return a

    def p_and(self):
        a = self.p_cmp()
        while self.is_op("&&"):
            self.i += 1
            a = ("and", a, self.p_cmp())
        return a

    def p_cmp(self):
        a = self.p_range()
        while self.cur.kind == "op" and self.cur.val in ("==", "!=", "<", "<=", ">", ">="):
            op = self.cur.val
            self.i += 1
            a = ("bin", op, a, self.p_range())
        return a

    def p_range(self):
        a = self.p_add()
        if self.is_op(":") and not (self.in_index and self.peek().val in (",", "]")):
            self.i += 1
            a = ("range", a, self.p_add())
        return a

    def p_add(self):
        a = self.p_mul()
        while self.cur.kind == "op" and self.cur.val in ("+", "-"):
            op = self.cur.val
            self.i += 1
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_unary()
        while self.cur.kind == "op" and self.cur.val in ("*", "/", "%"):
            op = self.cur.val
            self.i += 1
            a = ("bin", op, a, self.p_unary())
        return a

    def p_unary(self):
        if self.cur.kind == "op" and self.cur.val in ("-", "+", "!"):
            op = self.cur.val
            self.i += 1
            a = self.p_unary()
            return a if op == "+" else (("neg", a) if op == "-" else ("not", a))
        return self.p_pow()

    def p_pow(self):
        a = self.p_postfix()
        if self.is_op("^"):
            self.i += 1
            a = ("pow", a, self.p_unary())
        return a

    def p_postfix(self):
        tok = self.cur
        a = self.p_primary()
        if tok.kind == "num" and self.cur.start == tok.end and (self.cur.kind == "id" or self.is_op("(")):
            # numeric-literal coefficient: 2μ == (2*μ), binds tighter than * and /
            return ("bin", "*", a, self.p_pow())
        while True:
            prev_end = self.t[self.i - 1].end
            if self.is_op("(") and self.cur.start == prev_end:
                a = ("call", a, *self.call_args(), False)
            elif self.is_op("[") and self.cur.start == prev_end:
                self.i += 1
                self.in_index += 1
                idx = []
                while not self.is_op("]"):
                    if self.is_op(":") and self.peek().val in (",", "]"):
                        self.i += 1
                        idx.append(("colon",))
                    else:
                        idx.append(self.expr())
                    if self.is_op(","):
                        self.i += 1
                self.in_index -= 1
                self.eat("]")
                a = ("index", a, idx)
            elif self.is_op(".") and self.peek().kind == "op" and self.peek().val == "(":
                self.i += 1
                a = ("call", a, *self.call_args(), True)
            elif self.is_op(".") and self.peek().kind == "id":
                self.i += 1
                a = ("attr", a, self.cur.val)
                self.i += 1
            elif self.is_op("'") and self.cur.start == prev_end:
                self.i += 1
                a = ("transpose", a)
            else:
                return a

    def call_args(self):
        self.eat("(")
        save, self.in_index = self.in_index, 0
        args, kwargs = [], {}
        while not self.is_op(")"):
            if self.is_op(";"):
                self.i += 1
                continue
            if self.cur.kind == "id" and self.peek().kind == "op" and self.peek().val == "=":
                name = self.cur.val
                self.i += 2
                kwargs[name] = self.expr()
            else:
                args.append(self.expr())
            if self.is_op(","):
                self.i += 1
        self.eat(")")
        self.in_index = save
        return args, kwargs

    def paren_item(self):
        line = self.cur.line
        e = self.expr()
        if self.is_op("="):
            self.i += 1
            return ("assign", [e], "=", [self.expr()], line)
        return ("expr", e, line)

    def p_primary(self):
        c = self.cur
        if c.kind == "num":
            self.i += 1
            return ("num", c.val)
        if c.kind == "str":
            self.i += 1
            return ("str", c.val)
        if c.kind == "id":
            self.i += 1
            if c.val == "end" and self.in_index:
                return ("endidx",)
            if c.val in ("true", "false"):
                return ("num", c.val == "true")
            return ("id", c.val)
        if c.kind == "macro":
            self.i += 1
            if self.is_op("(") and self.cur.start == c.end:
                args, kwargs = self.call_args()
                return ("macrocall", c.val, args)
            raise ParseError(f"macro @{c.val} without parentheses in an expression (line {c.line})")
        if self.is_op("("):
            self.i += 1
            save, self.in_index = self.in_index, 0
            items = [self.paren_item()]
            kind = "paren"
            while self.is_op(",") or self.is_op(";"):
                kind = "tuple" if self.is_op(",") else "blockexpr"
                self.i += 1
                if self.is_op(")"):
                    break
                items.append(self.paren_item())
            self.eat(")")
            self.in_index = save
            if kind == "blockexpr":
                return ("blockexpr", items)
            vals = []
            for st in items:
                if st[0] != "expr":
                    raise ParseError(f"assignment inside a tuple (line {c.line})")
                vals.append(st[1])
            return vals[0] if kind == "paren" else ("tuple", vals)
        if self.is_op("["):
            self.i += 1
            save, self.in_index = self.in_index, 0
            if self.is_op("]"):
                self.i += 1
                self.in_index = save
                return ("vect", [])
            first = self.expr()
            if self.is_kw("for"):
                self.i += 1
                gens = []
                while True:
                    var = self.cur.val
                    self.i += 1
                    self.eat("=")
                    gens.append((var, self.expr()))
                    if self.is_op(","):
                        self.i += 1
                        continue
                    break
                self.eat("]")
                self.in_index = save
                return ("comprehension", first, gens)
            items = [first]
            while self.is_op(","):
                self.i += 1
                items.append(self.expr())
            self.eat("]")
            self.in_index = save
            return ("vect", items)
        if self.is_op(":") and self.peek().kind == "op" and self.peek().val == "(":
            self.i += 1      # quote  :( expr )
            return ("quote", self.p_primary())
        if self.is_op(":") and self.peek().kind == "id" and self.peek().start == c.end:
            self.i += 2      # a symbol  :log10
            return ("sym", self.t[self.i - 1].val)
        raise ParseError(f"unexpected token {c!r}")


# ------------------------------------------------------------------------------------------------
# values
# ------------------------------------------------------------------------------------------------
class L:
    """A per-thread ("lane") value of an @parallel_indices launch or of a comprehension."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v


def _raw(x):
    return x.v if isinstance(x, L) else x


def _wrap(res, *ops):
    return L(res) if any(isinstance(o, L) for o in ops) else res


class JlType:
    def __init__(self, name):
        self.name = name


class JlError(Exception):
    pass


class _Break(Exception):
    pass


class _Return(Exception):
    def __init__(self, val):
        self.val = val


class Def:
    def __init__(self, kind, name, params, body, indices=None, lines=(0, 0)):
        self.kind, self.name, self.params, self.body, self.indices, self.lines = kind, name, params, body, indices, lines


def lin_range(start, stop, n):
    """Julia LinRange(start,stop,n)[i] = lerpi(i-1, max(n-1,1), start, stop) = (1-t)*start + t*stop, t = j/d."""
    d = max(n - 1, 1)
    t = np.arange(n, dtype=np.float64) / d
    return (1 - t) * start + t * stop


# ------------------------------------------------------------------------------------------------
# ImplicitGlobalGrid on several ranks: one thread per rank, the collectives are rendezvous points
# ------------------------------------------------------------------------------------------------
class Comm:
    """A Cartesian process grid `dims` (MPI order: the last dimension varies fastest).  Every rank is a thread
    interpreting its own copy of the script; `update_halo!`, `MPI.Allreduce` and `gather!` meet here."""

    def __init__(self, dims):
        self.dims = tuple(int(d) for d in dims)
        self.coords = [(cx, cy, cz) for cx in range(self.dims[0]) for cy in range(self.dims[1]) for cz in range(self.dims[2])]
        self.size = len(self.coords)
        self.barrier = threading.Barrier(self.size)
        self.slots = [None] * self.size

    def rank_of(self, c):
        return self.coords.index(tuple(c))

    def exchange(self, rank, payload):
        """All-gather of one Python object per rank."""
        self.slots[rank] = payload
        self.barrier.wait(timeout=120)
        got = list(self.slots)
        self.barrier.wait(timeout=120)
        return got


# ------------------------------------------------------------------------------------------------
# the script: definitions + evaluator
# ------------------------------------------------------------------------------------------------
class JuliaScript:
    """One reference script: its definitions parsed from the text, host statements runnable by line range."""
    parser_cls = Parser
    cont_ops = ()

    def __init__(self, text: str, name: str = "script"):
        self.text, self.name = text, name
        self.toks = tokenize(text, self.cont_ops)
        self.defs: dict[str, Def] = {}
        self.macros: dict[str, tuple] = {}
        self.grid = None            # (nx, ny, nz) after init_global_grid (single rank)
        self.launches = []          # (kernel name, ranges) in launch order, for inspection by tests
        self.frozen: dict = {}      # host names whose assignments in the text are ignored (nx = 255 -> test size)
        self.comm, self.rank = None, 0   # set by run_ranks: several ranks of an ImplicitGlobalGrid
        self.matwrites = []              # (file name, Dict) of every `matwrite` call (MAT.jl is not restated: recorded)
        self._builtins = self._make_builtins()
        self._scan_definitions()

    @classmethod
    def from_file(cls, path: str):
        with open(path, encoding="utf-8") as fh:
            return cls(fh.read(), name=path)

    # -- top level: find definitions -------------------------------------------------------------
    def _top_statements(self):
        """Token spans of the top-level statements (block depth 0)."""
        spans, start, depth, br = [], 0, 0, 0
        for k, t in enumerate(self.toks):
            if t.kind == "op":
                br += (t.val in "([{") - (t.val in ")]}") if len(t.val) == 1 else 0
            elif t.kind == "id" and br == 0:
                if t.val in _BLOCK_OPEN:
                    depth += 1
                elif t.val == "end":
                    depth -= 1
            elif t.kind in ("nl", "eof") and depth == 0 and br == 0:
                if k > start:
                    spans.append((start, k))
                start = k + 1
        return spans

    def _scan_definitions(self):
        for a, b in self._top_statements():
            ts = self.toks[a:b]
            vals = [t.val for t in ts]
            if ts[0].kind == "id" and ts[0].val == "macro":
                p = self.parser_cls(ts[1:] + [Tok("eof", None, 0, 0, ts[-1].line)])
                name = p.cur.val
                p.i += 1
                p.call_args()
                body = p.block()
                # macro NAME() esc(:( expr )) end   ->   the quoted expression
                st = body[0]
                assert st[0] == "expr" and st[1][0] == "call" and st[1][1] == ("id", "esc"), st
                q = st[1][2][0]
                assert q[0] == "quote"
                self.macros[name] = q[1]
                continue
            if "function" in vals and all(t.kind == "macro" or t.val in ("(", ")", ",") or t.kind == "id"
                                          for t in ts[:vals.index("function")]):
                f = vals.index("function")
                deco = ts[:f]
                kind, indices = "function", None
                if deco and deco[0].kind == "macro" and deco[0].val == "parallel":
                    kind = "ps_kernel"
                elif deco and deco[0].kind == "macro" and deco[0].val == "parallel_indices":
                    kind, indices = "pi_kernel", [t.val for t in deco[1:] if t.kind == "id"]
                p = self.parser_cls(ts[f + 1:] + [Tok("eof", None, 0, 0, ts[-1].line)])
                name = p.cur.val
                p.i += 1
                if any(t.kind == "op" and t.val == ";" for t in ts[f + 2:f + 4]):
                    continue   # keyword-only signature (the run functions): executed by line range instead
                try:
                    args, _ = p.call_args()
                    body = p.block()
                except ParseError:
                    continue   # run functions with plotting syntax are not parsed as a whole
                params = [x[1] for x in args]
                self.defs[name] = Def(kind, name, params, body, indices, (ts[0].line, ts[-1].line))
                continue
            # short form  [@inline] name(args) = expr
            k = 1 if ts[0].kind == "macro" and ts[0].val == "inline" else 0
            if len(ts) > k + 3 and ts[k].kind == "id" and ts[k + 1].val == "(" and ts[k + 1].start == ts[k].end:
                p = self.parser_cls(ts[k:] + [Tok("eof", None, 0, 0, ts[-1].line)])
                name = p.cur.val
                p.i += 1
                try:
                    args, kw = p.call_args()
                    if not p.is_op("=") or kw or not all(x[0] == "id" for x in args):
                        continue
                    p.i += 1
                    body = p.expr()
                except ParseError:
                    continue
                self.defs[name] = Def("short", name, [x[1] for x in args], body, None, (ts[0].line, ts[-1].line))

    # -- host statements by line range -----------------------------------------------------------
    def parse_lines(self, first: int, last: int, close_blocks: int = 0):
        ts = [t for t in self.toks if first <= t.line <= last and t.kind != "eof"]
        line = ts[-1].line
        for _ in range(close_blocks):
            ts += [Tok("nl", "\n", 0, 0, line), Tok("id", "end", 0, 0, line)]
        ts += [Tok("nl", "\n", 0, 0, line), Tok("eof", None, 0, 0, line)]
        p = self.parser_cls(ts)
        body = p.block()
        if p.cur.kind != "eof":
            raise ParseError(f"unbalanced block in lines {first}-{last}: {p.cur!r}")
        return body

    def run_lines(self, first: int, last: int, env: dict, close_blocks: int = 0):
        body = self.parse_lines(first, last, close_blocks)
        self.exec_block(body, env, host=True)
        return env

    def find_line(self, pattern: str, after: int = 0) -> int:
        rx = re.compile(pattern)
        for n, ln in enumerate(self.text.split("\n"), 1):
            if n > after and rx.search(ln):
                return n
        raise JlError(f"no line matches {pattern!r} after line {after}")

    # -- builtins ----------------------------------------------------------------------------------
    @property
    def dims(self):
        return self.comm.dims if self.comm else (1, 1, 1)

    @property
    def coords(self):
        return self.comm.coords[self.rank] if self.comm else (0, 0, 0)

    def _n_g(self, d):
        return self.dims[d] * (self.grid[d] - 2) + 2            # nx_g() = dims*(nx-overlap)+overlap, overlap 2

    def _x_g(self, d, i, dx, A):
        n = self.grid[d]
        x0 = 0.5 * (n - A.shape[d]) * dx
        res = (self.coords[d] * (n - 2) + _raw(i) - 1) * dx + x0
        return _wrap(res, i)

    def _update_halo(self, *fields):
        """update_halo!(A...): dimension by dimension; plane ol / s-ol+1 (1-based, ol = overlap + size(A,d) - n)
        goes to the lower / upper neighbour's last / first plane.  Nothing to do on a single rank."""
        if self.comm is None:
            return None
        c = self.coords
        for d in range(3):
            if self.dims[d] == 1:
                continue
            sends = []
            for A in fields:
                sz = A.shape[d]
                ol = 2 + (sz - self.grid[d])
                if ol < 2:
                    raise JlError(f"update_halo!: a field of size {A.shape} has no halo in dimension {d + 1}")
                sends.append((np.take(A, ol - 1, axis=d).copy(), np.take(A, sz - ol, axis=d).copy()))
            got = self.comm.exchange(self.rank, sends)
            for k, A in enumerate(fields):
                idx = [slice(None)] * 3
                if c[d] > 0:
                    lo = list(c)
                    lo[d] -= 1
                    idx[d] = 0
                    A[tuple(idx)] = got[self.comm.rank_of(lo)][k][1]
                if c[d] < self.dims[d] - 1:
                    hi = list(c)
                    hi[d] += 1
                    idx[d] = A.shape[d] - 1
                    A[tuple(idx)] = got[self.comm.rank_of(hi)][k][0]
        return None

    def _allreduce(self, x, op, comm):
        if self.comm is None:
            return x
        vals = self.comm.exchange(self.rank, x)
        if op != "max":
            raise JlError(f"MPI.Allreduce with {op}")
        return math.nan if any(math.isnan(v) for v in vals) else max(vals)

    def _gather(self, A, A_global, **kw):
        """gather!(A, A_global): block (cx,cy,cz) of the root's A_global is rank (cx,cy,cz)'s A."""
        if self.comm is None:
            if A_global.shape != A.shape:
                raise JlError(f"gather!: size(A_global) {A_global.shape} != dims .* size(A) {A.shape}")
            A_global[...] = A
            return None
        got = self.comm.exchange(self.rank, A.copy())
        if self.rank == 0:
            want = tuple(self.dims[d] * A.shape[d] for d in range(3))
            if A_global.shape != want:
                raise JlError(f"gather!: size(A_global) {A_global.shape} != dims .* size(A) {want}")
            for r, c in enumerate(self.comm.coords):
                A_global[tuple(slice(c[d] * A.shape[d], (c[d] + 1) * A.shape[d]) for d in range(3))] = got[r]
        return None

    def builtin(self, name):
        if name not in self._builtins:
            raise JlError(f"unknown name {name!r}")
        return self._builtins[name]

    def _make_builtins(self):
        return {
            "size": lambda A, d=None: A.shape if d is None else int(A.shape[d - 1]),
            "clamp": lambda x, lo, hi: _wrap(np.clip(_raw(x), lo, hi) if isinstance(x, L) else min(max(x, lo), hi), x),
            "floor": self._floor, "ceil": self._ceil,
            "min": lambda *a: min(a), "max": lambda *a: max(a),
            "sqrt": lambda x: math.sqrt(x),
            "sincos": lambda x: (math.sin(x), math.cos(x)),
            "abs": lambda x: _wrap(np.abs(_raw(x)), x),
            "maximum": lambda A: float(np.max(A)) if A.size else (_ for _ in ()).throw(JlError("maximum of empty")),
            "isfinite": lambda x: math.isfinite(x),
            "push!": lambda lst, v: lst.append(v),
            "println": lambda *a: None, "print": lambda *a: None,
            "update_halo!": self._update_halo,
            "gather!": self._gather,
            "checkbounds": self._checkbounds,
            "init_global_grid": self._init_global_grid,
            "finalize_global_grid": lambda: None,
            "nx_g": lambda: self._n_g(0), "ny_g": lambda: self._n_g(1), "nz_g": lambda: self._n_g(2),
            "x_g": lambda i, dx, A: self._x_g(0, i, dx, A),
            "y_g": lambda i, dx, A: self._x_g(1, i, dx, A),
            "z_g": lambda i, dx, A: self._x_g(2, i, dx, A),
            "LinRange": lambda a, b, n: lin_range(a, b, n),
            "Array": lambda x: np.array(x, order="F"),
            "zeros": lambda *s: np.zeros(s, dtype=np.float64, order="F"),
            "Int": JlType("Int"), "Bool": JlType("Bool"), "Float64": JlType("Float64"), "Float32": JlType("Float32"),
            "Inf": math.inf, "π": math.pi, "pi": math.pi,
            "Data": {"Array": lambda x: np.array(x, dtype=np.float64, order="F"), "Number": JlType("Float64")},
            "MPI": {"Allreduce": self._allreduce, "MAX": "max", "COMM_WORLD": "world"},
            "nothing": None,
            # the save path (M:27-30, 404-413, 515-523)
            "ispath": os.path.exists, "isdir": os.path.isdir, "mkdir": os.mkdir,
            "string": lambda *a: "".join(str(x) for x in a),
            "open": lambda name, mode="r": open(name, {"w": "wb", "r": "rb"}[mode]),
            "write": lambda out, A: out.write(np.asfortranarray(A).tobytes(order="F")),
            "close": lambda out: out.close(),
            "convert": lambda T, A: np.asarray(A).astype({"Float32": np.float32, "Float64": np.float64}[T.name]),
            "Dict": lambda *pairs: {k: v for k, v in pairs},      # a repeated key keeps its LAST pair, as in Julia (G:89)
            "matwrite": lambda fname, d: self.matwrites.append((fname, d)),
        }

    @staticmethod
    def _floor(*a):
        if len(a) == 2:                                  # floor(Int, x): InexactError outside Int64, like Julia
            x = a[1]
            r = np.floor(np.asarray(_raw(x), dtype=np.float64))
            if not (np.isfinite(r).all() and (np.abs(r) < 9.0e18).all()):
                raise JlError("InexactError: floor(Int, x) of a non-finite or out-of-range value (the run has diverged)")
            return _wrap(r.astype(np.int64) if isinstance(x, L) else int(r), x)
        return _wrap(np.floor(_raw(a[0])), a[0])

    @staticmethod
    def _ceil(*a):
        if len(a) == 2:
            return math.ceil(a[1])
        return float(math.ceil(a[0]))

    @staticmethod
    def _checkbounds(_bool, A, *idx):
        ok = True
        for d, i in enumerate(idx):
            r = _raw(i)
            ok = ok & (r >= 1) & (r <= A.shape[d])
        return _wrap(ok, *idx)

    def _init_global_grid(self, nx, ny, nz, **kw):
        self.grid = (nx, ny, nz)
        return (self.rank, self.dims)

    # -- evaluation -----------------------------------------------------------------------------
    def lookup(self, name, env):
        if name in env:
            return env[name]
        if name in self.defs:
            return self.defs[name]
        return self.builtin(name)

    def ev(self, e, env, ps=None):
        """Evaluate expression e.  ps = (n1,n2,n3) inside an @parallel-function statement (macros -> slices)."""
        k = e[0]
        if k == "num":
            return e[1]
        if k == "id":
            return self.lookup(e[1], env)
        if k == "str":
            text = e[1][3:-3] if e[1].startswith('"""') else e[1][1:-1]
            if "$" in text:                                 # "$name" interpolation ("$(expr)" occurs in plot titles only)
                text = re.sub(r"\$([^\W\d]\w*)", lambda m: str(self.lookup(m.group(1), env)), text)
            return text
        if k == "sym":
            return ("sym", e[1])
        if k == "transpose":
            return np.transpose(self.ev(e[1], env, ps))
        if k == "endidx":                               # `end` inside an index: size of that dimension
            return env["__end__"]
        if k == "bin":
            return self.binop(e[1], self.ev(e[2], env, ps), self.ev(e[3], env, ps))
        if k == "neg":
            a = self.ev(e[1], env, ps)
            return _wrap(-_raw(a), a)
        if k == "not":
            a = self.ev(e[1], env, ps)
            return _wrap(np.logical_not(_raw(a)) if isinstance(a, L) else (not a), a)
        if k == "pow":
            a = self.ev(e[1], env, ps)
            if e[2] == ("num", 2):                     # Base.literal_pow: x^2 -> x*x
                return _wrap(_raw(a) * _raw(a), a)
            if e[2] == ("num", 3):
                return _wrap(_raw(a) * _raw(a) * _raw(a), a)
            b = self.ev(e[2], env, ps)
            return _wrap(np.power(np.float64(_raw(a)) if not isinstance(_raw(a), np.ndarray) else _raw(a), _raw(b)), a, b)
        if k in ("and", "or"):
            return self.shortcircuit(k, e[1], e[2], env, ps)
        if k == "call":
            return self.call(e, env, ps)
        if k == "macrocall":
            return self.macrocall(e[1], e[2], env, ps)
        if k == "index":
            return self.index_load(e, env, ps)
        if k == "attr":
            return self.ev(e[1], env, ps)[e[2]]
        if k == "tuple":
            return tuple(self.ev(x, env, ps) for x in e[1])
        if k == "range":
            return ("range", self.ev(e[1], env, ps), self.ev(e[2], env, ps))
        if k == "vect":
            return [self.ev(x, env, ps) for x in e[1]]
        if k == "comprehension":
            return self.comprehension(e, env)
        if k == "blockexpr":
            val = None
            for st in e[1]:
                val = self.exec_stmt(st, env, host=False, want_value=True)
            return val
        raise JlError(f"cannot evaluate {k}")

    @staticmethod
    def binop(op, a, b):
        x, y = _raw(a), _raw(b)
        if op == "+":
            r = x + y
        elif op == "-":
            r = x - y
        elif op == "*":
            r = x * y
        elif op == "/":
            if isinstance(x, np.ndarray) or isinstance(y, np.ndarray):
                r = np.true_divide(x, y)
            else:
                if y == 0:
                    r = math.copysign(math.inf, x) * math.copysign(1.0, y) if x != 0 else math.nan
                else:
                    r = x / y
        elif op == "%":
            if isinstance(x, np.ndarray) or isinstance(y, np.ndarray):
                r = np.fmod(x, y)
            elif isinstance(x, int) and isinstance(y, int):
                r = int(math.fmod(x, y))
            else:
                r = math.fmod(x, y)
        elif op == "==":
            r = x == y
        elif op == "!=":
            r = x != y
        elif op == "<":
            r = x < y
        elif op == "<=":
            r = x <= y
        elif op == ">":
            r = x > y
        elif op == ">=":
            r = x >= y
        else:
            raise JlError(f"operator {op}")
        return _wrap(r, a, b)

    def shortcircuit(self, k, ea, eb, env, ps):
        a = self.ev(ea, env, ps)
        if not isinstance(a, L):
            if k == "and":
                return self.ev(eb, env, ps) if a else False
            return True if a else self.ev(eb, env, ps)
        need = a.v if k == "and" else ~a.v          # lanes on which the right operand is evaluated at all
        out = a.v.copy()
        if need.any():
            sub = {n: (L(v.v[need]) if isinstance(v, L) else v) for n, v in env.items()}
            b = self.ev(eb, sub, ps)
            out[need] = _raw(b)
        return L(out)

    def call(self, e, env, ps):
        fn = self.ev(e[1], env, ps)
        args = [self.ev(x, env, ps) for x in e[2]]
        kwargs = {n: self.ev(x, env, ps) for n, x in e[3].items()}
        if isinstance(fn, Def):
            return self.call_def(fn, args)
        if isinstance(fn, JlType):                   # Float64[] handled in index_load; Int(x) not used
            raise JlError(f"type call {fn.name}")
        return fn(*args, **kwargs)

    def call_def(self, d: Def, args):
        if len(args) != len(d.params):
            raise JlError(f"{d.name}: {len(args)} arguments for {len(d.params)} parameters")
        env = dict(zip(d.params, args))
        if d.kind == "short":
            return self.ev(d.body, env)
        if d.kind in ("ps_kernel", "pi_kernel"):
            raise JlError(f"{d.name} is a kernel: launch it with @parallel")
        try:
            self.exec_block(d.body, env, host=False)
        except _Return as r:
            return r.val
        return None

    # -- ParallelStencil.FiniteDifferences3D, restated ---------------------
    def macrocall(self, name, args, env, ps):
        if name in self.macros:
            return self.ev(self.macros[name], env, ps)
        if name == "zeros":
            return np.zeros(tuple(self.ev(a, env) for a in args), dtype=np.float64, order="F")
        if name == "sprintf":                               # C-style formats: the same in Julia's Printf and in Python
            return self.ev(args[0], env) % tuple(self.ev(a, env) for a in args[1:])
        if name in ("printf", "show"):
            return None
        if ps is None:
            raise JlError(f"@{name} outside an @parallel function")
        A = self.ev(args[0], env)
        n = ps

        def sl(ox, oy, oz):
            s = A[ox:ox + n[0], oy:oy + n[1], oz:oz + n[2]]
            if s.shape != tuple(n):
                raise JlError(f"@{name}({args[0][1]}): out of bounds for statement box {n}, array {A.shape}")
            return s
        if name == "all":
            return sl(0, 0, 0)
        if name == "inn":
            return sl(1, 1, 1)
        if name == "d_xa":
            return sl(1, 0, 0) - sl(0, 0, 0)
        if name == "d_ya":
            return sl(0, 1, 0) - sl(0, 0, 0)
        if name == "d_za":
            return sl(0, 0, 1) - sl(0, 0, 0)
        if name == "d_xi":
            return sl(1, 1, 1) - sl(0, 1, 1)
        if name == "d_yi":
            return sl(1, 1, 1) - sl(1, 0, 1)
        if name == "d_zi":
            return sl(1, 1, 1) - sl(1, 1, 0)
        if name == "d2_xi":
            return (sl(2, 1, 1) - sl(1, 1, 1)) - (sl(1, 1, 1) - sl(0, 1, 1))
        if name == "d2_yi":
            return (sl(1, 2, 1) - sl(1, 1, 1)) - (sl(1, 1, 1) - sl(1, 0, 1))
        if name == "d2_zi":
            return (sl(1, 1, 2) - sl(1, 1, 1)) - (sl(1, 1, 1) - sl(1, 1, 0))
        raise JlError(f"unknown macro @{name}")

    def launch(self, d: Def, args, ranges=None):
        """`@parallel [ranges] f!(args...)`."""
        arrays = [a for a in args if isinstance(a, np.ndarray) and a.ndim == 3]
        self.launches.append((d.name, ranges))
        env = dict(zip(d.params, args))
        if len(args) != len(d.params):
            raise JlError(f"{d.name}: {len(args)} arguments for {len(d.params)} parameters")
        if d.kind == "ps_kernel":
            if ranges is not None:
                raise JlError("ranges with an @parallel function are not used by the scripts")
            for st in d.body:
                if st[0] == "return":
                    break
                if not (st[0] == "assign" and st[2] == "=" and len(st[1]) == 1 and st[1][0][0] == "macrocall"
                        and st[1][0][1] in ("all", "inn")):
                    raise JlError(f"{d.name}: statement form not supported (line {st[-1]})")
                which = st[1][0][1]
                target = self.ev(st[1][0][2][0], env)
                n = tuple(s - (2 if which == "inn" else 0) for s in target.shape)   # the @within guard
                if min(n) <= 0:
                    continue
                val = self.ev(st[3][0], env, ps=n)
                o = 1 if which == "inn" else 0
                target[o:o + n[0], o:o + n[1], o:o + n[2]] = val
            return
        if d.kind != "pi_kernel":
            raise JlError(f"{d.name} is not a kernel")
        if ranges is None:      # 1:max over the array arguments of size(A, d)
            ranges = tuple(("range", 1, max(a.shape[k] for a in arrays)) for k in range(len(d.indices)))
        if len(ranges) != len(d.indices):
            raise JlError(f"{d.name}: {len(ranges)} ranges for indices {d.indices}")
        axes = [np.arange(r[1], r[2] + 1, dtype=np.int64) for r in ranges]
        grids = np.meshgrid(*axes, indexing="ij")
        for name, g in zip(d.indices, grids):
            env[name] = L(g.ravel(order="F"))
        try:
            self.exec_block(d.body, env, host=False)
        except _Return:
            pass

    # -- indexing -----------------------------------------------------------------------------
    def _np_index(self, A, idx_exprs, env, ps):
        if not isinstance(A, np.ndarray):
            raise JlError("indexing a non-array")
        if len(idx_exprs) != A.ndim:
            raise JlError(f"{len(idx_exprs)} indices for a {A.ndim}-d array")
        out, lanes = [], False
        for d, ie in enumerate(idx_exprs):
            if ie == ("colon",):
                out.append(slice(None))
                continue
            sub = dict(env)
            sub["__end__"] = int(A.shape[d])
            v = self.ev(ie, sub, ps)
            if isinstance(v, tuple) and v and v[0] == "range":
                lo, hi = v[1], v[2]
                if lo < 1 or hi > A.shape[d]:
                    raise JlError(f"range {lo}:{hi} out of bounds for dim {d + 1} of {A.shape}")
                out.append(slice(lo - 1, hi))
                continue
            r = _raw(v)
            if isinstance(r, np.ndarray):
                lanes = True
                if r.size and (r.min() < 1 or r.max() > A.shape[d]):
                    raise JlError(f"index out of bounds in dim {d + 1}: [{r.min()}, {r.max()}] for {A.shape}")
                out.append(r - 1)
            else:
                if not (1 <= r <= A.shape[d]):
                    raise JlError(f"index {r} out of bounds in dim {d + 1} of {A.shape}")
                out.append(int(r) - 1)
        return tuple(out), lanes

    def index_load(self, e, env, ps):
        base = self.ev(e[1], env, ps)
        if isinstance(base, JlType):
            return []                                  # Float64[]
        if isinstance(base, (tuple, list)):
            return base[self.ev(e[2][0], env, ps) - 1]
        idx, lanes = self._np_index(base, e[2], env, ps)
        val = base[idx]
        return L(val) if lanes else (float(val) if np.ndim(val) == 0 else val)

    def comprehension(self, e, env):
        gens = [(var, self.ev(r, env)) for var, r in e[2]]
        axes = [np.arange(r[1], r[2] + 1, dtype=np.int64) for _, r in gens]
        grids = np.meshgrid(*axes, indexing="ij")
        sub = dict(env)
        for (var, _), g in zip(gens, grids):
            sub[var] = L(g.ravel(order="F"))
        val = _raw(self.ev(e[1], sub))
        shape = tuple(len(a) for a in axes)
        return np.array(np.broadcast_to(val, (int(np.prod(shape)),)), dtype=np.float64).reshape(shape, order="F")

    # -- statements ---------------------------------------------------------------------------
    def exec_block(self, body, env, host):
        for st in body:
            self.exec_stmt(st, env, host)

    def exec_stmt(self, st, env, host, want_value=False):
        k = st[0]
        if k == "expr":
            return self.ev(st[1], env)
        if k == "assign":
            return self.assign(st, env, host)
        if k == "parallel":
            call = st[2]
            if call[0] != "call":
                raise JlError(f"@parallel without a call (line {st[-1]})")
            d = self.ev(call[1], env)
            args = [self.ev(x, env) for x in call[2]]
            ranges = None
            if st[1] is not None:
                r = self.ev(st[1], env)
                ranges = (r,) if r and r[0] == "range" else tuple(r)
            self.launch(d, args, ranges)
            return None
        if k == "if":
            cond = self.ev(st[1], env)
            if isinstance(cond, L):
                self.masked_if(cond.v.astype(bool), st[2], st[3], env, host)
            elif cond:
                self.exec_block(st[2], env, host)
            else:
                self.exec_block(st[3], env, host)
            return None
        if k == "for":
            r = self.ev(st[2], env)
            for v in range(r[1], r[2] + 1):
                env[st[1]] = v
                try:
                    self.exec_block(st[3], env, host)
                except _Break:
                    break
            return None
        if k == "break":
            raise _Break()
        if k == "return":
            raise _Return(None if st[1] is None else self.ev(st[1], env))
        raise JlError(f"statement {k}")

    def masked_if(self, m, body, orelse, env, host):
        for mask, blk in ((m, body), (~m, orelse)):
            if not blk or not mask.any():
                continue
            sub, before = {}, {}
            for n, v in env.items():
                if isinstance(v, L):
                    sub[n] = before[n] = L(v.v[mask])
                else:
                    sub[n] = v
            self.exec_block(blk, sub, host)
            for n, v in sub.items():
                if n in before and v is before[n]:
                    continue                              # thread variable untouched
                if not isinstance(v, (L, int, float, bool, np.integer, np.floating)):
                    continue                              # arrays, functions: never rebound inside a kernel
                if not isinstance(v, L) and n in env and env[n] is v:
                    continue                              # thread-uniform value untouched
                r = np.asarray(_raw(v))
                if n in env and isinstance(env[n], L):
                    full = env[n].v.copy()
                    if r.dtype.kind == "f" and full.dtype.kind != "f":
                        full = full.astype(np.float64)
                else:                                     # first defined under the condition
                    full = np.full(mask.shape, np.nan if r.dtype.kind == "f" else 0, dtype=r.dtype)
                full[mask] = r
                env[n] = L(full)

    def assign(self, st, env, host):
        targets, op, values, line = st[1], st[2], st[3], st[4]
        if targets[0][0] == "call" and op == "=":
            return None                               # short-form definition: taken at scan time
        if len(values) == 1:
            val = self.ev(values[0], env)
            vals = list(val) if len(targets) > 1 else [val]
        else:
            vals = [self.ev(v, env) for v in values]
        if len(values) == 1 and len(vals) > len(targets) > 1:
            vals = vals[:len(targets)]             # a, b = f()  takes the first two of a longer tuple
        if len(vals) != len(targets):
            raise JlError(f"line {line}: {len(targets)} targets, {len(vals)} values")
        for t, v in zip(targets, vals):
            if t[0] == "id":
                name = t[1]
                if op == ".=":
                    np.copyto(env[name], _raw(v))
                elif op == "=":
                    if host and name in self.frozen:
                        env[name] = self.frozen[name]
                    else:
                        env[name] = v
                else:
                    env[name] = self.binop(op[0], env[name], v)
            elif t[0] == "index":
                A = self.ev(t[1], env)
                idx, _ = self._np_index(A, t[2], env, None)
                if op in ("=", ".="):
                    A[idx] = _raw(v)
                else:
                    raise JlError(f"line {line}: {op} on an indexed target")
            else:
                raise JlError(f"line {line}: cannot assign to {t[0]}")
        return vals[-1]
