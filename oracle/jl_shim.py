"""jl_shim.py -- executes julia/NS3DNative.jl (the `ccall` shim) and the re-pointed run scripts
(scripts/NavierStokes3D_gpu_b200.jl, scripts/NavierStokes3D_b200.jl) with the interpreter of
oracle/jl_interp.py, every `ccall` going into the real C ABI of libns3d.so (or, on the CPU, into the
emulated library of tests/emu).

TEST INFRASTRUCTURE ONLY.  Julia is not in the image, so the Julia side of the drop-in boundary has
never run under Julia.  What this module does instead: it extends the interpreter by the part of the
language the shim is written in -- `module`, `struct`, `const`, typed signatures with defaults /
varargs / keyword arguments and dispatch on arity and annotated types, `Ref{T}`, `Ptr{T}`, `x.field`,
ternaries, lambdas, typed comprehensions -- and implements `ccall((:sym, LIB), Ret, (Types...), args...)`
through ctypes on an UNTYPED handle of the library: every argument is converted by the Julia type the
ccall declares, exactly as Julia would, so a wrong argument order, a wrong type or a wrong arity in the
shim's text reaches the library as such.  The three look-alike macros (`@init_ns3d`, `@zeros`,
`@parallel`) are interpreted by their documented expansion (their `quote` bodies are not executed).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .jl_interp import Def, JlError, JlType, JuliaScript, ParseError, Parser, Tok, _Break, _Return

CONT_OPS = ("=", ",", "&&", "||", "+", "-", "*", "/", "?", "->", "==", ".=", "+=", "-=")
_SKIP_STATEMENTS = ("export", "using", "import", "include")    # top-level statements that define nothing


# ------------------------------------------------------------------------------------------------
# parser: the additional syntax
# ------------------------------------------------------------------------------------------------
class ShimParser(Parser):
    no_range = 0

    # -- signatures ------------------------------------------------------------------------------
    def signature(self):
        """`(a, b::T, c::T=1, d::T...; k=v)` [where {...}] -> list of parameter dicts."""
        self.eat("(")
        save, self.in_index = self.in_index, 0
        params, kw = [], False
        while not self.is_op(")"):
            if self.is_op(";"):
                kw = True
                self.i += 1
                continue
            p = {"name": None, "type": None, "default": None, "vararg": False, "kw": kw}
            at = self.i
            if self.cur.kind == "id":
                p["name"] = self.cur.val
                self.i += 1
            if self.is_op("::"):
                self.i += 1
                p["type"] = self.type_expr()
            if self.is_op("..."):
                self.i += 1
                p["vararg"] = True
            if self.is_op("="):
                self.i += 1
                p["default"] = self.expr()
            params.append(p)
            if self.is_op(","):
                self.i += 1
            if self.i == at:
                raise ParseError(f"not a parameter list (line {self.cur.line})")
        self.eat(")")
        self.in_index = save
        if self.is_kw("where"):
            self.i += 1
            self.type_expr()
        return params

    def type_expr(self):
        if self.is_op("{"):                       # where {T<:...}
            return self.curly(("id", "where"))
        t = self.cur
        a = self.p_primary()
        while True:
            prev_end = self.t[self.i - 1].end
            if self.is_op("{") and self.cur.start == prev_end:
                a = self.curly(a)
            elif self.is_op(".") and self.peek().kind == "id":
                self.i += 1
                a = ("attr", a, self.cur.val)
                self.i += 1
            elif self.is_op("(") and self.cur.start == prev_end and t.val == "typeof":
                a = ("call", a, *self.call_args(), False)
            else:
                return a

    def curly(self, base):
        self.eat("{")
        save, self.in_index = self.in_index, 0
        items = []
        while not self.is_op("}"):
            if self.is_op("<:"):
                self.i += 1
                items.append(("subtype", self.type_expr()))
            else:
                e = self.type_expr() if self.cur.kind == "id" else self.expr()
                if self.is_op("<:"):
                    self.i += 1
                    e = ("subtype", self.type_expr())
                items.append(e)
            if self.is_op(","):
                self.i += 1
        self.eat("}")
        self.in_index = save
        return ("curly", base, items)

    # -- statements --------------------------------------------------------------------------------
    def statement(self):
        c = self.cur
        if c.kind == "id" and c.val == "const":
            self.i += 1
            return self.statement()
        if c.kind == "id" and (c.val in ("export", "using", "import") or (c.val == "include" and self.peek().val == "(")):
            while self.cur.kind not in ("nl", "eof"):
                self.i += 1
            return ("expr", ("num", None), c.line)
        if c.kind == "id" and c.val == "return":
            self.i += 1
            if self.at_stmt_end():
                return ("return", None, c.line)
            vals = self.expr_list()
            return ("return", vals[0] if len(vals) == 1 else ("tuple", vals), c.line)
        if c.kind == "id" and c.val == "for" and (self.peek().val == "(" or self.peek(2).val == "in"):
            self.i += 1
            if self.is_op("("):
                self.i += 1
                names = []
                while not self.is_op(")"):
                    names.append(self.cur.val)
                    self.i += 1
                    if self.is_op(","):
                        self.i += 1
                self.i += 1
            else:
                names = self.cur.val
                self.i += 1
            self.eat("in")
            it = self.expr()
            body = self.block()
            self.eat("end")
            return ("forin", names, it, body, c.line)
        return super().statement()

    # -- expressions ---------------------------------------------------------------------------------
    def expr(self):
        a = self.p_or()
        if self.is_op("?"):
            self.i += 1
            self.no_range += 1
            b = self.expr()
            self.no_range -= 1
            self.eat(":")
            return ("ternary", a, b, self.expr())
        if self.is_op("->"):
            self.i += 1
            params = [a[1]] if a[0] == "id" else [x[1] for x in a[1]]
            return ("lambda", params, self.expr())
        if self.is_op("=>"):
            self.i += 1
            return ("tuple", [a, self.expr()])
        return a

    def p_cmp(self):
        a = self.p_range()
        while self.cur.kind == "op" and self.cur.val in ("==", "!=", "<", "<=", ">", ">=", "===", "!=="):
            op = self.cur.val
            self.i += 1
            a = ("bin", {"===": "is", "!==": "isnot"}.get(op, op), a, self.p_range())
        return a

    def p_range(self):
        if self.no_range:
            return self.p_add()
        return super().p_range()

    def p_mul(self):
        a = self.p_unary()
        while self.cur.kind == "op" and self.cur.val in ("*", "/", "%", "÷"):
            op = self.cur.val
            self.i += 1
            a = ("bin", op, a, self.p_unary())
        return a

    def p_postfix(self):
        tok = self.cur
        a = self.p_primary()
        if tok.kind == "num" and self.cur.start == tok.end and (self.cur.kind == "id" or self.is_op("(")):
            return ("bin", "*", a, self.p_pow())       # numeric-literal coefficient
        while True:
            prev_end = self.t[self.i - 1].end
            if self.is_op("{") and self.cur.start == prev_end:
                a = self.curly(a)
            elif self.is_op("::"):
                self.i += 1
                self.type_expr()                   # a type assertion: no effect on the value
            elif self.is_op("(") and self.cur.start == prev_end:
                a = ("call", a, *self.call_args(), False)
            elif self.is_op("[") and self.cur.start == prev_end:
                a = self.index_or_typed_comprehension(a)
            elif self.is_op(".") and self.peek().kind == "id":
                self.i += 1
                a = ("attr", a, self.cur.val)
                self.i += 1
            elif self.is_op(".") and self.peek().kind == "op" and self.peek().val == "(":
                self.i += 1
                a = ("call", a, *self.call_args(), True)
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
                a = self.expr()
                if self.is_op("..."):
                    self.i += 1
                    a = ("splat", a)
                args.append(a)
            if self.is_op(","):
                self.i += 1
        self.eat(")")
        self.in_index = save
        return args, kwargs

    def index_or_typed_comprehension(self, base):
        self.i += 1
        save = self.in_index
        self.in_index += 1
        idx = []
        while not self.is_op("]"):
            if self.is_op(":") and self.peek().val in (",", "]"):
                self.i += 1
                idx.append(("colon",))
            else:
                idx.append(self.expr())
            if self.is_kw("for"):
                self.in_index = 0
                gens = self.generators()
                self.eat("]")
                self.in_index = save
                return ("tcomp", base, idx[0], gens)
            if self.is_op(","):
                self.i += 1
        self.in_index = save
        self.eat("]")
        return ("index", base, idx)

    def generators(self):
        self.eat("for")
        gens = []
        while True:
            var = self.cur.val
            self.i += 1
            scalar = self.is_kw("in")
            self.i += 1                               # `=` or `in`
            gens.append((var, self.expr(), scalar))
            if self.is_op(","):
                self.i += 1
                continue
            return gens

    def p_primary(self):
        c = self.cur
        if c.kind == "op" and c.val == ":" and self.peek().kind == "id" and self.peek().start == c.end:
            self.i += 2
            return ("sym", self.t[self.i - 1].val)
        if c.kind == "macro" and not (self.peek().val == "(" and self.peek().start == c.end):
            self.i += 1
            if c.val.startswith("__"):
                return ("str", '"' + c.val + '"')
            if c.val in ("info", "warn", "show", "printf"):         # logging: the rest of the line is its argument
                while self.cur.kind not in ("nl", "eof"):
                    self.i += 1
                return ("num", None)
            if c.val == "parallel":
                first = self.expr()
                if self.at_stmt_end() or self.cur.kind == "op":
                    return ("parallelexpr", None, first)
                return ("parallelexpr", first, self.expr())
            raise ParseError(f"macro @{c.val} in an expression (line {c.line})")
        if c.kind == "id" and c.val == "break":
            self.i += 1
            return ("breakexpr",)
        if self.is_op("(") and self.peek().kind == "op" and self.peek().val == ")":
            self.i += 2
            return ("tuple", [])
        if self.is_op("["):
            # untyped comprehension with `in` generators, or a vector literal
            save_i = self.i
            self.i += 1
            save, self.in_index = self.in_index, 0
            if not self.is_op("]"):
                first = self.expr()
                if self.is_kw("for"):
                    gens = self.generators()
                    self.eat("]")
                    self.in_index = save
                    if any(g[2] for g in gens):
                        return ("tcomp", None, first, gens)
                    return ("comprehension", first, [(v, r) for v, r, _ in gens])
            self.i, self.in_index = save_i, save
        return super().p_primary()


# ------------------------------------------------------------------------------------------------
# values
# ------------------------------------------------------------------------------------------------
class JPtr:
    """A C pointer value (`Ptr{T}`)."""
    __slots__ = ("addr",)

    def __init__(self, addr):
        self.addr = int(addr or 0)

    def __eq__(self, o):
        return isinstance(o, JPtr) and o.addr == self.addr

    def __hash__(self):
        return hash(self.addr)


class RefCell:
    """`Ref{T}(v)`; `r[]` reads, `r[] = v` writes."""

    def __init__(self, T, v):
        self.T, self.v = T, v


class JCurly:
    """A parametrised type value: `Ptr{Float64}`, `Ref{Cint}`, `Array{Float64,3}`."""

    def __init__(self, base, params):
        self.base, self.params = base, params

    def __repr__(self):
        return f"{self.base}{{{','.join(map(tname, self.params))}}}"


class JStructType:
    def __init__(self, name, fields, mutable, owner):
        self.name, self.fields, self.mutable, self.owner = name, fields, mutable, owner   # fields: [(name, type ast)]


class JStruct:
    def __init__(self, T: JStructType, values):
        self.T = T
        self.f = dict(zip([n for n, _ in T.fields], values))


class Generic:
    """A function name of a script: its methods are selected by arity and annotated types."""

    def __init__(self, name, owner):
        self.name, self.owner = name, owner


class Module:
    def __init__(self, name, owner, prefix=""):
        self.name, self.owner, self.prefix = name, owner, prefix


def tname(t):
    if isinstance(t, JlType):
        return t.name
    if isinstance(t, JStructType):
        return t.name
    return repr(t)


_SCALAR = {"Cint": C.c_int, "Cdouble": C.c_double, "Csize_t": C.c_size_t, "Clonglong": C.c_longlong, "Float64": C.c_double,
           "Int": C.c_longlong, "UInt8": C.c_uint8, "Cchar": C.c_char}
_NP = {"Cint": np.int32, "Cdouble": np.float64, "Float64": np.float64, "Float32": np.float32, "UInt8": np.uint8, "Int": np.int64}


# ------------------------------------------------------------------------------------------------
# the script
# ------------------------------------------------------------------------------------------------
class ShimScript(JuliaScript):
    parser_cls = ShimParser
    cont_ops = CONT_OPS

    def __init__(self, text, name="script", lib: C.CDLL | None = None, imports=()):
        self.lib = lib                       # UNTYPED ctypes handle of libns3d.so (or of the emulated library)
        self.imports = list(imports)
        self.methods: dict[str, list] = {}
        self.structs: dict[str, JStructType] = {}
        self.globals: dict = {}
        self.finalizers = []
        self.ccalls = []                     # (symbol, number of arguments) in call order
        self._ctypes_structs = {}
        self._pending_consts = []
        super().__init__(text, name)

    # -- scanning ---------------------------------------------------------------------------------
    def _scan_definitions(self):
        self._scan_tokens(self.toks, prefix="", depth=0)

    def _statements(self, toks):
        spans, start, depth, br = [], 0, 0, 0
        block_open = {"function", "macro", "if", "for", "while", "begin", "let", "struct", "try", "quote", "module", "do"}
        for k, t in enumerate(toks):
            if t.kind == "op" and len(t.val) == 1:
                br += (t.val in "([{") - (t.val in ")]}")
            elif t.kind == "id" and br == 0:
                if t.val in block_open and not (k and toks[k - 1].kind == "op" and toks[k - 1].val == "."):
                    depth += 1
                elif t.val == "end":
                    depth -= 1
            elif t.kind in ("nl", "eof") and depth == 0 and br == 0:
                if k > start:
                    spans.append(toks[start:k])
                start = k + 1
        return spans

    def _scan_tokens(self, toks, prefix, depth):
        eof = Tok("eof", None, 0, 0, toks[-1].line if toks else 0)
        for ts in self._statements(toks):
            vals = [t.val for t in ts]
            head = ts[0]
            if head.kind == "str" and len(ts) == 1:
                continue                                           # a docstring
            if head.kind == "id" and head.val == "module":
                name = ts[1].val
                inner = "" if depth == 0 else prefix + name + "."      # the file's own module is the global scope
                self.globals[name] = Module(name, self, inner)
                self._scan_tokens(ts[2:-1] + [Tok("nl", "\n", 0, 0, ts[-1].line)], inner, depth + 1)
                continue
            if head.kind == "id" and head.val in _SKIP_STATEMENTS or head.kind == "id" and head.val == "macro":
                continue
            k = 1 if head.kind == "id" and head.val == "mutable" else 0
            if ts[k].kind == "id" and ts[k].val == "struct":
                self._scan_struct(ts[k + 1:], mutable=bool(k), prefix=prefix)
                continue
            if "function" in vals[:6]:
                f = vals.index("function")
                p = self.parser_cls(ts[f + 1:] + [eof])
                try:
                    name = self._def_name(p)
                    sig = p.signature()
                    body = p.block()
                except ParseError:
                    continue                      # the run functions (plotting / `do` blocks) run by line range
                self._add_method(prefix + name, Def("function", name, [q["name"] for q in sig], body, None, (ts[0].line, ts[-1].line)), sig)
                continue
            # short form  [Base.]name(sig) [where {..}] = expr      /     const a, b = ...    /  other top-level code
            p = self.parser_cls(ts + [Tok("nl", "\n", 0, 0, ts[-1].line), eof])
            k0 = 1 if head.kind == "macro" and head.val == "inline" else 0
            p.i = k0
            try:
                save = p.i
                name = self._def_name(p)
                if p.is_op("(") and p.cur.start == p.t[p.i - 1].end:
                    sig = p.signature()
                    if p.is_op("="):
                        p.i += 1
                        body = p.expr()
                        self._add_method(prefix + name, Def("short", name, [q["name"] for q in sig], body, None, (ts[0].line, ts[-1].line)), sig)
                        continue
                p.i = save
            except (ParseError, IndexError, TypeError):
                pass
            if head.kind == "id" and head.val == "const":
                p = self.parser_cls(ts + [Tok("nl", "\n", 0, 0, ts[-1].line), eof])
                st = p.statement()
                self._pending_consts.append((st, prefix))

    def _def_name(self, p):
        name = p.cur.val
        if p.cur.kind != "id":
            raise ParseError("not a definition")
        p.i += 1
        while p.is_op(".") and p.peek().kind == "id":          # Base.size -> size
            p.i += 1
            name = p.cur.val
            p.i += 1
        return name

    def _add_method(self, name, d: Def, sig):
        d.sig, d.owner = sig, self
        self.methods.setdefault(name, []).append(d)
        self.defs[name] = d

    def _scan_struct(self, ts, mutable, prefix):
        name = ts[0].val
        fields, i = [], 1
        body = ts[1:-1]
        eof = Tok("eof", None, 0, 0, ts[-1].line)
        i = 0
        while i < len(body):
            t = body[i]
            if t.kind == "id" and i + 1 < len(body) and body[i + 1].kind == "op" and body[i + 1].val == "::":
                j = i + 2
                depth = 0
                while j < len(body) and not (depth == 0 and (body[j].kind == "nl" or (body[j].kind == "op" and body[j].val == ";"))):
                    depth += (body[j].val in ("{", "(")) - (body[j].val in ("}", ")")) if body[j].kind == "op" else 0
                    j += 1
                p = self.parser_cls(body[i + 2:j] + [eof])
                fields.append((t.val, p.type_expr()))
                i = j
            else:
                i += 1
        self.structs[prefix + name] = JStructType(name, fields, mutable, self)

    def finish_loading(self):
        """Evaluate the module-level constants in text order (after every definition is known)."""
        for st, prefix in self._pending_consts:
            self.exec_stmt(st, self.globals, host=False)
        self._pending_consts = []

    # -- names -------------------------------------------------------------------------------------
    def lookup(self, name, env):
        if name in env:
            return env[name]
        for s in [self] + self.imports:
            if name in s.globals:
                return s.globals[name]
            if name in s.structs or name in s.methods:
                return Generic(name, s)
        return self.builtin(name)

    def _make_builtins(self):
        B = super()._make_builtins()
        for n in ("Cint", "Cdouble", "Csize_t", "Cstring", "Cvoid", "Clonglong", "UInt8", "Cchar", "Integer", "Real", "Any", "Nothing",
                  "UnitRange", "Vector", "Union", "Type", "NTuple", "AbstractString", "String"):
            B[n] = JlType(n)
        B.update({
            "C_NULL": JPtr(0), "undef": JlType("undef"), "ENV": {}, "true": True, "false": False,
            "error": self._error, "length": self._length, "prod": lambda t: int(np.prod(t)),
            "first": lambda r: r[1], "last": lambda r: r[2], "sum": lambda x: int(np.sum(x)),
            "isempty": lambda x: len(x) == 0, "enumerate": lambda x: list(enumerate(x, 1)), "collect": lambda *a: a[-1],
            "unsafe_string": lambda b: b.decode() if isinstance(b, bytes) else str(b),
            "finalizer": lambda f, obj: self.finalizers.append((f, obj)),
            "pointer": lambda a: JPtr(a.ctypes.data), "fill": lambda v, n: [v] * n, "string": lambda *a: "".join(map(str, a)),
            "get": lambda d, k, default: d.get(k, default), "parse": lambda T, s: int(s), "map": lambda f, xs: [f(x) for x in xs],
            "Ptr": JlType("Ptr"), "Ref": JlType("Ref"),
            "abs": B["abs"], "Int": B["Int"], "zeros": self._zeros, "Base": {"Array": JlType("Array")},
            "isfinite": B["isfinite"], "typeof": lambda x: ("typeof", x),
        })
        return B

    @staticmethod
    def _error(*parts):
        raise JlError("".join(str(p) for p in parts))

    @staticmethod
    def _length(x):
        if isinstance(x, tuple) and x and x[0] == "range":
            return x[2] - x[1] + 1
        if isinstance(x, np.ndarray):
            return int(x.size)
        return len(x)

    def _zeros(self, *a):
        if a and isinstance(a[0], JlType):
            return np.zeros(tuple(int(x) for x in a[1:]), dtype=_NP[a[0].name], order="F")
        return np.zeros(tuple(int(x) for x in a), dtype=np.float64, order="F")

    # -- evaluation: the additional node kinds -------------------------------------------------------
    def ev(self, e, env, ps=None):
        k = e[0]
        if k == "sym":
            return ("sym", e[1])
        if k == "str":
            return e[1][1:-1] if not e[1].startswith('"""') else e[1][3:-3]
        if k == "ternary":
            return self.ev(e[2], env, ps) if self.ev(e[1], env, ps) else self.ev(e[3], env, ps)
        if k == "lambda":
            return lambda *a, _e=e, _env=env: self.ev(_e[2], {**_env, **dict(zip(_e[1], a))})
        if k == "breakexpr":
            raise _Break()
        if k == "curly":
            base = e[1][1] if e[1][0] == "id" else (e[1][2] if e[1][0] == "attr" else repr(self.ev(e[1], env, ps)))
            if base in self.globals and isinstance(self.globals[base], JCurly):
                base = repr(self.globals[base])
            params = [self.ev(x[1] if x[0] == "subtype" else x, env, ps) for x in e[2]]
            return JCurly(base, params)
        if k == "attr":
            obj = self.ev(e[1], env, ps)
            return self.getattr(obj, e[2])
        if k == "tcomp":
            return self.scalar_comprehension(e, env)
        if k == "parallelexpr":
            return self.parallel_call(e[2], env)
        if k == "bin" and e[1] in ("is", "isnot"):
            a, b = self.ev(e[2], env, ps), self.ev(e[3], env, ps)
            same = (a is b) or (type(a) is type(b) and not isinstance(a, (np.ndarray, JStruct)) and a == b)
            return same if e[1] == "is" else not same
        if k == "bin" and e[1] == "÷":
            return int(self.ev(e[2], env, ps)) // int(self.ev(e[3], env, ps))
        if k == "bin" and e[1] == "==":
            a, b = self.ev(e[2], env, ps), self.ev(e[3], env, ps)
            if isinstance(a, (tuple, JlType, JCurly, JPtr, str)) or isinstance(b, (tuple, JlType, JCurly, JPtr, str)):
                if isinstance(a, JlType) and isinstance(b, JlType):
                    return a.name == b.name
                return a == b
            return self.binop("==", a, b)
        if k == "macrocall" and e[1] in ("zeros", "init_ns3d", "sprintf", "printf", "info", "__DIR__", "__FILE__"):
            return self.shim_macro(e[1], e[2], env)
        return super().ev(e, env, ps)

    def getattr(self, obj, name):
        if isinstance(obj, JStruct):
            return obj.f[name]
        if isinstance(obj, Module):
            o, full = obj.owner, obj.prefix + name
            return o.lookup(full if (full in o.methods or full in o.globals or full in o.structs) else name, {})
        if isinstance(obj, dict):
            return obj[name]
        return getattr(obj, name.rstrip("!") + ("_b" if name.endswith("!") else ""))

    def comprehension(self, e, env):
        """Vectorised over the index box where every callee takes lane vectors; element by element otherwise
        (e.g. `z_g(iz, dz, C)` of the shim dispatches on `iz::Integer`)."""
        try:
            return super().comprehension(e, env)
        except JlError as exc:
            if "MethodError" not in str(exc):
                raise
        gens = [(var, self.ev(r, env)) for var, r in e[2]]
        shape = tuple(r[2] - r[1] + 1 for _, r in gens)
        out = np.empty(shape, dtype=np.float64, order="F")
        scope = dict(env)
        for idx in np.ndindex(*shape):
            for (var, r), i in zip(gens, idx):
                scope[var] = r[1] + i
            out[idx] = self.ev(e[1], scope)
        return out

    def scalar_comprehension(self, e, env):
        base = self.ev(e[1], env) if e[1] is not None else None
        out = []

        def rec(gens, scope):
            if not gens:
                out.append(self.ev(e[2], scope))
                return
            var, it, _ = gens[0]
            seq = self.ev(it, scope)
            if isinstance(seq, tuple) and seq and seq[0] == "range":
                seq = range(seq[1], seq[2] + 1)
            for v in seq:
                rec(gens[1:], {**scope, var: v})
        rec(list(e[3]), dict(env))
        if base is None:
            return out
        if isinstance(base, JCurly) and base.base == "Ptr":
            return np.array([v.addr for v in out], dtype=np.uint64)
        return np.array(out, dtype=_NP[tname(base)])

    # -- calls ------------------------------------------------------------------------------------------
    def call(self, e, env, ps):
        if e[1] == ("id", "ccall"):
            return self.ccall(e[2], env)
        fn = self.ev(e[1], env, ps)
        args = []
        for x in e[2]:
            if x[0] == "splat":
                args.extend(self.ev(x[1], env, ps))
            else:
                args.append(self.ev(x, env, ps))
        kwargs = {n: self.ev(x, env, ps) for n, x in e[3].items()}
        return self.apply(fn, args, kwargs, dotted=e[4])

    def apply(self, fn, args, kwargs, dotted=False):
        if isinstance(fn, Generic):
            return fn.owner.dispatch(fn.name, args, kwargs)
        if isinstance(fn, Def):
            owner = getattr(fn, "owner", self)
            env = owner.bind(fn, args, kwargs) if hasattr(fn, "sig") else None
            if env is None:
                return owner.call_def(fn, args)
            return owner.run_def(fn, env)
        if isinstance(fn, JCurly):
            return self.construct_curly(fn, args)
        if isinstance(fn, JlType):
            if dotted:                                                # Cint.(v)
                return np.array(args[0], dtype=_NP[fn.name])
            if fn.name in ("Cint", "Int", "Csize_t"):
                return int(args[0])
            if fn.name in ("Cdouble", "Float64"):
                return float(args[0])
            raise JlError(f"cannot call type {fn.name}")
        return fn(*args, **kwargs)

    def construct_curly(self, T: JCurly, args):
        if T.base == "Ref":
            return RefCell(T.params[0], args[0] if args else None)
        if T.base == "Ptr":
            a = args[0]
            return a if isinstance(a, JPtr) else JPtr(a)
        if T.base == "Array":
            shape = []
            for a in args[1:]:
                shape += list(a) if isinstance(a, tuple) else [a]
            return np.empty(tuple(int(x) for x in shape), dtype=_NP[tname(T.params[0])], order="F")
        raise JlError(f"cannot construct {T!r}")

    def isa(self, v, t) -> bool:
        if t is None:
            return True
        if t[0] == "id":
            n = t[1]
            if n in ("Integer", "Int", "Cint"):
                return isinstance(v, (int, np.integer)) and not isinstance(v, bool)
            if n == "Real":
                return isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool)
            for s in [self] + self.imports:
                if n in s.structs:
                    return isinstance(v, JStruct) and v.T is s.structs[n]
            if n == "UnitRange":
                return isinstance(v, tuple) and v and v[0] == "range"
            return True
        if t[0] == "curly":
            b = t[1][1] if t[1][0] == "id" else t[1][2]
            if b in ("Vector", "Array"):
                return isinstance(v, (np.ndarray, list))
            if b == "Ptr":
                return isinstance(v, JPtr)
            if b == "Type":
                return isinstance(v, (JlType, JCurly, JStructType))
            return True
        if t[0] == "call":                        # ::typeof(abs)
            return callable(v) and not isinstance(v, (JStruct, Generic))
        if t[0] == "attr":
            return True
        return True

    def bind(self, d: Def, args, kwargs):
        """Match a call against a method's signature -> the callee's environment, or None."""
        pos = [p for p in d.sig if not p["kw"]]
        kws = [p for p in d.sig if p["kw"]]
        env, i = {}, 0
        for p in pos:
            if p["vararg"]:
                rest = tuple(args[i:])
                if not all(self.isa(v, p["type"]) for v in rest):
                    return None
                env[p["name"]] = rest
                i = len(args)
                continue
            if i < len(args):
                v = args[i]
                if not self.isa(v, p["type"]):
                    return None
                i += 1
            elif p["default"] is not None:
                v = self.ev(p["default"], {**env})
            else:
                return None
            if p["name"]:
                env[p["name"]] = v
            t = p["type"]
            if t is not None and t[0] == "curly" and t[1] == ("id", "Type") and t[2] and t[2][0][0] == "id":
                env[t[2][0][1]] = v                                   # ::Type{T} binds T
        if i != len(args):
            return None
        names = {p["name"] for p in kws}
        if set(kwargs) - names:
            return None
        for p in kws:
            env[p["name"]] = kwargs[p["name"]] if p["name"] in kwargs else self.ev(p["default"], {**env})
        return env

    def dispatch(self, name, args, kwargs):
        for d in self.methods.get(name, []):
            env = self.bind(d, args, kwargs)
            if env is not None:
                return self.run_def(d, env)
        if name in self.structs:                  # the default constructor
            T = self.structs[name]
            if len(args) == len(T.fields) and not kwargs:
                return JStruct(T, list(args))
        if name in ("size", "length", "Array") and name in self._builtins:     # Base functions the shim only EXTENDS
            return self._builtins[name](*args, **kwargs)
        sigs = [", ".join((p["name"] or "") + ("::…" if p["type"] else "") for p in d.sig) for d in self.methods.get(name, [])]
        raise JlError(f"MethodError: no method matching {name}({', '.join(self.describe(a) for a in args)}"
                      f"{'; ' + ', '.join(kwargs) if kwargs else ''}); methods: {sigs}")

    @staticmethod
    def describe(v):
        if isinstance(v, JStruct):
            return "::" + v.T.name
        return "::" + type(v).__name__

    def run_def(self, d: Def, env):
        if d.kind == "short":
            return self.ev(d.body, env)
        try:
            self.exec_block(d.body, env, host=False)
        except _Return as r:
            return r.val
        return None

    # -- statements ------------------------------------------------------------------------------------
    def exec_stmt(self, st, env, host, want_value=False):
        k = st[0]
        if k == "forin":
            seq = self.ev(st[2], env)
            if isinstance(seq, tuple) and seq and seq[0] == "range":
                seq = range(seq[1], seq[2] + 1)
            for v in seq:
                if isinstance(st[1], list):
                    env.update(dict(zip(st[1], v)))
                else:
                    env[st[1]] = v
                try:
                    self.exec_block(st[3], env, host)
                except _Break:
                    break
            return None
        if k == "parallel":
            return self.parallel_call(st[2], env)
        if k == "assign":
            t = st[1][0]
            if t[0] == "call" and st[2] == "=" and t[1][0] == "id":           # a local function  f(a,b) = expr
                d = Def("short", t[1][1], [a[1] for a in t[2]], st[3][0])
                d.closure = env
                env[t[1][1]] = d
                return None
            if t[0] == "index" and not t[2]:                                    # r[] = v
                cell = self.ev(t[1], env)
                cell.v = self.ev(st[3][0], env)
                return cell.v
        return super().exec_stmt(st, env, host, want_value)

    def call_def(self, d: Def, args):
        if getattr(d, "closure", None) is not None:
            return self.ev(d.body, {**d.closure, **dict(zip(d.params, args))})
        return super().call_def(d, args)

    def index_load(self, e, env, ps):
        base = self.ev(e[1], env, ps)
        if isinstance(base, RefCell) and not e[2]:
            return base.v
        if isinstance(base, np.ndarray) and base.ndim == 1 and len(e[2]) == 1 and e[2][0] != ("colon",):
            i = self.ev(e[2][0], {**env, "__end__": base.shape[0]}, ps)
            if isinstance(i, tuple) and i and i[0] == "range":
                return base[i[1] - 1:i[2]]
            if isinstance(i, (int, np.integer)):
                if not 1 <= i <= base.shape[0]:
                    raise JlError(f"BoundsError: index {i} of a vector of {base.shape[0]}")
                return base[int(i) - 1].item()
        if isinstance(base, (tuple, list)):
            i = self.ev(e[2][0], {**env, "__end__": len(base)}, ps)
            return base[i - 1]
        return super().index_load(e, env, ps)

    # -- the look-alike macros, by their documented expansion ----------------------------------------------
    def default_ctx(self):
        return self.apply(self.lookup("default_ctx", {}), [], {})

    def shim_macro(self, name, args, env):
        if name == "zeros":
            return self.apply(self.lookup("zeros3", {}), [self.default_ctx()] + [self.ev(a, env) for a in args], {})
        if name == "init_ns3d":                                # NS3DNative.DEFAULT[] = NS3DNative.Ctx(args...)
            ctx = self.apply(self.lookup("Ctx", {}), [self.ev(a, env) for a in args], {})
            self.lookup("DEFAULT", {}).v = ctx
            return ctx
        return None

    def parallel_call(self, call, env):
        """`@parallel [ranges] f!(args...)`  ->  `f!(default_ctx(), args...)` (the ranges are implied by the shapes)."""
        if call[0] != "call":
            raise JlError("@parallel expects a function call")
        fn = self.ev(call[1], env)
        if isinstance(fn, Def) and fn.kind in ("ps_kernel", "pi_kernel"):
            raise JlError("a ParallelStencil kernel in a script that is bound to the library")
        args = [self.ev(a, env) for a in call[2]]
        return self.apply(fn, [self.default_ctx()] + args, {})

    # -- ccall ---------------------------------------------------------------------------------------------
    def ctypes_struct(self, T: JStructType):
        if T.name not in self._ctypes_structs:
            fields = []
            for n, t in T.fields:
                tv = T.owner.ev(t, {})
                if isinstance(tv, JStructType) or isinstance(tv, Generic):
                    st = T.owner.structs[tv.name]
                    fields.append((n, self.ctypes_struct(st)))
                elif isinstance(tv, JCurly):
                    fields.append((n, C.c_void_p))
                else:
                    fields.append((n, _SCALAR[tv.name]))
            self._ctypes_structs[T.name] = type("C_" + T.name, (C.Structure,), {"_fields_": fields})
        return self._ctypes_structs[T.name]

    def to_cstruct(self, v: JStruct):
        cls = self.ctypes_struct(v.T)
        obj = cls()
        for (n, ctype) in cls._fields_:
            x = v.f[n]
            if isinstance(x, JStruct):
                setattr(obj, n, self.to_cstruct(x))
            elif isinstance(x, JPtr):
                setattr(obj, n, x.addr)
            elif ctype in (C.c_int, C.c_longlong, C.c_size_t):
                if isinstance(x, float) and x != int(x):
                    raise JlError(f"InexactError: {v.T.name}.{n} = {x} is not an integer")
                setattr(obj, n, int(x))
            else:
                setattr(obj, n, float(x))
        return obj

    def ccall(self, arg_asts, env):
        if self.lib is None:
            raise JlError("ccall without a library handle")
        target = self.ev(arg_asts[0], env)
        sym = target[0][1]
        ret = self.ev(arg_asts[1], env)
        types = self.ev(arg_asts[2], env)
        types = list(types) if isinstance(types, tuple) else [types]
        vals = [self.ev(a, env) for a in arg_asts[3:]]
        if len(vals) != len(types):
            raise JlError(f"ccall {sym}: {len(types)} argument types, {len(vals)} arguments")
        cargs, writeback, keep = [], [], []
        for k, (t, v) in enumerate(zip(types, vals)):
            tn = tname(t)
            if isinstance(t, JlType) and tn in _SCALAR:
                if tn in ("Cint", "Csize_t", "Clonglong"):
                    if isinstance(v, (float, np.floating)) and float(v) != int(v):
                        raise JlError(f"ccall {sym}: InexactError converting {v} to {tn} (argument {k + 1})")
                    if not isinstance(v, (int, float, bool, np.integer, np.floating)):
                        raise JlError(f"ccall {sym}: argument {k + 1} is {self.describe(v)}, declared {tn}")
                    cargs.append(_SCALAR[tn](int(v)))
                else:
                    if not isinstance(v, (int, float, np.integer, np.floating)) or isinstance(v, bool):
                        raise JlError(f"ccall {sym}: argument {k + 1} is {self.describe(v)}, declared {tn}")
                    cargs.append(C.c_double(float(v)))
            elif isinstance(t, JCurly) and t.base == "Ptr":
                if isinstance(v, JPtr):
                    cargs.append(C.c_void_p(v.addr))
                elif isinstance(v, np.ndarray):
                    want = t.params[0]
                    if isinstance(want, JlType) and want.name in _NP and v.dtype != _NP[want.name]:
                        raise JlError(f"ccall {sym}: argument {k + 1} is an array of {v.dtype}, declared {t!r}")
                    if not (v.flags.f_contiguous or v.flags.c_contiguous):
                        raise JlError(f"ccall {sym}: argument {k + 1} is not contiguous")
                    keep.append(v)
                    cargs.append(C.c_void_p(v.ctypes.data))
                else:
                    raise JlError(f"ccall {sym}: argument {k + 1} is {self.describe(v)}, declared {t!r}")
            elif isinstance(t, JCurly) and t.base == "Ref":
                inner = t.params[0]
                if isinstance(v, JStruct):
                    obj = self.to_cstruct(v)
                    keep.append(obj)
                    cargs.append(C.byref(obj))
                elif isinstance(v, RefCell):
                    if isinstance(inner, JCurly):                    # Ref{Ptr{..}}
                        obj = C.c_void_p(v.v.addr if isinstance(v.v, JPtr) else 0)
                        writeback.append((v, obj, "ptr"))
                    else:
                        obj = _SCALAR[tname(inner)](v.v or 0)
                        writeback.append((v, obj, "val"))
                    cargs.append(C.byref(obj))
                else:
                    raise JlError(f"ccall {sym}: argument {k + 1} is {self.describe(v)}, declared {t!r}")
            elif isinstance(t, JlType) and tn == "Cstring":
                cargs.append(C.c_char_p(v.encode() if isinstance(v, str) else v))
            else:
                raise JlError(f"ccall {sym}: unsupported argument type {t!r}")
        fn = getattr(self.lib, sym)
        fn.argtypes = None
        rn = tname(ret)
        fn.restype = {"Cint": C.c_int, "Cstring": C.c_char_p, "Csize_t": C.c_size_t, "Clonglong": C.c_longlong}.get(rn, C.c_void_p)
        self.ccalls.append((sym, len(cargs)))
        out = fn(*cargs)
        for cell, obj, kind in writeback:
            cell.v = JPtr(obj.value) if kind == "ptr" else obj.value
        if isinstance(ret, JCurly):
            return JPtr(out)
        return out


def load_shim(path: str, lib: C.CDLL) -> ShimScript:
    with open(path, encoding="utf-8") as fh:
        s = ShimScript(fh.read(), name=path, lib=lib)
    s.globals["LIB"] = lib
    s._pending_consts = [(st, p) for st, p in s._pending_consts
                         if not (st[0] == "assign" and st[1][0] == ("id", "LIB"))]
    s.finish_loading()
    return s


def load_script(path: str, shim: ShimScript) -> ShimScript:
    with open(path, encoding="utf-8") as fh:
        s = ShimScript(fh.read(), name=path, lib=shim.lib, imports=[shim])
    s._pending_consts = []
    return s


# ------------------------------------------------------------------------------------------------
# the two re-pointed run scripts
# ------------------------------------------------------------------------------------------------
def _finalize(*scripts):
    for s in scripts:
        for f, obj in s.finalizers:
            f(obj)
        s.finalizers = []


def run_gpu_b200(lib, shim_path, script_path, nx, nt, use_fused=True, mode=0):
    """scripts/NavierStokes3D_gpu_b200.jl: the file's top-level statements, then the body of `runme` (physics block to the
    end of the time loop; the `.mat` branch is not executed), with the literal `nx = 255` replaced and the context
    created in `mode` (0 = PARITY).  Returns (fields on the host, iterations per step, err history per step, scripts)."""
    shim = load_shim(shim_path, lib)
    scr = load_script(script_path, shim)
    scr.frozen = {"nx": nx, "USE_FUSED": use_fused}
    text = scr.text.split("\n")
    with np.errstate(all="ignore"):
        for n, ln in enumerate(text, 1):
            if ln.startswith("const "):
                scr.run_lines(n, n, scr.globals)
        # `@init_ns3d(0, FAST)` as written, then the arithmetic mode under test
        n_init = scr.find_line(r"^@init_ns3d")
        scr.run_lines(n_init, n_init, scr.globals)
        ctx = shim.globals["DEFAULT"].v
        scr.apply(shim.lookup("set_mode!", {}), [ctx, mode], {})
        head = scr.find_line(r"function runme\(")
        first = scr.find_line(r"^\s*for it = 1:nt", head)
        last = scr.find_line(r"^\s*if do_save && it % nsave == 0", first) - 1
        env = {"do_vis": False, "do_save": False, "nt": nt}
        scr.run_lines(head + 1, first - 1, env)
        body = scr.parse_lines(first, last, close_blocks=1)
        iters, errs = [], []
        for it in range(1, nt + 1):
            env["it"] = it
            scr.exec_block(body[0][3], env, host=True)
            iters.append(int(env["iters"] if use_fused else env["iter"]))
            errs.append([float(e) for e in env["err_evo"]])
        fields = {k: scr.apply(shim.lookup("to_host", {}), [ctx, env[j]], {})
                  for k, j in (("Pr", "Pr"), ("Vx", "Vx"), ("Vy", "Vy"), ("Vz", "Vz"), ("C", "C"))}
    _finalize(shim, scr)
    return fields, iters, errs, (shim, scr)


class SingleRankMPI:
    """Stand-in for the MPI.jl module on one rank (`import MPI` of scripts/NavierStokes3D_b200.jl)."""
    COMM_WORLD = "COMM_WORLD"

    def __init__(self):
        self.calls = []

    def Initialized(self):
        return True

    def Init(self):
        self.calls.append("Init")

    def Comm_rank(self, comm):
        return 0

    def Comm_size(self, comm):
        return 1

    def Bcast_b(self, buf, root, comm):
        self.calls.append("Bcast!")
        return buf

    def Finalize(self):
        self.calls.append("Finalize")


def run_multi_b200(lib, shim_path, script_path, nx, nt, use_fused=True, mode=0, mpi=None, literals=None):
    """scripts/NavierStokes3D_b200.jl on one rank: the body of `run_navierstokes3D` from its first line to the end of the
    time loop (the `do_save` branch is not executed) and the return statement; `Ctx(...; mode=FAST)` is followed by
    `set_mode!(ctx, mode)`.  Returns (the five returned interiors C, Pr, Vx, Vy, Vz; local fields; iterations; scripts)."""
    shim = load_shim(shim_path, lib)
    scr = load_script(script_path, shim)
    scr.frozen = {"USE_FUSED_PT": use_fused, **(literals or {})}
    mpi = mpi or SingleRankMPI()
    scr.globals["MPI"] = mpi
    with np.errstate(all="ignore"):
        for n, ln in enumerate(scr.text.split("\n"), 1):
            if ln.startswith("const "):
                scr.run_lines(n, n, scr.globals)
        head = scr.find_line(r"function run_navierstokes3D\(")
        first = scr.find_line(r"^\s*for it = 1:nt", head)
        last = scr.find_line(r"^\s*if do_save && it % 10 == 0", first) - 1
        ret_first = scr.find_line(r"^\s*np_c = fill", last)
        ret_last = scr.find_line(r"^end", ret_first) - 1
        env = {"do_vis": False, "do_save": False, "do_print": False, "nx": nx, "nt": nt}
        n_ctx = scr.find_line(r"^\s*ctx = Ctx\(", head)
        scr.run_lines(head + 1, n_ctx, env)
        scr.apply(shim.lookup("set_mode!", {}), [env["ctx"], mode], {})
        scr.run_lines(n_ctx + 1, first - 1, env)
        body = scr.parse_lines(first, last, close_blocks=1)
        iters = []
        for it in range(1, nt + 1):
            env["it"] = it
            scr.exec_block(body[0][3], env, host=True)
            iters.append(int(env["iters"] if use_fused else env["iter"]))
        local = {k: scr.apply(shim.lookup("to_host", {}), [env["ctx"], env[j]], {})
                 for k, j in (("Pr", "Pr"), ("Vx", "Vx"), ("Vy", "Vy"), ("Vz", "Vz"), ("C", "C"), ("dPrdtau", "dPrdτ"), ("divV", "∇V"))}
        try:
            scr.run_lines(ret_first, ret_last, env)
            returned = None
        except _Return as r:
            returned = r.val
    _finalize(shim, scr)
    return returned, local, iters, (shim, scr, mpi)


class PipeMPI(SingleRankMPI):
    """MPI.jl stand-in for rank processes of a test: rank / size fixed, `Bcast!` from rank 0 over pipes."""

    def __init__(self, rank, size, pipes):
        super().__init__()
        self.rank, self.size, self.pipes = rank, size, pipes   # rank 0: the write ends; others: their read end

    def Comm_rank(self, comm):
        return self.rank

    def Comm_size(self, comm):
        return self.size

    def Bcast_b(self, buf, root, comm):
        self.calls.append("Bcast!")
        if self.rank == root:
            for w in self.pipes:
                w.send(buf.tobytes())
        else:
            buf[...] = np.frombuffer(self.pipes.recv(), dtype=buf.dtype).reshape(buf.shape)
        return buf


def run_multi_gpu_lookalike_b200(lib, shim_path, script_path, nx, nt, use_fused=True, mode=0, mpi=None, literals=None):
    """scripts/NavierStokes3D_multi_gpu_b200.jl (the M script's text on the look-alike surface) on one rank: the body of
    `run_navierstokes3D` up to the end of the time loop, the local fields read back, then the gathers, `finalize_global_grid`
    and the return statement.  Returns (returned interiors, local fields, iterations per step, err histories, scripts)."""
    shim = load_shim(shim_path, lib)
    scr = load_script(script_path, shim)
    scr.frozen = {"USE_FUSED": use_fused, **(literals or {})}
    mpi = mpi or SingleRankMPI()
    scr.globals["MPI"] = mpi
    with np.errstate(all="ignore"):
        for n, ln in enumerate(scr.text.split("\n"), 1):
            if ln.startswith("const "):
                scr.run_lines(n, n, scr.globals)
        head = scr.find_line(r"function run_navierstokes3D\(")
        n_grid = scr.find_line(r"init_global_grid\(nx, ny, nz", head)
        first = scr.find_line(r"^\s*for it = 1:nt", head)
        last = scr.find_line(r"^\s*# gather the interiors", first) - 1
        ret_last = scr.find_line(r"^\s*return C_v", last)
        env = {"do_vis": False, "do_save": False, "do_print": False, "nx": nx, "nt": nt}
        scr.run_lines(head + 1, n_grid, env)
        ctx = shim.globals["DEFAULT"].v
        scr.apply(shim.lookup("set_mode!", {}), [ctx, mode], {})
        scr.run_lines(n_grid + 1, first - 1, env)
        body = scr.parse_lines(first, last)
        iters, errs = [], []
        for it in range(1, nt + 1):
            env["it"] = it
            scr.exec_block(body[0][3], env, host=True)
            iters.append(int(env["iters"] if use_fused else env["iter"]))
            errs.append([float(e) for e in env["err_evo"]])
        local = {k: scr.apply(shim.lookup("to_host", {}), [ctx, env[j]], {})
                 for k, j in (("Pr", "Pr"), ("Vx", "Vx"), ("Vy", "Vy"), ("Vz", "Vz"), ("C", "C"), ("dPrdtau", "dPrdτ"), ("divV", "∇V"))}
        try:
            scr.run_lines(last + 1, ret_last, env)
            returned = None
        except _Return as r:
            returned = r.val
    _finalize(shim, scr)
    return returned, local, iters, errs, (shim, scr, mpi)
