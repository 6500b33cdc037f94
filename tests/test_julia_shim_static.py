"""Static checks of the Julia side (julia/NS3DNative.jl, scripts/*_b200.jl), which cannot run here (no Julia).

What a machine without Julia can still establish, with the tokenizer of oracle/jl_interp.py:

* the three files lex, their brackets and their block keywords / `end` balance;
* every `ccall((:sym, LIB), Ret, (ArgTypes...), args...)` of the shim names a symbol that include/ns3d.h declares,
  passes exactly as many argument types and as many arguments as the C prototype has parameters (`ccall` takes
  no splats -- the defect ADVICE r1 found), and each Julia type is one that may carry the C parameter's type;
* every name the shim exports is defined in it, and every shim name the two run scripts call is exported;
* the struct mirrors (`PtParams`, `Fields`, `StepParams`) repeat the C structs field for field (names, order, types).
"""
import os
import re

import pytest

from oracle.jl_interp import _BLOCK_OPEN, tokenize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "julia", "NS3DNative.jl")
SCRIPTS = [os.path.join(ROOT, "scripts", "NavierStokes3D_b200.jl"), os.path.join(ROOT, "scripts", "NavierStokes3D_gpu_b200.jl"),
           os.path.join(ROOT, "scripts", "NavierStokes3D_multi_gpu_b200.jl")]
HEADER = os.path.join(ROOT, "include", "ns3d.h")


def read(p):
    with open(p, encoding="utf-8") as fh:
        return fh.read()


@pytest.mark.parametrize("path", [SHIM] + SCRIPTS, ids=os.path.basename)
def test_lexes_and_balances(path):
    toks = tokenize(read(path))
    stack, blocks = [], 0
    pairs = {")": "(", "]": "[", "}": "{"}
    prev = None
    for t in toks:
        if t.kind == "op" and t.val in "([{" and len(t.val) == 1:
            stack.append((t.val, t.line))
        elif t.kind == "op" and t.val in ")]}" and len(t.val) == 1:
            assert stack and stack[-1][0] == pairs[t.val], f"unbalanced {t.val!r} at line {t.line}"
            stack.pop()
        elif t.kind == "id" and not any(b[0] == "[" for b in stack):
            # `struct` after `mutable`, `for`/`if` inside a comprehension or generator (brackets/parens) open no block
            in_paren_generator = t.val in ("for", "if") and stack
            if t.val in _BLOCK_OPEN and not in_paren_generator and not (prev is not None and prev.kind == "op" and prev.val == "."):
                blocks += 1
            elif t.val == "end":
                blocks -= 1
                assert blocks >= 0, f"`end` without a block at line {t.line}"
        prev = t
    assert not stack, f"unclosed {stack[-1]}"
    assert blocks == 0, f"{blocks} block(s) without `end`"


def c_prototypes():
    text = re.sub(r"/\*.*?\*/", " ", read(HEADER), flags=re.S)
    protos = {}
    for m in re.finditer(r"NS3D_API\s+([\w\s\*]+?)\s*\b(ns3d_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        types = []
        for p in plist:
            arr = "[" in p
            p = re.sub(r"\[[^\]]*\]", "", p)
            ty = " ".join(p.split()[:-1]) if not p.rstrip().endswith("*") else p
            ty = ty.replace("const", "").strip()
            if re.search(r"\*\s*\w+$", p):          # "double* A" / "double *A": the star belongs to the type
                ty = re.sub(r"\s*\w+$", "", p).replace("const", "").strip()
            types.append(re.sub(r"\s+", "", ty) + ("*" if arr else ""))
        protos[name] = (re.sub(r"\s+", "", ret.replace("const", "")), types)
    return protos


def split_top(tokens):
    """Token lists between the top-level commas of a parenthesised argument list (tokens exclude the outer parens)."""
    items, cur, depth = [], [], 0
    for t in tokens:
        if t.kind == "op" and len(t.val) == 1 and t.val in "([{":
            depth += 1
        elif t.kind == "op" and len(t.val) == 1 and t.val in ")]}":
            depth -= 1
        if t.kind == "op" and t.val == "," and depth == 0:
            items.append(cur)
            cur = []
        else:
            cur.append(t)
    if cur:
        items.append(cur)
    return items


def ccalls(path):
    toks = [t for t in tokenize(read(path)) if t.kind != "nl"]
    out = []
    for i, t in enumerate(toks):
        if t.kind == "id" and t.val == "ccall" and toks[i + 1].val == "(":
            depth, j = 0, i + 1
            while True:
                if toks[j].kind == "op" and toks[j].val == "(":
                    depth += 1
                elif toks[j].kind == "op" and toks[j].val == ")":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            items = split_top(toks[i + 2:j])
            sym = next(x.val for x in items[0] if x.kind == "id" and x.val.startswith("ns3d_"))
            ret = "".join(str(x.val) for x in items[1])
            tt = items[2]
            assert tt[0].val == "(" and tt[-1].val == ")", f"line {t.line}: the argument types must be a tuple literal"
            types = ["".join(str(x.val) for x in it) for it in split_top(tt[1:-1])]
            args = items[3:]
            out.append((t.line, sym, ret, types, args))
    return out


def compatible(ctype: str, jtype: str) -> bool:
    scalar = {"int": {"Cint"}, "double": {"Cdouble"}, "size_t": {"Csize_t"}, "longlong": {"Clonglong"}}
    if ctype in scalar:
        return jtype in scalar[ctype]
    if ctype.endswith("*"):
        if not (jtype.startswith("Ptr{") or jtype.startswith("Ref{") or jtype == "Cstring"):
            return False
        base = ctype.rstrip("*")
        if ctype == "double*":
            return jtype in ("Ptr{Float64}", "Ref{Cdouble}", "Ref{Float64}", "Ptr{Cdouble}")
        if ctype == "double**":
            return jtype in ("Ref{Ptr{Float64}}", "Ptr{Ptr{Float64}}")
        if ctype == "double**" or ctype == "double*const*":
            return jtype in ("Ptr{Ptr{Float64}}", "Ref{Ptr{Float64}}")
        if ctype == "int*":
            return jtype in ("Ref{Cint}", "Ptr{Cint}")
        if ctype == "ns3d_ctx*":
            return jtype == "Ptr{Cvoid}"
        if ctype == "ns3d_ctx**":
            return jtype == "Ref{Ptr{Cvoid}}"
        if base == "char":
            return jtype in ("Cstring", "Ptr{UInt8}", "Ptr{Cchar}", "Ref{NTuple{128,UInt8}}", "Ptr{Cvoid}")
        if base == "void":
            return jtype == "Ptr{Cvoid}"
        return True      # struct pointers: Ref{PtParams} etc. (field counts checked separately)
    return False


def type_aliases():
    """`const P = Ptr{Float64}` and the like."""
    return dict(re.findall(r"^const\s+(\w+)\s*=\s*((?:Ptr|Ref)\{[\w{}]+\}|Cint|Cdouble)\s*$", read(SHIM), flags=re.M))


def test_every_ccall_matches_its_c_prototype():
    protos = c_prototypes()
    assert len(protos) > 60
    alias = type_aliases()
    def sub(t):
        return re.sub(r"\b(" + "|".join(alias) + r")\b", lambda m: alias[m.group(1)], t) if alias else t
    calls = [(ln, sym, sub(ret), [sub(t) for t in types], args) for ln, sym, ret, types, args in ccalls(SHIM)]
    assert len(calls) >= 40
    seen = set()
    for line, sym, ret, types, args in calls:
        assert sym in protos, f"line {line}: {sym} is not declared in include/ns3d.h"
        cret, ctypes_ = protos[sym]
        seen.add(sym)
        assert len(types) == len(ctypes_), f"line {line}: {sym} takes {len(ctypes_)} parameters, {len(types)} types given"
        assert len(args) == len(types), f"line {line}: {sym}: {len(types)} argument types, {len(args)} arguments"
        for a in args:
            assert not any(x.kind == "op" and x.val == "..." for x in a), f"line {line}: ccall takes no splatted arguments"
        for k, (c, j) in enumerate(zip(ctypes_, types)):
            assert compatible(c, j), f"line {line}: {sym} parameter {k + 1} is `{c}` in C, `{j}` in the ccall"
        want_ret = {"int": "Cint", "char*": "Cstring", "size_t": "Csize_t", "longlong": "Clonglong", "void*": "Ptr{Cvoid}"}[cret]
        assert ret == want_ret, f"line {line}: {sym} returns {cret}, ccall says {ret}"
    # the hot path's entry points are all bound
    for must in ("ns3d_create", "ns3d_zeros", "ns3d_h2d", "ns3d_d2h", "ns3d_update_tau", "ns3d_predict_V", "ns3d_set_cylinder_M",
                 "ns3d_set_cylinder_G", "ns3d_update_divV", "ns3d_update_dPrdtau", "ns3d_update_Pr", "ns3d_compute_res",
                 "ns3d_correct_V", "ns3d_advect", "ns3d_set_bc_Vel_M", "ns3d_set_bc_Pr_M", "ns3d_pt_solve", "ns3d_step",
                 "ns3d_update_halo", "ns3d_comm_init", "ns3d_box_d2h", "ns3d_gather_box"):
        assert must in seen, f"{must} is not bound by the shim"


def defined_names(text):
    names = set(re.findall(r"^\s*(?:function|macro)\s+(?:Base\.)?([\w!∇τ@]+)", text, flags=re.M))
    names |= set(re.findall(r"^\s*(?:mutable\s+)?struct\s+(\w+)", text, flags=re.M))
    names |= set(re.findall(r"^\s*(?:baremodule|module)\s+(\w+)", text, flags=re.M))
    names |= set(re.findall(r"^([\w!∇τ]+)\([^=\n]*\)\s*(?:where\s*\{[^}]*\}\s*)?=(?!=)", text, flags=re.M))
    for m in re.finditer(r"^\s*const\s+([^=\n]+)=", text, flags=re.M):
        names |= set(x.strip() for x in m.group(1).split(","))
    return names


def test_exports_are_defined_and_scripts_use_exported_names():
    text = read(SHIM)
    m = re.search(r"^export\s+(.*?)\n\n", text, flags=re.S | re.M)
    exports = [x.strip() for x in m.group(1).replace("\n", " ").split(",")]
    defs = defined_names(text)
    for e in exports:
        name = e[1:] if e.startswith("@") else e
        assert name in defs, f"exported but not defined: {e}"
    plain = {e for e in exports if not e.startswith("@")}
    macros = {e[1:] for e in exports if e.startswith("@")}
    for sp in SCRIPTS:
        stoks = tokenize(read(sp))
        local = defined_names(read(sp))
        for i, t in enumerate(stoks):
            if t.kind == "macro" and t.val in ("zeros", "parallel", "init_ns3d"):
                assert t.val in macros
            if t.kind == "id" and t.val.endswith("!") and stoks[i + 1].val == "(" and t.val not in ("push!", "copy!"):   # Base
                assert t.val in plain or t.val in local, f"{os.path.basename(sp)} line {t.line}: {t.val} is neither exported by the shim nor local"


def c_struct_fields(name):
    """[(field name, kind)] of a C struct of include/ns3d.h, kind in {int, double, ptr, <struct name>}."""
    text = re.sub(r"/\*.*?\*/", " ", read(HEADER), flags=re.S)
    m = re.search(r"typedef\s+struct\s+\w*\s*\{([^{}]*)\}\s*" + name + r"\s*;", text, flags=re.S)
    out = []
    for decl in re.sub(r"//[^\n]*", "", m.group(1)).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        ty, rest = decl.split(" ", 1)
        for item in rest.split(","):
            item = item.strip()
            out.append((item.lstrip("*"), "ptr" if item.startswith("*") else ty))
    return out


@pytest.mark.parametrize("jl,c", [("PtParams", "ns3d_pt_params"), ("Fields", "ns3d_fields"), ("StepParams", "ns3d_step_params")])
def test_struct_mirrors_match_the_c_structs_field_for_field(jl, c):
    text = read(SHIM)
    alias = type_aliases()
    m = re.search(r"^\s*(?:mutable\s+)?struct\s+" + jl + r"\b(.*?)^end", text, flags=re.S | re.M)
    assert m, f"struct {jl} not found"
    body = re.sub(r"#[^\n]*", "", m.group(1))
    got = [(n, alias.get(t, t)) for n, t in re.findall(r"([\w∇τ]+)\s*::\s*([\w{}]+)", body)]
    want = c_struct_fields(c)
    assert [n for n, _ in got] == [n for n, _ in want], f"{jl} vs {c}: field names / order differ"
    kinds = {"int": "Cint", "double": "Cdouble", "ptr": "Ptr{Float64}", "ns3d_pt_params": "PtParams"}
    for (n, jt), (_, ck) in zip(got, want):
        assert jt == kinds[ck], f"{jl}.{n}: {jt} in Julia, {ck} in C"
