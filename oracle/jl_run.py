"""jl_run.py -- runs the reference's two run functions from their source text (oracle/jl_interp.py).

TEST INFRASTRUCTURE ONLY.  `run_M` executes, line by line as written in
scripts/NavierStokes3D_multi_gpu.jl, the parameter / allocation / initialisation block of
`run_navierstokes3D` (M:288-373) and then the body of its time loop (M:446-477) `nt` times on a
single rank; `run_G` does the same for `runme` of scripts/NavierStokes3D_gpu.jl (G:13-88,
G:119-142).  The plotting / saving / gathering code between and behind those ranges is not
executed (it does not touch the fields).  The line ranges are found by anchors in the text and
reported, so a test can assert that they are the cited ones.
"""
from __future__ import annotations

import os

from .jl_interp import JuliaScript

REFERENCE_ROOT = os.environ.get("NS3D_REFERENCE_ROOT", "/root/reference")
M_PATH = os.path.join(REFERENCE_ROOT, "scripts", "NavierStokes3D_multi_gpu.jl")
G_PATH = os.path.join(REFERENCE_ROOT, "scripts", "NavierStokes3D_gpu.jl")

FIELDS = ("Pr", "dPrdτ", "C", "C_o", "τxx", "τyy", "τzz", "τxy", "τxz", "τyz", "Vx", "Vy", "Vz",
          "Vx_o", "Vy_o", "Vz_o", "∇V", "Rp")
# the oracle's ASCII names for the same arrays
ASCII = {"dPrdτ": "dPrdtau", "τxx": "txx", "τyy": "tyy", "τzz": "tzz", "τxy": "txy", "τxz": "txz", "τyz": "tyz",
         "∇V": "divV"}


def reference_available() -> bool:
    return os.path.exists(M_PATH) and os.path.exists(G_PATH)


def _run(script: JuliaScript, prefix, loop, env, nt, on_step=None):
    script.run_lines(prefix[0], prefix[1], env)
    body = script.parse_lines(loop[0], loop[1], close_blocks=1)
    assert len(body) == 1 and body[0][0] == "for" and body[0][1] == "it", "time loop not found"
    iters, errs = [], []
    for it in range(1, nt + 1):
        env["it"] = it
        script.exec_block(body[0][3], env, host=True)
        iters.append(int(env["iter"]))
        errs.append([float(e) for e in env["err_evo"]])
        if on_step is not None:
            on_step(it, env)
    return iters, errs


def run_M(nx: int, nt: int, path: str = M_PATH, on_step=None, literals: dict | None = None):
    """`run_navierstokes3D(; nx, nt)` on one rank.  Returns (env, iters per step, err history per step, line ranges).
    `literals` replaces literals of the text (every assignment to such a name keeps the given value): what one
    edits in the source to run another case."""
    s = JuliaScript.from_file(path)
    s.frozen = dict(literals or {})
    head = s.find_line(r"function run_navierstokes3D\(")
    prefix = (head + 1, s.find_line(r"# Initialization for saving results", head) - 1)
    first = s.find_line(r"^\s*for it = 1:nt", prefix[1])
    loop = (first, s.find_line(r"^\s*# Visualization", first) - 1)
    env = {"nx": nx, "nt": nt, "do_vis": False, "do_save": False, "do_print": False}
    iters, errs = _run(s, prefix, loop, env, nt, on_step)
    return env, iters, errs, {"prefix": prefix, "loop": loop, "script": s}


def run_G(nx: int, nt: int, path: str = G_PATH, on_step=None, literals: dict | None = None):
    """`runme()` with the literals `nx = 255` (G:44) and `nt = 10000` (G:51) replaced by the arguments."""
    s = JuliaScript.from_file(path)
    s.frozen = {"nx": nx, "nt": nt, **(literals or {})}
    head = s.find_line(r"function runme\(")
    prefix = (head + 1, s.find_line(r"^\s*if do_save !ispath", head) - 1)
    first = s.find_line(r"^\s*for it = 1:nt", prefix[1])
    loop = (first, s.find_line(r"^\s*if do_vis && it % nvis == 0", first) - 1)
    env = {"do_vis": False, "do_save": False}
    iters, errs = _run(s, prefix, loop, env, nt, on_step)
    return env, iters, errs, {"prefix": prefix, "loop": loop, "script": s}
