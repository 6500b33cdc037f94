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
import threading

import numpy as np

from .jl_interp import Comm, JuliaScript

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
    with np.errstate(all="ignore"):     # IEEE semantics, like Julia: overflow / invalid produce Inf / NaN silently
        if prefix is not None:
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


def _anchors_M(s: JuliaScript):
    head = s.find_line(r"function run_navierstokes3D\(")
    prefix = (head + 1, s.find_line(r"# Initialization for saving results", head) - 1)
    first = s.find_line(r"^\s*for it = 1:nt", prefix[1])
    loop = (first, s.find_line(r"^\s*# Visualization", first) - 1)
    # global arrays for the return value (M:375-403) and the gathers behind the loop (M:528-532)
    out_alloc = (prefix[1] + 1, s.find_line(r"^\s*if do_save", prefix[1]) - 1)
    ret_first = s.find_line(r"# gather local arrays without halo for return call", loop[1]) + 1
    ret = (ret_first, s.find_line(r"^\s*C_v = Array\(C_v\)", ret_first) - 1)
    return {"prefix": prefix, "loop": loop, "out_alloc": out_alloc, "ret": ret}


def run_M(nx: int, nt: int, path: str = M_PATH, on_step=None, literals: dict | None = None, returns: bool = False):
    """`run_navierstokes3D(; nx, nt)` on one rank.  Returns (env, iters per step, err history per step, line ranges).
    `literals` replaces literals of the text (every assignment to such a name keeps the given value): what one
    edits in the source to run another case.  With `returns`, the allocation of the global arrays (M:375-403)
    and the gathers behind the loop (M:528-532) are executed too: env["C_v"], ["Pr_v"], ["Vx_v"], ["Vy_v"],
    ["Vz_v"] are then what the function returns (M:535) -- what test/test3D.jl:6 receives."""
    s = JuliaScript.from_file(path)
    s.frozen = dict(literals or {})
    a = _anchors_M(s)
    env = {"nx": nx, "nt": nt, "do_vis": False, "do_save": False, "do_print": False}
    if returns:
        s.run_lines(a["prefix"][0], a["prefix"][1], env)
        s.run_lines(a["out_alloc"][0], a["out_alloc"][1], env)
        iters, errs = _run(s, None, a["loop"], env, nt, on_step)
        s.run_lines(a["ret"][0], a["ret"][1], env)
    else:
        iters, errs = _run(s, a["prefix"], a["loop"], env, nt, on_step)
    return env, iters, errs, {**a, "script": s}


def run_M_ranks(nx: int, nt: int, dims, path: str = M_PATH, literals: dict | None = None):
    """The same text on a `dims` process grid of ImplicitGlobalGrid, one interpreter thread per rank (as if
    `init_global_grid(nx, ny, nz; dimx, dimy, dimz)` had been given): `update_halo!` at the text's call sites,
    `max_g` through `MPI.Allreduce`.  Returns [(env, iters, errs)] per rank in MPI rank order."""
    comm = Comm(dims)
    out, errors = [None] * comm.size, []

    def body(rank):
        try:
            s = JuliaScript.from_file(path)
            s.frozen = dict(literals or {})
            s.comm, s.rank = comm, rank
            a = _anchors_M(s)
            env = {"nx": nx, "nt": nt, "do_vis": False, "do_save": False, "do_print": False}
            iters, errs = _run(s, a["prefix"], a["loop"], env, nt)
            out[rank] = (env, iters, errs)
        except BaseException as e:                      # noqa: BLE001 -- release the other ranks, then report
            errors.append((rank, e))
            comm.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(comm.size)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        first = [e for e in errors if not isinstance(e[1], threading.BrokenBarrierError)] or errors
        raise RuntimeError(f"rank {first[0][0]}: {first[0][1]!r}") from first[0][1]
    return out


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


def run_M_whole(nx: int, nt: int, path: str = M_PATH, literals: dict | None = None, do_save: bool = False):
    """The WHOLE body of `run_navierstokes3D` (M:288-535) executed in one piece -- parameter block, allocation, initial
    gathers, the `do_save` frames (M:404-413, 515-523: `save_array` writes `out_save/out_<A>_v_%04d.bin` into the current
    directory), the time loop with its `if (do_vis && ...) || (do_save && ...)` branch, the final gathers and the `return`
    (plotting is parsed but, with do_vis=false, never reached).  Returns the function's return value."""
    from .jl_interp import _Return
    s = JuliaScript.from_file(path)
    s.frozen = dict(literals or {})
    head = s.find_line(r"function run_navierstokes3D\(")
    last = s.find_line(r"^\s*return C_v,Pr_v,Vx_v,Vy_v,Vz_v", head)
    env = {"nx": nx, "nt": nt, "do_vis": False, "do_save": do_save, "do_print": False}
    with np.errstate(all="ignore"):
        try:
            s.run_lines(head + 1, last, env)
        except _Return as r:
            return r.val, env, (head + 1, last)
    raise RuntimeError("the function body ended without `return`")


def run_G_whole(nx: int, nt: int, path: str = G_PATH, literals: dict | None = None, do_save: bool = False):
    """The WHOLE body of `runme` (G:13-172) in one piece, `nx`/`nt` literals replaced; with `do_save` the `.mat` dumps
    (G:89, 168-170) are recorded as (file name, Dict) in `script.matwrites` (`out_save/` is created in the current
    directory, as the script does)."""
    from .jl_interp import _Return
    s = JuliaScript.from_file(path)
    s.frozen = {"nx": nx, "nt": nt, **(literals or {})}
    head = s.find_line(r"function runme\(")
    last = s.find_line(r"^\s*return\s*$", head)
    env = {"do_vis": False, "do_save": do_save}
    with np.errstate(all="ignore"):
        try:
            s.run_lines(head + 1, last, env)
        except _Return:
            pass
    return env, s, (head + 1, last)
