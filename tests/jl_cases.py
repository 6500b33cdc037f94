"""Cases that pin the CPU oracle (and, through the fixtures, the CUDA library) to the reference's own
source text as executed by oracle/jl_interp.py.  Test infrastructure, shared by
tests/golden/make_jl_fixtures.py (which needs /root/reference) and tests/test_jl_reference.py (which
needs only the committed fixtures, and re-derives them live when the reference tree is present).

Two kinds of case:

* KERNEL_CASES -- one launch of one reference kernel (or of `set_bc_Vel!` / `set_bc_Pr!`) on seeded random
  fields: ragged grids, the smallest legal grid, back-tracking that clamps at every face, exact-integer
  displacements, a rotated elliptic obstacle, the float-equality guards on and off, both scripts.
* RUN_CASES -- whole runs of `run_navierstokes3D` (M) / `runme` (G) from the scripts' own parameter block,
  initial conditions and time loop.
"""
from __future__ import annotations

import hashlib

import numpy as np

# (kernel, variant, grid, option)
KERNEL_CASES = []
for _v in ("M", "G"):
    for _g in ((3, 3, 3), (7, 6, 5), (13, 9, 8)):
        for _k in ("update_τ!", "predict_V!", "update_∇V!", "update_dPrdτ!", "update_Pr!", "compute_res!", "correct_V!",
                   "set_bc_Vel!", "set_bc_Pr!"):
            KERNEL_CASES.append((_k, _v, _g, None))
        for _vs in (0.0, 0.3, 3.0):
            KERNEL_CASES.append(("advect!", _v, _g, _vs))
    KERNEL_CASES.append(("set_cylinder!", _v, (31, 19, 4), "script"))
    KERNEL_CASES.append(("set_cylinder!", _v, (40, 24, 3), "rotated"))
KERNEL_CASES.append(("set_bc_Vel!", "M", (7, 6, 5), "guard_off"))
KERNEL_CASES.append(("set_bc_Pr!", "M", (7, 6, 5), "guard_off"))

OUTPUTS = {
    "update_τ!": ["txx", "tyy", "tzz", "txy", "txz", "tyz"], "predict_V!": ["Vx", "Vy", "Vz"], "update_∇V!": ["divV"],
    "update_dPrdτ!": ["dPrdtau"], "update_Pr!": ["Pr"], "compute_res!": ["Rp"], "correct_V!": ["Vx", "Vy", "Vz"],
    "set_bc_Vel!": ["Vx", "Vy", "Vz"], "set_bc_Pr!": ["Pr"], "advect!": ["Vx", "Vy", "Vz", "C"],
    "set_cylinder!": ["C", "Vx", "Vy", "Vz"],
}

# (id, variant, nx, nt, literals replaced in the script text / passed to the oracle's parameter block)
RUN_CASES = [
    ("M31", "M", 31, 3, {}, {}),
    ("M63", "M", 63, 3, {}, {}),      # test/test3D.jl's size (its nt=1 is the degenerate first step: three steps here)
    ("G20", "G", 20, 2, {}, {}),
    ("G40", "G", 40, 2, {}, {}),
    # "what one edits in the source to run another case": a rotated, wider ellipse further downstream
    ("M40rot", "M", 40, 3, {"β": 0.3, "a_lx": 0.1, "ox_lx": -0.2}, {"beta": 0.3, "a_lx": 0.1, "ox_lx": -0.2}),
    # explicit ny, nz with dx = dy = dz, the way BASELINE's configs D / E name their grids (1023x511x511, 511^3)
    ("M20x14x9", "M", 20, 3, {"ny": 14, "nz": 9, "ly_lx": 14 / 20, "lz_lx": 9 / 20}, {"ny": 14, "nz": 9, "ly": 14 / 20, "lz": 9 / 20}),
]
RUN_FIELDS = ("Pr", "Vx", "Vy", "Vz", "C")

# Several ranks of ImplicitGlobalGrid: (id, local nx, ny, nz, dims, nt, literals of the text, kwargs
# of oracle.VirtualRanks).  The script's own rule keeps lz fixed while nz_g grows, which makes the grid anisotropic and
# the PT loop (dτ from max(dx,dy,dz)) diverge: the domain lengths are edited so that dx = dy = dz, as for the weak-scaling runs.
RANK_CASES = [
    ("z2", 24, 15, 15, (1, 1, 2), 2, {"nz": 15, "lz_lx": 28 / 24}, {"lz": 28 / 24}),
    ("z3", 24, 15, 15, (1, 1, 3), 2, {"nz": 15, "lz_lx": 41 / 24}, {"lz": 41 / 24}),
    ("x2", 24, 15, 15, (2, 1, 1), 1, {}, {}),
    ("y2z2", 24, 15, 15, (1, 2, 2), 1, {"ny": 15, "nz": 15, "ly_lx": 28 / 24, "lz_lx": 28 / 24}, {"ly": 28 / 24, "lz": 28 / 24}),
]
RANK_FIELDS = ("Pr", "dPrdtau", "C", "C_o", "Vx", "Vy", "Vz", "Vx_o", "Vy_o", "Vz_o", "divV", "txx", "tyy", "tzz", "txy", "txz", "tyz")
JL_NAME = {"dPrdtau": "dPrdτ", "divV": "∇V", "txx": "τxx", "tyy": "τyy", "tzz": "τzz", "txy": "τxy", "txz": "τxz", "tyz": "τyz"}
FULL_ARRAYS = {"M31", "G20", "M20x14x9"}      # the other runs are stored as digests + the test3D.jl samples


def case_id(case) -> str:
    k, v, g, opt = case
    return f"{v}.{k}.{g[0]}x{g[1]}x{g[2]}" + ("" if opt is None else f".{opt}")


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.asfortranarray(a).tobytes(order="F")).hexdigest()


def params_of(O, case):
    k, v, g, opt = case
    lit = {}
    if k == "set_cylinder!" and opt == "rotated":
        lit = dict(beta=0.4, a_lx=0.12, b_lx=0.05, ox_lx=-0.1, oy_lx=0.05)
    p = O.params_M(*g, **lit) if v == "M" else O.params_G(*g, **lit)
    if opt == "guard_off":
        p.inlet_guard = p.outlet_guard = False
    return p


def inputs_of(O, case):
    """Seeded random fields for a kernel case (inputs are regenerated, only outputs are stored)."""
    k, v, g, opt = case
    p = params_of(O, case)
    seed = int(hashlib.sha256(case_id(case).encode()).hexdigest()[:8], 16)
    rng = np.random.default_rng(seed)
    f = {name: np.asfortranarray(rng.uniform(-1.0, 1.0, size=shape)) for name, shape in O.shapes(p.nx, p.ny, p.nz).items()}
    if k == "advect!":
        for n in ("Vx", "Vy", "Vz"):
            f[n] *= opt
        if opt:   # exact-integer displacements (quirk 7, M:196): dt*v/dx == 1 and -2 exactly
            f["Vx"][1:3, :, :] = p.dx / p.dt
            f["Vy"][:, 1:2, :] = -2 * p.dy / p.dt
        for n in ("Vx", "Vy", "Vz", "C"):
            f[n + "_o"][...] = f[n]
    return p, f


def run_oracle(O, case, p, f):
    k = case[0]
    {"update_τ!": O.update_tau, "predict_V!": O.predict_V, "update_∇V!": O.update_divV, "update_dPrdτ!": O.update_dPrdtau,
     "update_Pr!": O.update_Pr, "compute_res!": O.compute_res, "correct_V!": O.correct_V, "set_bc_Vel!": O.set_bc_Vel,
     "set_bc_Pr!": O.set_bc_Pr, "advect!": O.advect, "set_cylinder!": O.set_cylinder}[k](p, f)


def run_interp(script, case, p, f):
    """The same launch through the reference's source text; `script` is the JuliaScript of the case's variant."""
    k, v, g, opt = case
    S = script
    if k == "update_τ!":
        args = [f[n] for n in ("txx", "tyy", "tzz", "txy", "txz", "tyz", "Vx", "Vy", "Vz")] + [p.mu, p.dx, p.dy, p.dz]
    elif k == "predict_V!":
        args = [f[n] for n in ("Vx", "Vy", "Vz", "txx", "tyy", "tzz", "txy", "txz", "tyz")] + [p.rho, p.g, p.dt, p.dx, p.dy, p.dz]
    elif k == "update_∇V!":
        args = [f["divV"], f["Vx"], f["Vy"], f["Vz"], p.dx, p.dy, p.dz]
    elif k == "update_dPrdτ!":
        args = [f["Pr"], f["dPrdtau"], f["divV"], p.rho, p.dt, p.dtau, p.damp, p.dx, p.dy, p.dz]
    elif k == "update_Pr!":
        args = [f["Pr"], f["dPrdtau"], p.dtau]
    elif k == "compute_res!":
        args = [f["Rp"], f["Pr"], f["divV"], p.rho, p.dt, p.dx, p.dy, p.dz]
    elif k == "correct_V!":
        args = [f["Vx"], f["Vy"], f["Vz"], f["Pr"], p.dt, p.rho, p.dx, p.dy, p.dz]
    elif k == "advect!":
        args = [f[n] for n in ("Vx", "Vx_o", "Vy", "Vy_o", "Vz", "Vz_o", "C", "C_o")] + [p.dt, p.dx, p.dy, p.dz]
    elif k == "set_cylinder!":
        head = [f["C"], f["Vx"], f["Vy"], f["Vz"], p.a2, p.b2, p.ox, p.oy, p.sinb, p.cosb]
        args = head + ([p.xco_g, p.yco_g, p.zco_g] if v == "M" else []) + [p.lx, p.ly, p.lz, p.dx, p.dy, p.dz]
    elif k == "set_bc_Vel!":
        if v == "M":   # the guard `xvo_g == -lx/2` (M:164) is evaluated by the script's own text
            xvo = p.xvo_g if opt != "guard_off" else p.xvo_g + p.dx
            args = [f["Vx"], f["Vy"], f["Vz"], xvo, p.lx, p.vin]
        else:
            args = [f["Vx"], f["Vy"], f["Vz"], np.linspace(0.0, 1.0, p.nz)]
        return S.call_def(S.defs[k], args)
    elif k == "set_bc_Pr!":
        if v == "M":
            xve = p.xve_g if opt != "guard_off" else p.xve_g - p.dx
            args = [f["Pr"], xve, p.lx, 0.0]
        else:
            args = [f["Pr"], p.dz, p.nz, p.g, p.rho]
        return S.call_def(S.defs[k], args)
    else:
        raise KeyError(k)
    S.launch(S.defs[k], args)


def run_case_interp(jl_run, rc):
    """A whole run through the scripts' text -> (fields dict with the oracle's names, iters, errs)."""
    rid, variant, nx, nt, text_lit, _ = rc
    fn = jl_run.run_M if variant == "M" else jl_run.run_G
    env, iters, errs, info = fn(nx, nt, literals=text_lit)
    return {n: env[n] for n in RUN_FIELDS}, iters, errs, env, info


def run_case_oracle(O, rc):
    rid, variant, nx, nt, _, lit = rc
    p = O.params_M(nx, **lit) if variant == "M" else O.params_G(nx, **lit)
    f, iters, errs = O.run(p, nt)
    return {n: f[n] for n in RUN_FIELDS}, iters, errs, p


def run_ranks_interp(jl_run, case):
    rid, nx, ny, nz, dims, nt, text_lit, _ = case
    res = jl_run.run_M_ranks(nx, nt, dims, literals=text_lit)
    return [({n: env[JL_NAME.get(n, n)] for n in RANK_FIELDS}, iters, errs) for env, iters, errs in res]


def run_ranks_oracle(O, case):
    rid, nx, ny, nz, dims, nt, _, kw = case
    V = O.VirtualRanks(nx, ny, nz, dims, **kw)
    for _ in range(nt):
        V.step()
    return [({n: f[n] for n in RANK_FIELDS}, V.iters, V.errs) for f in V.f]


def save_path_records(jl_run):
    """The scripts' `do_save` output from their text (whole function bodies, M:288-535 / G:13-172): M at nx = 31, nt = 3,
    nsave = 1 -> the Float32 frames `out_save/out_<A>_v_%04d.bin` (M:404-413, 515-523) as SHA-256 of the file bytes;
    G at nx = 20, nt = 2, nsave = 1 -> every `matwrite(name, Dict(...))` (G:89, 168-170) as keys + digests."""
    import os
    import tempfile
    cwd = os.getcwd()
    rec = {}
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            ret, env, lines = jl_run.run_M_whole(31, 3, literals={"nsave": 1}, do_save=True)
            files = {}
            for fn in sorted(os.listdir("out_save")):
                with open(os.path.join("out_save", fn), "rb") as fh:
                    files[fn] = hashlib.sha256(fh.read()).hexdigest()
            rec["M31"] = {"lines": list(lines), "files": files, "returned": [digest(a) for a in ret]}
            env, s, lines = jl_run.run_G_whole(20, 2, literals={"nsave": 1}, do_save=True)
            rec["G20"] = [{"file": f, "keys": sorted(dd), "digest": {k: (digest(v) if hasattr(v, "shape") else v) for k, v in dd.items()}}
                          for f, dd in s.matwrites]
        finally:
            os.chdir(cwd)
    return rec
