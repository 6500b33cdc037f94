"""The Julia side of the drop-in boundary, EXECUTED (not by Julia -- there is none here -- but by oracle/jl_shim.py).

oracle/jl_shim.py interprets the text of julia/NS3DNative.jl (its structs, its typed methods, `Ref`/`Ptr`, every
`ccall((:sym, LIB), Ret, (Types...), args...)` converted by the declared Julia types and sent into the C ABI through
an untyped ctypes handle) and the text of the two re-pointed run scripts.  What is checked:

* scripts/NavierStokes3D_gpu_b200.jl and scripts/NavierStokes3D_b200.jl, through the shim, into the library, reproduce
  the fixtures that the REFERENCE's own text yields (tests/golden/jl_reference_fixtures.npz) bit for bit -- PT iteration
  counts, every residual, the fields, the returned interiors -- with the level-1 loop and with the fused entry points;
* every one of the shim's ccall sites is reached by some test (compared with the static list), incl. the level-2
  struct-passing calls (`step!`, `predictor!` ...), the device-side initialisers, the output path and the error path;
* a swapped argument in the shim's text is caught.

CPU suite: the emulated library (tests/emu).  `-m gpu`: the same against libns3d.so on the B200.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import jl_shim
from oracle.jl_interp import JlError
from tests import jl_cases as J
from tests.test_julia_shim_static import SHIM, ccalls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G_B200 = os.path.join(ROOT, "scripts", "NavierStokes3D_gpu_b200.jl")
M_B200 = os.path.join(ROOT, "scripts", "NavierStokes3D_b200.jl")
M_LOOKALIKE = os.path.join(ROOT, "scripts", "NavierStokes3D_multi_gpu_b200.jl")
REACHED = set()          # C symbols the shim's ccalls reached in this session (per library kind)


@pytest.fixture(scope="module")
def fx():
    z = np.load(os.path.join(ROOT, "tests", "golden", "jl_reference_fixtures.npz"))
    return z, json.loads(str(z["meta"]))


def emu_lib():
    from tests.emu import build_lib
    return C.CDLL(build_lib.build())


def gpu_lib():
    from navierstokes3d_b200 import native
    native.load()
    return C.CDLL(native.lib_path())


def note(shim):
    REACHED.update(s for s, _ in shim.ccalls)


# ---- the checks, independent of which library executes them -----------------------------------------
def check_gpu_script(lib, fx, fused):
    z, meta = fx
    f, iters, errs, (shim, scr) = jl_shim.run_gpu_b200(lib, SHIM, G_B200, 20, 2, use_fused=fused)
    note(shim)
    m = meta["run"]["G20"]
    assert iters == m["iters"] and errs == m["errs"]
    for n in J.RUN_FIELDS:
        assert J.digest(f[n]) == m["digest"][n], n
    return shim


def check_multi_script(lib, fx, fused):
    z, meta = fx
    ret, f, iters, (shim, scr, mpi) = jl_shim.run_multi_b200(lib, SHIM, M_B200, 31, 3, use_fused=fused)
    note(shim)
    m = meta["run"]["M31"]
    assert iters == m["iters"]
    for n in J.RUN_FIELDS:
        assert J.digest(f[n]) == m["digest"][n], n
    assert mpi.calls == ["Bcast!"]                   # the one use of MPI that remains: the 128-byte NCCL id
    assert len(ret) == 5                             # M:535: C_v, Pr_v, Vx_v, Vy_v, Vz_v
    for got, n in zip(ret, ("C", "Pr", "Vx", "Vy", "Vz")):
        assert np.array_equal(got, z[f"run/M31/{n}"][1:-1, 1:-1, 1:-1]), n
    return shim


def check_multi_lookalike_script(lib, fx, fused):
    """scripts/NavierStokes3D_multi_gpu_b200.jl: the M script's own text on the look-alike surface (`init_global_grid`,
    `update_halo!`, `max_g`, `gather!`, `finalize_global_grid` of the shim)."""
    z, meta = fx
    ret, f, iters, errs, (shim, scr, mpi) = jl_shim.run_multi_gpu_lookalike_b200(lib, SHIM, M_LOOKALIKE, 31, 3, use_fused=fused)
    note(shim)
    m = meta["run"]["M31"]
    assert iters == m["iters"] and errs == m["errs"]
    for n in J.RUN_FIELDS:
        assert J.digest(f[n]) == m["digest"][n], n
    assert mpi.calls == ["Bcast!", "Finalize"]
    for got, n in zip(ret, ("C", "Pr", "Vx", "Vy", "Vz")):
        assert np.array_equal(got, z[f"run/M31/{n}"][1:-1, 1:-1, 1:-1]), n
    n_halo = sum(1 for s, _ in shim.ccalls if s == "ns3d_update_halo")
    if not fused:       # the text's ten call sites: 2 at start-up, 5 + 3 per PT iteration in every time step
        assert n_halo == 2 + 3 * 5 + 3 * sum(m["iters"])
    return shim


def snippet(shim, src, env):
    s = jl_shim.ShimScript(src, name="snippet", lib=shim.lib, imports=[shim])
    s._pending_consts = []
    s.run_lines(1, src.count("\n") + 1, env)
    return env


def check_level2(lib, fx):
    """`step!` and the four groups with `Fields` / `StepParams` passed by reference, on the state the M script sets up."""
    z, meta = fx
    m = meta["run"]["M31"]
    for how in ("step", "groups"):
        shim = jl_shim.load_shim(SHIM, lib)
        scr = jl_shim.load_script(M_B200, shim)
        scr.globals["MPI"] = jl_shim.SingleRankMPI()
        env = {"do_vis": False, "do_save": False, "do_print": False, "nx": 31, "nt": 3}
        head = scr.find_line(r"function run_navierstokes3D\(")
        n_ctx = scr.find_line(r"^\s*ctx = Ctx\(", head)
        scr.run_lines(head + 1, n_ctx, env)
        scr.apply(shim.lookup("set_mode!", {}), [env["ctx"], 0], {})
        scr.run_lines(n_ctx + 1, scr.find_line(r"^\s*for it = 1:nt", head) - 1, env)
        src = ("f = Fields(Pr, dPrdτ, C, C_o, τxx, τyy, τzz, τxy, τxz, τyz, Vx, Vy, Vz, Vx_o, Vy_o, Vz_o, ∇V, Rp)\n"
               "sp = StepParams(pt, μ, vin, a2, b2, ox, oy, sinβ, cosβ, xco_g, yco_g, lx, ly, xvo_g == -lx/2, 0)\n"
               "all_iters = Float64[]\n"
               "for it = 1:nt\n")
        if how == "step":
            src += "    iters, hist = step!(ctx, f, sp)\n"
        else:
            src += ("    predictor!(ctx, f, sp)\n    iters, hist = pt_solve!(ctx, Pr, dPrdτ, ∇V, pt)\n"
                    "    corrector!(ctx, f, sp)\n    advect_swap!(ctx, f, sp)\n")
        src += "    push!(all_iters, iters)\nend\nhPr = to_host(ctx, Pr); hVx = to_host(ctx, Vx); hVy = to_host(ctx, Vy); hVz = to_host(ctx, Vz); hC = to_host(ctx, C)"
        snippet(shim, src, env)
        note(shim)
        assert env["all_iters"] == m["iters"], how
        for n in J.RUN_FIELDS:
            assert J.digest(env["h" + n]) == m["digest"][n], (how, n)
        jl_shim._finalize(shim, scr)


def check_surface(lib, O):
    """The rest of the shim: initialisers, output path, face kernels, error path, the look-alike grid functions."""
    shim = jl_shim.load_shim(SHIM, lib)
    rng = np.random.default_rng(11)
    n = (6, 5, 4)
    env = {"prof": rng.uniform(-1, 1, 4), "addy": rng.uniform(-1, 1, 5), "addz": rng.uniform(-1, 1, 4)}
    snippet(shim, "@init_ns3d(0, PARITY)\nc = default_ctx()\nA = @zeros(6,5,4)\nfill_profile_z!(c, A, prof)\nhA = Array(A)\n"
                  "B = @zeros(6,5,4)\nfill_profile_zy!(c, B, prof, addy, addz)\nhB = Array(B)\nfill_plane_x!(c, B, 2, 7.5)\nhB2 = Array(B)\n"
                  "in64 = inner(c, B)\nin32 = inner32(c, B)\npxy = plane_xy(c, B, 2)\npxz = plane_xz(c, B, 3)\nm = max_g(abs, B)\n"
                  "sz = size(B, 2)\nlen = length(B)", env)
    want = np.broadcast_to(env["prof"][None, None, :], n)
    assert np.array_equal(env["hA"], want)
    wantB = (env["prof"][None, None, :] + env["addy"][None, :, None]) + env["addz"][None, None, :]
    assert np.array_equal(env["hB"], np.broadcast_to(wantB, n))
    b2 = np.array(np.broadcast_to(wantB, n))
    b2[1, :, :] = 7.5
    assert np.array_equal(env["hB2"], b2)
    assert np.array_equal(env["in64"], b2[1:-1, 1:-1, 1:-1])
    assert env["in32"].dtype == np.float32 and np.array_equal(env["in32"], b2[1:-1, 1:-1, 1:-1].astype(np.float32))
    assert np.array_equal(env["pxy"], b2[1:-1, 1:-1, 2]) and np.array_equal(env["pxz"], b2[1:-1, 3, 1:-1])
    assert env["m"] == np.abs(b2).max() and env["sz"] == 5 and env["len"] == 120

    # face kernels of both scripts against the oracle
    p = O.params_G(7, 6, 5)
    f = {k: np.asfortranarray(rng.uniform(-1, 1, s)) for k, s in O.shapes(7, 6, 5).items()}
    env.update({"h" + k: f[k].copy(order="F") for k in ("Vx", "Vy", "Vz", "Pr")}, dz=p.dz, nz=p.nz, g=p.g, ρ=p.rho,
               xvo_g=-0.5, xve_g=0.5, lx=1.0, vin=1.0)
    snippet(shim, "Vx = Data.Array(hVx); Vy = Data.Array(hVy); Vz = Data.Array(hVz); Pr = Data.Array(hPr)\n"
                  "set_bc_Vel_G!(c, Vx, Vy, Vz, Pr)\nset_bc_Pr_G!(c, Pr, dz, nz, g, ρ)\n"
                  "gVx = Array(Vx); gVy = Array(Vy); gVz = Array(Vz); gPr = Array(Pr)\n"
                  "Vx2 = Data.Array(hVx); Vy2 = Data.Array(hVy); Vz2 = Data.Array(hVz); Pr2 = Data.Array(hPr)\n"
                  "set_bc_Vel_M!(c, Vx2, Vy2, Vz2, xvo_g, lx, vin, Pr2)\nset_bc_Pr_M!(c, Pr2, xve_g, lx, 0.0)\n"
                  "mVx = Array(Vx2); mVy = Array(Vy2); mVz = Array(Vz2); mPr = Array(Pr2)\n"
                  "Vx3 = Data.Array(hVx); Pr3 = Data.Array(hPr)\nbc_x_Vx!(c, Vx3, 1.5)\nbc_x_Pr!(c, Pr3, 0.25)\nxVx = Array(Vx3); xPr = Array(Pr3)", env)
    fg = {k: v.copy(order="F") for k, v in f.items()}
    O.set_bc_Vel(p, fg)
    O.set_bc_Pr(p, fg)
    for k in ("Vx", "Vy", "Vz", "Pr"):
        assert np.array_equal(env["g" + k], fg[k]), k
    pm = O.params_M(7, 6, 5)
    fm = {k: v.copy(order="F") for k, v in f.items()}
    O.set_bc_Vel(pm, fm)
    O.set_bc_Pr(pm, fm)
    for k in ("Vx", "Vy", "Vz", "Pr"):
        assert np.array_equal(env["m" + k], fm[k]), k
    fx_ = {k: v.copy(order="F") for k, v in f.items()}
    O.bc("x_Vx", fx_["Vx"], 1.5)
    O.bc("x_Pr", fx_["Pr"], 0.25)
    assert np.array_equal(env["xVx"], fx_["Vx"]) and np.array_equal(env["xPr"], fx_["Pr"])

    # the error path: `check` raises with the library's own message (ns3d_last_error)
    with pytest.raises(JlError, match="shape mismatch"):
        snippet(shim, "set!(c, A, zeros(2,2,2))", env)
    with pytest.raises(JlError, match="libns3d: .*bad shape"):
        snippet(shim, "Z = zeros3(c, -1, 2, 2)", env)

    # the ImplicitGlobalGrid look-alike on one rank
    snippet(shim, "finalize_global_grid_was = 0\nme, dims, nprocs, coords, comm = init_global_grid(9, 6, 5; mode=PARITY, quiet=true)\n"
                  "ng = (nx_g(), ny_g(), nz_g())\nQ = @zeros(10,6,5)\nxg = x_g(10, 0.5, Q)\nupdate_halo!(Q)\n"
                  "qv = gather!(Q)\nfinalize_global_grid()", env)
    assert (env["me"], env["dims"], env["nprocs"], env["coords"]) == (0, (1, 1, 1), 1, (0, 0, 0))
    assert env["ng"] == (9, 6, 5) and env["xg"] == (10 - 1) * 0.5 + 0.5 * (9 - 10) * 0.5 and env["qv"].shape == (8, 4, 3)
    note(shim)
    jl_shim._finalize(shim)


def check_swapped_argument_is_caught(lib, fx):
    """`correct_V!`'s ccall with `dt` and `ρ` exchanged: same types, same arity -- only an execution sees it."""
    z, meta = fx
    with open(SHIM, encoding="utf-8") as fh:
        text = fh.read()
    old = "c.h, Vx.p, Vy.p, Vz.p, Pr.p, dt, ρ, dx, dy, dz, nx, ny, nz))"
    assert text.count(old) == 1
    bad = os.path.join(ROOT, "tests", "emu", "_build", "NS3DNative_swapped.jl")
    with open(bad, "w", encoding="utf-8") as fh:
        fh.write(text.replace(old, old.replace("dt, ρ", "ρ, dt")))
    f, iters, errs, _ = jl_shim.run_gpu_b200(lib, bad, G_B200, 20, 1, use_fused=True)
    assert J.digest(f["Vx"]) != meta["run"]["G20"]["digest"]["Vx"]


# ---- CPU: the emulated library ---------------------------------------------------------------------------
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "level1"])
def test_gpu_script_through_the_shim(fx, fused):
    check_gpu_script(emu_lib(), fx, fused)


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "level1"])
def test_multi_gpu_script_through_the_shim(fx, fused):
    check_multi_script(emu_lib(), fx, fused)


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "level1"])
def test_multi_gpu_script_with_its_text_kept_through_the_shim(fx, fused):
    check_multi_lookalike_script(emu_lib(), fx, fused)


def test_level2_structs_through_the_shim(fx):
    check_level2(emu_lib(), fx)


def test_rest_of_the_shim_surface(O):
    check_surface(emu_lib(), O)


def test_swapped_ccall_argument_is_caught(fx):
    check_swapped_argument_is_caught(emu_lib(), fx)


def test_zz_every_ccall_site_was_reached(fx, O):
    """The union of the C symbols reached by the checks above is the set the shim's text binds (when this test runs
    alone the checks are executed here first)."""
    bound = {c[1] for c in ccalls(SHIM)}
    if bound - REACHED:
        lib = emu_lib()
        check_gpu_script(lib, fx, False)
        check_multi_script(lib, fx, True)
        check_level2(lib, fx)
        check_surface(lib, O)
    assert bound - REACHED == set(), f"ccall sites never executed: {sorted(bound - REACHED)}"


# ---- GPU: libns3d.so itself ----------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "level1"])
def test_gpu_script_through_the_shim_on_the_device(fx, fused):
    check_gpu_script(gpu_lib(), fx, fused)


@pytest.mark.gpu
def test_multi_gpu_script_through_the_shim_on_the_device(fx):
    try:
        check_multi_script(gpu_lib(), fx, True)
    except JlError as e:
        if "NCCL" in str(e) or "nccl" in str(e):
            pytest.skip(f"no NCCL in this process: {e}")
        raise


@pytest.mark.gpu
def test_multi_gpu_script_with_its_text_kept_on_the_device(fx):
    try:
        check_multi_lookalike_script(gpu_lib(), fx, True)
    except JlError as e:
        if "NCCL" in str(e) or "nccl" in str(e):
            pytest.skip(f"no NCCL in this process: {e}")
        raise


@pytest.mark.gpu
def test_level2_and_surface_through_the_shim_on_the_device(fx, O):
    check_surface(gpu_lib(), O)
    try:
        check_level2(gpu_lib(), fx)
    except JlError as e:
        if "NCCL" in str(e) or "nccl" in str(e):
            pytest.skip(f"no NCCL in this process: {e}")
        raise


# ---- the interpreter's additional language features on snippets with known values (no library involved) ----------
SRC_LANG = '''
module Demo
const A, B = Cint(3), 2.5
const P = Ptr{Float64}
struct Pt
    x::Cdouble; y::Cdouble
    tag::Cint
end
mutable struct Box
    p::Pt
    n::Int
end
Pt(x::Real) = Pt(x, x, 0)
norm1(p::Pt) = abs(p.x) + abs(p.y)
norm1(b::Box) = b.n * norm1(b.p)
pick(a::Integer, b=10; scale=1) = (a + b) * scale
pick(a::Pt, rest::Integer...) = length(rest)
kind(::Type{T}) where {T<:Union{Float64,Float32}} = T == Float32 ? 1 : 0
function total(xs::Vector{Float64}; start=0.0)
    acc = start
    for (i, x) in enumerate(xs)
        acc += i * x
        i == 3 && break
    end
    return acc, length(xs)
end
cell() = Ref{Cint}(7)
end # module
'''


def test_language_features_of_the_shim_interpreter():
    s = jl_shim.ShimScript(SRC_LANG)
    s.finish_loading()
    env = {}
    snippet_src = ("p = Pt(1.5, -2.0, 4)\nq = Pt(3)\nb = Box(p, 2)\nn1 = norm1(p); n2 = norm1(b); n3 = norm1(q)\n"
                   "k1 = pick(1); k2 = pick(1, 2); k3 = pick(1, 2; scale=3); k4 = pick(p, 1, 2, 3)\n"
                   "t64 = kind(Float64); t32 = kind(Float32)\nacc, n = total([1.0, 2.0, 3.0, 4.0]; start=0.5)\n"
                   "r = cell(); v0 = r[]; r[] = 9; v1 = r[]\nsq = map(x -> x * x, [1, 2, 3])\nd = 7 ÷ 2\n"
                   "same = p.tag === 4 ? :yes : :no\nz = [i * 2 for i in 1:3]\nty = Cint[i for i in 1:2]\nconsts = (A, B)")
    sn = jl_shim.ShimScript(snippet_src, imports=[s])
    sn.run_lines(1, snippet_src.count("\n") + 1, env)
    assert (env["n1"], env["n2"], env["n3"]) == (3.5, 7.0, 6.0)
    assert (env["k1"], env["k2"], env["k3"], env["k4"]) == (11, 3, 9, 3)
    assert (env["t64"], env["t32"]) == (0, 1)
    assert env["acc"] == 0.5 + 1 * 1.0 + 2 * 2.0 + 3 * 3.0 and env["n"] == 4
    assert (env["v0"], env["v1"]) == (7, 9) and env["sq"] == [1, 4, 9] and env["d"] == 3
    assert env["same"] == ("sym", "yes") and env["z"] == [2, 4, 6] and env["consts"] == (3, 2.5)
    assert env["ty"].dtype == np.int32 and list(env["ty"]) == [1, 2]
    with pytest.raises(JlError, match="MethodError"):
        sn2 = jl_shim.ShimScript("bad = norm1(1.0)", imports=[s])
        sn2.run_lines(1, 1, {})
    with pytest.raises(JlError, match="ccall without a library"):
        sn3 = jl_shim.ShimScript('rc = ccall((:ns3d_version, LIB), Cstring, ())', imports=[s])
        sn3.run_lines(1, 1, {"LIB": None})
