"""GPU parity, level 2: the fused pseudo-transient loop and whole time steps vs the CPU oracle.

Bars
* PARITY mode: identical PT iteration counts, identical err history, BIT-EXACT fields.
* FAST mode (Markstein-corrected reciprocal division): identical iteration counts; fields within
  1e-12 of the common scale (bit-equal in practice -- the test reports any differing value).
* FASTEST mode (1/(dx*dx) multiply + FMA): identical iteration counts; fields within the
  north-star tolerance 1e-10 relative to the field's scale (velocities: common velocity scale,
  SURVEY.md "Hard parts": Vz is rounding noise in variant M).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_FAST = 1e-12
TOL_FASTEST = 1e-10   # north_star: "fields within a stated FP64 relative tolerance (e.g. 1e-10 after N steps)"


def setup_for(ns, variant, nx, **kw):
    return ns.setup_multi_gpu(nx, **kw) if variant == "M" else ns.setup_gpu(nx, **kw)


def oracle_params(O, variant, nx, **kw):
    return O.params_M(nx, **kw) if variant == "M" else O.params_G(nx, **kw)


def rel_inf(a, b, scale=None):
    scale = np.abs(b).max() if scale is None else scale
    return np.abs(a - b).max() / max(scale, 1e-300)


def pt_problem(O, variant, grid, seed):
    """A seeded random PT state: smooth-ish Pr, random dPrdtau and divV."""
    nx, ny, nz = grid
    p = oracle_params(O, variant, nx, ny=ny, nz=nz)
    rng = np.random.default_rng(seed)
    f = O.alloc_fields(p)
    f["Pr"][...] = rng.uniform(-1, 1, size=f["Pr"].shape)
    f["dPrdtau"][...] = rng.uniform(-1, 1, size=f["dPrdtau"].shape)
    f["divV"][...] = rng.uniform(-1e-3, 1e-3, size=f["divV"].shape)
    return p, f


@pytest.mark.parametrize("variant", ["M", "G"])
@pytest.mark.parametrize("grid", [(3, 3, 3), (5, 4, 3), (37, 23, 19), (63, 38, 38)])
@pytest.mark.parametrize("zchunk", [0, 1, 5])
def test_fused_iteration_bit_exact(O, ns, ctx, variant, grid, zchunk):
    """n fused iterations == n x (update_dPrdτ!, update_Pr!, set_bc_Pr!) of the oracle, bit for bit."""
    p, f = pt_problem(O, variant, grid, 11)
    s = setup_for(ns, variant, grid[0], ny=grid[1], nz=grid[2])
    d = {k: ctx.from_host(f[k]) for k in ("Pr", "dPrdtau", "divV")}
    done = 0
    for n in (1, 2, 37):   # odd and even counts exercise both ping-pong outcomes
        ctx.pt_iterate(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params(zchunk), n)
        for _ in range(n):
            O.update_dPrdtau(p, f)
            O.update_Pr(p, f)
            O.set_bc_Pr(p, f)
        done += n
        for name in ("Pr", "dPrdtau"):
            got = d[name].to_host()
            assert (got == f[name]).all(), f"{name} differs after {done} iterations ({(got != f[name]).sum()} values)"


PTV_CONFIGS = {
    # name: options of the pitched-layout kernel (ptv_kernel): iterations per launch, launch-bounds variant,
    # thread columns / rows of a tile (0 = chosen by the library), staging slots, TMA or plain-load staging
    "auto": {},
    "k1": {"ptv_k": 1},
    "k2_lb0": {"ptv_k": 2, "ptv_lb": 0},
    "k2_lb4_ns3": {"ptv_k": 2, "ptv_lb": 4, "ptv_ns": 3},
    "k3_lb0": {"ptv_k": 3, "ptv_lb": 0, "ptv_ns": 5},
    "k2_tiles": {"ptv_k": 2, "ptv_pxt": 5, "ptv_bty": 6},      # several small tiles in x and y
    "k3_tiles": {"ptv_k": 3, "ptv_pxt": 6, "ptv_bty": 7},
    "k2_coop_nographs": {"ptv_k": 2, "ptv_tma": 0, "graphs": 0, "serpentine": 1},
    "k1_coop_tiles": {"ptv_k": 1, "ptv_tma": 0, "ptv_pxt": 4, "ptv_bty": 3},
    # the PERSISTENT launch (ptv_flow_kernel, option ptv_flow: all full passes of a call are ONE kernel, work queue +
    # per-chunk completion counters between dependent passes); off by default -- measured slower, DESIGN.md section 3.4
    "k2_flow": {"ptv_k": 2, "ptv_flow": 1},
    "k2_flow_lb0": {"ptv_k": 2, "ptv_lb": 0, "ptv_flow": 1},
    "k3_flow_tiles": {"ptv_k": 3, "ptv_lb": 0, "ptv_pxt": 6, "ptv_bty": 7, "ptv_flow": 1},
    "k1_flow": {"ptv_k": 1, "ptv_flow": 1},
    # z-bands (ptv_run_banded: a pass as several launches on as many streams, band b waiting for bands b-1, b, b+1 of the
    # previous pass; on by default for grids of this size -- here with explicit band counts, and switched off)
    "k2_bands3": {"ptv_k": 2, "ptv_bands": 3},
    "k3_bands5_tiles": {"ptv_k": 3, "ptv_lb": 0, "ptv_pxt": 6, "ptv_bty": 7, "ptv_bands": 5},
    "k2_bands16_nographs": {"ptv_k": 2, "ptv_bands": 16, "graphs": 0},
    "k2_nobands": {"ptv_k": 2, "ptv_bands": 0},
}


@pytest.mark.parametrize("variant", ["M", "G"])
@pytest.mark.parametrize("grid", [(3, 3, 3), (5, 4, 3), (4, 3, 6), (16, 9, 8), (37, 23, 19), (63, 38, 38), (70, 47, 41)])
@pytest.mark.parametrize("zchunk", [0, 1, 2, 7])
@pytest.mark.parametrize("cfg", sorted(PTV_CONFIGS))
def test_ptv_kernel_bit_exact(O, ns, ctx, variant, grid, zchunk, cfg):
    """The default fused loop: ptv_kernel, K PT iterations per launch on the library's pitched copies (two columns
    per thread, 128-bit accesses, K-stage pipeline over z with the intermediate iterates in registers and shared
    memory).  Same per-cell arithmetic as the reference's K5 + K6 + set_bc_Pr! -> bit-equal to the oracle for every
    K, rows per thread, tile shape (one tile, several tiles with rims, even and odd nx) and chunking; counts that
    are not multiples of K end with shorter launches; 40 iterations replay a captured CUDA graph."""
    p, f = pt_problem(O, variant, grid, 15)
    s = setup_for(ns, variant, grid[0], ny=grid[1], nz=grid[2])
    for name, val in PTV_CONFIGS[cfg].items():
        ctx.set_option(name, val)
    d = {k: ctx.from_host(f[k]) for k in ("Pr", "dPrdtau", "divV")}
    done = 0
    for n in (2, 1, 5, 40):
        ctx.pt_iterate(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params(zchunk), n)
        for _ in range(n):
            O.update_dPrdtau(p, f)
            O.update_Pr(p, f)
            O.set_bc_Pr(p, f)
        done += n
        for name in ("Pr", "dPrdtau"):
            got = d[name].to_host()
            bad = np.argwhere(got != f[name])
            assert len(bad) == 0, f"{name} differs after {done} iterations: {len(bad)} values, first {bad[:3].tolist()}"


@pytest.mark.parametrize("variant,nx,nt", [("M", 63, 4), ("G", 40, 2)])
def test_time_steps_with_two_iterations_per_launch(O, ns, variant, nx, nt):
    p = oracle_params(O, variant, nx)
    f, iters_o, errs_o = O.run(p, nt)
    c = ns.Context(0, ns.PARITY)
    sim = ns.Simulation(setup_for(ns, variant, nx), c)
    for _ in range(nt):
        sim.step()
    assert sim.iters == iters_o and sim.err_hist == errs_o
    for name in ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C"):
        assert (sim.host(name) == f[name]).all(), name
    c.close()


def test_outlet_guard_off(O, ns, ctx):
    """Variant M with the float == guard false (quirk 5): plain Neumann outlet."""
    p, f = pt_problem(O, "M", (20, 12, 12), 12)
    p.outlet_guard = False
    s = setup_for(ns, "M", 20, ny=12, nz=12)
    s.outlet_guard = False
    d = {k: ctx.from_host(f[k]) for k in ("Pr", "dPrdtau", "divV")}
    ctx.pt_iterate(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params(), 7)
    for _ in range(7):
        O.update_dPrdtau(p, f)
        O.update_Pr(p, f)
        O.set_bc_Pr(p, f)
    assert (d["Pr"].to_host() == f["Pr"]).all()


@pytest.mark.parametrize("variant", ["M", "G"])
def test_pt_solve_matches_oracle(O, ns, ctx, variant):
    """Full loop with residual checks on a converging problem: same iterations, same err history."""
    grid = (31, 19, 19)
    p, f = pt_problem(O, variant, grid, 13)
    s = setup_for(ns, variant, grid[0], ny=grid[1], nz=grid[2])
    f["Pr"][...] = 0.0
    f["dPrdtau"][...] = 0.0
    if variant == "G":   # start from the hydrostatic state so that the loop converges
        f["Pr"][...] = O.initial_fields(p)["Pr"]
    d = {k: ctx.from_host(f[k]) for k in ("Pr", "dPrdtau", "divV")}
    it_o, hist_o = O.pt_solve(p, f)
    it_g, hist_g = ctx.pt_solve(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params())
    assert it_g == it_o and it_o >= p.nchk
    assert hist_g == hist_o
    assert (d["Pr"].to_host() == f["Pr"]).all()
    assert (d["dPrdtau"].to_host() == f["dPrdtau"]).all()


def test_pt_solve_nonfinite_breaks(O, ns, ctx):
    """!isfinite(err) leaves the loop at the first check (M:469)."""
    p, f = pt_problem(O, "M", (12, 9, 9), 14)
    f["Pr"][5, 4, 4] = np.nan
    s = setup_for(ns, "M", 12, ny=9, nz=9)
    d = {k: ctx.from_host(f[k]) for k in ("Pr", "dPrdtau", "divV")}
    it_g, hist_g = ctx.pt_solve(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params())
    it_o, hist_o = O.pt_solve(p, f)
    assert it_g == it_o == p.nchk
    assert len(hist_g) == 1 and np.isnan(hist_g[0]) and np.isnan(hist_o[0])


@pytest.mark.parametrize("variant,nx,nt", [("M", 63, 6), ("G", 63, 3), ("M", 40, 4), ("G", 40, 2), ("M", 31, 6)])
@pytest.mark.parametrize("level1", [False, True])
def test_time_steps_parity_mode(O, ns, variant, nx, nt, level1):
    """Config A (test/test3D.jl's grid, nt extended: step 1 of variant M is degenerate) end to end.

    40x24x24 (dx=dy=dz) is a second stable grid.  31x19x19 is UNSTABLE in the reference's own
    numerics (the PT loop hits niter at step 4 and the fields overflow to Inf/NaN by step 5):
    it is kept to pin that kernel and oracle agree bit-for-bit even on a diverged state
    (saturating floor(Int, .) in backtrack!, NaN-propagating maximum)."""
    if level1 and nx == 63:
        pytest.skip("the call-by-call level-1 loop is exercised on the small grids only (launch bound)")
    p = oracle_params(O, variant, nx)
    f, iters_o, errs_o = O.run(p, nt)
    s = setup_for(ns, variant, nx)
    sim = ns.Simulation(s, ns.Context(0, ns.PARITY))
    for _ in range(nt):
        sim.step_level1() if level1 else sim.step()
    assert sim.iters == iters_o
    assert np.array_equal(np.concatenate(sim.err_hist), np.concatenate(errs_o), equal_nan=True)
    for name in ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C", "divV"):
        got = sim.host(name)
        same = (got == f[name]) | (np.isnan(got) & np.isnan(f[name]))
        assert same.all(), f"{name}: {(~same).sum()} values differ"
    if nx != 31:
        assert all(np.isfinite(f[v]).all() for v in ("Pr", "Vx", "Vy", "Vz", "C"))
    if variant == "M":
        assert iters_o[0] == p.nchk and errs_o[0] == [0.0]   # quirk 2: step 1 is degenerate, Pr == 0
    sim.ctx.close()


@pytest.mark.parametrize("variant", ["M", "G"])
@pytest.mark.parametrize("mode,tol", [("FAST", TOL_FAST), ("FASTEST", TOL_FASTEST)])
def test_time_steps_fast_modes(O, ns, variant, mode, tol):
    nx, nt = 63, 4
    p = oracle_params(O, variant, nx)
    f, iters_o, errs_o = O.run(p, nt)
    s = setup_for(ns, variant, nx)
    sim = ns.Simulation(s, ns.Context(0, getattr(ns, mode)))
    for _ in range(nt):
        sim.step()
    assert sim.iters == iters_o, "PT iteration counts must be identical in every mode"
    vscale = max(np.abs(f[v]).max() for v in ("Vx", "Vy", "Vz"))
    report = {}
    for name in ("Pr", "C"):
        report[name] = rel_inf(sim.host(name), f[name])
    for name in ("Vx", "Vy", "Vz"):
        report[name] = rel_inf(sim.host(name), f[name], vscale)
    print(f"{variant} {mode}: max rel diff after {nt} steps: {report}")
    assert max(report.values()) <= tol, report
    sim.ctx.close()


def test_large_grid_linearity(ns):
    """BASELINE config B size (255x153x153), where the oracle is too slow for many iterations:
    a size-independent property instead.  The variant-M iteration map (Neumann faces, outlet
    value 0) is linear in (Pr, dPrdtau, divV), so T(a) - T(b) == T(a - b) up to rounding.  A
    wrong index, a missed boundary mirror or a chunk seam would break it at O(1)."""
    ctx = ns.Context(0, ns.PARITY)
    sm = ns.setup_multi_gpu(255)
    rng = np.random.default_rng(21)
    shp = sm.shapes()
    a = {k: np.asfortranarray(rng.uniform(-1, 1, size=shp[k])) for k in ("Pr", "dPrdtau", "divV")}
    b = {k: np.asfortranarray(rng.uniform(-1, 1, size=shp[k])) for k in ("Pr", "dPrdtau", "divV")}

    def T(x):
        d = {k: ctx.from_host(v) for k, v in x.items()}
        ctx.pt_iterate(d["Pr"], d["dPrdtau"], d["divV"], sm.pt_params(), 3)
        out = d["Pr"].to_host(), d["dPrdtau"].to_host()
        for v in d.values():
            ctx.free(v)
        return out

    Ta, Tb, Tab = T(a), T(b), T({k: a[k] - b[k] for k in a})
    for q in range(2):
        assert rel_inf(Ta[q] - Tb[q], Tab[q]) < 1e-9
    ctx.close()


def test_large_grid_oracle_spot_check(O, ns):
    """Config B size, 2 fused iterations vs the oracle (about a second of CPU work), bit-exact."""
    p, f = pt_problem(O, "G", (255, 153, 153), 22)
    s = ns.setup_gpu(255)
    ctx = ns.Context(0, ns.PARITY)
    d = {k: ctx.from_host(f[k]) for k in ("Pr", "dPrdtau", "divV")}
    ctx.pt_iterate(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params(), 2)
    for _ in range(2):
        O.update_dPrdtau(p, f)
        O.update_Pr(p, f)
        O.set_bc_Pr(p, f)
    assert (d["Pr"].to_host() == f["Pr"]).all()
    assert (d["dPrdtau"].to_host() == f["dPrdtau"]).all()
    ctx.close()


@pytest.mark.parametrize("variant,grid", [("M", (511, 511, 9)), ("G", (1023, 511, 7)), ("G", (255, 153, 20))])
def test_default_tiles_on_the_benchmark_planes(O, ns, variant, grid):
    """Grids with the x-y extent of the reference scripts' and BASELINE.json's configurations (255x153, 511x511,
    1023x511), thin in z so that the oracle finishes in seconds: the default launch configuration (compile-time
    32 x 16-cell tiles, TMA staging) and the general-geometry instantiation; 5 iterations = two double launches +
    one single.  Bit-exact, and identical to each other."""
    p, f = pt_problem(O, variant, grid, 23)
    s = setup_for(ns, variant, grid[0], ny=grid[1], nz=grid[2])
    ctx = ns.Context(0, ns.PARITY)
    out = {}
    for generic in (0, 1):
        if generic:
            ctx.set_option("ptv_pxt", 17)   # no compile-time instantiation for 34-cell-wide tiles
            ctx.set_option("ptv_bty", 15)
        d = {k: ctx.from_host(f[k]) for k in ("Pr", "dPrdtau", "divV")}
        ctx.pt_iterate(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params(), 5)
        out[generic] = (d["Pr"].to_host(), d["dPrdtau"].to_host())
    for _ in range(5):
        O.update_dPrdtau(p, f)
        O.update_Pr(p, f)
        O.set_bc_Pr(p, f)
    assert (out[0][0] == f["Pr"]).all() and (out[0][1] == f["dPrdtau"]).all()
    assert (out[0][0] == out[1][0]).all() and (out[0][1] == out[1][1]).all()
    ctx.close()


def test_config_B_whole_time_steps_vs_oracle(O, ns):
    """BASELINE configs[1] at FULL size -- 255x153x153, variant G (scripts/NavierStokes3D_gpu.jl as
    shipped, G:119-142), nt = 2, about 6 500 PT iterations: about 20 s of oracle on 16 host threads.
    PARITY: bit-exact fields, identical PT iteration counts and identical `err` history.
    FAST: the same, bit for bit (it is the bench default).  FASTEST: identical counts, fields within
    1e-10 of the common scale after these 2 steps (north_star's tolerance, horizon stated)."""
    nt = 2
    O.lib().ns3d_oracle_set_num_threads(len(os.sched_getaffinity(0)))
    p = O.params_G(255)
    f, iters_o, errs_o = O.run(p, nt)
    assert (p.nx, p.ny, p.nz) == (255, 153, 153) and all(it > 1000 for it in iters_o)
    fields = ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C")
    vscale = max(np.abs(f[v]).max() for v in ("Vx", "Vy", "Vz"))
    for mode in ("PARITY", "FAST", "FASTEST"):
        c = ns.Context(0, getattr(ns, mode))
        sim = ns.Simulation(ns.setup_gpu(255), c)
        for _ in range(nt):
            sim.step()
        assert sim.iters == iters_o, (mode, sim.iters, iters_o)
        if mode != "FASTEST":
            assert sim.err_hist == errs_o, mode
            for name in fields:
                got = sim.host(name)
                assert (got == f[name]).all(), f"{mode}: {name}: {(got != f[name]).sum()} values differ"
        else:
            # the solution fields; dPrdtau is the loop's internal rate (about 1e-3 of Pr/dtau near convergence),
            # for which a tolerance relative to its own magnitude says nothing about the solution
            for name in ("Pr", "Vx", "Vy", "Vz", "C"):
                scale = vscale if name[0] == "V" else None
                assert rel_inf(sim.host(name), f[name], scale) <= TOL_FASTEST, (mode, name)
        c.close()


def test_config_B_whole_time_steps_vs_the_shipped_scripts_text(ns):
    """The benchmark configuration against the reference's own TEXT: `runme()` of scripts/NavierStokes3D_gpu.jl as
    shipped, executed line by line by oracle/jl_interp.py (tests/golden/make_jl_config_B.py; record under
    tests/golden/jl_reference_config_B.json -- no oracle, no library involved).  PARITY and FAST (the bench default):
    identical PT iteration counts, every residual of every check and the five fields bit for bit after each recorded
    time step."""
    import hashlib
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jl_reference_config_B.json")) as fh:
        rec = json.load(fh)
    assert rec["grid"] == [255, 153, 153] and rec["steps"]
    for mode in ("PARITY", "FAST"):
        c = ns.Context(0, getattr(ns, mode))
        sim = ns.Simulation(ns.setup_gpu(255), c)
        for s in rec["steps"]:
            it, hist = sim.step()
            assert it == s["iters"], (mode, s["it"], it)
            assert hist == s["errs"], (mode, s["it"])
            for name in ("Pr", "Vx", "Vy", "Vz", "C"):
                got = hashlib.sha256(np.asfortranarray(sim.host(name)).tobytes(order="F")).hexdigest()
                assert got == s["digest"][name], f"{mode}: step {s['it']}: {name} differs from the script's text"
        c.close()


def test_config_B_variant_M_vs_the_multi_gpu_scripts_text(ns):
    """The same grid with the multi-GPU script (the base of the weak-scaling curve): `run_navierstokes3D(nx=255)` of
    scripts/NavierStokes3D_multi_gpu.jl on one rank, three time steps executed from its text
    (tests/golden/jl_reference_config_B_M.json: 152, 2 280, 2 280 PT iterations; the CPU oracle reproduces the record bit
    for bit, checked when it was made).  PARITY: counts, residuals, five fields bit for bit after every step."""
    import hashlib
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jl_reference_config_B_M.json")) as fh:
        rec = json.load(fh)
    assert rec["grid"] == [255, 153, 153] and len(rec["steps"]) == 3
    c = ns.Context(0, ns.PARITY)
    sim = ns.Simulation(ns.setup_multi_gpu(255), c)
    for s in rec["steps"]:
        it, hist = sim.step()
        assert it == s["iters"], (s["it"], it)
        assert hist == s["errs"], s["it"]
        for name in ("Pr", "Vx", "Vy", "Vz", "C"):
            got = hashlib.sha256(np.asfortranarray(sim.host(name)).tobytes(order="F")).hexdigest()
            assert got == s["digest"][name], f"step {s['it']}: {name} differs from the script's text"
    c.close()

