"""GPU parity, level 1: every reference kernel, CUDA (through the C ABI) vs the CPU oracle on the
same seeded random fields.  Bar: BIT-EXACT in PARITY mode (same IEEE operations in the same
order, no FMA contraction) -- compared with ``==`` so that +0.0/-0.0 are the only tolerated
difference, and none is expected.

Edge cases covered: the smallest legal grid (3^3: every interior point touches every face),
ragged sizes that are not multiples of the CTA tile (37x23x19), backtracking that clamps at
every array face (random velocities up to ~3 cells per step), exact-integer displacements
(quirk 7 of SURVEY.md), NaN/Inf propagation through maximum(abs.(A)).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GRIDS = [(3, 3, 3), (4, 5, 3), (37, 23, 19), (63, 38, 38)]


def rand_fields(O, p, seed):
    rng = np.random.default_rng(seed)
    f = {}
    for name, shape in O.shapes(p.nx, p.ny, p.nz).items():
        f[name] = np.asfortranarray(rng.uniform(-1.0, 1.0, size=shape))
    return f


def upload(ctx, f):
    return {k: ctx.from_host(v) for k, v in f.items() if k != "absRp"}


def assert_same(dev, host, names):
    for n in names:
        got = dev[n].to_host()
        assert got.shape == host[n].shape
        bad = np.flatnonzero(~(got == host[n]).ravel(order="F"))
        assert bad.size == 0, f"{n}: {bad.size} of {got.size} values differ, first at flat index {bad[:5]}"


def mk_params(O, variant, grid):
    nx, ny, nz = grid
    p = O.params_M(nx, ny, nz) if variant == "M" else O.params_G(nx, ny, nz)
    return p


@pytest.mark.parametrize("grid", GRIDS)
def test_update_tau(O, ctx, grid):
    p = mk_params(O, "M", grid)
    f = rand_fields(O, p, 1)
    d = upload(ctx, f)
    O.update_tau(p, f)
    ctx.call("ns3d_update_tau", d["txx"], d["tyy"], d["tzz"], d["txy"], d["txz"], d["tyz"], d["Vx"], d["Vy"], d["Vz"],
             p.mu, p.dx, p.dy, p.dz, p.nx, p.ny, p.nz)
    assert_same(d, f, ["txx", "tyy", "tzz", "txy", "txz", "tyz"])


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("variant", ["M", "G"])
def test_predict_V(O, ctx, grid, variant):
    p = mk_params(O, variant, grid)   # G has g = 9.81: the -rho*g term is live
    f = rand_fields(O, p, 2)
    d = upload(ctx, f)
    O.predict_V(p, f)
    ctx.call("ns3d_predict_V", d["Vx"], d["Vy"], d["Vz"], d["txx"], d["tyy"], d["tzz"], d["txy"], d["txz"], d["tyz"],
             p.rho, p.g, p.dt, p.dx, p.dy, p.dz, p.nx, p.ny, p.nz)
    assert_same(d, f, ["Vx", "Vy", "Vz"])


@pytest.mark.parametrize("grid", GRIDS)
def test_update_divV(O, ctx, grid):
    p = mk_params(O, "M", grid)
    f = rand_fields(O, p, 3)
    d = upload(ctx, f)
    O.update_divV(p, f)
    ctx.call("ns3d_update_divV", d["divV"], d["Vx"], d["Vy"], d["Vz"], p.dx, p.dy, p.dz, p.nx, p.ny, p.nz)
    assert_same(d, f, ["divV"])


@pytest.mark.parametrize("grid", GRIDS)
def test_pt_level1_kernels(O, ctx, grid):
    p = mk_params(O, "M", grid)
    f = rand_fields(O, p, 4)
    d = upload(ctx, f)
    n = (p.nx, p.ny, p.nz)
    O.update_dPrdtau(p, f)
    ctx.call("ns3d_update_dPrdtau", d["Pr"], d["dPrdtau"], d["divV"], p.rho, p.dt, p.dtau, p.damp, p.dx, p.dy, p.dz, *n)
    assert_same(d, f, ["dPrdtau"])
    O.update_Pr(p, f)
    ctx.call("ns3d_update_Pr", d["Pr"], d["dPrdtau"], p.dtau, *n)
    assert_same(d, f, ["Pr"])
    O.compute_res(p, f)
    ctx.call("ns3d_compute_res", d["Rp"], d["Pr"], d["divV"], p.rho, p.dt, p.dx, p.dy, p.dz, *n)
    assert_same(d, f, ["Rp"])
    assert ctx.max_abs(d["Rp"]) == O.max_abs(f["Rp"])


def test_max_abs_nan_inf(O, ctx):
    a = np.asfortranarray(np.random.default_rng(0).uniform(-5, 5, size=(33, 7, 5)))
    d = ctx.from_host(a)
    assert ctx.max_abs(d) == np.abs(a).max()
    a[3, 2, 1] = -np.inf
    assert ctx.max_abs(d.set(a)) == np.inf
    a[30, 6, 4] = np.nan   # Julia's maximum propagates NaN; fmax would not
    assert np.isnan(ctx.max_abs(d.set(a)))
    assert np.isnan(O.max_abs(a))


@pytest.mark.parametrize("grid", GRIDS)
def test_correct_V(O, ctx, grid):
    p = mk_params(O, "G", grid)
    f = rand_fields(O, p, 5)
    d = upload(ctx, f)
    O.correct_V(p, f)
    ctx.call("ns3d_correct_V", d["Vx"], d["Vy"], d["Vz"], d["Pr"], p.dt, p.rho, p.dx, p.dy, p.dz, p.nx, p.ny, p.nz)
    assert_same(d, f, ["Vx", "Vy", "Vz"])


@pytest.mark.parametrize("shape", [(3, 3, 3), (38, 24, 20), (64, 38, 39)])
def test_bc_kernels(O, ctx, shape):
    rng = np.random.default_rng(6)
    for name, scal in [("x", ()), ("y", ()), ("z", ()), ("zV", ()), ("x_Vx", (1.25,)), ("x_Pr", (0.0,)),
                       ("xhydstatic", (0.004, shape[2], 9.81, 1000.0))]:
        a = np.asfortranarray(rng.uniform(-1, 1, size=shape))
        d = ctx.from_host(a)
        O.bc(name, a, *scal)
        if name == "xhydstatic":
            dz, nz, g, rho = scal
            ctx.call("ns3d_bc_xhydstatic", d, dz, nz, g, rho, *shape)
        else:
            ctx.call("ns3d_bc_" + name, d, *scal, *shape)
        assert (d.to_host() == a).all(), name


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("variant", ["M", "G"])
def test_set_bc_wrappers(O, ctx, grid, variant):
    p = mk_params(O, variant, grid)
    f = rand_fields(O, p, 7)
    d = upload(ctx, f)
    n = (p.nx, p.ny, p.nz)
    O.set_bc_Vel(p, f)
    O.set_bc_Pr(p, f)
    if variant == "M":
        ctx.call("ns3d_set_bc_Vel_M", d["Vx"], d["Vy"], d["Vz"], int(p.inlet_guard), p.vin, *n)
        ctx.call("ns3d_set_bc_Pr_M", d["Pr"], int(p.outlet_guard), 0.0, *n)
    else:
        ctx.call("ns3d_set_bc_Vel_G", d["Vx"], d["Vy"], d["Vz"], *n)
        ctx.call("ns3d_set_bc_Pr_G", d["Pr"], p.dz, p.nz, p.g, p.rho, *n)
    assert_same(d, f, ["Vx", "Vy", "Vz", "Pr"])


@pytest.mark.parametrize("grid", GRIDS)
@pytest.mark.parametrize("vscale", [0.0, 0.3, 3.0])
def test_advect(O, ctx, grid, vscale):
    """vscale*dx/dt ~ cells travelled per step: 3.0 clamps at every face of every array."""
    p = mk_params(O, "M", grid)
    f = rand_fields(O, p, 8)
    for v in ("Vx", "Vy", "Vz"):
        f[v] *= vscale
    # exact-integer displacements (weight quirk, M:196): dt*v/dx == +-1 and 2 exactly
    f["Vx"][1:3, :, :] = p.dx / p.dt
    f["Vy"][:, 1:2, :] = -2 * p.dy / p.dt
    for v in ("Vx", "Vy", "Vz", "C"):
        f[v + "_o"][...] = f[v]
    d = upload(ctx, f)
    O.advect(p, f)
    ctx.call("ns3d_advect", d["Vx"], d["Vx_o"], d["Vy"], d["Vy_o"], d["Vz"], d["Vz_o"], d["C"], d["C_o"], p.dt, p.dx,
             p.dy, p.dz, p.nx, p.ny, p.nz)
    assert_same(d, f, ["Vx", "Vy", "Vz", "C"])
    if vscale == 0.0 and grid == (3, 3, 3):
        pass


def test_advect_zero_velocity_is_identity(O, ctx):
    p = mk_params(O, "M", (20, 12, 12))
    f = O.alloc_fields(p)
    f["C_o"][...] = np.random.default_rng(9).uniform(0, 1, size=f["C"].shape)
    d = upload(ctx, f)
    ctx.call("ns3d_advect", d["Vx"], d["Vx_o"], d["Vy"], d["Vy_o"], d["Vz"], d["Vz_o"], d["C"], d["C_o"], p.dt, p.dx,
             p.dy, p.dz, p.nx, p.ny, p.nz)
    assert (d["C"].to_host() == f["C_o"]).all()


@pytest.mark.parametrize("variant", ["M", "G"])
@pytest.mark.parametrize("nx", [31, 63, 255])
def test_set_cylinder(O, ctx, variant, nx):
    """The mask must be bit-exact at the rim (a flipped cell changes the solution)."""
    p = O.params_M(nx, nz=6) if variant == "M" else O.params_G(nx, nz=6)
    f = rand_fields(O, p, 10)
    d = upload(ctx, f)
    O.set_cylinder(p, f)
    if variant == "M":
        ctx.call("ns3d_set_cylinder_M", d["C"], d["Vx"], d["Vy"], d["Vz"], p.a2, p.b2, p.ox, p.oy, p.sinb, p.cosb,
                 p.xco_g, p.yco_g, p.dx, p.dy, p.nx, p.ny, p.nz)
    else:
        ctx.call("ns3d_set_cylinder_G", d["C"], d["Vx"], d["Vy"], d["Vz"], p.a2, p.b2, p.ox, p.oy, p.sinb, p.cosb,
                 p.lx, p.ly, p.dx, p.dy, p.nx, p.ny, p.nz)
    assert_same(d, f, ["C", "Vx", "Vy", "Vz"])
    assert (f["C"] == 1.0).sum() > 0


# ---- level 1 against the reference's own TEXT (tests/golden/jl_reference_fixtures.npz) ------------------------------
def _library_launch(ctx, case, p, d):
    """One launch of tests/jl_cases.KERNEL_CASES through the level-1 entry point that replaces the reference call."""
    k, v = case[0], case[1]
    n = (p.nx, p.ny, p.nz)
    if k == "update_τ!":
        ctx.call("ns3d_update_tau", d["txx"], d["tyy"], d["tzz"], d["txy"], d["txz"], d["tyz"], d["Vx"], d["Vy"], d["Vz"],
                 p.mu, p.dx, p.dy, p.dz, *n)
    elif k == "predict_V!":
        ctx.call("ns3d_predict_V", d["Vx"], d["Vy"], d["Vz"], d["txx"], d["tyy"], d["tzz"], d["txy"], d["txz"], d["tyz"],
                 p.rho, p.g, p.dt, p.dx, p.dy, p.dz, *n)
    elif k == "update_∇V!":
        ctx.call("ns3d_update_divV", d["divV"], d["Vx"], d["Vy"], d["Vz"], p.dx, p.dy, p.dz, *n)
    elif k == "update_dPrdτ!":
        ctx.call("ns3d_update_dPrdtau", d["Pr"], d["dPrdtau"], d["divV"], p.rho, p.dt, p.dtau, p.damp, p.dx, p.dy, p.dz, *n)
    elif k == "update_Pr!":
        ctx.call("ns3d_update_Pr", d["Pr"], d["dPrdtau"], p.dtau, *n)
    elif k == "compute_res!":
        ctx.call("ns3d_compute_res", d["Rp"], d["Pr"], d["divV"], p.rho, p.dt, p.dx, p.dy, p.dz, *n)
    elif k == "correct_V!":
        ctx.call("ns3d_correct_V", d["Vx"], d["Vy"], d["Vz"], d["Pr"], p.dt, p.rho, p.dx, p.dy, p.dz, *n)
    elif k == "set_bc_Vel!":
        if v == "M":
            ctx.call("ns3d_set_bc_Vel_M", d["Vx"], d["Vy"], d["Vz"], int(p.inlet_guard), p.vin, *n)
        else:
            ctx.call("ns3d_set_bc_Vel_G", d["Vx"], d["Vy"], d["Vz"], *n)
    elif k == "set_bc_Pr!":
        if v == "M":
            ctx.call("ns3d_set_bc_Pr_M", d["Pr"], int(p.outlet_guard), 0.0, *n)
        else:
            ctx.call("ns3d_set_bc_Pr_G", d["Pr"], p.dz, p.nz, p.g, p.rho, *n)
    elif k == "advect!":
        ctx.call("ns3d_advect", d["Vx"], d["Vx_o"], d["Vy"], d["Vy_o"], d["Vz"], d["Vz_o"], d["C"], d["C_o"], p.dt, p.dx, p.dy,
                 p.dz, *n)
    elif k == "set_cylinder!":
        if v == "M":
            ctx.call("ns3d_set_cylinder_M", d["C"], d["Vx"], d["Vy"], d["Vz"], p.a2, p.b2, p.ox, p.oy, p.sinb, p.cosb,
                     p.xco_g, p.yco_g, p.dx, p.dy, *n)
        else:
            ctx.call("ns3d_set_cylinder_G", d["C"], d["Vx"], d["Vy"], d["Vz"], p.a2, p.b2, p.ox, p.oy, p.sinb, p.cosb,
                     p.lx, p.ly, p.dx, p.dy, *n)
    else:
        raise KeyError(k)


def test_level1_launches_equal_the_reference_text(O, ctx):
    """All 78 single launches of tests/jl_cases.py (every kernel of both scripts on seeded random fields: 3^3, ragged
    grids, back-tracking clamped at every face, exact-integer displacements, rotated ellipse, guards on and off)
    through the C ABI, against what the reference scripts' own text computes for the same inputs
    (oracle/jl_interp.py; the oracle only supplies shapes and scalars here).  Bit for bit, outputs and bystanders."""
    import json
    import os

    from tests import jl_cases as J
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jl_reference_fixtures.npz"))
    meta = json.loads(str(z["meta"]))["kernel"]
    for case in J.KERNEL_CASES:
        p, f = J.inputs_of(O, case)
        d = upload(ctx, f)
        _library_launch(ctx, case, p, d)
        cid = J.case_id(case)
        for name in f:
            if name == "absRp":
                continue
            got = d[name].to_host()
            if name in J.OUTPUTS[case[0]]:
                assert J.digest(got) == meta[cid][name], f"{cid}: {name} differs from the reference text's result"
            else:
                assert (got == f[name]).all(), f"{cid}: {name} was modified"
        for a in d.values():
            ctx.free(a)
