"""The GPU parity tests, run on the CPU against the EMULATED library.

tests/emu/build_lib.py compiles libns3d.so's own translation units -- kernels and host code -- with
g++ against a fake CUDA runtime (device memory = host memory, streams execute in order, stream
capture records and a graph launch replays, one host thread per CUDA thread where a kernel
synchronises).  With the ctypes binding pointed at that library the ordinary test functions of
test_gpu_kernels.py / test_gpu_solver.py / test_gpu_zz_output.py run unchanged: same Context, same
entry points, same oracle comparisons, bit for bit -- including the host logic a kernel-only
emulation cannot see (ping-pong bookkeeping across captured chunks, copy-back after odd counts, the
residual loop, ns3d_step's composition).  The big cases stay GPU-only; what runs here is sized for
a CPU.  This is test infrastructure: the product has no switch that selects the emulated library.
"""
import numpy as np
import pytest

import tests.test_gpu_driver as D
import tests.test_gpu_kernels as K
import tests.test_gpu_solver as S
import tests.test_gpu_zz_output as Z
from tests import emu


@pytest.fixture(scope="module", autouse=True)
def _emulated_library():
    with emu.use_emulated_library():
        yield


# ---- level 1: every kernel test of test_gpu_kernels.py, unchanged ------------------------------------
for _name in dir(K):
    if _name.startswith("test_"):
        globals()[_name] = getattr(K, _name)

# ---- output path: the box tests of test_gpu_zz_output.py --------------------------------------------
test_box_matches_numpy_slicing = Z.test_box_matches_numpy_slicing
test_box_rejects_a_box_outside_the_array = Z.test_box_rejects_a_box_outside_the_array


# ---- device-side initialisers (signed zeros of M:370 included) ----------------------------------------
test_device_side_initialisers = D.test_device_side_initialisers_match_the_scripts_arrays


# ---- whole runs against what the reference's own source text computes (jl_reference_fixtures.npz) ----
@pytest.mark.parametrize("rid,path", [("M31", "fused"), ("M31", "level1"), ("M31", "groups"), ("G20", "fused"), ("G20", "level1"),
                                      ("M40rot", "fused"), ("M20x14x9", "fused")])
def test_library_runs_equal_the_reference_text(ns, rid, path):
    D.test_library_runs_equal_the_reference_text(ns, rid, path)


# ---- the scripts' do_save output against their own text --------------------------------------------------
test_do_save_frames_equal_the_multi_gpu_scripts_text = Z.test_do_save_frames_equal_the_multi_gpu_scripts_text
test_do_save_mat_dumps_equal_the_single_gpu_scripts_text = Z.test_do_save_mat_dumps_equal_the_single_gpu_scripts_text


# ---- level 2 on CPU-sized grids ---------------------------------------------------------------------
@pytest.mark.parametrize("variant", ["M", "G"])
@pytest.mark.parametrize("grid", [(3, 3, 3), (5, 4, 3), (20, 12, 9)])
@pytest.mark.parametrize("zchunk", [0, 5])
def test_fused_iteration(O, ns, ctx, variant, grid, zchunk):
    S.test_fused_iteration_bit_exact(O, ns, ctx, variant, grid, zchunk)


@pytest.mark.parametrize("variant,grid,zchunk,cfg", [
    ("M", (3, 3, 3), 0, "auto"), ("G", (4, 3, 6), 2, "auto"), ("M", (37, 23, 19), 7, "auto"), ("G", (5, 4, 3), 0, "k1"),
    ("G", (20, 19, 11), 0, "k2_lb0"), ("M", (20, 19, 11), 1, "k2_lb4_ns3"), ("G", (20, 12, 12), 0, "k3_lb0"),
    ("M", (20, 12, 12), 2, "k2_tiles"), ("G", (20, 19, 11), 0, "k3_tiles"), ("M", (16, 9, 8), 3, "k2_coop_nographs"),
    ("G", (16, 9, 8), 2, "k1_coop_tiles"), ("M", (4, 3, 6), 1, "k3_tiles"), ("G", (3, 3, 3), 0, "k3_lb0"),
    ("M", (20, 12, 12), 2, "k2_flow"), ("G", (20, 19, 11), 3, "k3_flow_tiles"), ("M", (16, 9, 8), 0, "k1_flow"),
    ("G", (37, 23, 19), 7, "k2_flow_lb0"),
    ("M", (20, 12, 19), 2, "k2_bands3"), ("G", (20, 19, 23), 3, "k3_bands5_tiles"), ("M", (16, 9, 38), 2, "k2_bands16_nographs"),
    ("G", (20, 12, 19), 2, "k2_nobands"),
])
def test_ptv_kernel(O, ns, ctx, variant, grid, zchunk, cfg):
    """The fused loop's kernel in every configuration behind the options -- iterations per launch, launch bounds,
    staging slots, one tile and several tiles with rims -- through ns3d_pt_iterate: 2, 1, 5 and 40 iterations (graph
    capture and replay from 8 iterations up).  On the CPU the z-plane tiles are staged by plain loads (the TMA unit
    is the one thing the emulation cannot execute; the GPU suite runs the same cases through it)."""
    S.test_ptv_kernel_bit_exact(O, ns, ctx, variant, grid, zchunk, cfg)


def test_outlet_guard_off(O, ns, ctx):
    S.test_outlet_guard_off(O, ns, ctx)


def test_pt_solve_nonfinite_breaks(O, ns, ctx):
    S.test_pt_solve_nonfinite_breaks(O, ns, ctx)


@pytest.mark.parametrize("variant", ["M", "G"])
def test_pt_solve_matches_oracle(O, ns, ctx, variant):
    """The full loop with residual checks (test_gpu_solver.test_pt_solve_matches_oracle on a smaller grid):
    same iteration count, same err history, bit-exact fields."""
    grid = (16, 10, 10)
    p, f = S.pt_problem(O, variant, grid, 13)
    s = S.setup_for(ns, variant, grid[0], ny=grid[1], nz=grid[2])
    f["Pr"][...] = 0.0
    f["dPrdtau"][...] = 0.0
    if variant == "G":   # start from the hydrostatic state so that the loop converges
        f["Pr"][...] = O.initial_fields(p)["Pr"]
    d = {k: ctx.from_host(f[k]) for k in ("Pr", "dPrdtau", "divV")}
    it_o, hist_o = O.pt_solve(p, f)
    it_g, hist_g = ctx.pt_solve(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params())
    assert it_g == it_o and it_o >= p.nchk and hist_g == hist_o
    assert (d["Pr"].to_host() == f["Pr"]).all() and (d["dPrdtau"].to_host() == f["dPrdtau"]).all()


@pytest.mark.parametrize("variant,nx,nt,how", [("M", 20, 3, "step"), ("G", 16, 1, "step"), ("M", 24, 2, "groups"),
                                               ("G", 16, 1, "groups"), ("G", 16, 1, "level1")])
def test_whole_time_steps(O, ns, variant, nx, nt, how):
    """ns3d_step, its four level-2 groups, and the call-by-call level-1 loop: identical PT iteration counts
    and err history, bit-exact fields."""
    p = O.params_M(nx) if variant == "M" else O.params_G(nx)
    f, iters_o, errs_o = O.run(p, nt)
    sim = ns.Simulation(ns.setup_multi_gpu(nx) if variant == "M" else ns.setup_gpu(nx), ns.Context(0, ns.PARITY))
    for _ in range(nt):
        {"step": sim.step, "groups": sim.step_groups, "level1": sim.step_level1}[how]()
    assert sim.iters == iters_o
    assert np.array_equal(np.concatenate(sim.err_hist), np.concatenate(errs_o), equal_nan=True)
    # the snapshots of M:475 too: the fused step works ON the `_o` arrays (predictor V -> V_o, corrector in place,
    # advection V_o -> V) and must leave in them what `A_o .= A` leaves there
    for name in ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C", "divV", "Vx_o", "Vy_o", "Vz_o", "C_o"):
        got = sim.host(name)
        assert ((got == f[name]) | (np.isnan(got) & np.isnan(f[name]))).all(), name
    sim.ctx.close()


def test_driver_returns_the_reference_test_call(O, ns):
    """test/test3D.jl:6 through the drop-in driver (nx=63, nt=1: the degenerate first step, 37 iterations)."""
    C, Pr, Vx, Vy, Vz = ns.run_navierstokes3D(do_vis=False, do_save=False, do_print=False, nx=63, nt=1, mode=ns.PARITY)
    assert Pr.shape == (61, 36, 36) and (Pr == 0.0).all()
    f, _, _ = O.run(O.params_M(63), 1)
    for got, name in ((C, "C"), (Pr, "Pr"), (Vx, "Vx"), (Vy, "Vy"), (Vz, "Vz")):
        assert np.array_equal(got, O.interior(f[name])), name


@pytest.mark.parametrize("variant,nx,nt", [("M", 20, 3), ("G", 16, 1)])
def test_other_physics_than_the_scripts_literals(O, ns, variant, nx, nt):
    """`Physics` replaces the literals of M:290-335 / G:15-56: a rotated elliptic obstacle somewhere else,
    another Reynolds number, inflow speed, gravity and CFL numbers -- kernel arguments the scripts'
    defaults never exercise (sin(beta) != 0, a != b).  Same setup on the oracle: bit-exact."""
    lit = dict(rho=900.0, vin=0.8, mu=2e-3, a_lx=0.11, b_lx=0.06, ox_lx=-0.15, oy_lx=0.04, beta=np.pi / 6,
               g=(0.3 if variant == "M" else 9.0), cfl_tau=0.5, cfl_visc=0.2, cfl_adv=0.9)
    p = O.params_M(nx, **lit) if variant == "M" else O.params_G(nx, **lit)
    f, iters_o, errs_o = O.run(p, nt)
    s = (ns.setup_multi_gpu if variant == "M" else ns.setup_gpu)(nx, physics=ns.Physics(**lit))
    for k in ("dt", "dtau", "damp", "a2", "b2", "ox", "oy", "sinb", "cosb", "g", "rho", "mu", "vin", "psc"):
        assert getattr(s, k) == getattr(p, k), k
    sim = ns.Simulation(s, ns.Context(0, ns.PARITY))
    for _ in range(nt):
        sim.step()
    assert (f["C"] > 0.5).sum() > 8                  # the obstacle is there (script G masks inside the loop only)
    assert sim.iters == iters_o
    assert np.array_equal(np.concatenate(sim.err_hist), np.concatenate(errs_o), equal_nan=True)
    for name in ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C", "divV"):
        assert np.array_equal(sim.host(name), f[name]), name
    assert all(np.isfinite(f[k]).all() for k in ("Pr", "Vx", "C"))
    sim.ctx.close()
