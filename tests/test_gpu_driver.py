"""GPU tests of the drop-in drivers (the reference's run-script surface) and of bench.py's output."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_run_navierstokes3D_is_the_reference_test_call(O, ns):
    """test/test3D.jl:6 -- run_navierstokes3D(do_vis=false, do_save=false, do_print=true, nx=63, nt=1)
    returns the interior arrays C,Pr,Vx,Vy,Vz.  With the script as shipped Pr is exactly 0 after one
    step (the stale literals of test3D.jl:12-27 are the recorded expected mismatch)."""
    C, Pr, Vx, Vy, Vz = ns.run_navierstokes3D(do_vis=False, do_save=False, do_print=False, nx=63, nt=1, mode=ns.PARITY)
    assert Pr.shape == C.shape == (61, 36, 36) and Vx.shape == (62, 36, 36)
    assert Vy.shape == (61, 37, 36) and Vz.shape == (61, 36, 37)
    assert (Pr == 0.0).all()
    p = O.params_M(63)
    f, _, _ = O.run(p, 1)
    for got, name in ((C, "C"), (Pr, "Pr"), (Vx, "Vx"), (Vy, "Vy"), (Vz, "Vz")):
        assert np.array_equal(got, O.interior(f[name])), name
    assert (C == 1.0).sum() > 0


def test_run_navierstokes3D_matches_oracle_after_six_steps(O, ns):
    out = ns.run_navierstokes3D(nx=63, nt=6, mode=ns.PARITY)
    p = O.params_M(63)
    f, iters, _ = O.run(p, 6)
    for got, name in zip(out, ("C", "Pr", "Vx", "Vy", "Vz")):
        assert np.array_equal(got, O.interior(f[name])), name


def test_runme_single_gpu_script(O, ns):
    sim = ns.runme(do_vis=False, do_save=False, nx=40, nt=2, mode=ns.PARITY, do_print=False, return_sim=True)
    p = O.params_G(40)
    f, iters, _ = O.run(p, 2)
    assert sim.iters == iters
    for name in ("Pr", "Vx", "Vy", "Vz", "C"):
        assert np.array_equal(sim.host(name), f[name]), name
    sim.ctx.close()


def test_do_save_writes_float32_bins(ns, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    out, sim = ns.run_navierstokes3D(do_save=True, nx=40, nt=2, nsave=2, return_sim=True)
    s = sim.s
    for name, arr in zip(("C", "Pr", "Vx", "Vy", "Vz"), out):
        # frame 1 = the first saved time step (here step 2 = the returned state), M:27-30,515-523
        raw = np.fromfile(tmp_path / "out_save" / f"out_{name}_v_0001.bin", dtype=np.float32)
        assert raw.size == arr.size
        assert np.array_equal(raw, arr.astype(np.float32).ravel(order="F"))
        # frame 0 = initial conditions (M:404-413): C and V carry the cylinder mask, Vy the (sic) inlet plane
        ic = np.fromfile(tmp_path / "out_save" / f"out_{name}_v_0000.bin", dtype=np.float32)
        assert ic.size == arr.size and np.isfinite(ic).all()
    c0 = np.fromfile(tmp_path / "out_save" / "out_C_v_0000.bin", dtype=np.float32).reshape((s.nx - 2, s.ny - 2, s.nz - 2), order="F")
    assert (c0 == 1.0).sum() > 0 and set(np.unique(c0)) <= {0.0, 1.0}
    assert not (tmp_path / "out_save" / "out_C_v_0002.bin").exists()
    sim.ctx.close()


def test_bench_line_has_the_contract_keys():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "A", "--steps", "2", "--warmup", "3",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-3000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "parity_check"):
        assert k in line, k
    assert line["dtype"] == "f64" and line["unit"] == "GB/s" and line["vs_baseline"] is None
    assert line["gpu_launches"] > 0 and line["value"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0
    assert line["parity_check"]["pt_iters_identical"] and line["parity_check"]["within_tolerance"]


@pytest.mark.parametrize("variant,nx", [("M", 40), ("G", 40), ("G", 63)])
def test_device_side_initialisers_match_the_scripts_arrays(ns, variant, nx):
    """M:369-373 / G:86-88 built on the device (ns3d_fill_profile_z, ns3d_fill_plane_x: nz host values cross PCIe)
    against the same comprehensions evaluated as 3-D host arrays: bit-equal, signs of zeros included."""
    from navierstokes3d_b200.driver import initial_host_fields
    s = ns.setup_multi_gpu(nx) if variant == "M" else ns.setup_gpu(nx)
    dev = ns.Simulation(s, ns.Context(0, ns.PARITY))
    host = ns.Simulation(s, ns.Context(0, ns.PARITY), host_fields=initial_host_fields(s))
    if variant == "M":
        host.set_cylinder()      # host_fields replace only the arrays; the masking call of M:372 follows them
    for name in ("Pr", "Vx", "Vy", "Vz", "C"):
        a, b = dev.host(name), host.host(name)
        assert np.array_equal(a, b) and np.array_equal(np.signbit(a), np.signbit(b)), name
    dev.ctx.close()
    host.ctx.close()


# ---- the library against the reference's own source text (tests/golden/jl_reference_fixtures.npz) ----------
def _text_fixtures():
    z = np.load(os.path.join(ROOT, "tests", "golden", "jl_reference_fixtures.npz"))
    return z, json.loads(str(z["meta"]))


@pytest.mark.parametrize("rid,path", [("M31", "fused"), ("M31", "level1"), ("M31", "groups"), ("G20", "fused"), ("G20", "level1"),
                                      ("M40rot", "fused"), ("G40", "fused"), ("M63", "fused"), ("M20x14x9", "fused")])
def test_library_runs_equal_the_reference_text(ns, rid, path):
    """Whole runs of the two scripts through the C ABI (PARITY arithmetic) against what the scripts' own TEXT
    computes when oracle/jl_interp.py executes it (tests/golden/make_jl_fixtures.py; nothing of the C oracle is
    involved in the expected values): identical PT iteration counts, every residual of every check, and the
    final Pr, Vx, Vy, Vz, C bit for bit -- through ns3d_step, through the level-2 groups and call site by call
    site through level 1."""
    from tests import jl_cases as J
    z, meta = _text_fixtures()
    _, variant, nx, nt, _, lit = next(rc for rc in J.RUN_CASES if rc[0] == rid)
    geo = {k: lit[k] for k in ("ny", "nz", "ly", "lz") if k in lit}            # explicit dims (configs D / E style)
    phys = {k: v for k, v in lit.items() if k not in geo}
    ph = ns.Physics(**phys) if phys else None
    s = ns.setup_multi_gpu(nx, physics=ph, **geo) if variant == "M" else ns.setup_gpu(nx, physics=ph)
    sim = ns.Simulation(s, mode=ns.PARITY, device=0)
    for _ in range(nt):
        {"fused": sim.step, "level1": sim.step_level1, "groups": sim.step_groups}[path]()
    m = meta["run"][rid]
    assert sim.iters == m["iters"]
    assert sim.err_hist == m["errs"]
    for name in J.RUN_FIELDS:
        got = sim.host(name)
        assert J.digest(got) == m["digest"][name], f"{rid}/{path}: {name} differs from the reference text's result"
        if rid in J.FULL_ARRAYS:
            assert np.array_equal(got, z[f"run/{rid}/{name}"])
    sim.ctx.close()
