"""CPU tests of the host side and of the C-ABI boundary (no GPU, no compute calls)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ns3d.h")).read()
    return sorted(set(re.findall(r"NS3D_API[^;(]*?\b(ns3d_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(ns):
    """libns3d.so loads without a GPU and exports exactly what include/ns3d.h declares."""
    lib = ns.native.load()
    declared = header_symbols()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ns3d.h but not exported"
    assert sorted(ns.native.SIGNATURES) == declared, "native.py and ns3d.h disagree on the ABI surface"
    out = subprocess.run(["nm", "-D", "--defined-only", ns.native.lib_path()], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (ns3d_\w+)", out)))
    assert exported == declared, "extra or missing exported symbols"
    assert b"sm_100a" in lib.ns3d_version()


def test_struct_layouts_match_header(ns):
    """The ctypes mirrors must have the C layout (checked against a tiny C program's sizeof/offsetof)."""
    prog = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "ns3d.h"
    int main(void) {
        printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(ns3d_pt_params), offsetof(ns3d_pt_params, eps_it),
               offsetof(ns3d_pt_params, outlet_val), offsetof(ns3d_pt_params, zchunk), sizeof(ns3d_fields),
               sizeof(ns3d_step_params), offsetof(ns3d_step_params, inlet_guard));
        return 0;
    }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    N = ns.native
    want = [ctypes.sizeof(N.PtParams), N.PtParams.eps_it.offset, N.PtParams.outlet_val.offset, N.PtParams.zchunk.offset,
            ctypes.sizeof(N.Fields), ctypes.sizeof(N.StepParams), N.StepParams.inlet_guard.offset]
    assert got == want


def test_no_gpu_means_loud_failure(ns):
    """There is no CPU fallback: on a box without CUDA the context cannot be created."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(ns.NS3DError, match="no CUDA device"):
        ns.Context(0)


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "navierstokes3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|from\s+\.+oracle|libns3d_oracle|oracle/|ns3d_oracle\.", text), \
                    f"{fn} references the oracle"
    code = ("import sys; sys.path.insert(0, %r); import navierstokes3d_b200; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


@pytest.mark.parametrize("nx", [20, 40, 63, 255, 511, 1023])
def test_setup_matches_the_oracle_derivation(O, ns, nx):
    """params.py (product) and oracle.py (checker) derive the scripts' scalars independently."""
    names = ["nx", "ny", "nz", "lx", "ly", "lz", "dx", "dy", "dz", "dt", "dtau", "damp", "rho", "mu", "g", "vin", "psc",
             "a2", "b2", "ox", "oy", "sinb", "cosb", "eps_it", "niter", "nchk"]
    for s, p in ((ns.setup_multi_gpu(nx), O.params_M(nx)), (ns.setup_gpu(nx), O.params_G(nx))):
        for n in names:
            assert getattr(s, n) == getattr(p, n), n
    s, p = ns.setup_multi_gpu(nx), O.params_M(nx)
    assert (s.xco_g, s.yco_g, s.zco_g, s.inlet_guard, s.outlet_guard) == (p.xco_g, p.yco_g, p.zco_g, p.inlet_guard, p.outlet_guard)
    assert s.shapes() == {k: v for k, v in O.shapes(p.nx, p.ny, p.nz).items() if k != "absRp"}


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_slab_setup_matches_igg_emulation(O, ns, nranks):
    """z-slab ranks: global sizes, dz, niter/nchk from global sizes, damp from the LOCAL nx (quirk 4)."""
    nx, ny, nz = 40, 24, 14
    for rank in range(nranks):
        s = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=rank, nranks=nranks)
        p = O.params_M(nx, ny=ny, nz=nz, dims=(1, 1, nranks), coords=(0, 0, rank))
        assert s.grid.nz_g == nranks * (nz - 2) + 2 == O.n_g(nz, nranks)
        for n in ("dx", "dy", "dz", "dt", "dtau", "damp", "niter", "nchk", "xco_g", "yco_g", "zco_g", "inlet_guard",
                  "outlet_guard"):
            assert getattr(s, n) == getattr(p, n), n
        assert s.damp == 2 / nx
        lo, hi = s.grid.z_range_global()
        assert hi - lo == nz and lo == rank * (nz - 2)


def test_initial_conditions_match(O, ns):
    from navierstokes3d_b200.driver import initial_host_fields
    for s, p in ((ns.setup_multi_gpu(40), O.params_M(40)), (ns.setup_gpu(40), O.params_G(40))):
        mine = initial_host_fields(s)
        ref = O.alloc_fields(p)
        yc = O.linrange(-(p.ly - p.dy) / 2, (p.ly - p.dy) / 2, p.ny)
        full = O.initial_fields(p)
        if p.variant == "M":
            # the oracle's initial_fields already applied set_cylinder!; compare the pre-mask arrays
            assert (mine["Vy"][0] == p.vin).all() and (mine["Vy"][1:] == 0).all()
            assert (mine["Pr"] == 0).all()          # +-0.0: g = 0
        else:
            assert np.array_equal(mine["Vx"], full["Vx"]) and np.array_equal(mine["Pr"], full["Pr"])
        assert set(mine) <= set(ref) and yc.shape == (p.ny,)


def test_bench_algorithmic_bytes():
    sys.path.insert(0, ROOT)
    import bench
    n = 255 * 153 * 153
    assert bench.a_eff_bytes(n, 0, 0) == 168 * n          # once-per-step part: 21 passes
    assert bench.a_eff_bytes(n, 1, 0) - bench.a_eff_bytes(n, 0, 0) == 40 * n   # one PT iteration: 5 passes
    assert bench.a_eff_bytes(n, 0, 1) - bench.a_eff_bytes(n, 0, 0) == 16 * n   # one residual check: 2 passes


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: the CPU arm (oracle port on the host cores) with the native
    arm's metric / unit / workload naming, `impl`, `cpu_baseline` and a zero-copy `e2e`."""
    import json
    import subprocess
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "A",
                          "--steps", "1", "--warmup", "0", "--sample-seconds", "0.3"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "T_eff" and line["unit"] == "GB/s"
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["value"] > 0
    assert line["config"]["workload"].startswith("A: cylinder flow 63x38x38")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0
