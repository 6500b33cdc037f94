"""The CUDA source of the fused PT kernels, executed on the CPU and compared with the oracle.

tests/emu compiles navierstokes3d_b200/csrc/ns3d_pt_kernels.cuh (the file nvcc compiles into
libns3d.so) with g++ behind a CUDA-on-host shim: one host thread per CUDA thread, barriers for
__syncthreads.  Bars: PARITY mode BIT-EXACT against the oracle's update_dPrdτ! + update_Pr! +
set_bc_Pr! sequence; the two two-iterations-per-launch kernels bit-identical to each other in
every arithmetic mode.  (The GPU suite repeats the oracle comparison on the device, -m gpu.)
"""
import numpy as np
import pytest

import navierstokes3d_b200 as ns
from tests import emu


def oracle_params(O, variant, grid):
    nx, ny, nz = grid
    return O.params_M(nx, ny=ny, nz=nz) if variant == "M" else O.params_G(nx, ny=ny, nz=nz)


def setup_for(variant, grid):
    nx, ny, nz = grid
    return ns.setup_multi_gpu(nx, ny=ny, nz=nz) if variant == "M" else ns.setup_gpu(nx, ny=ny, nz=nz)


def problem(O, variant, grid, seed):
    p = oracle_params(O, variant, grid)
    rng = np.random.default_rng(seed)
    f = O.alloc_fields(p)
    f["Pr"][...] = rng.uniform(-1, 1, size=f["Pr"].shape)
    f["dPrdtau"][...] = rng.uniform(-1, 1, size=f["dPrdtau"].shape)
    f["divV"][...] = rng.uniform(-1e-3, 1e-3, size=f["divV"].shape)
    return p, f


def oracle_iterations(O, p, f, n):
    for _ in range(n):
        O.update_dPrdtau(p, f)
        O.update_Pr(p, f)
        O.set_bc_Pr(p, f)


CASES = [
    # (variant, grid, zchunk, tile height, iteration counts)
    ("M", (3, 3, 3), 0, 16, (2, 1, 2)),
    ("G", (3, 3, 3), 0, 16, (2, 3)),
    ("M", (5, 4, 3), 1, 16, (2, 3)),
    ("G", (4, 3, 6), 2, 16, (4, 1)),
    ("M", (37, 23, 19), 0, 16, (2, 5)),    # 2x2 tiles, one chunk
    ("G", (37, 23, 19), 7, 16, (6, 1)),    # three chunks, serpentine both ways
    ("M", (40, 20, 12), 3, 8, (4,)),       # tile height 8: 2x3 tiles
    ("G", (33, 31, 9), 4, 32, (4, 1)),     # tile height 32
    ("M", (63, 38, 38), 0, 16, (2,)),      # the reference test's grid (config A)
]


@pytest.mark.parametrize("kernel", ["pt_tb2s", "pt_tb2d", "pt_tb2s_pb", "pt_tb2", "pt_iter"])
@pytest.mark.parametrize("variant,grid,zchunk,ty,counts", CASES)
def test_emulated_kernel_bit_exact_vs_oracle(O, kernel, variant, grid, zchunk, ty, counts):
    if kernel == "pt_iter" and (ty != 16 or grid == (63, 38, 38)):
        pytest.skip("tile height is a parameter of the two-iteration kernels only")
    if kernel == "pt_tb2s_pb" and ty == 32:
        pytest.skip("pairwise row barriers: 16 named barriers cover at most 16 tile rows")
    p, f = problem(O, variant, grid, 31)
    s = setup_for(variant, grid)
    g = {k: f[k].copy(order="F") for k in ("Pr", "dPrdtau", "divV")}
    done = 0
    for n in counts:
        emu.pt_iterate(kernel, ns.PARITY, s.pt_params(zchunk), g["Pr"], g["dPrdtau"], g["divV"], n, ty=ty)
        oracle_iterations(O, p, f, n)
        done += n
        for name in ("Pr", "dPrdtau"):
            bad = np.argwhere(g[name] != f[name])
            assert len(bad) == 0, f"{kernel} {name} differs after {done} iterations: {len(bad)} values, first {bad[:3].tolist()}"


@pytest.mark.parametrize("mode", ["FAST", "FASTEST"])
@pytest.mark.parametrize("variant,grid,zchunk", [("M", (37, 23, 19), 5), ("G", (35, 17, 11), 0)])
def test_slim_kernel_identical_to_first_tb2_kernel_in_fast_modes(O, mode, variant, grid, zchunk):
    """pt_tb2s_kernel is a re-write of pt_tb2_kernel for instruction count: same operations in the
    same order, hence bit-identical in the modes the oracle does not define, too."""
    _, f = problem(O, variant, grid, 32)
    s = setup_for(variant, grid)
    out = {}
    for kernel in ("pt_tb2", "pt_tb2s", "pt_tb2d"):
        g = {k: f[k].copy(order="F") for k in ("Pr", "dPrdtau", "divV")}
        emu.pt_iterate(kernel, getattr(ns, mode), s.pt_params(zchunk), g["Pr"], g["dPrdtau"], g["divV"], 6)
        out[kernel] = g
    for name in ("Pr", "dPrdtau"):
        assert np.array_equal(out["pt_tb2"][name], out["pt_tb2s"][name]), name
        assert np.array_equal(out["pt_tb2"][name], out["pt_tb2d"][name]), name
    assert np.isfinite(out["pt_tb2s"]["Pr"]).all()


def test_outlet_guard_off_emulated(O):
    """Variant M with the float == guard false (quirk 5): plain Neumann outlet (xfix stays off)."""
    p, f = problem(O, "M", (20, 12, 12), 33)
    p.outlet_guard = False
    s = setup_for("M", (20, 12, 12))
    s.outlet_guard = False
    g = {k: f[k].copy(order="F") for k in ("Pr", "dPrdtau", "divV")}
    emu.pt_iterate("pt_tb2s", ns.PARITY, s.pt_params(), g["Pr"], g["dPrdtau"], g["divV"], 4)
    oracle_iterations(O, p, f, 4)
    assert np.array_equal(g["Pr"], f["Pr"]) and np.array_equal(g["dPrdtau"], f["dPrdtau"])


@pytest.mark.parametrize("variant,grid,klo,khi,ty,zc", [("M", (37, 23, 30), 9, 21, 8, 5), ("G", (20, 19, 26), 9, 17, 16, 12),
                                                         ("M", (35, 12, 22), 9, 13, 8, 12)])
def test_split_launch_composition(O, variant, grid, klo, khi, ty, zc):
    """Slabs update the chunks next to their interfaces and the planes in between with different
    launches (and different kernels): plane ranges must compose without a seam."""
    p, f = problem(O, variant, grid, 34)
    s = setup_for(variant, grid)
    g = {k: f[k].copy(order="F") for k in ("Pr", "dPrdtau", "divV")}
    emu.pt_tb2_split("pt_tb2s", ns.PARITY, s.pt_params(), g["Pr"], g["dPrdtau"], g["divV"], 2, klo, khi, ty_mid=ty,
                     zchunk_mid=zc)
    oracle_iterations(O, p, f, 4)
    for name in ("Pr", "dPrdtau"):
        bad = np.argwhere(g[name] != f[name])
        assert len(bad) == 0, f"{name}: {len(bad)} values differ, first {bad[:3].tolist()}"


def test_compile_time_stride_instantiation(O):
    """pt_tb2s_kernel<., 8, 1, true, 255, 153>: row/plane strides as immediates (the x-y extent of
    BASELINE configs[1]); a thin 255x153x5 slab keeps the emulation short."""
    grid = (255, 153, 5)
    p, f = problem(O, "G", grid, 35)
    s = setup_for("G", grid)
    g = {k: f[k].copy(order="F") for k in ("Pr", "dPrdtau", "divV")}
    emu.pt_iterate("pt_tb2s", ns.PARITY, s.pt_params(), g["Pr"], g["dPrdtau"], g["divV"], 2, ty=8)
    oracle_iterations(O, p, f, 2)
    assert np.array_equal(g["Pr"], f["Pr"]) and np.array_equal(g["dPrdtau"], f["dPrdtau"])


def test_random_small_grids(O):
    """Seeded sweep over odd shapes: grids from 3^3 up (narrower than a tile, one-row last tiles,
    nz below the chunk length), every chunking, the three tile heights, both BC variants, outlet guard
    on and off -- the default kernel and the two candidates, bit-exact against the oracle."""
    rng = np.random.default_rng(2024)
    for case in range(36):
        variant = ["M", "G"][case % 2]
        grid = (int(rng.integers(3, 70)), int(rng.integers(3, 40)), int(rng.integers(3, 30)))
        zchunk = int(rng.choice([0, 1, 2, 3, 5, 9]))
        kernel = ["pt_tb2s", "pt_tb2d", "pt_tb2s_pb"][case % 3]
        ty = int(rng.choice([8, 16] if kernel == "pt_tb2s_pb" else [8, 16, 32]))
        n = int(rng.choice([2, 3, 4]))
        p, f = problem(O, variant, grid, 100 + case)
        s = setup_for(variant, grid)
        if variant == "M" and case % 4 == 0:
            p.outlet_guard = False
            s.outlet_guard = False
        g = {k: f[k].copy(order="F") for k in ("Pr", "dPrdtau", "divV")}
        emu.pt_iterate(kernel, ns.PARITY, s.pt_params(zchunk), g["Pr"], g["dPrdtau"], g["divV"], n, ty=ty)
        oracle_iterations(O, p, f, n)
        assert np.array_equal(g["Pr"], f["Pr"]) and np.array_equal(g["dPrdtau"], f["dPrdtau"]), \
            (case, variant, grid, zchunk, kernel, ty, n)
