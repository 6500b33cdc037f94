"""The multi-GPU path of the fused PT loop, emulated on the CPU with one PROCESS per rank.

On the device a slab's face CTAs store their new planes straight into the neighbour's halo plane
over NVLink (CUDA IPC mappings), read one plane of the neighbour's current iterate for the
two-iteration kernel, and hand over with acquire/release mailbox flags.  Here
the same kernel source runs in N processes whose fields and mailboxes live in shared memory mapped
by the neighbours -- the launches of run_direct()/pt_iteration() on slabs, the real protocol, no
GPU.  Truth is the oracle's ImplicitGlobalGrid emulation (N virtual ranks, update_halo! at the
script's call sites).  Bit-exact, halo planes included.
"""
import multiprocessing as mp

import numpy as np
import pytest

import navierstokes3d_b200 as ns
from tests import emu


def truth_iterations(O, vr, n):
    for _ in range(n):
        vr.each(O.update_dPrdtau)       # M:459
        vr.each(O.update_Pr)            # M:461
        vr.update_halo("Pr")            # M:462
        vr.each(O.set_bc_Pr)            # M:463
        vr.update_halo("Pr")            # M:182


@pytest.mark.parametrize("world,grid,n_iter,kernel,kernel_mid,ty_mid", [
    (2, (20, 12, 26), 4, "pt_tb2", "pt_tb2s", 8),     # split launches: interface chunks + slim interior (the default path)
    (2, (20, 12, 26), 5, "pt_tb2", "pt_tb2", 16),     # + odd tail: one peer-store pt_iter_kernel launch
    (3, (14, 10, 9), 6, "pt_tb2", "pt_tb2s", 8),      # thin slabs: unsplit peer launches; the middle rank has two neighbours
    (2, (14, 10, 9), 3, "pt_iter", "pt_iter", 8),     # one-iteration kernel with peer stores only
    (4, (9, 7, 23), 6, "pt_tb2", "pt_tb2d", 16),      # four ranks, split launches, the dual-row candidate in the interior
])
def test_slab_ranks_in_processes_match_igg_emulation(O, world, grid, n_iter, kernel, kernel_mid, ty_mid):
    nx, ny, nz = grid
    lz = (world * (nz - 2) + 2) / nx        # dz == dx for any rank count
    vr = O.VirtualRanks(nx, ny, nz, (1, 1, world), lz=lz)
    rng = np.random.default_rng(51)
    for f in vr.f:
        f["Pr"][...] = rng.uniform(-1, 1, size=f["Pr"].shape)
        f["dPrdtau"][...] = rng.uniform(-1, 1, size=f["dPrdtau"].shape)
        f["divV"][...] = rng.uniform(-1e-3, 1e-3, size=f["divV"].shape)
    vr.update_halo("Pr")
    vr.update_halo("divV")
    emu.lib()                               # build before forking
    mems = [emu.SlabMemory(nx, ny, nz, create=True) for _ in range(world)]
    try:
        for r, m in enumerate(mems):
            m.view(0)[...] = vr.f[r]["Pr"]
            m.view(2)[...] = vr.f[r]["dPrdtau"]
            m.view(4)[...] = vr.f[r]["divV"]
        truth_iterations(O, vr, n_iter)
        names = [m.shm.name for m in mems]
        ctx = mp.get_context("spawn")     # the pytest process is multi-threaded (OpenMP oracle): no fork
        queue = ctx.Queue()
        procs = []
        for r in range(world):
            s = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=r, nranks=world, lz=lz)
            procs.append(ctx.Process(target=emu.slab_rank_main,
                                     args=(r, world, names, grid, kernel, ns.PARITY, bytes(s.pt_params()), n_iter,
                                           kernel_mid, ty_mid, queue)))
        for p in procs:
            p.start()
        results = {}
        for _ in range(world):
            rank, rc, wp, wd = queue.get(timeout=180)
            results[rank] = (rc, wp, wd)
        for p in procs:
            p.join(timeout=30)
            assert p.exitcode == 0
        for r in range(world):
            rc, wp, wd = results[r]
            assert rc == 0, f"rank {r}: {rc} {wp}"
            for name, which in (("Pr", wp), ("dPrdtau", 2 + wd)):
                got, want = mems[r].view(which).copy(), vr.f[r][name]
                bad = np.argwhere(got != want)
                assert len(bad) == 0, (f"rank {r} {name}: {len(bad)} values differ, planes "
                                       f"{sorted(set(bad[:, 2].tolist()))}, first {bad[:3].tolist()}")
    finally:
        for m in mems:
            m.close(unlink=True)
