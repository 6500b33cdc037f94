"""N > 1 host logic on CPU: world_size-2 and -3 gloo runs of the z-slab decomposition (no GPU)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,grid,nt,lz", [(2, (40, 24, 13), 4, None), (3, (40, 24, 10), 3, 26 / 40)])
def test_gloo_slabs_match_igg_emulation(world, grid, nt, lz):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "gloo_slab_worker.py"), *map(str, grid), str(nt)] + ([repr(lz)] if lz else [])
    env = dict(os.environ, OMP_NUM_THREADS="2")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert f"GLOO_SLAB_OK world={world}" in res.stdout


def test_halo_plane_rule(ns):
    from navierstokes3d_b200.params import halo_planes
    assert halo_planes(10, 10) == (1, 0, 8, 9)      # cell-centred: planes 2 / n-1 (1-based) -> n / 1
    assert halo_planes(11, 10) == (2, 0, 8, 10)     # staggered:    planes 3 / n-1 (1-based)
    with pytest.raises(ValueError):
        halo_planes(8, 10)                          # dPrdtau-shaped: overlap 0, never exchanged


def test_decomposed_vs_single_domain(O):
    """z-slab run == single-domain run on the global grid: identical PT iteration counts, fields to
    rounding (backtrack! works in LOCAL indices, so ix-δx rounds differently per rank: ~1e-15)."""
    nx, ny, nzl, N, nt = 40, 24, 13, 2, 5
    vr = O.VirtualRanks(nx, ny, nzl, (1, 1, N))
    pg = O.params_M(nx, ny, N * (nzl - 2) + 2)
    fg = O.initial_fields(pg)
    for _ in range(nt):
        a = vr.step()
        b = O.step(pg, fg)
        assert a[0] == b[0]
    vs = max(np.abs(fg[v]).max() for v in ("Vx", "Vy", "Vz"))
    for k in ("Pr", "Vx", "Vy", "Vz", "C"):
        scale = vs if k[0] == "V" else np.abs(fg[k]).max()
        assert np.abs(vr.assemble(k) - fg[k]).max() / scale < 1e-12, k
