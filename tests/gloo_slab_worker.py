"""World-size-N CPU worker (gloo): the z-slab host logic of the product, driven with the oracle's
kernels as the compute, must reproduce the oracle's ImplicitGlobalGrid emulation bit for bit.

What is under test (product code): params.setup_multi_gpu(rank, nranks) -- per-rank scalars,
global sizes, guards --, params.halo_planes -- which planes update_halo! moves --, and the
exchange / max-reduction pattern of the time loop.  Launched by tests/test_slab_gloo.py with
``python -m torch.distributed.run --nproc-per-node N``.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import navierstokes3d_b200 as ns  # noqa: E402
from navierstokes3d_b200.params import halo_planes  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    nx, ny, nz, nt = (int(v) for v in sys.argv[1:5])
    lz = float(sys.argv[5]) if len(sys.argv) > 5 else None   # explicit lz keeps dz == dx for any rank count
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    s = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=rank, nranks=world, lz=lz)
    # oracle-side parameter object for the kernels, filled from the PRODUCT's set-up
    p = O.params_M(nx, ny, nz, dims=(1, 1, world), coords=(0, 0, rank), lz=lz)
    for k in ("dx", "dy", "dz", "dt", "dtau", "damp", "xco_g", "yco_g", "niter", "nchk", "inlet_guard", "outlet_guard"):
        assert getattr(p, k) == getattr(s, k), k
        setattr(p, k, getattr(s, k))
    f = O.initial_fields(p)

    def update_halo(*names):
        for name in names:
            a = f[name]
            send_lo, recv_lo, send_hi, recv_hi = halo_planes(a.shape[2], s.nz)
            ops, bufs = [], []
            if rank > 0:
                out = torch.from_numpy(np.ascontiguousarray(a[:, :, send_lo]))
                inp = torch.empty_like(out)
                ops += [dist.P2POp(dist.isend, out, rank - 1), dist.P2POp(dist.irecv, inp, rank - 1)]
                bufs.append((recv_lo, inp))
            if rank < world - 1:
                out = torch.from_numpy(np.ascontiguousarray(a[:, :, send_hi]))
                inp = torch.empty_like(out)
                ops += [dist.P2POp(dist.isend, out, rank + 1), dist.P2POp(dist.irecv, inp, rank + 1)]
                bufs.append((recv_hi, inp))
            if ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
            for plane, t in bufs:
                a[:, :, plane] = t.numpy()

    def max_g(x):   # NaN-propagating MPI.MAX of Julia's maximum
        t = torch.tensor([np.inf if np.isnan(x) else x, 1.0 if np.isnan(x) else 0.0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float("nan") if t[1] > 0 else float(t[0])

    update_halo("Pr")
    update_halo("C", "Vx", "Vy", "Vz")
    iters_all = []
    for _ in range(nt):
        O.update_tau(p, f); O.predict_V(p, f); O.set_cylinder(p, f)
        update_halo("C", "Vx", "Vy", "Vz")
        O.update_divV(p, f)
        update_halo("divV")
        iters = 0
        for it in range(1, s.niter + 1):
            O.update_dPrdtau(p, f); O.update_Pr(p, f); O.set_bc_Pr(p, f)
            update_halo("Pr")          # M:462 + M:182 collapse into one exchange after the BCs
            iters = it
            if it % s.nchk == 0:
                O.compute_res(p, f)
                err = max_g(O.max_abs(f["Rp"])) * (s.ly * s.ly) / s.psc
                if err < s.eps_it or not np.isfinite(err):
                    break
        O.correct_V(p, f); O.set_cylinder(p, f); O.set_bc_Vel(p, f)
        update_halo("Vx", "Vy", "Vz")
        for a in ("Vx", "Vy", "Vz", "C"):
            f[a + "_o"][...] = f[a]
        O.advect(p, f)
        update_halo("Vx", "Vy", "Vz")
        iters_all.append(iters)

    truth = O.VirtualRanks(nx, ny, nz, (1, 1, world), lz=lz)
    for _ in range(nt):
        truth.step()
    assert iters_all == truth.iters, (iters_all, truth.iters)
    for name in ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C", "divV"):
        assert np.isfinite(f[name]).all(), f"{name}: the test grid must be a stable one"
        assert np.array_equal(f[name], truth.f[rank][name]), f"rank {rank}: {name} differs from the IGG emulation"
    if rank == 0:
        print(f"GLOO_SLAB_OK world={world} iters={iters_all}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
