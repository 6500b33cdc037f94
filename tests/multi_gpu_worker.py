"""One rank per GPU (torchrun, NCCL): z-slab run of libns3d.so vs the oracle's IGG emulation.

PARITY mode: every rank's local arrays must equal the emulation's arrays for that rank bit for bit
(halo planes included) and the PT iteration counts must be identical.  Launched by
tests/test_gpu_multi.py when the box has >= 2 GPUs.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import navierstokes3d_b200 as ns  # noqa: E402
from navierstokes3d_b200.driver import attach_communicator, gather_interior  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    nx, ny, nz, nt = (int(v) for v in sys.argv[1:5])
    lz = float(sys.argv[5]) if len(sys.argv) > 5 and sys.argv[5] != "None" else None
    level1 = len(sys.argv) > 6 and sys.argv[6] == "level1"
    tb2 = len(sys.argv) > 6 and sys.argv[6] == "tb2"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=rank, nranks=world, lz=lz)
    ctx = ns.Context(local, ns.PARITY)
    attach_communicator(ctx, rank, world)
    pt_random = len(sys.argv) > 6 and sys.argv[6] == "pt_random"
    ctx.set_option("ptv_k", 2 if (tb2 or pt_random) else 1)   # "fused": one iteration per launch with peer stores
    sim = ns.Simulation(s, ctx)
    truth = O.VirtualRanks(nx, ny, nz, (1, 1, world), lz=lz)
    names = ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C", "divV")
    if pt_random:
        # The script's flow is invariant along z for many steps, which would hide a wrong plane offset
        # across a slab interface: the fused loop alone on random fields (same seed on every rank ->
        # the same global state; halos made consistent), nt iterations (tests/emu runs the same on the CPU).
        rng = np.random.default_rng(4321)
        for f in truth.f:
            f["Pr"][...] = rng.uniform(-1, 1, size=f["Pr"].shape)
            f["dPrdtau"][...] = rng.uniform(-1, 1, size=f["dPrdtau"].shape)
            f["divV"][...] = rng.uniform(-1e-3, 1e-3, size=f["divV"].shape)
        truth.update_halo("Pr")
        truth.update_halo("divV")
        names = ("Pr", "dPrdtau")
        for name in names + ("divV",):
            sim.f[name].set(truth.f[rank][name])
        ctx.pt_iterate(sim.f["Pr"], sim.f["dPrdtau"], sim.f["divV"], s.pt_params(), nt)
        for _ in range(nt):
            truth.each(O.update_dPrdtau)   # M:459
            truth.each(O.update_Pr)        # M:461
            truth.update_halo("Pr")        # M:462
            truth.each(O.set_bc_Pr)        # M:463
            truth.update_halo("Pr")        # M:182
        sim.iters, truth.iters = [nt], [nt]
    else:
        for _ in range(nt):
            sim.step_level1() if level1 else sim.step()
        for _ in range(nt):
            truth.step()
    problems = []
    for name in names:
        got = sim.host(name)
        bad = (got != truth.f[rank][name])
        if bad.any() or not np.isfinite(got).all():
            problems.append(f"rank {rank}: {name}: {bad.sum()} values differ (planes {sorted(set(np.argwhere(bad)[:, 2].tolist()))})")
    assert sim.iters == truth.iters and not problems, (sim.iters, truth.iters, problems)
    # gather!(A_inn, A_v): interior of the global field on rank 0
    for name in (() if pt_random else ("Pr", "Vz", "C")):
        g = gather_interior(sim, name)
        if rank == 0:
            want = truth.assemble(name)[1:-1, 1:-1, 1:-1]
            assert g.shape == want.shape and np.array_equal(g, want), name
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} level1={level1} iters={sim.iters} launches={ctx.launch_count}")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
