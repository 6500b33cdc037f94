"""2-rank debug: n PT iterations with tb2 on slabs vs the IGG emulation's level-1 loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import navierstokes3d_b200 as ns
from navierstokes3d_b200.driver import attach_communicator
from oracle import oracle as O
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nx, ny, nz = 40, 24, 13
niter = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tb2 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
vr = O.VirtualRanks(nx, ny, nz, (1, 1, world))
rng = np.random.default_rng(3)
# a consistent random global state: fill per-rank arrays from global random fields
g = {"Pr": rng.uniform(-1, 1, size=(nx, ny, world * (nz - 2) + 2)), "divV": rng.uniform(-1e-3, 1e-3, size=(nx, ny, world * (nz - 2) + 2)),
     "dPrdtau": rng.uniform(-1, 1, size=(nx - 2, ny - 2, world * (nz - 2)))}
for r in range(world):
    lo = r * (nz - 2)
    vr.f[r]["Pr"][...] = g["Pr"][:, :, lo:lo + nz]
    vr.f[r]["divV"][...] = g["divV"][:, :, lo:lo + nz]
    vr.f[r]["dPrdtau"][...] = g["dPrdtau"][:, :, lo:lo + nz - 2]
s = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=rank, nranks=world)
ctx = ns.Context(local, ns.PARITY); attach_communicator(ctx, rank, world)
ctx.set_option("tb2", tb2)
ctx.set_option("graphs", int(sys.argv[3]) if len(sys.argv) > 3 else 1)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
d = {k: ctx.from_host(vr.f[rank][k]) for k in ("Pr", "dPrdtau", "divV")}
ctx.update_halo([d["Pr"]], nz)   # connect peers (uncaptured)
for _ in range(reps):
    ctx.pt_iterate(d["Pr"], d["dPrdtau"], d["divV"], s.pt_params(), niter); ctx.sync()
for _ in range(niter * reps):
    vr.each(O.update_dPrdtau); vr.each(O.update_Pr); vr.each(O.set_bc_Pr); vr.update_halo("Pr")
for k in ("Pr", "dPrdtau"):
    got = d[k].to_host(); bad = np.argwhere(got != vr.f[rank][k])
    print(f"rank {rank} {k}: {len(bad)} differ; planes {sorted(set(bad[:,2].tolist()))}; maxabs {np.abs(got - vr.f[rank][k]).max():.3e}", flush=True)
dist.barrier(); ctx.close(); dist.destroy_process_group()
