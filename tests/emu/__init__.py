"""The CUDA library, executed on the CPU (test infrastructure).

``build_lib.py`` compiles libns3d.so's own translation units -- kernels and host code, the very source nvcc
compiles -- with g++ behind ``cuda_host_shim.h`` (CUDA threads = host threads, ``__syncthreads`` = barrier) and a
fake CUDA runtime / NCCL.  It lets the CPU test suite execute the kernels' index arithmetic, boundary folding,
tiling and ping-pong logic bit for bit against the oracle without a GPU.  It is NOT a product path (nothing in
navierstokes3d_b200/ can reach it).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


# ---- the whole library on the CPU ------------------------------------------------------------------
def emulated_library() -> C.CDLL:
    """tests/emu/_build/libns3d_emu.so (see build_lib.py): libns3d.so's translation units compiled by
    g++ against a fake CUDA runtime, typed like the real one (native.SIGNATURES)."""
    from navierstokes3d_b200 import native
    from . import build_lib
    lib = C.CDLL(build_lib.build())
    for name, (res, args) in native.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


class use_emulated_library:
    """``with use_emulated_library():`` -- inside the block ``navierstokes3d_b200.native`` drives the
    CPU-emulated library instead of libns3d.so, so Context / Simulation / the drivers run without a
    GPU.  Test-only: the product has no switch for this; the binding's module global is patched
    from the outside and restored on exit."""

    def __enter__(self):
        from navierstokes3d_b200 import native
        self._native, self._saved = native, native._lib
        native._lib = emulated_library()
        return native._lib

    def __exit__(self, *exc):
        self._native._lib = self._saved
        return False


def multi_rank_main(rank, world, grid, nt, lz, how, options, uid_pipes, queue):
    """Body of one rank PROCESS of an emulated multi-rank run (spawned with NS3D_EMU_SHARED_ARENA=1 and
    LD_LIBRARY_PATH = tests/emu/_build): the library's own multi-GPU host path -- NCCL bootstrap, CUDA IPC
    peer mappings, stream/event protocol, graph replay, halo exchanges, residual all-reduce -- over the fake
    runtime, checked against the oracle's ImplicitGlobalGrid emulation for this rank."""
    try:
        import navierstokes3d_b200 as ns
        from oracle import oracle as O
        nx, ny, nz = grid
        with use_emulated_library():
            ctx = ns.Context(0, ns.PARITY)
            if rank == 0:
                uid = ctx.unique_id()
                for p in uid_pipes:
                    p.send(uid)
            else:
                uid = uid_pipes.recv()
            ctx.comm_init(rank, world, uid)
            for name, val in options.items():
                ctx.set_option(name, val)
            s = ns.setup_multi_gpu(nx, ny=ny, nz=nz, rank=rank, nranks=world, lz=lz)
            sim = ns.Simulation(s, ctx)
            truth = O.VirtualRanks(nx, ny, nz, (1, 1, world), lz=lz)
            names = ("Pr", "dPrdtau", "Vx", "Vy", "Vz", "C", "divV")
            if how == "pt_random":
                # The script's flow is invariant along z for many steps, which would hide a wrong plane
                # offset across a slab interface: the fused loop alone on random fields (same seed in
                # every rank process -> the same global state; halos made consistent), nt iterations.
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
                    {"step": sim.step, "level1": sim.step_level1, "groups": sim.step_groups}[how]()
                for _ in range(nt):
                    truth.step()
            problems = []
            for name in names:
                got = sim.host(name)
                bad = got != truth.f[rank][name]
                if bad.any() or not np.isfinite(got).all():
                    problems.append(f"{name}: {int(bad.sum())} values differ (planes {sorted(set(np.argwhere(bad)[:, 2].tolist()))})")
            if sim.iters != truth.iters:
                problems.append(f"iterations {sim.iters} != {truth.iters}")
            if how != "pt_random":   # gather!(A_inn, A_v): the global interior on rank 0, over the (fake) NCCL
                from navierstokes3d_b200.driver import gather_interior
                for name in ("Pr", "Vz", "C"):
                    for dtype in (np.float64, np.float32):
                        g = gather_interior(sim, name, dtype)
                        if rank == 0:
                            want = truth.assemble(name)[1:-1, 1:-1, 1:-1].astype(dtype)
                            if g is None or g.shape != want.shape or not np.array_equal(g, want):
                                problems.append(f"gather {name} {np.dtype(dtype).name}")
                        elif g is not None:
                            problems.append(f"gather {name}: rank {rank} got an array")
            p2p = bool(ctx.lib.ns3d_comm_size(ctx.h) == world)
            queue.put((rank, problems, sim.iters, int(ctx.launch_count), p2p))
            ctx.close()
    except BaseException as exc:  # noqa: BLE001
        import traceback
        queue.put((rank, ["exception: " + repr(exc) + "\n" + traceback.format_exc()], None, 0, False))


def julia_rank_main(rank, world, nx, nt, literals, fused, case_id, pipes, queue, script="lookalike"):
    """One rank PROCESS of the Julia multi-GPU script: scripts/NavierStokes3D_multi_gpu_b200.jl (the reference script's text on
    the shim's look-alike surface) interpreted by oracle/jl_shim.py, every ccall into the emulated library of this process,
    MPI.jl replaced by a stand-in that broadcasts the NCCL id over pipes.  Checked against the per-rank digests that the
    REFERENCE script's text yields on the same process grid (tests/golden/jl_reference_fixtures.npz, "ranks")."""
    try:
        import ctypes
        import json
        from oracle import jl_shim
        from tests import jl_cases as J
        from . import build_lib
        lib = ctypes.CDLL(build_lib.build())
        z = np.load(os.path.join(ROOT, "tests", "golden", "jl_reference_fixtures.npz"))
        want = json.loads(str(z["meta"]))["ranks"][case_id][rank]
        mpi = jl_shim.PipeMPI(rank, world, pipes)
        if script == "lookalike":
            ret, local, iters, errs, (shim, scr, _) = jl_shim.run_multi_gpu_lookalike_b200(
                lib, os.path.join(ROOT, "julia", "NS3DNative.jl"), os.path.join(ROOT, "scripts", "NavierStokes3D_multi_gpu_b200.jl"),
                nx, nt, use_fused=fused, mpi=mpi, literals=literals)
        else:                                   # the explicit-context script (it keeps no residual history)
            ret, local, iters, (shim, scr, _) = jl_shim.run_multi_b200(
                lib, os.path.join(ROOT, "julia", "NS3DNative.jl"), os.path.join(ROOT, "scripts", "NavierStokes3D_b200.jl"),
                nx, nt, use_fused=fused, mpi=mpi, literals=literals)
            errs = want["errs"]
        problems = []
        if iters != want["iters"]:
            problems.append(f"iterations {iters} != {want['iters']}")
        if errs != want["errs"]:
            problems.append("residual history differs")
        for n in ("Pr", "Vx", "Vy", "Vz", "C", "dPrdtau", "divV"):
            if J.digest(local[n]) != want["digest"][n]:
                problems.append(f"{n} differs from the reference text's rank {rank}")
        shapes = None if ret is None or ret[0] is None else [tuple(a.shape) for a in ret]
        halos = sum(1 for s, _ in shim.ccalls if s == "ns3d_update_halo")
        queue.put((rank, problems, iters, shapes, halos, [np.asarray(a) for a in ret] if rank == 0 else None))
        return
    except BaseException as exc:  # noqa: BLE001
        import traceback
        queue.put((rank, ["exception: " + repr(exc) + "\n" + traceback.format_exc()], None, None, 0, None))
