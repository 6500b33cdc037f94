"""The library's multi-GPU host path on the CPU: one PROCESS per rank, the emulated library in each.

"Device" memory is a POSIX shared-memory arena per process, so the fake runtime's cudaIpcGetMemHandle /
cudaIpcOpenMemHandle give a rank a real mapping of its neighbours' fields and mailboxes; a fake
libnccl.so.2 (tests/emu/fake_nccl.cpp) carries ncclSend/Recv/AllReduce over shared-memory rings.
Everything else is libns3d.so's own code: ns3d_comm_init (NCCL bootstrap, mailbox exchange, the
min-all-reduce that makes all ranks agree on the peer-memory path), peer_prepare, the split
interface/interior launches with their event protocol, graph capture and replay, the once-per-step
halo exchanges over NCCL, the residual max-all-reduce.  Truth: the oracle's ImplicitGlobalGrid
emulation (tests/test_gpu_multi.py runs the same comparison on real GPUs).  Bit-exact.
"""
import multiprocessing as mp
import os

import pytest

from tests import emu
from tests.emu import build_lib


def run_ranks(world, grid, nt, lz, how="step", options=None):
    build_lib.build()
    saved = {k: os.environ.get(k) for k in ("NS3D_EMU_SHARED_ARENA", "LD_LIBRARY_PATH")}
    os.environ["NS3D_EMU_SHARED_ARENA"] = "1"
    os.environ["LD_LIBRARY_PATH"] = build_lib.OUT + os.pathsep + (saved["LD_LIBRARY_PATH"] or "")
    try:
        ctx = mp.get_context("spawn")     # fresh processes: the dynamic loader reads LD_LIBRARY_PATH at start-up
        queue = ctx.Queue()
        pipes = [ctx.Pipe(duplex=False) for _ in range(world - 1)]   # rank 0 -> the others: the NCCL unique id
        procs = []
        for r in range(world):
            ends = [w for _, w in pipes] if r == 0 else pipes[r - 1][0]
            procs.append(ctx.Process(target=emu.multi_rank_main,
                                     args=(r, world, grid, nt, lz, how, options or {}, ends, queue)))
        for p in procs:
            p.start()
        results = {}
        for _ in range(world):
            rank, problems, iters, launches, _ = queue.get(timeout=600)
            results[rank] = (problems, iters, launches)
        for p in procs:
            p.join(timeout=60)
        return results
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("world,grid,nt,lz,how,options", [
    (2, (20, 12, 9), 2, None, "step", {}),                        # thin slabs: unsplit peer launches
    (2, (16, 10, 26), 2, 50 / 16, "step", {}),                     # split launches: interface chunks + interior
    (3, (16, 10, 8), 2, 20 / 16, "step", {"ptv_k": 1}),            # one iteration per launch, peer stores, three ranks
    (2, (16, 10, 9), 2, None, "step", {"p2p_halo": 0}),            # NCCL send/recv halo path instead of peer memory
    (2, (16, 10, 9), 1, None, "level1", {}),                       # the call-by-call level-1 loop over NCCL
    # the fused loop alone on RANDOM fields (the script's flow is z-invariant and would hide a wrong plane offset)
    (2, (16, 10, 26), 12, 50 / 16, "pt_random", {}),               # split launches, graph replay (>= 8 iterations)
    (3, (14, 10, 9), 7, 23 / 14, "pt_random", {}),                 # unsplit peer launches + odd tail, three ranks
    (2, (14, 10, 9), 5, None, "pt_random", {"ptv_k": 1}),          # one iteration per launch with peer stores
    (2, (14, 10, 9), 5, None, "pt_random", {"p2p_halo": 0}),       # NCCL send/recv halo exchange
    (4, (12, 9, 23), 6, 86 / 12, "pt_random", {"ptv_pxt": 4, "ptv_bty": 5}),   # four ranks, several tiles per plane
    (2, (16, 10, 26), 12, 50 / 16, "pt_random", {"ptv_k": 3}),     # K = 3 asked for: slabs fall back to 2
    (2, (16, 10, 26), 2, 50 / 16, "step", {"graphs": 0}),          # whole time steps without graph replay
    (2, (16, 10, 26), 12, 50 / 16, "pt_random", {"p2p_split": 0}), # thick slabs, ONE launch per pass (interface chunks inside it)
    (3, (16, 10, 26), 2, 76 / 16, "step", {"p2p_split": 0}),
    # tall slabs: the planes between the interface chunks run as z-BANDS of launches on their own streams (band b of a pass
    # waits for bands b-1, b, b+1 of the previous one), the interface chunks as one more band on the high-priority stream
    (2, (12, 9, 50), 12, 98 / 12, "pt_random", {}),
    (2, (12, 9, 50), 13, 98 / 12, "pt_random", {"ptv_bands": 3, "graphs": 0}),
    (2, (12, 9, 50), 2, 98 / 12, "step", {}),
    # the two z-slab configurations at which the IGG emulation itself is pinned to the reference script's TEXT
    # (tests/jl_cases.py RANK_CASES "z2", "z3"; tests/test_jl_reference.py): library == emulation == text
    (2, (24, 15, 15), 2, 28 / 24, "step", {}),
    (3, (24, 15, 15), 2, 41 / 24, "step", {}),
])
def test_rank_processes_match_igg_emulation(world, grid, nt, lz, how, options):
    results = run_ranks(world, grid, nt, lz, how, options)
    for r in range(world):
        problems, iters, launches = results[r]
        assert not problems, f"rank {r}: " + "; ".join(problems)
        assert launches > 0 and iters is not None
    assert len({tuple(results[r][1]) for r in range(world)}) == 1   # every rank saw the same iteration counts


# ---- the Julia multi-GPU script itself on several ranks -------------------------------------------------------------
def run_julia_ranks(world, nx, nt, literals, fused, case_id, script="lookalike"):
    build_lib.build()
    saved = {k: os.environ.get(k) for k in ("NS3D_EMU_SHARED_ARENA", "LD_LIBRARY_PATH", "NS3D_EMU_DEVICES")}
    os.environ["NS3D_EMU_SHARED_ARENA"] = "1"
    os.environ["NS3D_EMU_DEVICES"] = str(world)   # init_global_grid selects device = rank, like ImplicitGlobalGrid on one node
    os.environ["LD_LIBRARY_PATH"] = build_lib.OUT + os.pathsep + (saved["LD_LIBRARY_PATH"] or "")
    try:
        ctx = mp.get_context("spawn")
        queue = ctx.Queue()
        pipes = [ctx.Pipe(duplex=False) for _ in range(world - 1)]
        procs = []
        for r in range(world):
            ends = [w for _, w in pipes] if r == 0 else pipes[r - 1][0]
            procs.append(ctx.Process(target=emu.julia_rank_main, args=(r, world, nx, nt, literals, fused, case_id, ends, queue, script)))
        for p in procs:
            p.start()
        results = {}
        for _ in range(world):
            res = queue.get(timeout=600)
            results[res[0]] = res[1:]
        for p in procs:
            p.join(timeout=60)
        return results
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("case_id,fused", [("z2", True), ("z2", False), ("z3", True)])
def test_julia_multi_gpu_script_on_several_ranks(O, case_id, fused):
    """scripts/NavierStokes3D_multi_gpu_b200.jl -- the reference's multi-GPU script with its text kept -- interpreted in one
    process per rank (oracle/jl_shim.py), through julia/NS3DNative.jl's own text (`init_global_grid(...; MPI=MPI)`,
    `comm_init_mpi!`, `update_halo!`, `max_g`, `gather!`), every ccall into the emulated library: each rank's local arrays,
    iteration counts and residuals equal what the REFERENCE script's text yields on the same ImplicitGlobalGrid ranks
    (fixtures "z2" / "z3"), and rank 0's return value is the global interior."""
    import numpy as np
    from tests import jl_cases as J
    case = next(c for c in J.RANK_CASES if c[0] == case_id)
    _, nx, ny, nz, dims, nt, literals, kw = case
    world = dims[2]
    results = run_julia_ranks(world, nx, nt, literals, fused, case_id)
    for r in range(world):
        problems = results[r][0]
        assert not problems, f"rank {r}: " + "; ".join(problems)
    truth = O.VirtualRanks(nx, ny, nz, dims, **kw)
    for _ in range(nt):
        truth.step()
    ret = results[0][4]
    for got, name in zip(ret, ("C", "Pr", "Vx", "Vy", "Vz")):
        want = truth.assemble(name)[1:-1, 1:-1, 1:-1]
        assert got.shape == want.shape and np.array_equal(got, want), name
    for r in range(1, world):
        assert results[r][2] == results[0][2]            # off the root the `zeros(...)` of M:386-390 come back, same shapes
    if not fused:
        assert results[0][3] == 2 + nt * 5 + 3 * sum(results[0][1])   # the text's ten update_halo! call sites


def test_julia_explicit_context_script_on_two_ranks(O):
    """scripts/NavierStokes3D_b200.jl (the same run written against the explicit-context API: `comm_init_mpi!`,
    `update_halo!(ctx, nz, ...)`, `gather_inner`) on two rank processes: every rank equals the reference text's rank, rank 0
    returns the global interior, the other rank `nothing`."""
    import numpy as np
    from tests import jl_cases as J
    _, nx, ny, nz, dims, nt, literals, kw = next(c for c in J.RANK_CASES if c[0] == "z2")
    lit = {"nz": literals["nz"], "lz": literals["lz_lx"]}        # this script spells `ly, lz = 0.6 * lx, 0.6 * lx`
    results = run_julia_ranks(2, nx, nt, lit, True, "z2", script="explicit")
    for r in range(2):
        assert not results[r][0], f"rank {r}: " + "; ".join(results[r][0])
    truth = O.VirtualRanks(nx, ny, nz, dims, **kw)
    for _ in range(nt):
        truth.step()
    for got, name in zip(results[0][4], ("C", "Pr", "Vx", "Vy", "Vz")):
        want = truth.assemble(name)[1:-1, 1:-1, 1:-1]
        assert got.shape == want.shape and np.array_equal(got, want), name
    assert results[1][2] is None
