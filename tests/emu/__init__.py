"""Host emulation of the CUDA kernels of the fused PT loop (test infrastructure).

``tests/emu/pt_emu.cpp`` compiles ``navierstokes3d_b200/csrc/ns3d_pt_kernels.cuh`` -- the very
source nvcc compiles into libns3d.so -- with g++ behind ``cuda_host_shim.h`` (CUDA threads = host
threads, ``__syncthreads`` = barrier).  It lets the CPU test suite execute the kernels' index
arithmetic, boundary folding, tiling and ping-pong logic bit for bit against the oracle without
a GPU.  It is NOT a product path (nothing in navierstokes3d_b200/ can reach it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libpt_emu.so")
DEPS = [os.path.join(HERE, "pt_emu.cpp"), os.path.join(HERE, "cuda_host_shim.h"),
        os.path.join(ROOT, "navierstokes3d_b200", "csrc", "ns3d_pt_kernels.cuh"),
        os.path.join(ROOT, "navierstokes3d_b200", "csrc", "ns3d_shared.cuh"),
        os.path.join(ROOT, "include", "ns3d.h")]
_lib = None

KERNELS = {"pt_iter": 0, "pt_tb2": 1, "pt_tb2s": 2, "pt_tb2d": 3, "pt_tb2s_pb": 4}


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    # -ffp-contract=off: like nvcc --fmad=false, FMA only where fma() is written
    cmd = ["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-pthread",
           os.path.join(HERE, "pt_emu.cpp"), "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("g++ failed building the kernel emulation:\n" + res.stderr)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_pt_iterate.restype = C.c_int
        _lib.emu_pt_iterate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
        _lib.emu_pt_tb2_split.restype = C.c_int
        _lib.emu_pt_tb2_split.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.emu_pt_slab_iterate.restype = C.c_int
        _lib.emu_pt_slab_iterate.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                             C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return _lib


class EmuSlab(C.Structure):
    """``EmuSlab`` of pt_emu.cpp: this rank's buffers and the neighbours' as 16 raw addresses."""
    _fields_ = [("pr", C.c_void_p * 2), ("dp", C.c_void_p * 2), ("divV", C.c_void_p),
                ("lo", C.c_void_p * 4), ("hi", C.c_void_p * 4),
                ("mbox", C.c_void_p), ("lo_mbox", C.c_void_p), ("hi_mbox", C.c_void_p)]


class SlabMemory:
    """The shared-memory image of one rank of an emulated z-slab run: Pr, its shadow, dPrdτ, its
    shadow, ∇V (each padded by three planes like ns3d_zeros pads) and a 16-word mailbox, in ONE
    ``multiprocessing.shared_memory`` block that the neighbouring rank processes map as well --
    the stand-in for the CUDA IPC mappings of the peer-memory halo path."""

    def __init__(self, nx, ny, nz, name=None, create=False):
        from multiprocessing import shared_memory
        self.shape, self.dshape = (nx, ny, nz), (nx - 2, ny - 2, nz - 2)
        n, nd = nx * ny * nz, (nx - 2) * (ny - 2) * (nz - 2)
        pn, pd = n + 3 * nx * ny + 32, nd + 3 * (nx - 2) * (ny - 2) + 32
        self.counts = [pn, pn, pd, pd, pn, 16]           # Pr, Pr shadow, dP, dP shadow, divV, mailbox
        self.offsets = np.concatenate([[0], np.cumsum(self.counts)[:-1]]) * 8
        size = int(sum(self.counts) * 8)
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=size)
        self._anchor = C.c_char.from_buffer(self.shm.buf)
        self.base = C.addressof(self._anchor)
        if create:
            whole = np.frombuffer(self.shm.buf, dtype=np.float64, count=sum(self.counts))
            whole[:] = np.nan                            # padding and shadows: a value USED from there poisons the result
            whole[-16:] = 0.0                            # mailbox words start at zero
            del whole

    def addr(self, which: int) -> int:
        return self.base + int(self.offsets[which])

    def view(self, which: int) -> np.ndarray:
        """A copy-free view while alive; callers drop it before close()."""
        shape = self.dshape if which in (2, 3) else self.shape
        return np.frombuffer(self.shm.buf, dtype=np.float64, count=int(np.prod(shape)),
                             offset=int(self.offsets[which])).reshape(shape, order="F")

    def close(self, unlink=False):
        self._anchor = None
        try:
            self.shm.close()
        except BufferError:      # a view is still alive somewhere: the mapping goes away with the process
            pass
        if unlink:
            self.shm.unlink()


def slab_rank_main(rank, nranks, names, grid, kernel, mode, pt_bytes, n_iter, kernel_mid, ty_mid, queue):
    """Body of one rank process: map own and neighbours' memory, run the launches, report."""
    try:
        mem = {r: SlabMemory(*grid, name=names[r]) for r in (rank - 1, rank, rank + 1) if 0 <= r < nranks}
        me = mem[rank]
        b = EmuSlab()
        b.pr[0], b.pr[1], b.dp[0], b.dp[1], b.divV = me.addr(0), me.addr(1), me.addr(2), me.addr(3), me.addr(4)
        b.mbox = me.addr(5)
        for side, r in (("lo", rank - 1), ("hi", rank + 1)):
            if r in mem:
                arr = getattr(b, side)
                for q in range(4):
                    arr[q] = mem[r].addr(q)
                setattr(b, side + "_mbox", mem[r].addr(5))
        pt = (C.c_char * len(pt_bytes)).from_buffer_copy(pt_bytes)
        wp, wd = C.c_int(-1), C.c_int(-1)
        rc = lib().emu_pt_slab_iterate(KERNELS[kernel], mode, C.addressof(pt), rank, nranks, C.addressof(b), n_iter,
                                       KERNELS[kernel_mid], ty_mid, C.byref(wp), C.byref(wd))
        queue.put((rank, rc, wp.value, wd.value))
        del b
        for m in mem.values():
            m.close()
    except BaseException as exc:  # noqa: BLE001
        queue.put((rank, -99, repr(exc), 0))


def pt_tb2_split(kernel_mid: str, mode: int, pt_params, Pr, dP, divV, n_pairs: int, klo: int, khi: int,
                 ty_mid: int = 16, zchunk_mid: int = 16) -> None:
    """n_pairs double iterations, each as three launches over the plane ranges [1,klo), [klo,khi),
    [khi,nz-1) -- the outer two with pt_tb2_kernel, the middle one with `kernel_mid` (the way slabs
    split every launch into interface chunks and the rest)."""
    rc = lib().emu_pt_tb2_split(KERNELS[kernel_mid], mode, ty_mid, C.addressof(pt_params), Pr.ctypes.data,
                                dP.ctypes.data, divV.ctypes.data, n_pairs, klo, khi, zchunk_mid)
    if rc != 0:
        raise RuntimeError(f"emu_pt_tb2_split failed ({rc})")


def pt_iterate(kernel: str, mode: int, pt_params, Pr: np.ndarray, dP: np.ndarray, divV: np.ndarray, n: int,
               ty: int = 16, serpentine: bool = True, zlo_halo: bool = False, zhi_halo: bool = False) -> int:
    """n fused PT iterations on host arrays (Fortran order, updated in place) through the emulated
    kernels, driven like one rank's run_direct() in ns3d_pt.cu.  Returns the number of launches."""
    for a in (Pr, dP, divV):
        assert a.dtype == np.float64 and a.flags.f_contiguous
    nl = C.c_longlong(0)
    rc = lib().emu_pt_iterate(KERNELS[kernel], mode, ty, C.addressof(pt_params), int(zlo_halo), int(zhi_halo),
                              int(serpentine), Pr.ctypes.data, dP.ctypes.data, divV.ctypes.data, n, C.byref(nl))
    if rc != 0:
        raise RuntimeError(f"emu_pt_iterate failed ({rc})")
    return nl.value


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
