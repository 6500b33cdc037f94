"""Host drivers with the reference's run-script surface, on top of libns3d.so.

* ``run_navierstokes3D(; do_vis, do_save, do_print, nx, nt)`` -- scripts/NavierStokes3D_multi_gpu.jl
  (M:287-536): returns the interior arrays ``C_v, Pr_v, Vx_v, Vy_v, Vz_v`` (M:528-535).
* ``runme(; do_vis, do_save)`` -- scripts/NavierStokes3D_gpu.jl (G:12-173).

The time loop is the scripts' (M:446-477 / G:119-142): ``Simulation.step()`` runs it through
the fused level-2 entry point ``ns3d_step`` and ``Simulation.step_level1()`` line by line
through the level-1 operators, i.e. exactly what a Julia driver re-pointed at the C ABI does.
There is no CPU path: every field lives in device memory allocated by the library.
"""
from __future__ import annotations

import os

import numpy as np

from . import native
from .params import Setup, setup_gpu, setup_multi_gpu


def _linrange(start: float, stop: float, n: int) -> np.ndarray:
    """Julia ``LinRange(start, stop, n)``: element i is (1-t)*start + t*stop with t = (i-1)/(n-1)."""
    t = np.arange(n, dtype=np.float64) / max(n - 1, 1)
    return (1 - t) * start + t * stop


def initial_host_fields(s: Setup) -> dict:
    """Host arrays that differ from zero at t = 0 (before set_cylinder!).

    Variant M (M:369-370): ``Vy[1,:,:] .= vin`` (sic) and the hydrostatic ``Pr`` (= +-0.0, g = 0).
    Variant G (G:86-87): 1/7-power-law ``Vx`` profile and hydrostatic ``Pr``.
    """
    nx, ny, nz = s.nx, s.ny, s.nz
    yc = _linrange(-(s.ly - s.dy) / 2, (s.ly - s.dy) / 2, ny)
    zc = _linrange(-(s.lz - s.dz) / 2, (s.lz - s.dz) / 2, nz)
    out = {}
    if s.variant == native.VARIANT_M:
        vy = np.zeros((nx, ny + 1, nz), order="F")
        vy[0, :, :] = s.vin
        zg = np.array([s.grid.x_g(iz, s.dz, nz, 2) for iz in range(1, nz + 1)])
        pr = (-(zg - s.dz / 2) * s.rho * s.g)[None, None, :] + (0 * yc)[None, :, None] + (0 * zc)[None, None, :]
        out["Vy"] = vy
        out["Pr"] = np.asfortranarray(np.broadcast_to(pr, (nx, ny, nz)))
    else:
        xc = _linrange(-(s.lx - s.dx) / 2, (s.lx - s.dx) / 2, nx)
        xv = _linrange(-s.lx / 2, s.lx / 2, nx + 1)
        prof = s.vin * (7.0 / 6.0) * np.power((zc + s.lz / 2) / s.lz, 1.0 / 6.0)
        out["Vx"] = np.asfortranarray(prof[None, None, :] + (0 * yc)[None, :, None] + (0 * xv)[:, None, None])
        out["Pr"] = np.asfortranarray((-(zc - s.lz / 2) * s.rho * s.g)[None, None, :] + (0 * yc)[None, :, None]
                                      + (0 * xc)[:, None, None])
    return out


def initial_profiles(s: Setup) -> dict:
    """The z-profiles behind ``initial_host_fields`` -- every initial array of the scripts depends on iz alone:
    ``{"Pr": (nz values, ny zeros, nz zeros)}`` (M:370) or ``{"Vx": ..., "Pr": ...}`` (G:86-87).  The ``+ 0*yc[iy] + 0*xv[ix]`` terms of
    the comprehensions are kept (they turn a -0.0 into +0.0)."""
    nz = s.nz
    zc = _linrange(-(s.lz - s.dz) / 2, (s.lz - s.dz) / 2, nz)
    if s.variant == native.VARIANT_M:
        # M:370 `-(z_g(iz,dz,C)-dz/2)*ρ*g + 0*yc[iy] + 0*zc[iz]`: with g = 0 three signed zeros, so the sign of the sum
        # depends on iy too -- the three terms stay apart (ns3d_fill_profile_zy adds them on the device)
        zg = np.array([s.grid.x_g(iz, s.dz, nz, 2) for iz in range(1, nz + 1)])
        yc = _linrange(-(s.ly - s.dy) / 2, (s.ly - s.dy) / 2, s.ny)
        return {"Pr": (-(zg - s.dz / 2) * s.rho * s.g, 0 * yc, 0 * zc)}
    prof = s.vin * (7.0 / 6.0) * np.power((zc + s.lz / 2) / s.lz, 1.0 / 6.0)      # host pow: the reference's `^`
    return {"Vx": prof + 0.0 + 0.0, "Pr": (-(zc - s.lz / 2) * s.rho * s.g) + 0.0 + 0.0}


class Simulation:
    """Device state of one rank + the time step."""

    def __init__(self, setup: Setup, ctx: native.Context | None = None, *, device: int | None = None,
                 mode: int = native.FAST, host_fields: dict | None = None, zchunk: int = 0):
        self.s = setup
        if ctx is None:
            if device is None:
                device = int(os.environ.get("LOCAL_RANK", "0"))
            ctx = native.Context(device, mode)
        self.ctx = ctx
        self.zchunk = zchunk
        self.f = {name: ctx.zeros(*shape) for name, shape in setup.shapes().items()}   # M:343-360
        if host_fields is None:
            # device-side initialisers: nz profile values cross PCIe, not 3-D arrays (M:369-370 / G:86-87)
            for name, prof in initial_profiles(setup).items():
                if isinstance(prof, tuple):
                    ctx.fill_profile_zy(self.f[name], *prof)
                else:
                    ctx.fill_profile_z(self.f[name], prof)
            if setup.variant == native.VARIANT_M:
                ctx.fill_plane_x(self.f["Vy"], 0, setup.vin)                          # M:369 (sic: Vy)
        else:
            for name, arr in host_fields.items():
                self.f[name].set(arr)
        if host_fields is None:
            if setup.variant == native.VARIANT_M:
                self.update_halo("Pr")                                                 # M:371
                self.set_cylinder()                                                    # M:372
                self.update_halo("C", "Vx", "Vy", "Vz")                                # M:373
        self._fields_struct = native.Fields()
        for name in native.FIELD_NAMES:
            setattr(self._fields_struct, name, self.f[name].ptr)
        self.iters: list[int] = []
        self.err_hist: list[list[float]] = []

    # -- pieces of the time loop -----------------------------------------------------------------
    def update_halo(self, *names):
        if self.s.grid.nranks > 1:
            self.ctx.update_halo([self.f[n] for n in names], self.s.nz)

    def set_cylinder(self):
        s, f, c = self.s, self.f, self.ctx
        if s.variant == native.VARIANT_M:
            c.call("ns3d_set_cylinder_M", f["C"], f["Vx"], f["Vy"], f["Vz"], s.a2, s.b2, s.ox, s.oy, s.sinb, s.cosb,
                   s.xco_g, s.yco_g, s.dx, s.dy, s.nx, s.ny, s.nz)
        else:
            c.call("ns3d_set_cylinder_G", f["C"], f["Vx"], f["Vy"], f["Vz"], s.a2, s.b2, s.ox, s.oy, s.sinb, s.cosb,
                   s.lx, s.ly, s.dx, s.dy, s.nx, s.ny, s.nz)

    def set_bc_Pr(self):
        s, f, c = self.s, self.f, self.ctx
        if s.variant == native.VARIANT_M:
            c.call("ns3d_set_bc_Pr_M", f["Pr"], int(s.outlet_guard), 0.0, s.nx, s.ny, s.nz)
        else:
            c.call("ns3d_set_bc_Pr_G", f["Pr"], s.dz, s.nz, s.g, s.rho, s.nx, s.ny, s.nz)

    def set_bc_Vel(self):
        s, f, c = self.s, self.f, self.ctx
        if s.variant == native.VARIANT_M:
            c.call("ns3d_set_bc_Vel_M", f["Vx"], f["Vy"], f["Vz"], int(s.inlet_guard), s.vin, s.nx, s.ny, s.nz)
        else:
            c.call("ns3d_set_bc_Vel_G", f["Vx"], f["Vy"], f["Vz"], s.nx, s.ny, s.nz)

    def step(self):
        """One time step through the fused entry point ``ns3d_step``."""
        it, hist = self.ctx.step(self._fields_struct, self.s.step_params(self.zchunk))
        self.iters.append(it)
        self.err_hist.append(hist)
        return it, hist

    def step_groups(self):
        """The same time step through the four level-2 groups: ``ns3d_predictor``, ``ns3d_pt_solve``,
        ``ns3d_corrector``, ``ns3d_advect_swap`` -- what ``ns3d_step`` composes, for a driver that
        wants to act between them (the script's ``println`` of the residuals sits there, M:468)."""
        sp = self.s.step_params(self.zchunk)
        self.ctx.predictor(self._fields_struct, sp)
        it, hist = self.ctx.pt_solve(self.f["Pr"], self.f["dPrdtau"], self.f["divV"], sp.pt)
        self.ctx.corrector(self._fields_struct, sp)
        self.ctx.advect_swap(self._fields_struct, sp)
        self.iters.append(it)
        self.err_hist.append(hist)
        return it, hist

    def step_level1(self):
        """The same time step, call site by call site (M:449-477 / G:121-142) through level 1."""
        s, f, c = self.s, self.f, self.ctx
        n = (s.nx, s.ny, s.nz)
        c.call("ns3d_update_tau", f["txx"], f["tyy"], f["tzz"], f["txy"], f["txz"], f["tyz"], f["Vx"], f["Vy"],
               f["Vz"], s.mu, s.dx, s.dy, s.dz, *n)                                            # M:449
        self.update_halo("txx", "tyy", "tzz")                                                  # M:450
        c.call("ns3d_predict_V", f["Vx"], f["Vy"], f["Vz"], f["txx"], f["tyy"], f["tzz"], f["txy"], f["txz"],
               f["tyz"], s.rho, s.g, s.dt, s.dx, s.dy, s.dz, *n)                               # M:451
        self.set_cylinder()                                                                    # M:452
        self.update_halo("C", "Vx", "Vy", "Vz")                                                # M:453
        c.call("ns3d_update_divV", f["divV"], f["Vx"], f["Vy"], f["Vz"], s.dx, s.dy, s.dz, *n)  # M:454
        self.update_halo("divV")                                                               # M:455
        iters, hist = 0, []
        for it in range(1, s.niter + 1):                                                       # M:458
            c.call("ns3d_update_dPrdtau", f["Pr"], f["dPrdtau"], f["divV"], s.rho, s.dt, s.dtau, s.damp, s.dx,
                   s.dy, s.dz, *n)                                                             # M:459
            c.call("ns3d_update_Pr", f["Pr"], f["dPrdtau"], s.dtau, *n)                        # M:461
            self.set_bc_Pr()                                                                   # M:463 (halo inside)
            iters = it
            if it % s.nchk == 0:                                                               # M:464
                c.call("ns3d_compute_res", f["Rp"], f["Pr"], f["divV"], s.rho, s.dt, s.dx, s.dy, s.dz, *n)
                err = c.max_abs(f["Rp"]) * (s.ly * s.ly) / s.psc                               # M:466
                hist.append(err)
                if err < s.eps_it or not np.isfinite(err):                                     # M:469
                    break
        c.call("ns3d_correct_V", f["Vx"], f["Vy"], f["Vz"], f["Pr"], s.dt, s.rho, s.dx, s.dy, s.dz, *n)   # M:472
        self.set_cylinder()                                                                    # M:473
        self.set_bc_Vel()                                                                      # M:474
        for a in ("Vx", "Vy", "Vz", "C"):                                                      # M:475
            c.copy(f[a + "_o"], f[a])
        c.call("ns3d_advect", f["Vx"], f["Vx_o"], f["Vy"], f["Vy_o"], f["Vz"], f["Vz_o"], f["C"], f["C_o"], s.dt,
               s.dx, s.dy, s.dz, *n)                                                           # M:476
        self.update_halo("Vx", "Vy", "Vz")                                                     # M:477
        self.iters.append(iters)
        self.err_hist.append(hist)
        return iters, hist

    # -- results ---------------------------------------------------------------------------------
    def host(self, name: str) -> np.ndarray:
        return self.f[name].to_host()

    def interior(self, name: str, dtype=np.float64, drop_last_z: bool = False) -> np.ndarray:
        """``Array(A)[2:end-1,2:end-1,2:end-1]`` (M:399-403, 528-532), extracted on the device: only
        the interior crosses PCIe.  ``dtype=np.float32`` is ``convert.(Float32, .)`` (M:408);
        ``drop_last_z`` leaves out the last interior plane (z-slab gather of ``Vz``)."""
        a = self.f[name]
        sx, sy, sz = a.shape
        return self.ctx.box(a, (1, sx - 1), (1, sy - 1), (1, sz - 1 - int(drop_last_z)), dtype)

    def slice_xy(self, name: str, dtype=np.float64) -> np.ndarray:
        """The horizontal heat-map plane of the visualisation block, ``A_v[:, :, ceil(Int, nz_g()/2)]``
        (M:422-426), of the LOCAL interior: one x-y plane leaves the device."""
        a = self.f[name]
        sx, sy, sz = a.shape
        k = -(-self.s.nz // 2)            # ceil(nz/2), 1-based index into the interior = 0-based index into A
        return self.ctx.box(a, (1, sx - 1), (1, sy - 1), (k, k + 1), dtype)[:, :, 0]

    def slice_xz(self, name: str, dtype=np.float64) -> np.ndarray:
        """The vertical heat-map plane ``A_v[:, ceil(Int, ny_g()/2), :]`` (M:428-432)."""
        a = self.f[name]
        sx, sy, sz = a.shape
        j = -(-self.s.ny // 2)
        return self.ctx.box(a, (1, sx - 1), (j, j + 1), (1, sz - 1), dtype)[:, 0, :]


class StreamedSteps:
    """Time steps fed from and drained to HOST memory every step, with the traffic hidden behind the computation.

    Two sets of device fields on one context: while step n computes on one set, the input state of step n+1 is
    uploaded into the other (``ns3d_h2d_async``, the context's upload stream) and the result of step n-1 is
    downloaded from it (``ns3d_d2h_async``, the download stream); ``ns3d_stream_wait`` orders the three streams.
    Host buffers must be page-locked: ``inputs`` / ``outputs`` map field names to ``(address, count)``.
    Nothing in the reference corresponds to this (its ``Data.Array(x)`` / ``Array(A)`` are blocking); it is what
    bench.py's end-to-end leg calls."""

    def __init__(self, setup: Setup, ctx: native.Context, zchunk: int = 0):
        self.ctx = ctx
        self.sims = [Simulation(setup, ctx, zchunk=zchunk), Simulation(setup, ctx, zchunk=zchunk)]

    def run(self, steps: int, inputs: dict, outputs: dict):
        c, N = self.ctx, native
        results = []

        def upload(sim):
            for name, (addr, count) in inputs.items():
                c.h2d_async(sim.f[name].ptr, addr, count)

        def download(sim):
            for name, (addr, count) in outputs.items():
                c.d2h_async(addr, sim.f[name].ptr, count)

        c.stream_wait(N.STREAM_H2D, N.STREAM_COMPUTE)      # whatever still computes on set 0 finishes first
        upload(self.sims[0])
        for n in range(steps):
            cur, nxt = self.sims[n % 2], self.sims[(n + 1) % 2]
            c.stream_wait(N.STREAM_COMPUTE, N.STREAM_H2D)  # the input of this step has arrived
            if n + 1 < steps:
                c.stream_wait(N.STREAM_H2D, N.STREAM_D2H)  # the other set's previous result has left
                upload(nxt)                                # ... and its next input travels while this step computes
            results.append(cur.step())
            c.stream_wait(N.STREAM_D2H, N.STREAM_COMPUTE)  # corrector and advection of this step are done
            download(cur)
        c.stream_sync(N.STREAM_H2D)
        c.stream_sync(N.STREAM_D2H)
        return results

    def close(self):
        for sim in self.sims:
            for a in sim.f.values():
                self.ctx.free(a)


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def attach_communicator(ctx: native.Context, rank: int, world: int):
    """Bootstraps the library's NCCL communicator: rank 0 creates the id, torch.distributed
    carries the 128 bytes to the other ranks (what MPI.bcast does in the Julia shim)."""
    if world == 1:
        ctx.comm_init(0, 1, b"\0" * 128)
        return
    import torch.distributed as dist
    if not dist.is_initialized():
        raise native.NS3DError("torch.distributed must be initialised before attach_communicator()")
    box = [ctx.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(rank, world, box[0])


def run_navierstokes3D(*, do_vis: bool = False, do_save: bool = False, do_print: bool = False, nx: int = 255,
                       nt: int = 10, ny: int | None = None, nz: int | None = None, mode: int = native.FAST,
                       level1: bool = False, return_sim: bool = False, nsave: int | None = None):
    """Drop-in for ``run_navierstokes3D`` (M:287): same keywords, same return value.

    Under ``torchrun`` (WORLD_SIZE > 1, torch.distributed initialised by the caller) the domain is
    split into z-slabs, one rank per GPU, with the script's local ``nx, ny, nz`` per rank.
    ``do_vis`` is accepted and ignored (plotting is out of scope); ``do_save`` writes the
    script's Float32 frames ``out_save/out_<A>_v_%04d.bin`` on rank 0: frame 0 holds the initial
    conditions (M:404-413), then one frame every ``nsave`` (M:332: 10) time steps (M:515-523).
    """
    nsave = NSAVE if nsave is None else nsave
    rank, world, local = _dist_env()
    s = setup_multi_gpu(nx, ny=ny, nz=nz, rank=rank, nranks=world)
    ctx = native.Context(local, mode)
    attach_communicator(ctx, rank, world)
    sim = Simulation(s, ctx)
    iframe = 0
    if do_save:
        save_frame(sim, iframe)                                          # initial conditions, M:404-413
    for it in range(1, nt + 1):
        if rank == 0 and do_print:
            print(f"#it = {it}")                                         # M:456
        iters, hist = sim.step_level1() if level1 else sim.step()
        if rank == 0 and do_print:
            for c, err in enumerate(hist, 1):
                print("  #iter = %d, err = %1.3e" % (min(c * s.nchk, iters), err))   # M:468
        if do_save and it % nsave == 0:                                  # M:479-523
            iframe += 1
            save_frame(sim, iframe)
    out = tuple(gather_interior(sim, name) for name in ("C", "Pr", "Vx", "Vy", "Vz"))   # M:528-535
    if return_sim:
        return out, sim
    ctx.close()
    return out


NSAVE = 10   # M:332 `nsave`: time steps between two saved frames


def save_array(aname: str, a: np.ndarray) -> str:
    """``save_array(Aname, A)`` (M:27-30): the raw column-major bytes of A in ``Aname.bin``."""
    fname = aname + ".bin"
    with open(fname, "wb") as fh:
        fh.write(np.asfortranarray(a).tobytes(order="F"))
    return fname


def save_frame(sim: Simulation, iframe: int, outdir: str = "out_save"):
    """One frame of the script's ``do_save`` output (M:404-413, 515-523): the gathered interior of
    C, Pr, Vx, Vy, Vz as Float32 in ``out_save/out_<A>_v_%04d.bin`` on rank 0.  The conversion to
    Float32 happens on the device, so half the bytes cross PCIe."""
    for name in ("C", "Pr", "Vx", "Vy", "Vz"):
        g = gather_interior(sim, name, dtype=np.float32)
        if sim.s.grid.rank == 0:
            os.makedirs(outdir, exist_ok=True)
            save_array(os.path.join(outdir, "out_%s_v_%04d" % (name, iframe)), g)


def gather_interior(sim: Simulation, name: str, dtype=np.float64) -> np.ndarray | None:
    """``gather!(A_inn, A_v)`` (M:399-403): interior blocks concatenated along z on rank 0.

    The reference's own multi-rank gather of the staggered fields is shape-inconsistent
    (SURVEY.md quirk 10); here each rank contributes its interior planes and, for the field
    staggered along the split dimension (Vz), the last rank contributes the extra plane.
    """
    s = sim.s
    world, rank = s.grid.nranks, s.grid.rank
    if world == 1:
        return sim.interior(name, dtype)
    # every rank packs its interior planes on its device; the blocks go to rank 0 over NCCL
    # (ns3d_gather_box) and arrive concatenated along z
    a = sim.f[name]
    sx, sy, sz = a.shape
    nplanes = [(sz - 2) - int(name == "Vz" and r < world - 1) for r in range(world)]
    return sim.ctx.gather_box(a, (1, sx - 1), (1, sy - 1), (1, 1 + nplanes[rank]), nplanes, dtype)


def save_mat(fname: str, sim: Simulation, initial: bool = False):
    """The single-GPU script's ``.mat`` dumps.

    Periodic frames, ``matwrite("out_save/step_$it.mat", Dict("Pr"=>Array(Pr), "Vx"=>Array(Vx), "Vy"=>Array(Vy),
    "Vz"=>Array(Vz), "C"=>Array(C), "dx"=>dx, "dy"=>dy, "dz"=>dz))`` (G:169): eight distinct keys.
    ``initial=True`` is ``step_0.mat`` (G:89), whose Dict literal names the key "Vy" twice -- Julia keeps
    the last pair, so that file holds ``Vz`` under "Vy" and has neither the true ``Vy`` nor a "Vz" key
    (SURVEY.md quirk 9); reproduced as written.  Neither call compresses."""
    from scipy.io import savemat
    s = sim.s
    d = {"Pr": sim.host("Pr"), "Vx": sim.host("Vx"), "C": sim.host("C"), "dx": s.dx, "dy": s.dy, "dz": s.dz}
    if initial:
        d["Vy"] = sim.host("Vz")
    else:
        d["Vy"], d["Vz"] = sim.host("Vy"), sim.host("Vz")
    savemat(fname, d)


def runme(*, do_vis: bool = True, do_save: bool = False, nx: int = 255, nt: int = 10000, mode: int = native.FAST,
          do_print: bool = True, return_sim: bool = False, nsave: int | None = None):
    """Drop-in for ``runme`` of the single-GPU script (G:12); ``nx``/``nt``/``nsave`` are literals there (G:44,51,52)."""
    nsave = NSAVE if nsave is None else nsave
    s = setup_gpu(nx)
    sim = Simulation(s, native.Context(_dist_env()[2], mode))
    if do_save:                                                          # G:89
        os.makedirs("out_save", exist_ok=True)
        save_mat("out_save/step_0.mat", sim, initial=True)
    for it in range(1, nt + 1):
        if do_print:
            print(f"#it = {it}")                                         # G:125
        iters, hist = sim.step()
        if do_print:
            for c, err in enumerate(hist, 1):
                print("  #iter = %d, err = %1.3e" % (min(c * s.nchk, iters), err))   # G:134
        if do_save and it % nsave == 0:                                  # G:168-170
            os.makedirs("out_save", exist_ok=True)
            save_mat(f"out_save/step_{it}.mat", sim)
    if return_sim:
        return sim
    sim.ctx.close()
    return None
