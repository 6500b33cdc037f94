"""Host-side set-up of a run: the scalars the reference scripts derive before their time loop.

``setup_multi_gpu`` follows scripts/NavierStokes3D_multi_gpu.jl (M:290-341, 363-367) for one
rank of a z-slab decomposition -- the native replacement of ``init_global_grid(nx,ny,nz;
dimx=1,dimy=1,dimz=N)`` -- and ``setup_gpu`` follows scripts/NavierStokes3D_gpu.jl (G:15-61).
All expressions are evaluated in IEEE double in the scripts' association order, because two
of them feed float ``==`` guards (M:164, M:179) whose outcome selects boundary conditions.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

from . import native


@dataclass(frozen=True)
class SlabGrid:
    """ImplicitGlobalGrid-compatible geometry for dims = (1, 1, nranks), overlap 2, halo width 1."""
    nx: int
    ny: int
    nz: int          # local sizes (what the reference passes to init_global_grid)
    rank: int = 0
    nranks: int = 1

    @property
    def dims(self):
        return (1, 1, self.nranks)

    @property
    def coords(self):
        return (0, 0, self.rank)

    # nx_g() = dims*(nx-overlap)+overlap
    @property
    def nx_g(self) -> int:
        return self.nx

    @property
    def ny_g(self) -> int:
        return self.ny

    @property
    def nz_g(self) -> int:
        return self.nranks * (self.nz - 2) + 2

    def x_g(self, i: int, d: float, size_a: int, dim: int) -> float:
        """IGG ``x_g/y_g/z_g(i, d, A)`` for 1-based index i of an array with size_a points along dim."""
        n = (self.nx, self.ny, self.nz)[dim]
        coord = self.coords[dim]
        x0 = 0.5 * (n - size_a) * d
        return (coord * (n - 2) + i - 1) * d + x0

    def z_range_global(self):
        """0-based global z-plane indices [lo, hi) that this rank's local planes map to."""
        lo = self.rank * (self.nz - 2)
        return lo, lo + self.nz


def halo_planes(sz: int, nz: int):
    """Planes ``update_halo!`` moves along a split dimension (IGG, overlap 2, halo width 1).

    For a field with ``sz`` points along the dimension whose local cell count is ``nz`` the overlap
    is ol = 2 + (sz - nz).  Returns 0-based ``(send_lo, recv_lo, send_hi, recv_hi)``: plane
    ``send_lo`` goes to the lower neighbour's ``recv_hi``, ``send_hi`` to the upper neighbour's
    ``recv_lo`` (cell-centred: planes 1 / nz-2; face-staggered sz = nz+1: planes 2 / nz-2).
    The C library applies the same rule in ``ns3d_update_halo``.
    """
    ol = 2 + (sz - nz)
    if ol < 2:
        raise ValueError(f"a field with {sz} planes on {nz} cells has overlap {ol} < 2 and cannot be exchanged")
    return ol - 1, 0, sz - ol, sz - 1


@dataclass
class Setup:
    variant: int                 # native.VARIANT_M | native.VARIANT_G
    grid: SlabGrid
    lx: float
    ly: float
    lz: float
    dx: float
    dy: float
    dz: float
    dt: float
    dtau: float
    damp: float
    rho: float
    mu: float
    g: float
    vin: float
    psc: float
    a2: float
    b2: float
    ox: float
    oy: float
    sinb: float
    cosb: float
    xco_g: float
    yco_g: float
    zco_g: float
    eps_it: float
    niter: int
    nchk: int
    inlet_guard: bool
    outlet_guard: bool

    @property
    def nx(self):
        return self.grid.nx

    @property
    def ny(self):
        return self.grid.ny

    @property
    def nz(self):
        return self.grid.nz

    def shapes(self) -> dict:
        """Allocation shapes of the 18 fields (M:343-360)."""
        nx, ny, nz = self.nx, self.ny, self.nz
        cell, edge, inner = (nx, ny, nz), (nx - 1, ny - 1, nz - 1), (nx - 2, ny - 2, nz - 2)
        return {"Pr": cell, "dPrdtau": inner, "C": cell, "C_o": cell, "txx": cell, "tyy": cell, "tzz": cell,
                "txy": edge, "txz": edge, "tyz": edge, "Vx": (nx + 1, ny, nz), "Vy": (nx, ny + 1, nz),
                "Vz": (nx, ny, nz + 1), "Vx_o": (nx + 1, ny, nz), "Vy_o": (nx, ny + 1, nz),
                "Vz_o": (nx, ny, nz + 1), "divV": cell, "Rp": inner}

    def pt_params(self, zchunk: int = 0) -> native.PtParams:
        p = native.PtParams()
        p.nx, p.ny, p.nz, p.variant = self.nx, self.ny, self.nz, self.variant
        p.rho, p.dt, p.dtau, p.damp = self.rho, self.dt, self.dtau, self.damp
        p.dx, p.dy, p.dz = self.dx, self.dy, self.dz
        p.eps_it = self.eps_it
        p.err_num, p.err_den = self.ly * self.ly, self.psc      # err = max*ly^2/psc  (M:466)
        p.niter, p.nchk = self.niter, self.nchk
        p.outlet_guard, p.outlet_val = int(self.outlet_guard), 0.0
        p.g = self.g
        p.zchunk = zchunk
        return p

    def step_params(self, zchunk: int = 0) -> native.StepParams:
        s = native.StepParams()
        s.pt = self.pt_params(zchunk)
        s.mu, s.vin = self.mu, self.vin
        s.a2, s.b2, s.ox, s.oy, s.sinb, s.cosb = self.a2, self.b2, self.ox, self.oy, self.sinb, self.cosb
        s.xco_g, s.yco_g, s.lx, s.ly = self.xco_g, self.yco_g, self.lx, self.ly
        s.inlet_guard = int(self.inlet_guard)
        return s


@dataclass(frozen=True)
class Physics:
    """The literals the reference's run functions hard-code (M:290-335 / G:15-56), as a config
    (SURVEY.md 8f row 3).  ``None`` = the script's value, which differs between the two scripts for
    ``ox_lx`` (M: -0.4, G: -0.3) and gravity (M: Fr = Inf, i.e. g = 0; G: g = 9.81)."""
    rho: float = 1000.0                  # density                       M:291
    vin: float = 1.0                     # inflow velocity               M:292
    mu: float = 0.001                    # dynamic viscosity             M:293  (Re = rho*vin*lx/mu)
    a_lx: float = 0.05                   # obstacle half-axes / lx       M:304-305
    b_lx: float = 0.05
    ox_lx: float | None = None           # obstacle centre / lx          M:307-308 / G:29-30
    oy_lx: float = 0.0
    beta: float = 0 * math.pi / 6        # obstacle rotation             M:309
    g: float | None = None               # gravity                       M:316 / G:38
    cfl_tau: float = 1.0 / math.sqrt(3.1)   # pseudo-time step           M:333
    cfl_visc: float = 1 / 4.1            # viscous time-step limit       M:334
    cfl_adv: float = 1.0                 # advective time-step limit     M:335


def _common(ph: Physics):
    lx = 1.0                                            # M:290 / G:15
    psc = ph.rho * ph.vin * ph.vin                      # M:296
    return lx, ph.rho, ph.vin, ph.mu, psc, math.sin(ph.beta), math.cos(ph.beta)


def _time_steps(dx, dy, dz, ph: Physics):
    dmax = max(dx, dy, dz)
    dt = min(ph.cfl_visc * (dmax * dmax) * ph.rho / ph.mu, ph.cfl_adv * dmax / ph.vin)   # M:339
    return dt, ph.cfl_tau * dmax                                                         # M:341


def setup_multi_gpu(nx: int = 255, *, ny: int | None = None, nz: int | None = None, rank: int = 0,
                    nranks: int = 1, eps_it: float = 1e-3, niter: int | None = None, nchk: int | None = None,
                    ly: float | None = None, lz: float | None = None, physics: Physics | None = None) -> Setup:
    """Scalars of ``run_navierstokes3D`` (M:290-341) on one rank of a (1,1,nranks) process grid.

    ``ny, nz, ly, lz, niter, nchk, eps_it`` default to the script's rules; the overrides exist for
    the benchmark configurations that name explicit sizes (511^3, 1023x511x511) or fixed work;
    ``physics`` replaces the literals of M:290-335 (density, viscosity, obstacle, CFL numbers).
    """
    ph = physics or Physics()
    lx, rho, vin, mu, psc, sinb, cosb = _common(ph)
    ly_lx = lz_lx = 0.6
    ly = ly_lx * lx if ly is None else ly
    lz = lz_lx * lx if lz is None else lz
    ox, oy = (-0.4 if ph.ox_lx is None else ph.ox_lx) * lx, ph.oy_lx * lx   # M:307-308,314-315
    g = 1 / (math.inf * math.inf) * (vin * vin) / lx if ph.g is None else ph.g   # Fr = Inf -> 0.0  (M:301,316)
    a2 = (ph.a_lx * lx) * (ph.a_lx * lx)                # M:317
    b2 = (ph.b_lx * lx) * (ph.b_lx * lx)
    if ny is None:
        ny = math.ceil(nx * ly_lx)                      # M:323
    if nz is None:
        nz = math.ceil(nx * lz_lx)                      # M:324
    grid = SlabGrid(nx, ny, nz, rank, nranks)
    if niter is None:
        niter = 50 * max(grid.nx_g, grid.ny_g, grid.nz_g)   # M:328
    if nchk is None:
        nchk = grid.ny_g - 1                                # M:329
    dx, dy, dz = lx / grid.nx_g, ly / grid.ny_g, lz / grid.nz_g   # M:338
    dt, dtau = _time_steps(dx, dy, dz, ph)
    damp = 2 / nx                                       # M:340: the LOCAL nx
    xco_g = grid.x_g(1, dx, nx, 0) - (lx - dx) / 2      # M:363
    yco_g = grid.x_g(1, dy, ny, 1) - (ly - dy) / 2      # M:364
    zco_g = grid.x_g(1, dz, nz, 2) - (lz - dz) / 2      # M:365
    xvo_g = grid.x_g(1, dx, nx + 1, 0) - (lx - dx) / 2  # M:366
    xve_g = grid.x_g(nx + 1, dx, nx + 1, 0) - (lx - dx) / 2   # M:367
    return Setup(variant=native.VARIANT_M, grid=grid, lx=lx, ly=ly, lz=lz, dx=dx, dy=dy, dz=dz, dt=dt, dtau=dtau,
                 damp=damp, rho=rho, mu=mu, g=g, vin=vin, psc=psc, a2=a2, b2=b2, ox=ox, oy=oy, sinb=sinb,
                 cosb=cosb, xco_g=xco_g, yco_g=yco_g, zco_g=zco_g, eps_it=eps_it, niter=niter, nchk=nchk,
                 inlet_guard=(xvo_g == -lx / 2),        # M:164, float == as written
                 outlet_guard=(xve_g == lx / 2))        # M:179


def setup_gpu(nx: int = 255, *, ny: int | None = None, nz: int | None = None, eps_it: float = 1e-3,
              niter: int | None = None, nchk: int | None = None, physics: Physics | None = None) -> Setup:
    """Scalars of ``runme`` (G:15-61); the script hard-codes nx = 255 (G:44)."""
    ph = physics or Physics()
    lx, rho, vin, mu, psc, sinb, cosb = _common(ph)
    ly, lz = 0.6 * lx, 0.6 * lx                         # G:34-35
    ox, oy = (-0.3 if ph.ox_lx is None else ph.ox_lx) * lx, ph.oy_lx * lx   # G:29-30,36-37
    g = 9.81 if ph.g is None else ph.g                  # G:38
    a2 = (ph.a_lx * lx) * (ph.a_lx * lx)
    b2 = (ph.b_lx * lx) * (ph.b_lx * lx)
    if ny is None:
        ny = math.ceil(nx * 0.6)                        # G:45
    if nz is None:
        nz = math.ceil(nx * 0.6)                        # G:46
    if niter is None:
        niter = 50 * max(ny, nz)                        # G:48
    if nchk is None:
        nchk = ny - 1                                   # G:49
    dx, dy, dz = lx / nx, ly / ny, lz / nz              # G:58
    dt, dtau = _time_steps(dx, dy, dz, ph)              # G:59,61
    return Setup(variant=native.VARIANT_G, grid=SlabGrid(nx, ny, nz), lx=lx, ly=ly, lz=lz, dx=dx, dy=dy, dz=dz,
                 dt=dt, dtau=dtau, damp=2 / nx, rho=rho, mu=mu, g=g, vin=vin, psc=psc, a2=a2, b2=b2, ox=ox,
                 oy=oy, sinb=sinb, cosb=cosb, xco_g=0.0, yco_g=0.0, zco_g=0.0, eps_it=eps_it, niter=niter,
                 nchk=nchk, inlet_guard=False, outlet_guard=False)
