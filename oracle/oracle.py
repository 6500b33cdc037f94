"""ctypes front end of the CPU oracle (oracle/ns3d_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never by the product package
``navierstokes3d_b200``.  PARITY UNPINNED by a run of the reference (Julia + un-vendored
ParallelStencil/ImplicitGlobalGrid cannot run in the build container; see the header of
ns3d_oracle.c) and pinned to the reference's SOURCE TEXT instead: ``jl_interp.py`` /
``jl_run.py`` execute the scripts' own text, and this oracle agrees with that bit for bit
(tests/test_jl_reference.py).

Everything host-side that the reference scripts do around the kernels is restated here
literally as well, so that oracle runs need nothing from the product package:

* parameter derivation            M:290-341 (``params_M``)  /  G:15-61 (``params_G``)
* ImplicitGlobalGrid arithmetic   ``nx_g``, ``x_g`` (IGG source is not in the reference tree;
                                   restated from its documented behaviour)
* initial conditions              M:369-373 / G:85-88
* allocation shapes               M:343-360

(M:n = scripts/NavierStokes3D_multi_gpu.jl line n, G:n = scripts/NavierStokes3D_gpu.jl line n.)
Arrays are numpy float64, Fortran order, shape (sx, sy, sz) == Julia ``Array{Float64,3}``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libns3d_oracle.so")

c_double_p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "ns3d_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Params(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("variant", C.c_int),
        ("lx", C.c_double), ("ly", C.c_double), ("lz", C.c_double),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
        ("dt", C.c_double), ("dtau", C.c_double), ("damp", C.c_double),
        ("rho", C.c_double), ("mu", C.c_double), ("g", C.c_double), ("vin", C.c_double), ("psc", C.c_double),
        ("a2", C.c_double), ("b2", C.c_double), ("ox", C.c_double), ("oy", C.c_double),
        ("sinb", C.c_double), ("cosb", C.c_double), ("xco_g", C.c_double), ("yco_g", C.c_double),
        ("eps_it", C.c_double),
        ("niter", C.c_int), ("nchk", C.c_int), ("inlet_guard", C.c_int), ("outlet_guard", C.c_int),
    ]


_FIELD_NAMES = ["Pr", "dPrdtau", "C", "C_o", "txx", "tyy", "tzz", "txy", "txz", "tyz",
                "Vx", "Vy", "Vz", "Vx_o", "Vy_o", "Vz_o", "divV", "Rp", "absRp"]


class _Fields(C.Structure):
    _fields_ = [(n, c_double_p) for n in _FIELD_NAMES]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        # Spinning OpenMP barriers collapse on oversubscribed / SMT hosts (measured here: 8 threads
        # slower than 1); passive waiting restores scaling.  Must be set before libgomp starts.
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_max_abs.restype = C.c_double
        _lib.oracle_sizeof_params.restype = C.c_size_t
        _lib.oracle_sizeof_fields.restype = C.c_size_t
        assert _lib.oracle_sizeof_params() == C.sizeof(_Params)
        assert _lib.oracle_sizeof_fields() == C.sizeof(_Fields)
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.f_contiguous, "oracle arrays are float64, Fortran order"
    return a.ctypes.data_as(c_double_p)


def _d(x):
    return C.c_double(float(x))


def zeros(*shape) -> np.ndarray:
    return np.zeros(shape, dtype=np.float64, order="F")


# ----------------------------------------------------------------------------------------------
# ImplicitGlobalGrid arithmetic (non-periodic, overlap 2)
# ----------------------------------------------------------------------------------------------
def n_g(n: int, dim: int) -> int:
    """nx_g() = dims*(nx-overlap)+overlap with overlap 2."""
    return dim * (n - 2) + 2


def x_g(i: int, dx: float, size_A: int, n: int, coord: int) -> float:
    """IGG ``x_g(ix,dx,A)``: x0 = 0.5*(nx-size(A,1))*dx; x = (coord*(nx-2) + ix-1)*dx + x0."""
    x0 = 0.5 * (n - size_A) * dx
    return (coord * (n - 2) + i - 1) * dx + x0


def linrange(start: float, stop: float, n: int) -> np.ndarray:
    """Julia ``LinRange(start,stop,n)[i] = lerpi(i-1, max(n-1,1), start, stop) = (1-t)*a + t*b``."""
    d = max(n - 1, 1)
    t = np.arange(n, dtype=np.float64) / d
    return (1 - t) * start + t * stop


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
@dataclass
class Params:
    variant: str  # "M" or "G"
    nx: int
    ny: int
    nz: int
    dims: tuple = (1, 1, 1)
    coords: tuple = (0, 0, 0)
    lx: float = 1.0
    ly: float = 0.6
    lz: float = 0.6
    dx: float = 0.0
    dy: float = 0.0
    dz: float = 0.0
    dt: float = 0.0
    dtau: float = 0.0
    damp: float = 0.0
    rho: float = 1000.0
    mu: float = 0.001
    g: float = 0.0
    vin: float = 1.0
    psc: float = 1000.0
    a2: float = 0.0
    b2: float = 0.0
    ox: float = 0.0
    oy: float = 0.0
    sinb: float = 0.0
    cosb: float = 1.0
    xco_g: float = 0.0
    yco_g: float = 0.0
    zco_g: float = 0.0
    xvo_g: float = 0.0
    xve_g: float = 0.0
    eps_it: float = 1e-3
    niter: int = 0
    nchk: int = 0
    inlet_guard: bool = True
    outlet_guard: bool = True
    extra: dict = field(default_factory=dict)

    def cstruct(self) -> _Params:
        s = _Params()
        for name, _ in _Params._fields_:
            if name == "variant":
                s.variant = 0 if self.variant == "M" else 1
            elif name in ("inlet_guard", "outlet_guard"):
                setattr(s, name, int(getattr(self, name)))
            else:
                setattr(s, name, getattr(self, name))
        return s


def params_M(nx: int = 255, ny: int | None = None, nz: int | None = None, dims=(1, 1, 1),
             coords=(0, 0, 0), eps_it: float = 1e-3, niter: int | None = None,
             nchk: int | None = None, ly: float | None = None, lz: float | None = None, **lit) -> Params:
    """Literal restatement of M:290-341 + M:363-367 for one rank of an IGG grid.

    ``ny``/``nz``/``ly``/``lz``/``niter``/``nchk`` overrides exist only for the BASELINE configs
    that give explicit dims (511^3, 1023x511x511); left as None they follow the script.  ``lit``
    replaces literals of the script (rho, vin, mu, a_lx, b_lx, ox_lx, oy_lx, beta, g, cfl_tau,
    cfl_visc, cfl_adv) -- what one edits in the source to run another case.
    """
    unknown = set(lit) - {"rho", "vin", "mu", "a_lx", "b_lx", "ox_lx", "oy_lx", "beta", "g", "cfl_tau", "cfl_visc", "cfl_adv"}
    assert not unknown, unknown
    lx = 1.0
    rho = lit.get("rho", 1000.0)
    vin = lit.get("vin", 1.0)
    mu = lit.get("mu", 0.001)
    psc = rho * vin * vin                      # M:296  ρ*vin^2
    Fr = math.inf                              # M:301
    ly_lx, lz_lx = 0.6, 0.6                    # M:302-303
    a_lx, b_lx = lit.get("a_lx", 0.05), lit.get("b_lx", 0.05)   # M:304-305
    ox_lx, oy_lx = lit.get("ox_lx", -0.4), lit.get("oy_lx", 0.0)   # M:307-308
    beta = lit.get("beta", 0 * math.pi / 6)    # M:309
    ly_ = ly_lx * lx if ly is None else ly     # M:312
    lz_ = lz_lx * lx if lz is None else lz     # M:313
    ox = ox_lx * lx
    oy = oy_lx * lx
    g = lit["g"] if "g" in lit else 1 / (Fr * Fr) * (vin * vin) / lx   # M:316 -> 0.0
    a2 = (a_lx * lx) * (a_lx * lx)             # M:317
    b2 = (b_lx * lx) * (b_lx * lx)
    sinb, cosb = math.sin(beta), math.cos(beta)
    if ny is None:
        ny = math.ceil(nx * ly_lx)             # M:323
    if nz is None:
        nz = math.ceil(nx * lz_lx)             # M:324
    nxg, nyg, nzg = n_g(nx, dims[0]), n_g(ny, dims[1]), n_g(nz, dims[2])
    if niter is None:
        niter = 50 * max(nxg, nyg, nzg)        # M:328
    if nchk is None:
        nchk = 1 * (nyg - 1)                   # M:329
    CFLtau = lit.get("cfl_tau", 1.0 / math.sqrt(3.1))   # M:333
    CFL_visc = lit.get("cfl_visc", 1 / 4.1)    # M:334
    CFL_adv = lit.get("cfl_adv", 1.0)          # M:335
    dx, dy, dz = lx / nxg, ly_ / nyg, lz_ / nzg                    # M:338
    dmax = max(dx, dy, dz)
    dt = min(CFL_visc * (dmax * dmax) * rho / mu, CFL_adv * dmax / vin)   # M:339
    damp = 2 / nx                              # M:340 (LOCAL nx)
    dtau = CFLtau * dmax                       # M:341
    xco_g = x_g(1, dx, nx, nx, coords[0]) - (lx - dx) / 2          # M:363
    yco_g = x_g(1, dy, ny, ny, coords[1]) - (ly_ - dy) / 2         # M:364
    zco_g = x_g(1, dz, nz, nz, coords[2]) - (lz_ - dz) / 2         # M:365
    xvo_g = x_g(1, dx, nx + 1, nx, coords[0]) - (lx - dx) / 2      # M:366
    xve_g = x_g(nx + 1, dx, nx + 1, nx, coords[0]) - (lx - dx) / 2  # M:367
    return Params(variant="M", nx=nx, ny=ny, nz=nz, dims=tuple(dims), coords=tuple(coords), lx=lx, ly=ly_,
                  lz=lz_, dx=dx, dy=dy, dz=dz, dt=dt, dtau=dtau, damp=damp, rho=rho, mu=mu, g=g, vin=vin,
                  psc=psc, a2=a2, b2=b2, ox=ox, oy=oy, sinb=sinb, cosb=cosb, xco_g=xco_g, yco_g=yco_g,
                  zco_g=zco_g, xvo_g=xvo_g, xve_g=xve_g, eps_it=eps_it, niter=niter, nchk=nchk,
                  inlet_guard=(xvo_g == -lx / 2),   # M:164
                  outlet_guard=(xve_g == lx / 2))   # M:179


def params_G(nx: int = 255, ny: int | None = None, nz: int | None = None, eps_it: float = 1e-3,
             niter: int | None = None, nchk: int | None = None, **lit) -> Params:
    """Literal restatement of G:15-61 (nx is hard-coded to 255 in the script, G:44); ``lit`` as in params_M."""
    unknown = set(lit) - {"rho", "vin", "mu", "a_lx", "b_lx", "ox_lx", "oy_lx", "beta", "g", "cfl_tau", "cfl_visc", "cfl_adv"}
    assert not unknown, unknown
    lx = 1.0
    rho = lit.get("rho", 1000.0)
    vin = lit.get("vin", 1.0)
    mu = lit.get("mu", 0.001)
    psc = rho * vin * vin
    ly_lx, lz_lx = 0.6, 0.6
    a_lx, b_lx = lit.get("a_lx", 0.05), lit.get("b_lx", 0.05)
    ox_lx, oy_lx = lit.get("ox_lx", -0.3), lit.get("oy_lx", 0.0)   # G:29-30
    beta = lit.get("beta", 0 * math.pi / 6)
    ly = ly_lx * lx
    lz = lz_lx * lx
    ox = ox_lx * lx
    oy = oy_lx * lx
    g = lit.get("g", 9.81)                     # G:38
    a2 = (a_lx * lx) * (a_lx * lx)
    b2 = (b_lx * lx) * (b_lx * lx)
    sinb, cosb = math.sin(beta), math.cos(beta)
    if ny is None:
        ny = math.ceil(nx * ly_lx)             # G:45
    if nz is None:
        nz = math.ceil(nx * lz_lx)             # G:46
    if niter is None:
        niter = 50 * max(ny, nz)               # G:48
    if nchk is None:
        nchk = 1 * (ny - 1)                    # G:49
    CFLtau = lit.get("cfl_tau", 1.0 / math.sqrt(3.1))
    CFL_visc = lit.get("cfl_visc", 1 / 4.1)
    CFL_adv = lit.get("cfl_adv", 1.0)
    dx, dy, dz = lx / nx, ly / ny, lz / nz     # G:58
    dmax = max(dx, dy, dz)
    dt = min(CFL_visc * (dmax * dmax) * rho / mu, CFL_adv * dmax / vin)   # G:59
    damp = 2 / nx                              # G:60
    dtau = CFLtau * dmax                       # G:61
    return Params(variant="G", nx=nx, ny=ny, nz=nz, lx=lx, ly=ly, lz=lz, dx=dx, dy=dy, dz=dz, dt=dt,
                  dtau=dtau, damp=damp, rho=rho, mu=mu, g=g, vin=vin, psc=psc, a2=a2, b2=b2, ox=ox, oy=oy,
                  sinb=sinb, cosb=cosb, eps_it=eps_it, niter=niter, nchk=nchk)


# ----------------------------------------------------------------------------------------------
# allocation + initial conditions
# ----------------------------------------------------------------------------------------------
def shapes(nx: int, ny: int, nz: int) -> dict:
    """Allocation shapes M:343-360 (= G:65-82)."""
    return {
        "Pr": (nx, ny, nz), "dPrdtau": (nx - 2, ny - 2, nz - 2), "C": (nx, ny, nz), "C_o": (nx, ny, nz),
        "txx": (nx, ny, nz), "tyy": (nx, ny, nz), "tzz": (nx, ny, nz),
        "txy": (nx - 1, ny - 1, nz - 1), "txz": (nx - 1, ny - 1, nz - 1), "tyz": (nx - 1, ny - 1, nz - 1),
        "Vx": (nx + 1, ny, nz), "Vy": (nx, ny + 1, nz), "Vz": (nx, ny, nz + 1),
        "Vx_o": (nx + 1, ny, nz), "Vy_o": (nx, ny + 1, nz), "Vz_o": (nx, ny, nz + 1),
        "divV": (nx, ny, nz), "Rp": (nx - 2, ny - 2, nz - 2), "absRp": (nx - 2, ny - 2, nz - 2),
    }


def alloc_fields(p: Params) -> dict:
    return {k: zeros(*s) for k, s in shapes(p.nx, p.ny, p.nz).items()}


def initial_fields(p: Params) -> dict:
    """Initial state: M:369-372 or G:85-88 (no halo update: single rank / caller's business)."""
    f = alloc_fields(p)
    nx, ny, nz = p.nx, p.ny, p.nz
    yc = linrange(-(p.ly - p.dy) / 2, (p.ly - p.dy) / 2, ny)
    zc = linrange(-(p.lz - p.dz) / 2, (p.lz - p.dz) / 2, nz)
    if p.variant == "M":
        f["Vy"][0, :, :] = p.vin                                         # M:369 (sic: Vy)
        zg = np.array([x_g(iz, p.dz, nz, nz, p.coords[2]) for iz in range(1, nz + 1)])
        pr = (-(zg - p.dz / 2) * p.rho * p.g)[None, None, :] + (0 * yc)[None, :, None] + (0 * zc)[None, None, :]
        f["Pr"][...] = np.broadcast_to(pr, (nx, ny, nz))                 # M:370
        set_cylinder(p, f)                                               # M:372
    else:
        xc = linrange(-(p.lx - p.dx) / 2, (p.lx - p.dx) / 2, nx)
        xv = linrange(-p.lx / 2, p.lx / 2, nx + 1)
        prof = p.vin * (7.0 / 6.0) * np.power((zc + p.lz / 2) / p.lz, 1.0 / 6.0)
        vx = prof[None, None, :] + (0 * yc)[None, :, None] + (0 * xv)[:, None, None]
        f["Vx"][...] = vx                                                # G:86
        pr = (-(zc - p.lz / 2) * p.rho * p.g)[None, None, :] + (0 * yc)[None, :, None] + (0 * xc)[:, None, None]
        f["Pr"][...] = pr                                                # G:87
    return f


# ----------------------------------------------------------------------------------------------
# kernel wrappers (one per reference kernel)
# ----------------------------------------------------------------------------------------------
def update_tau(p: Params, f: dict):
    lib().oracle_update_tau(_p(f["txx"]), _p(f["tyy"]), _p(f["tzz"]), _p(f["txy"]), _p(f["txz"]), _p(f["tyz"]),
                            _p(f["Vx"]), _p(f["Vy"]), _p(f["Vz"]), _d(p.mu), _d(p.dx), _d(p.dy), _d(p.dz),
                            p.nx, p.ny, p.nz)


def predict_V(p: Params, f: dict):
    lib().oracle_predict_V(_p(f["Vx"]), _p(f["Vy"]), _p(f["Vz"]), _p(f["txx"]), _p(f["tyy"]), _p(f["tzz"]),
                           _p(f["txy"]), _p(f["txz"]), _p(f["tyz"]), _d(p.rho), _d(p.g), _d(p.dt), _d(p.dx),
                           _d(p.dy), _d(p.dz), p.nx, p.ny, p.nz)


def update_divV(p: Params, f: dict):
    lib().oracle_update_divV(_p(f["divV"]), _p(f["Vx"]), _p(f["Vy"]), _p(f["Vz"]), _d(p.dx), _d(p.dy), _d(p.dz),
                             p.nx, p.ny, p.nz)


def update_dPrdtau(p: Params, f: dict):
    lib().oracle_update_dPrdtau(_p(f["Pr"]), _p(f["dPrdtau"]), _p(f["divV"]), _d(p.rho), _d(p.dt), _d(p.dtau),
                                _d(p.damp), _d(p.dx), _d(p.dy), _d(p.dz), p.nx, p.ny, p.nz)


def update_Pr(p: Params, f: dict):
    lib().oracle_update_Pr(_p(f["Pr"]), _p(f["dPrdtau"]), _d(p.dtau), p.nx, p.ny, p.nz)


def compute_res(p: Params, f: dict):
    lib().oracle_compute_res(_p(f["Rp"]), _p(f["Pr"]), _p(f["divV"]), _d(p.rho), _d(p.dt), _d(p.dx), _d(p.dy),
                             _d(p.dz), p.nx, p.ny, p.nz)


def max_abs(a: np.ndarray) -> float:
    return float(lib().oracle_max_abs(_p(a), None, C.c_size_t(a.size)))


def correct_V(p: Params, f: dict):
    lib().oracle_correct_V(_p(f["Vx"]), _p(f["Vy"]), _p(f["Vz"]), _p(f["Pr"]), _d(p.dt), _d(p.rho), _d(p.dx),
                           _d(p.dy), _d(p.dz), p.nx, p.ny, p.nz)


def bc(name: str, a: np.ndarray, *scalars):
    """bc('x'|'y'|'z'|'zV', A) ; bc('x_Vx', A, V) ; bc('x_Pr', A, val) ; bc('xhydstatic', A, dz, nz, g, rho)."""
    sx, sy, sz = a.shape
    L = lib()
    if name in ("x", "y", "z", "zV"):
        getattr(L, "oracle_bc_" + name)(_p(a), sx, sy, sz)
    elif name in ("x_Vx", "x_Pr"):
        getattr(L, "oracle_bc_" + name)(_p(a), _d(scalars[0]), sx, sy, sz)
    elif name == "xhydstatic":
        dz, nz, g, rho = scalars
        L.oracle_bc_xhydstatic(_p(a), _d(dz), int(nz), _d(g), _d(rho), sx, sy, sz)
    else:
        raise ValueError(name)


def set_bc_Vel(p: Params, f: dict):
    if p.variant == "M":
        lib().oracle_set_bc_Vel_M(_p(f["Vx"]), _p(f["Vy"]), _p(f["Vz"]), int(p.inlet_guard), _d(p.vin),
                                  p.nx, p.ny, p.nz)
    else:
        lib().oracle_set_bc_Vel_G(_p(f["Vx"]), _p(f["Vy"]), _p(f["Vz"]), p.nx, p.ny, p.nz)


def set_bc_Pr(p: Params, f: dict):
    if p.variant == "M":
        lib().oracle_set_bc_Pr_M(_p(f["Pr"]), int(p.outlet_guard), _d(0.0), p.nx, p.ny, p.nz)
    else:
        lib().oracle_set_bc_Pr_G(_p(f["Pr"]), _d(p.dz), p.nz, _d(p.g), _d(p.rho), p.nx, p.ny, p.nz)


def advect(p: Params, f: dict):
    lib().oracle_advect(_p(f["Vx"]), _p(f["Vx_o"]), _p(f["Vy"]), _p(f["Vy_o"]), _p(f["Vz"]), _p(f["Vz_o"]),
                        _p(f["C"]), _p(f["C_o"]), _d(p.dt), _d(p.dx), _d(p.dy), _d(p.dz), p.nx, p.ny, p.nz)


def set_cylinder(p: Params, f: dict):
    if p.variant == "M":
        lib().oracle_set_cylinder_M(_p(f["C"]), _p(f["Vx"]), _p(f["Vy"]), _p(f["Vz"]), _d(p.a2), _d(p.b2),
                                    _d(p.ox), _d(p.oy), _d(p.sinb), _d(p.cosb), _d(p.xco_g), _d(p.yco_g),
                                    _d(p.dx), _d(p.dy), p.nx, p.ny, p.nz)
    else:
        lib().oracle_set_cylinder_G(_p(f["C"]), _p(f["Vx"]), _p(f["Vy"]), _p(f["Vz"]), _d(p.a2), _d(p.b2),
                                    _d(p.ox), _d(p.oy), _d(p.sinb), _d(p.cosb), _d(p.lx), _d(p.ly),
                                    _d(p.dx), _d(p.dy), p.nx, p.ny, p.nz)


def _cfields(f: dict) -> _Fields:
    s = _Fields()
    for n in _FIELD_NAMES:
        setattr(s, n, _p(f[n]))
    return s


def pt_solve(p: Params, f: dict):
    """PT loop M:458-471 / G:126-137 -> (iterations, [err at each check])."""
    cap = p.niter // p.nchk + 2
    hist = (C.c_double * cap)()
    nchecks = C.c_int(0)
    ps, fs = p.cstruct(), _cfields(f)
    iters = lib().oracle_pt_solve(C.byref(ps), C.byref(fs), hist, cap, C.byref(nchecks))
    return iters, [hist[i] for i in range(nchecks.value)]


def step(p: Params, f: dict):
    """One time step M:449-477 (one rank) / G:121-142 -> (iterations, [err at each check])."""
    cap = p.niter // p.nchk + 2
    hist = (C.c_double * cap)()
    nchecks = C.c_int(0)
    ps, fs = p.cstruct(), _cfields(f)
    iters = lib().oracle_step(C.byref(ps), C.byref(fs), hist, cap, C.byref(nchecks))
    return iters, [hist[i] for i in range(nchecks.value)]


def run(p: Params, nt: int, f: dict | None = None):
    """nt time steps from the script's initial state -> (fields, iters per step, err history per step)."""
    if f is None:
        f = initial_fields(p)
    iters, errs = [], []
    for _ in range(nt):
        it, hist = step(p, f)
        iters.append(it)
        errs.append(hist)
    return f, iters, errs


def interior(a: np.ndarray) -> np.ndarray:
    """``A[2:end-1,2:end-1,2:end-1]`` -- what run_navierstokes3D returns (M:528-535)."""
    return a[1:-1, 1:-1, 1:-1]


# ----------------------------------------------------------------------------------------------
# ImplicitGlobalGrid emulation: N virtual ranks in one process (multi-rank truth)
# ----------------------------------------------------------------------------------------------
class VirtualRanks:
    """Runs script M on a dims = (dx,dy,dz) Cartesian process grid, all ranks in this process.

    Restates what the reference does around its kernels when more than one MPI rank runs
    (IGG source is not in the reference tree; semantics from SURVEY.md section 5 / 3.4):
    every rank holds local arrays of the script's shapes (overlap 2), ``update_halo!`` is applied
    at exactly the script's call sites (M:371,373,450,453,455,460,462,182,167,477), dimension
    by dimension x -> y -> z, sending plane ``ol`` / ``s-ol+1`` (1-based, ol = 2 + s - n) to the
    lower / upper neighbour's last / first plane; the residual is max-reduced over ranks
    (``max_g``, M:21).  ``damp`` uses the local nx (M:340), coordinates come from ``x_g`` per rank.
    """

    def __init__(self, nx, ny, nz, dims, **kw):
        self.dims = tuple(dims)
        self.coords = [(cx, cy, cz) for cx in range(dims[0]) for cy in range(dims[1]) for cz in range(dims[2])]
        self.p = [params_M(nx, ny, nz, dims=dims, coords=c, **kw) for c in self.coords]
        self.f = []
        for p in self.p:                              # M:343-373
            f = alloc_fields(p)
            init = initial_fields(p)                  # Vy[1,:,:] .= vin on EVERY rank (M:369, sic), Pr, cylinder
            for k in f:
                f[k][...] = init[k]
            self.f.append(f)
        # initial_fields applied set_cylinder! after the Pr assignment; the halo updates of
        # M:371 and M:373 follow (update_halo!(Pr) commutes with set_cylinder!, which does not touch Pr)
        self.update_halo("Pr")
        self.update_halo("C", "Vx", "Vy", "Vz")
        self.iters, self.errs = [], []

    def rank_of(self, c):
        return self.coords.index(tuple(c))

    def update_halo(self, *names):
        n_loc = (self.p[0].nx, self.p[0].ny, self.p[0].nz)
        for name in names:
            for d in range(3):
                if self.dims[d] == 1:
                    continue
                s = self.f[0][name].shape[d]
                ol = 2 + (s - n_loc[d])
                assert ol >= 2, f"{name} cannot be exchanged along dim {d}"
                sends = {}
                for r, c in enumerate(self.coords):   # pack every send buffer before any unpack
                    a = self.f[r][name]
                    sends[r] = (np.take(a, ol - 1, axis=d).copy(), np.take(a, s - ol, axis=d).copy())
                for r, c in enumerate(self.coords):
                    a = self.f[r][name]
                    idx = [slice(None)] * 3
                    if c[d] > 0:                      # from the lower neighbour: its plane s-ol+1 -> my plane 1
                        lo = list(c); lo[d] -= 1
                        idx[d] = 0
                        a[tuple(idx)] = sends[self.rank_of(lo)][1]
                    if c[d] < self.dims[d] - 1:       # from the upper neighbour: its plane ol -> my plane s
                        hi = list(c); hi[d] += 1
                        idx[d] = s - 1
                        a[tuple(idx)] = sends[self.rank_of(hi)][0]

    def each(self, fn):
        for p, f in zip(self.p, self.f):
            fn(p, f)

    def step(self):
        p0 = self.p[0]
        self.each(update_tau)                                         # M:449
        self.update_halo("txx", "tyy", "tzz")                         # M:450
        self.each(predict_V)                                          # M:451
        self.each(set_cylinder)                                       # M:452
        self.update_halo("C", "Vx", "Vy", "Vz")                       # M:453
        self.each(update_divV)                                        # M:454
        self.update_halo("divV")                                      # M:455
        iters, hist = 0, []
        for it in range(1, p0.niter + 1):                             # M:458
            self.each(update_dPrdtau)                                 # M:459
            self.update_halo("divV")                                  # M:460
            self.each(update_Pr)                                      # M:461
            self.update_halo("Pr")                                    # M:462
            self.each(set_bc_Pr)                                      # M:463 ...
            self.update_halo("Pr")                                    # ... M:182
            iters = it
            if it % p0.nchk == 0:                                     # M:464
                self.each(compute_res)
                m = max((max_abs(f["Rp"]) for f in self.f), key=lambda v: (math.isnan(v), v))   # max_g, NaN wins
                err = m * (p0.ly * p0.ly) / p0.psc
                hist.append(err)
                if err < p0.eps_it or not math.isfinite(err):
                    break
        self.each(correct_V)                                          # M:472
        self.each(set_cylinder)                                       # M:473
        self.each(set_bc_Vel)                                         # M:474 ...
        self.update_halo("Vx", "Vy", "Vz")                            # ... M:167
        for f in self.f:                                              # M:475
            for a in ("Vx", "Vy", "Vz", "C"):
                f[a + "_o"][...] = f[a]
        self.each(advect)                                             # M:476
        self.update_halo("Vx", "Vy", "Vz")                            # M:477
        self.iters.append(iters)
        self.errs.append(hist)
        return iters, hist

    def assemble(self, name):
        """Global array from the ranks' OWNED planes (local 2..n-1, plus the physical faces).

        Halo planes are skipped: right after advect! (M:476-477) only Vx,Vy,Vz are exchanged, so
        C's halo planes hold locally clamped values until the next step's M:453."""
        n_loc = (self.p[0].nx, self.p[0].ny, self.p[0].nz)
        shp = self.f[0][name].shape
        gshape = tuple(self.dims[d] * (n_loc[d] - 2) + 2 + (shp[d] - n_loc[d]) for d in range(3))
        out = np.full(gshape, np.nan, order="F")
        for c, f in zip(self.coords, self.f):
            src, dst = [], []
            for d in range(3):
                lo = 0 if c[d] == 0 else 1
                hi = shp[d] if c[d] == self.dims[d] - 1 else n_loc[d] - 1
                src.append(slice(lo, hi))
                dst.append(slice(c[d] * (n_loc[d] - 2) + lo, c[d] * (n_loc[d] - 2) + hi))
            out[tuple(dst)] = f[name][tuple(src)]
        assert not np.isnan(out).any() or np.isnan(f[name]).any()
        return out
