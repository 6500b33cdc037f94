"""Second, independently written restatement of the reference kernels -- vectorised numpy.

TEST INFRASTRUCTURE ONLY (see oracle/ns3d_oracle.c).  Written from the Julia sources
(scripts/NavierStokes3D_multi_gpu.jl = M, scripts/NavierStokes3D_gpu.jl = G) with whole-array
slices instead of the C oracle's index loops, so that an indexing or association mistake in
one of the two shows up as a bit difference between them (tests/test_oracle.py).  numpy
evaluates element-wise IEEE double operations without contraction, so agreement is bit-exact.

Slice dictionary for an array A of shape (sx,sy,sz), ParallelStencil.FiniteDifferences3D:
    @all(A)   = A                      @inn(A)   = A[1:-1,1:-1,1:-1]
    @d_xa(A)  = A[1:,:,:]-A[:-1,:,:]   @d_xi(A)  = A[1:,1:-1,1:-1]-A[:-1,1:-1,1:-1]
    @d2_xi(A) = (A[2:,1:-1,1:-1]-A[1:-1,1:-1,1:-1]) - (A[1:-1,1:-1,1:-1]-A[:-2,1:-1,1:-1])
"""
from __future__ import annotations

import numpy as np


def d_xa(A): return A[1:, :, :] - A[:-1, :, :]
def d_ya(A): return A[:, 1:, :] - A[:, :-1, :]
def d_za(A): return A[:, :, 1:] - A[:, :, :-1]
# The "inner" macros only shift by +1; the far end is cut by the LHS guard, see crop().
def d_xi(A): return A[1:, 1:, 1:] - A[:-1, 1:, 1:]
def d_yi(A): return A[1:, 1:, 1:] - A[1:, :-1, 1:]
def d_zi(A): return A[1:, 1:, 1:] - A[1:, 1:, :-1]
def inn(A): return A[1:-1, 1:-1, 1:-1]
def d2_xi(A): return (A[2:, 1:, 1:] - A[1:-1, 1:, 1:]) - (A[1:-1, 1:, 1:] - A[:-2, 1:, 1:])
def d2_yi(A): return (A[1:, 2:, 1:] - A[1:, 1:-1, 1:]) - (A[1:, 1:-1, 1:] - A[1:, :-2, 1:])
def d2_zi(A): return (A[1:, 1:, 2:] - A[1:, 1:, 1:-1]) - (A[1:, 1:, 1:-1] - A[1:, 1:, :-2])


def crop(E, shape):
    """The bounds guard ParallelStencil derives from the LHS: keep threads ix<=shape[0], ..."""
    assert all(e >= s for e, s in zip(E.shape, shape)), (E.shape, shape)
    return E[:shape[0], :shape[1], :shape[2]]


def divV_expr(p, f):   # macro @∇V  M:15
    return d_xa(f["Vx"]) / p.dx + d_ya(f["Vy"]) / p.dy + d_za(f["Vz"]) / p.dz


def update_tau(p, f):   # M:36-44
    Vx, Vy, Vz, mu = f["Vx"], f["Vy"], f["Vz"], p.mu
    dv = divV_expr(p, f)
    f["txx"][...] = 2 * mu * (d_xa(Vx) / p.dx - dv / 3.0)
    f["tyy"][...] = 2 * mu * (d_ya(Vy) / p.dy - dv / 3.0)
    f["tzz"][...] = 2 * mu * (d_za(Vz) / p.dz - dv / 3.0)
    # @all(τxy) is guarded by size(τxy) = (nx-1,ny-1,nz-1): the leading part of the @d_*i arrays
    e = f["txy"].shape
    f["txy"][...] = mu * (crop(d_yi(Vx), e) / p.dy + crop(d_xi(Vy), e) / p.dx)
    f["txz"][...] = mu * (crop(d_zi(Vx), e) / p.dz + crop(d_xi(Vz), e) / p.dx)
    f["tyz"][...] = mu * (crop(d_zi(Vy), e) / p.dz + crop(d_yi(Vz), e) / p.dy)


def predict_V(p, f):   # M:50-55
    nx, ny, nz = p.nx, p.ny, p.nz
    dtr = p.dt / p.rho
    txx, tyy, tzz, txy, txz, tyz = (f[k] for k in ("txx", "tyy", "tzz", "txy", "txz", "tyz"))
    # @inn(Vx): guard ix<=nx-1, iy<=ny-2, iz<=nz-2
    g = (nx - 1, ny - 2, nz - 2)
    rhs = crop(d_xi(txx), g) / p.dx + crop(d_ya(txy), g) / p.dy + crop(d_za(txz), g) / p.dz
    inn(f["Vx"])[...] = inn(f["Vx"]) + dtr * rhs
    g = (nx - 2, ny - 1, nz - 2)
    rhs = crop(d_yi(tyy), g) / p.dy + crop(d_xa(txy), g) / p.dx + crop(d_za(tyz), g) / p.dz
    inn(f["Vy"])[...] = inn(f["Vy"]) + dtr * rhs
    g = (nx - 2, ny - 2, nz - 1)
    rhs = crop(d_zi(tzz), g) / p.dz + crop(d_xa(txz), g) / p.dx + crop(d_ya(tyz), g) / p.dy - p.rho * p.g
    inn(f["Vz"])[...] = inn(f["Vz"]) + dtr * rhs


def update_divV(p, f):   # M:61-64
    f["divV"][...] = divV_expr(p, f)


def bracket(p, f):   # M:71 / M:89
    Pr, g = f["Pr"], f["dPrdtau"].shape
    return (crop(d2_xi(Pr), g) / p.dx / p.dx + crop(d2_yi(Pr), g) / p.dy / p.dy + crop(d2_zi(Pr), g) / p.dz / p.dz
            - p.rho / p.dt * inn(f["divV"]))


def update_dPrdtau(p, f):   # M:70-73
    f["dPrdtau"][...] = f["dPrdtau"] * (1.0 - p.damp) + p.dtau * bracket(p, f)


def update_Pr(p, f):   # M:79-82
    inn(f["Pr"])[...] = inn(f["Pr"]) + p.dtau * f["dPrdtau"]


def compute_res(p, f):   # M:88-91
    f["Rp"][...] = bracket(p, f)


def correct_V(p, f):   # M:97-102
    dtr, Pr = p.dt / p.rho, f["Pr"]
    for name, d, h in (("Vx", d_xi, p.dx), ("Vy", d_yi, p.dy), ("Vz", d_zi, p.dz)):
        V = inn(f[name])
        V[...] = V - dtr * crop(d(Pr), V.shape) / h


def bc_x(A): A[0, :, :] = A[1, :, :]; A[-1, :, :] = A[-2, :, :]      # M:108-112
def bc_y(A): A[:, 0, :] = A[:, 1, :]; A[:, -1, :] = A[:, -2, :]      # M:118-122
def bc_z(A): A[:, :, 0] = A[:, :, 1]; A[:, :, -1] = A[:, :, -2]      # M:128-132
def bc_zV(A): A[:, :, 0] = 0.0; A[:, :, -1] = A[:, :, -2]            # G:239-243


def set_bc_Pr(p, f):
    Pr = f["Pr"]
    if p.variant == "M":   # M:175-184
        bc_x(Pr); bc_y(Pr); bc_z(Pr)
        if p.outlet_guard:
            Pr[-1, :, :] = 0.0
    else:                  # G:281-286, bc_xhydstatic! G:257-261
        bc_y(Pr); bc_z(Pr)
        iz = np.arange(1, p.nz + 1)
        h = p.rho * p.g * ((p.nz - iz) + 0.5) * p.dz
        Pr[0, :, :] = (h + 100)[None, :]
        Pr[-1, :, :] = h[None, :]


def set_bc_Vel(p, f):
    Vx, Vy, Vz = f["Vx"], f["Vy"], f["Vz"]
    if p.variant == "M":   # M:156-169
        bc_x(Vx); bc_y(Vx); bc_z(Vx); bc_x(Vy); bc_z(Vy); bc_x(Vz); bc_y(Vz)
        if p.inlet_guard:
            Vx[0, :, :] = p.vin
    else:                  # G:264-279
        for A in (Vx, Vy, Vz):
            bc_x(A); bc_y(A); bc_zV(A)


def set_cylinder(p, f):   # M:249-281 / G:336-368
    nx, ny = p.nx, p.ny
    ix = np.arange(1, nx + 2)[:, None]
    iy = np.arange(1, ny + 2)[None, :]
    if p.variant == "M":
        xc = p.xco_g + (ix - 1) * p.dx
        yc = p.yco_g + (iy - 1) * p.dy
        xv, yv = xc - p.dx / 2, yc - p.dy / 2
    else:
        xv = (ix - 1) * p.dx - p.lx / 2
        yv = (iy - 1) * p.dy - p.ly / 2
        xc, yc = xv + p.dx / 2, yv + p.dx / 2

    def inside(X, Y, thr):
        xr = (X - p.ox) * p.cosb - (Y - p.oy) * p.sinb
        yr = (X - p.ox) * p.sinb + (Y - p.oy) * p.cosb
        return xr * xr / p.a2 + yr * yr / p.b2 < thr

    f["C"][inside(xc, yc, 1.05)[:nx, :ny]] = 1.0
    f["Vx"][inside(xv, yc, 1.0)[:nx + 1, :ny]] = 0.0
    f["Vy"][inside(xc, yv, 1.0)[:nx, :ny + 1]] = 0.0
    f["Vz"][inside(xc, yc, 1.0)[:nx, :ny]] = 0.0


def backtrack(p, A_o, vxc, vyc, vzc, I, J, K):
    """backtrack! M:190-205 for all target points at once (I,J,K: 1-based index grids)."""
    sx, sy, sz = A_o.shape
    dlx, dly, dlz = p.dt * vxc / p.dx, p.dt * vyc / p.dy, p.dt * vzc / p.dz
    i1 = np.clip(np.floor(I - dlx).astype(np.int64), 1, sx)
    j1 = np.clip(np.floor(J - dly).astype(np.int64), 1, sy)
    k1 = np.clip(np.floor(K - dlz).astype(np.int64), 1, sz)
    i2, j2, k2 = np.clip(i1 + 1, 1, sx), np.clip(j1 + 1, 1, sy), np.clip(k1 + 1, 1, sz)
    tx = (dlx > 0).astype(np.float64) - np.fmod(dlx, 1.0)
    ty = (dly > 0).astype(np.float64) - np.fmod(dly, 1.0)
    tz = (dlz > 0).astype(np.float64) - np.fmod(dlz, 1.0)

    def lerp(a, b, t): return b * t + a * (1 - t)       # M:211
    def at(i, j, k): return A_o[i - 1, j - 1, k - 1]
    f11 = lerp(at(i1, j1, k1), at(i2, j1, k1), tx)
    f12 = lerp(at(i1, j1, k2), at(i2, j1, k2), tx)
    f21 = lerp(at(i1, j2, k1), at(i2, j2, k1), tx)
    f22 = lerp(at(i1, j2, k2), at(i2, j2, k2), tx)
    return lerp(lerp(f11, f21, ty), lerp(f12, f22, ty), tz)


def advect(p, f):   # M:217-243
    nx, ny, nz = p.nx, p.ny, p.nz
    Vxo, Vyo, Vzo, Co = f["Vx_o"], f["Vy_o"], f["Vz_o"], f["C_o"]

    def grid(ix, iy, iz):
        return np.meshgrid(ix, iy, iz, indexing="ij")

    # branch 1: Vx at ix in 2..nx, all iy, iz
    I, J, K = grid(np.arange(2, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1))
    vx = Vxo[1:nx, :, :]
    vy = 0.25 * (Vyo[0:nx - 1, 0:ny, :] + Vyo[0:nx - 1, 1:ny + 1, :] + Vyo[1:nx, 0:ny, :] + Vyo[1:nx, 1:ny + 1, :])
    vz = 0.25 * (Vzo[0:nx - 1, :, 0:nz] + Vzo[0:nx - 1, :, 1:nz + 1] + Vzo[1:nx, :, 0:nz] + Vzo[1:nx, :, 1:nz + 1])
    f["Vx"][1:nx, :, :] = backtrack(p, Vxo, vx, vy, vz, I, J, K)
    # branch 2: Vy at iy in 2..ny, all ix, iz
    I, J, K = grid(np.arange(1, nx + 1), np.arange(2, ny + 1), np.arange(1, nz + 1))
    vx = 0.25 * (Vxo[0:nx, 0:ny - 1, :] + Vxo[1:nx + 1, 0:ny - 1, :] + Vxo[0:nx, 1:ny, :] + Vxo[1:nx + 1, 1:ny, :])
    vy = Vyo[:, 1:ny, :]
    vz = 0.25 * (Vzo[:, 0:ny - 1, 0:nz] + Vzo[:, 0:ny - 1, 1:nz + 1] + Vzo[:, 1:ny, 0:nz] + Vzo[:, 1:ny, 1:nz + 1])
    f["Vy"][:, 1:ny, :] = backtrack(p, Vyo, vx, vy, vz, I, J, K)
    # branch 3 (sic: targets Vy, M:234): iz in 2..nz, ix in 1..nx, iy in 1..ny -- overwrites branch 2
    I, J, K = grid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(2, nz + 1))
    vx = 0.25 * (Vxo[0:nx, :, 0:nz - 1] + Vxo[1:nx + 1, :, 0:nz - 1] + Vxo[0:nx, :, 1:nz] + Vxo[1:nx + 1, :, 1:nz])
    vy = 0.25 * (Vyo[:, 0:ny, 0:nz - 1] + Vyo[:, 1:ny + 1, 0:nz - 1] + Vyo[:, 0:ny, 1:nz] + Vyo[:, 1:ny + 1, 1:nz])
    vz = Vzo[:, :, 1:nz]
    f["Vy"][:, 0:ny, 1:nz] = backtrack(p, Vyo, vx, vy, vz, I, J, K)
    # branch 4: C everywhere
    I, J, K = grid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1))
    vx = 0.5 * (Vxo[0:nx, :, :] + Vxo[1:nx + 1, :, :])
    vy = 0.5 * (Vyo[:, 0:ny, :] + Vyo[:, 1:ny + 1, :])
    vz = 0.5 * (Vzo[:, :, 0:nz] + Vzo[:, :, 1:nz + 1])
    f["C"][...] = backtrack(p, Co, vx, vy, vz, I, J, K)


def step(p, f):
    """One time step M:449-477 / G:121-142 -> (iterations, err history)."""
    update_tau(p, f); predict_V(p, f); set_cylinder(p, f); update_divV(p, f)
    iters, hist = 0, []
    for it in range(1, p.niter + 1):
        update_dPrdtau(p, f); update_Pr(p, f); set_bc_Pr(p, f)
        iters = it
        if it % p.nchk == 0:
            compute_res(p, f)
            err = np.abs(f["Rp"]).max() * (p.ly * p.ly) / p.psc
            hist.append(float(err))
            if err < p.eps_it or not np.isfinite(err):
                break
    correct_V(p, f); set_cylinder(p, f); set_bc_Vel(p, f)
    for a in ("Vx", "Vy", "Vz", "C"):
        f[a + "_o"][...] = f[a]
    advect(p, f)
    return iters, hist
