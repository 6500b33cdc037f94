// libns3d.so -- level 2: the fused pseudo-transient (PT) pressure loop and the once-per-step groups.
//
// Reference (per PT iteration, M:459-463 / G:127-129): update_dPrdτ! (K5), update_Pr! (K6), set_bc_Pr! = 3-4 face
// kernels (K7) and up to three update_halo! calls: >= 5 synchronous launches and 7+ full-field passes.  The fused loop
// itself lives in ns3d_ptv.cu (host side: pitched working copies, TMA descriptors, CUDA-graph replay, the peer-memory
// halo exchange on z-slabs) and ns3d_ptv_kernels.cuh (device side: ptv_kernel, K iterations per launch); this file
// holds the entry points ns3d_pt_solve / ns3d_pt_iterate / ns3d_pt_describe, the three once-per-step groups
// (predictor, corrector, advection) and ns3d_step.
//
// Arithmetic: see NS3D_PARITY / NS3D_FAST / NS3D_FASTEST in ns3d.h.  The library is compiled with --fmad=false; FMA
// appears only where fma() is written explicitly.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ns3d_internal.cuh"

namespace {
int check_pt_params(ns3d_ctx* ctx, const ns3d_pt_params* p)
{
    if (!p) return ns3d_fail(ctx, NS3D_EINVAL, "pt: params is NULL");
    if (p->nx < 3 || p->ny < 3 || p->nz < 3) return ns3d_fail(ctx, NS3D_EINVAL, "pt: grid must be at least 3^3");
    if (p->variant != NS3D_VARIANT_M && p->variant != NS3D_VARIANT_G)
        return ns3d_fail(ctx, NS3D_EINVAL, "pt: unknown variant %d", p->variant);
    return NS3D_OK;
}
}  // namespace

extern "C" int ns3d_pt_solve(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV,
                             const ns3d_pt_params* p, int* h_iters, double* h_err_hist, int err_cap,
                             int* h_nchecks)
{
    NS3D_CHECK_CTX(ctx);
    if (!Pr || !dPrdtau || !divV) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_pt_solve: NULL field");
    NS3D_TRY(check_pt_params(ctx, p));
    if (p->nchk <= 0 || p->niter < 0) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_pt_solve: bad niter/nchk");
    return ns3d_internal_ptv_solve(ctx, Pr, dPrdtau, divV, p, h_iters, h_err_hist, err_cap, h_nchecks);
}

extern "C" int ns3d_pt_iterate(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV,
                               const ns3d_pt_params* p, int n_iter)
{
    NS3D_CHECK_CTX(ctx);
    if (!Pr || !dPrdtau || !divV || n_iter < 0) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_pt_iterate: bad argument");
    NS3D_TRY(check_pt_params(ctx, p));
    return ns3d_internal_ptv_iterate(ctx, Pr, dPrdtau, divV, p, n_iter);
}

extern "C" int ns3d_pt_describe(ns3d_ctx* ctx, const ns3d_pt_params* p, char* buf, int cap, int* iters_per_launch)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_TRY(check_pt_params(ctx, p));
    return ns3d_internal_ptv_describe(ctx, p, buf, cap, iters_per_launch);
}

// ---- level 2: the three once-per-step groups around the PT loop, and the whole step ---------------
namespace {
int step_cylinder(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* sp)
{
    const ns3d_pt_params& p = sp->pt;
    return p.variant == NS3D_VARIANT_M
               ? ns3d_set_cylinder_M(ctx, f->C, f->Vx, f->Vy, f->Vz, sp->a2, sp->b2, sp->ox, sp->oy, sp->sinb, sp->cosb,
                                     sp->xco_g, sp->yco_g, p.dx, p.dy, p.nx, p.ny, p.nz)
               : ns3d_set_cylinder_G(ctx, f->C, f->Vx, f->Vy, f->Vz, sp->a2, sp->b2, sp->ox, sp->oy, sp->sinb, sp->cosb,
                                     sp->lx, sp->ly, p.dx, p.dy, p.nx, p.ny, p.nz);
}
}  // namespace

// Chorin predictor, M:449-455 / G:121-124: update_τ!, predict_V!, set_cylinder!, update_∇V! and the
// halo updates between them (update_halo!(τxx,τyy,τzz) M:450 is redundant: τ is computed on the
// halo cells too).
extern "C" int ns3d_predictor(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* sp)
{
    NS3D_CHECK_CTX(ctx);
    if (!f || !sp) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_predictor: NULL argument");
    const ns3d_pt_params& p = sp->pt;
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    NS3D_TRY(ns3d_update_tau(ctx, f->txx, f->tyy, f->tzz, f->txy, f->txz, f->tyz, f->Vx, f->Vy, f->Vz, sp->mu, p.dx, p.dy,
                             p.dz, nx, ny, nz));                                                       // M:449
    NS3D_TRY(ns3d_predict_V(ctx, f->Vx, f->Vy, f->Vz, f->txx, f->tyy, f->tzz, f->txy, f->txz, f->tyz, p.rho, p.g, p.dt,
                            p.dx, p.dy, p.dz, nx, ny, nz));                                            // M:451
    NS3D_TRY(step_cylinder(ctx, f, sp));                                                               // M:452
    if (ctx->nranks > 1) {                                                                             // M:453
        double* h[4] = {f->C, f->Vx, f->Vy, f->Vz};
        const int sx[4] = {nx, nx + 1, nx, nx}, sy[4] = {ny, ny, ny + 1, ny}, sz[4] = {nz, nz, nz, nz + 1};
        NS3D_TRY(ns3d_internal_halo_z(ctx, ctx->stream, h, sx, sy, sz, 4, nz));
    }
    NS3D_TRY(ns3d_update_divV(ctx, f->divV, f->Vx, f->Vy, f->Vz, p.dx, p.dy, p.dz, nx, ny, nz));      // M:454
    if (ctx->nranks > 1) {                                                                             // M:455
        double* h[1] = {f->divV};
        NS3D_TRY(ns3d_internal_halo_z(ctx, ctx->stream, h, &nx, &ny, &nz, 1, nz));
    }
    return NS3D_OK;
}

// Pressure-gradient correction, M:472-474 / G:138-140: correct_V!, set_cylinder!, set_bc_Vel!.
extern "C" int ns3d_corrector(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* sp)
{
    NS3D_CHECK_CTX(ctx);
    if (!f || !sp) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_corrector: NULL argument");
    const ns3d_pt_params& p = sp->pt;
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    NS3D_TRY(ns3d_correct_V(ctx, f->Vx, f->Vy, f->Vz, f->Pr, p.dt, p.rho, p.dx, p.dy, p.dz, nx, ny, nz));   // M:472
    NS3D_TRY(step_cylinder(ctx, f, sp));                                                               // M:473
    if (p.variant == NS3D_VARIANT_M)
        return ns3d_set_bc_Vel_M(ctx, f->Vx, f->Vy, f->Vz, sp->inlet_guard, sp->vin, nx, ny, nz);      // M:474
    return ns3d_set_bc_Vel_G(ctx, f->Vx, f->Vy, f->Vz, nx, ny, nz);                                    // G:140
}

// Advection, M:475-477 / G:141-142: the four snapshots `A_o .= A`, advect! and update_halo!(Vx,Vy,Vz).
extern "C" int ns3d_advect_swap(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* sp)
{
    NS3D_CHECK_CTX(ctx);
    if (!f || !sp) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_advect_swap: NULL argument");
    const ns3d_pt_params& p = sp->pt;
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    NS3D_TRY(ns3d_copy(ctx, f->Vx_o, f->Vx, (size_t)(nx + 1) * ny * nz));                              // M:475
    NS3D_TRY(ns3d_copy(ctx, f->Vy_o, f->Vy, (size_t)nx * (ny + 1) * nz));
    NS3D_TRY(ns3d_copy(ctx, f->Vz_o, f->Vz, (size_t)nx * ny * (nz + 1)));
    NS3D_TRY(ns3d_copy(ctx, f->C_o, f->C, (size_t)nx * ny * nz));
    NS3D_TRY(ns3d_advect(ctx, f->Vx, f->Vx_o, f->Vy, f->Vy_o, f->Vz, f->Vz_o, f->C, f->C_o, p.dt, p.dx, p.dy, p.dz, nx,
                         ny, nz));                                                                     // M:476
    if (ctx->nranks > 1) {                                                                             // M:477
        double* h[3] = {f->Vx, f->Vy, f->Vz};
        const int sx[3] = {nx + 1, nx, nx}, sy[3] = {ny, ny + 1, ny}, sz[3] = {nz, nz, nz + 1};
        NS3D_TRY(ns3d_internal_halo_z(ctx, ctx->stream, h, sx, sy, sz, 3, nz));
    }
    return NS3D_OK;
}

// One time step, M:449-477 / G:121-142 = predictor, PT loop, corrector, advection.
extern "C" int ns3d_step(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* sp, int* h_iters,
                         double* h_err_hist, int err_cap, int* h_nchecks)
{
    NS3D_CHECK_CTX(ctx);
    if (!f || !sp) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_step: NULL argument");
    NS3D_TRY(ns3d_predictor(ctx, f, sp));                                                              // M:449-455
    NS3D_TRY(ns3d_pt_solve(ctx, f->Pr, f->dPrdtau, f->divV, &sp->pt, h_iters, h_err_hist, err_cap, h_nchecks));  // M:458-471
    NS3D_TRY(ns3d_corrector(ctx, f, sp));                                                              // M:472-474
    return ns3d_advect_swap(ctx, f, sp);                                                               // M:475-477
}
