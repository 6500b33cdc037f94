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
// z-slab halo exchange of the velocity triple / of cell fields (update_halo!, M:453,455,167,477)
namespace {
int halo_V(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, double* C, int nx, int ny, int nz)
{
    if (ctx->nranks == 1) return NS3D_OK;
    double* h[4] = {Vx, Vy, Vz, C};
    const int sx[4] = {nx + 1, nx, nx, nx}, sy[4] = {ny, ny + 1, ny, ny}, sz[4] = {nz, nz, nz + 1, nz};
    return ns3d_internal_halo_z(ctx, ctx->stream, h, sx, sy, sz, C ? 4 : 3, nz);
}
int halo_cell(ns3d_ctx* ctx, double* A, int nx, int ny, int nz)
{
    if (ctx->nranks == 1) return NS3D_OK;
    double* h[1] = {A};
    return ns3d_internal_halo_z(ctx, ctx->stream, h, &nx, &ny, &nz, 1, nz);
}
int set_bc_vel(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, const ns3d_step_params* sp)
{
    const ns3d_pt_params& p = sp->pt;
    if (p.variant == NS3D_VARIANT_M) return ns3d_set_bc_Vel_M(ctx, Vx, Vy, Vz, sp->inlet_guard, sp->vin, p.nx, p.ny, p.nz);  // M:474
    return ns3d_set_bc_Vel_G(ctx, Vx, Vy, Vz, p.nx, p.ny, p.nz);                                                             // G:140
}
}  // namespace

// Chorin predictor, M:449-455 / G:121-124: update_τ!, predict_V!, set_cylinder!, update_∇V! and the halo updates
// between them (update_halo!(τxx,τyy,τzz) M:450 is redundant: the stresses are computed on the halo cells too).
// The first three are ONE kernel (predictor_kernel, ns3d_step.cu) that writes the predicted velocity into the
// `_o` arrays (scratch at this point of a step: M:475 overwrites them before they are read again); as a stand-alone
// group it is copied back so that Vx, Vy, Vz hold it as after the reference's lines.  The stress arrays are not used.
extern "C" int ns3d_predictor(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* sp)
{
    NS3D_CHECK_CTX(ctx);
    if (!f || !sp) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_predictor: NULL argument");
    const ns3d_pt_params& p = sp->pt;
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_TRY(ns3d_internal_predict_fused(ctx, f->Vx_o, f->Vy_o, f->Vz_o, f->C, f->Vx, f->Vy, f->Vz, sp));  // M:449-452
    NS3D_TRY(ns3d_copy(ctx, f->Vx, f->Vx_o, (size_t)(nx + 1) * ny * nz));
    NS3D_TRY(ns3d_copy(ctx, f->Vy, f->Vy_o, (size_t)nx * (ny + 1) * nz));
    NS3D_TRY(ns3d_copy(ctx, f->Vz, f->Vz_o, (size_t)nx * ny * (nz + 1)));
    NS3D_TRY(halo_V(ctx, f->Vx, f->Vy, f->Vz, f->C, nx, ny, nz));                                      // M:453
    NS3D_TRY(ns3d_update_divV(ctx, f->divV, f->Vx, f->Vy, f->Vz, p.dx, p.dy, p.dz, nx, ny, nz));      // M:454
    return halo_cell(ctx, f->divV, nx, ny, nz);                                                        // M:455
}

// Pressure-gradient correction, M:472-474 / G:138-140: correct_V! + set_cylinder! in one kernel (in place), set_bc_Vel!.
extern "C" int ns3d_corrector(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* sp)
{
    NS3D_CHECK_CTX(ctx);
    if (!f || !sp) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_corrector: NULL argument");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_TRY(ns3d_internal_correct_fused(ctx, f->Vx, f->Vy, f->Vz, f->C, nullptr, f->Pr, sp));   // M:472-473
    return set_bc_vel(ctx, f->Vx, f->Vy, f->Vz, sp);                                             // M:474
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
    return halo_V(ctx, f->Vx, f->Vy, f->Vz, nullptr, nx, ny, nz);                                      // M:477
}

// One time step, M:449-477 / G:121-142.  The velocity makes one round trip per step: predictor V -> V_o, corrector and
// boundary conditions in place on V_o, advection V_o -> V (see ns3d_step.cu) -- 27 field passes around the PT loop
// instead of the 47 of the call-by-call sequence, no stress arrays (f->t** may be NULL), no `A_o .= A` copies; on return
// every other array holds what the reference's step leaves in it (V_o, C_o: the pre-advection snapshot of M:475).
extern "C" int ns3d_step(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* sp, int* h_iters,
                         double* h_err_hist, int err_cap, int* h_nchecks)
{
    NS3D_CHECK_CTX(ctx);
    if (!f || !sp) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_step: NULL argument");
    const ns3d_pt_params& p = sp->pt;
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_TRY(ns3d_internal_predict_fused(ctx, f->Vx_o, f->Vy_o, f->Vz_o, f->C, f->Vx, f->Vy, f->Vz, sp));  // M:449-452
    NS3D_TRY(halo_V(ctx, f->Vx_o, f->Vy_o, f->Vz_o, f->C, nx, ny, nz));                                     // M:453
    NS3D_TRY(ns3d_update_divV(ctx, f->divV, f->Vx_o, f->Vy_o, f->Vz_o, p.dx, p.dy, p.dz, nx, ny, nz));     // M:454
    NS3D_TRY(halo_cell(ctx, f->divV, nx, ny, nz));                                                          // M:455
    NS3D_TRY(ns3d_pt_solve(ctx, f->Pr, f->dPrdtau, f->divV, &sp->pt, h_iters, h_err_hist, err_cap, h_nchecks));  // M:458-471
    NS3D_TRY(ns3d_internal_correct_fused(ctx, f->Vx_o, f->Vy_o, f->Vz_o, f->C, f->C_o, f->Pr, sp));         // M:472-473, C_o .= C
    NS3D_TRY(set_bc_vel(ctx, f->Vx_o, f->Vy_o, f->Vz_o, sp));                                               // M:474
    NS3D_TRY(ns3d_internal_advect_all(ctx, f->Vx, f->Vx_o, f->Vy, f->Vy_o, f->Vz, f->Vz_o, f->C, f->C_o, p.dt, p.dx, p.dy, p.dz,
                                      nx, ny, nz));                                                         // M:475-476
    return halo_V(ctx, f->Vx, f->Vy, f->Vz, nullptr, nx, ny, nz);                                           // M:477
}
