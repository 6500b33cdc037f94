/*
 * ns3d.h -- C ABI of libns3d.so: the B200-native (sm_100a) implementation of the
 * per-timestep hot path of mattbuergler/NavierStokes3D.
 *
 * The reference has no FFI/plugin interface: its seam is the set of call sites in
 * the two Julia run scripts (M = scripts/NavierStokes3D_multi_gpu.jl,
 * G = scripts/NavierStokes3D_gpu.jl).  Every entry point below names the call site
 * it replaces.  A Julia driver keeps its structure and re-points each of those
 * lines at a `ccall` of the same-named function (see INTEGRATION.md and
 * julia/NS3DNative.jl); the Python ctypes binding in navierstokes3d_b200/native.py
 * is the executable proof of this ABI.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  `double*` arguments are DEVICE
 *    pointers obtained from ns3d_zeros() unless the name starts with `h_`.
 *  - Arrays are dense, column-major, x fastest, exactly the reference's shapes
 *    (M:343-360): cell fields (nx,ny,nz); Vx (nx+1,ny,nz); Vy (nx,ny+1,nz);
 *    Vz (nx,ny,nz+1); txy/txz/tyz (nx-1,ny-1,nz-1); dPrdtau/Rp (nx-2,ny-2,nz-2).
 *  - Argument order = the reference kernel's (arrays, then scalars), followed by
 *    the LOCAL grid size nx,ny,nz (Julia reads it from size(A); C cannot).
 *  - Every function returns 0 on success, a negative NS3D_E* code on failure and
 *    never throws/aborts across the boundary; ns3d_last_error() has the message.
 *  - Operators are asynchronous on the context's stream; functions that return a
 *    host scalar (ns3d_max_abs, ns3d_pt_solve, ns3d_step, ns3d_d2h) synchronise.
 *  - A context is bound to one GPU and is not thread-safe (Julia calls from one
 *    task).  There is NO CPU fallback: without a CUDA device ns3d_create() fails.
 */
#ifndef NS3D_H
#define NS3D_H

#include <stddef.h>

#if defined(__GNUC__)
#define NS3D_API __attribute__((visibility("default")))
#else
#define NS3D_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ns3d_ctx ns3d_ctx;

enum {
    NS3D_OK = 0,
    NS3D_EINVAL = -1,  /* bad argument */
    NS3D_ECUDA = -2,   /* CUDA runtime error (message has cudaGetErrorString) */
    NS3D_ENOMEM = -3,  /* device allocation failed */
    NS3D_ECOMM = -4,   /* NCCL error / communicator missing */
    NS3D_ENODEV = -5   /* no CUDA device */
};

/* Arithmetic mode of the floating-point kernels (ns3d_set_mode).
 *  PARITY : IEEE division, no FMA contraction, the reference's association order
 *           -> bit-equal to the CPU oracle (the reference's Threads backend).
 *  FAST   : the divisions `x/dx/dx` use a host-precomputed correctly-rounded
 *           reciprocal plus one FMA residual correction (Markstein); everything
 *           else as PARITY.  Equal to PARITY except for signed zeros / non-finite
 *           inputs (and, in theory, rare double-rounding cases).
 *  FASTEST: multiply by precomputed 1/(dx*dx), FMA contraction allowed.  Tolerance
 *           contract: identical PT iteration counts, <= 1e-10 relative drift.      */
enum { NS3D_PARITY = 0, NS3D_FAST = 1, NS3D_FASTEST = 2 };

/* Which script's boundary conditions / cylinder coordinates. */
enum { NS3D_VARIANT_M = 0, NS3D_VARIANT_G = 1 };

/* ---- lifecycle ----------------------------------------------------------- */
/* Replaces @init_parallel_stencil(CUDA, Float64, 3) (M:5) + IGG's device selection. */
NS3D_API int ns3d_create(int device, ns3d_ctx** out);
NS3D_API int ns3d_destroy(ns3d_ctx* ctx);
NS3D_API const char* ns3d_version(void);
NS3D_API const char* ns3d_last_error(const ns3d_ctx* ctx); /* ctx may be NULL: last create() error */
NS3D_API int ns3d_set_mode(ns3d_ctx* ctx, int mode);
NS3D_API int ns3d_get_mode(const ns3d_ctx* ctx);
/* Tuning / diagnostic knobs.  None of them changes a result (every kernel variant is bit-identical
 * in a given arithmetic mode); unknown names fail with NS3D_EINVAL.
 *   "tb2"         1* two PT iterations per launch, 0 = one-iteration kernel only
 *   "tb2_ty"      tile height of the two-iteration kernels: 0* = by grid size, 8, 16, 32
 *   "tb2_slim"    1* pt_tb2s_kernel for launches off the slab interfaces, 0 = pt_tb2_kernel everywhere
 *   "tb2_np"      1* in-plane neighbours of the next plane loaded one step ahead; 0 = no prefetch beyond the
 *                 three streams (the baseline the prefetches were measured against; ignores tb2_pf)
 *   "tb2_pf"      planes of software prefetch into L2 ahead of the register prefetch: 0, 1*, 2
 *   "tb2_spec"    1* compile-time-stride instantiation when the grid's x-y extent has one
 *   "tb2_dual"    0* | 2: pt_tb2d_kernel, two tile rows per thread (value = CTAs per SM)       [candidate]
 *   "tb2_pairbar" 0* | 1: pairwise row barriers instead of __syncthreads in pt_tb2s_kernel     [candidate]
 *   "tb2_slim_faces" 0* | 1: slab-interface chunks with pt_tb2sp_kernel (the slim pipeline with peer
 *                 loads/stores) instead of pt_tb2_kernel<.,16,true>                            [candidate]
 *   "pt_bands"    0* | 2..8: split every two-iteration launch into z-bands with band-to-band dependencies
 *                 so that consecutive launches overlap (single rank)                           [candidate]
 *   "pt_minb"     CTAs per SM the one-iteration kernel is compiled for: 0* = per mode, 3..6
 *   "serpentine"  -1* = by working-set size, 0, 1: alternate the z sweep direction between launches
 *   "graphs"      1* replay chunks of PT iterations as CUDA graphs
 *   "p2p_halo"    1* peer-memory halo exchange fused into the PT kernels on slabs, 0 = NCCL send/recv
 * (* = default) */
NS3D_API int ns3d_set_option(ns3d_ctx* ctx, const char* name, int value);
NS3D_API int ns3d_sync(ns3d_ctx* ctx);
/* Number of kernels this library launched on ctx since creation (bench bookkeeping). */
NS3D_API long long ns3d_launch_count(const ns3d_ctx* ctx);
/* The context's CUDA stream (a cudaStream_t), so callers can record events on it. */
NS3D_API void* ns3d_stream(ns3d_ctx* ctx);

/* Device-side initialisers.  The scripts build their initial arrays on the host and upload them:
 *   G:86   Vx = [vin*(7/6)*((zc[iz]+lz/2)/lz)^(1/6) + 0*yc[iy] + 0*xv[ix] ...]   G:87 / M:370   Pr = [-(z - ..)*ρ*g ...]
 *   M:369  Vy[1,:,:] .= vin
 * Every comprehension depends on iz alone: ns3d_fill_profile_z sets A[ix,iy,iz] = h_profile[iz] from sz host values
 * (the power law's `^(1/6)` is evaluated on the host, whose libm CUDA's pow does not match to the last bit), and
 * ns3d_fill_plane_x sets A[ix,:,:] = value (ix 0-based).  No 3-D array exists on the host.                        */
NS3D_API int ns3d_fill_profile_z(ns3d_ctx* ctx, double* A, int sx, int sy, int sz, const double* h_profile);
/* ... with the comprehension's other terms kept apart: A[ix,iy,iz] = (h_profile[iz] + h_add_y[iy]) + h_add_z[iz].  M:370 is
 * `-(z_g(iz,dz,C)-dz/2)*ρ*g + 0*yc[iy] + 0*zc[iz]` with g = 0 (M:316): three SIGNED zeros whose sum is -0.0 exactly where all
 * three are, so the sign of the initial Pr depends on iy as well (sy + 2 sz host values cross PCIe).                   */
NS3D_API int ns3d_fill_profile_zy(ns3d_ctx* ctx, double* A, int sx, int sy, int sz, const double* h_profile, const double* h_add_y,
                                  const double* h_add_z);
NS3D_API int ns3d_fill_plane_x(ns3d_ctx* ctx, double* A, int sx, int sy, int sz, int ix, double value);
/* Asynchronous host <-> device copies for a driver that overlaps the traffic of one time step with the computation
 * of another (the reference's `Data.Array(x)` / `Array(A)` are blocking; nothing in it corresponds to these).  Uploads
 * run on the context's upload stream, downloads on its download stream; host buffers must be page-locked and stay
 * valid until the copy has completed.  ns3d_stream_wait(ctx, waiter, signaller) makes everything enqueued on `waiter`
 * from now on wait for what has been enqueued on `signaller` so far; ns3d_stream_sync blocks the host.            */
enum { NS3D_STREAM_COMPUTE = 0, NS3D_STREAM_H2D = 1, NS3D_STREAM_D2H = 2 };
NS3D_API int ns3d_h2d_async(ns3d_ctx* ctx, double* dptr, const double* h_pinned_src, size_t count);
NS3D_API int ns3d_d2h_async(ns3d_ctx* ctx, double* h_pinned_dst, const double* dptr, size_t count);
NS3D_API int ns3d_stream_wait(ns3d_ctx* ctx, int waiter, int signaller);
NS3D_API int ns3d_stream_sync(ns3d_ctx* ctx, int which);
/* ---- device field allocator  (replaces @zeros M:343-360, Data.Array M:370, Array() M:399) */
NS3D_API int ns3d_zeros(ns3d_ctx* ctx, int sx, int sy, int sz, double** dptr);
NS3D_API int ns3d_free(ns3d_ctx* ctx, double* dptr);
NS3D_API int ns3d_h2d(ns3d_ctx* ctx, double* dptr, const double* h_src, size_t count);
NS3D_API int ns3d_d2h(ns3d_ctx* ctx, double* h_dst, const double* dptr, size_t count);
NS3D_API int ns3d_copy(ns3d_ctx* ctx, double* dst, const double* src, size_t count); /* `A_o .= A` M:475 */
NS3D_API int ns3d_fill(ns3d_ctx* ctx, double* dptr, double value, size_t count);
NS3D_API size_t ns3d_bytes_allocated(const ns3d_ctx* ctx);

/* ---- level 1: one entry point per reference kernel call site ---------------- */
/* update_τ!  M:36-44 / G:177-185, call site M:449 */
NS3D_API int ns3d_update_tau(ns3d_ctx* ctx, double* txx, double* tyy, double* tzz, double* txy, double* txz,
                    double* tyz, const double* Vx, const double* Vy, const double* Vz, double mu,
                    double dx, double dy, double dz, int nx, int ny, int nz);
/* predict_V!  M:50-55, call site M:451 */
NS3D_API int ns3d_predict_V(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, const double* txx,
                   const double* tyy, const double* tzz, const double* txy, const double* txz,
                   const double* tyz, double rho, double g, double dt, double dx, double dy, double dz,
                   int nx, int ny, int nz);
/* update_∇V!  M:61-64, call site M:454 */
NS3D_API int ns3d_update_divV(ns3d_ctx* ctx, double* divV, const double* Vx, const double* Vy, const double* Vz,
                     double dx, double dy, double dz, int nx, int ny, int nz);
/* update_dPrdτ!  M:70-73, call site M:459 */
NS3D_API int ns3d_update_dPrdtau(ns3d_ctx* ctx, const double* Pr, double* dPrdtau, const double* divV, double rho,
                        double dt, double dtau, double damp, double dx, double dy, double dz, int nx,
                        int ny, int nz);
/* update_Pr!  M:79-82, call site M:461 */
NS3D_API int ns3d_update_Pr(ns3d_ctx* ctx, double* Pr, const double* dPrdtau, double dtau, int nx, int ny, int nz);
/* compute_res!  M:88-91, call site M:465 */
NS3D_API int ns3d_compute_res(ns3d_ctx* ctx, double* Rp, const double* Pr, const double* divV, double rho,
                     double dt, double dx, double dy, double dz, int nx, int ny, int nz);
/* max_g(abs.(A))  M:21,466 / maximum(abs.(Rp)) G:132.  NaN-propagating like Julia's
 * `maximum`; max-allreduced over the communicator when one is attached.          */
NS3D_API int ns3d_max_abs(ns3d_ctx* ctx, const double* A, size_t count, double* h_out);
/* correct_V!  M:97-102, call site M:472 */
NS3D_API int ns3d_correct_V(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, const double* Pr, double dt,
                   double rho, double dx, double dy, double dz, int nx, int ny, int nz);
/* bc_x!/bc_y!/bc_z!  M:108-132 on an array of shape (sx,sy,sz) */
NS3D_API int ns3d_bc_x(ns3d_ctx* ctx, double* A, int sx, int sy, int sz);
NS3D_API int ns3d_bc_y(ns3d_ctx* ctx, double* A, int sx, int sy, int sz);
NS3D_API int ns3d_bc_z(ns3d_ctx* ctx, double* A, int sx, int sy, int sz);
/* bc_x_Vx!  M:138-141 ; bc_x_Pr!  M:147-150 */
NS3D_API int ns3d_bc_x_Vx(ns3d_ctx* ctx, double* A, double V, int sx, int sy, int sz);
NS3D_API int ns3d_bc_x_Pr(ns3d_ctx* ctx, double* A, double val, int sx, int sy, int sz);
/* bc_zV!  G:239-243 ; bc_xhydstatic!  G:257-261 */
NS3D_API int ns3d_bc_zV(ns3d_ctx* ctx, double* A, int sx, int sy, int sz);
NS3D_API int ns3d_bc_xhydstatic(ns3d_ctx* ctx, double* A, double dz, int nz_arg, double g, double rho, int sx,
                       int sy, int sz);
/* set_bc_Vel!  M:156-169 (inlet_guard = `xvo_g == -lx/2` evaluated by the caller,
 * halo update included when a communicator is attached) / G:264-279            */
NS3D_API int ns3d_set_bc_Vel_M(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, int inlet_guard, double vin,
                      int nx, int ny, int nz);
NS3D_API int ns3d_set_bc_Vel_G(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, int nx, int ny, int nz);
/* set_bc_Pr!  M:175-184 (outlet_guard = `xve_g == lx/2`) / G:281-286 */
NS3D_API int ns3d_set_bc_Pr_M(ns3d_ctx* ctx, double* Pr, int outlet_guard, double val, int nx, int ny, int nz);
NS3D_API int ns3d_set_bc_Pr_G(ns3d_ctx* ctx, double* Pr, double dz, int nz_arg, double g, double rho, int nx,
                     int ny, int nz);
/* advect! + backtrack! + lerp  M:190-243, call site M:476.  Vz is (faithfully) never written. */
NS3D_API int ns3d_advect(ns3d_ctx* ctx, double* Vx, const double* Vx_o, double* Vy, const double* Vy_o,
                double* Vz, const double* Vz_o, double* C, const double* C_o, double dt, double dx,
                double dy, double dz, int nx, int ny, int nz);
/* set_cylinder!  M:249-281, call sites M:372,452,473 (zco_g,lx,ly,lz,dz are dead arguments
 * in the reference and are dropped) / G:336-368, call sites G:123,139            */
NS3D_API int ns3d_set_cylinder_M(ns3d_ctx* ctx, double* C, double* Vx, double* Vy, double* Vz, double a2,
                        double b2, double ox, double oy, double sinb, double cosb, double xco_g,
                        double yco_g, double dx, double dy, int nx, int ny, int nz);
NS3D_API int ns3d_set_cylinder_G(ns3d_ctx* ctx, double* C, double* Vx, double* Vy, double* Vz, double a2,
                        double b2, double ox, double oy, double sinb, double cosb, double lx,
                        double ly, double dx, double dy, int nx, int ny, int nz);

/* ---- communication: z-slab decomposition, one rank per GPU ---------------------
 * Replaces init_global_grid(nx,ny,nz; dimx=1,dimy=1,dimz=N) (M:325), update_halo!
 * (10 call sites) and MPI.Allreduce(MAX) in max_g (M:21).
 * The 128-byte id is created on rank 0 and broadcast by the caller (MPI.bcast in
 * Julia, torch.distributed in the Python harness).                               */
NS3D_API int ns3d_comm_unique_id(char id[128]);
NS3D_API int ns3d_comm_init(ns3d_ctx* ctx, int rank, int nranks, const char id[128]);
NS3D_API int ns3d_comm_rank(const ns3d_ctx* ctx);
NS3D_API int ns3d_comm_size(const ns3d_ctx* ctx);
/* update_halo!(A1,...,An): fields[f] has shape (sx[f],sy[f],sz[f]); local cell count
 * along z is nz (overlap 2: a field with sz = nz+1 exchanges planes 3 / nz-1). */
NS3D_API int ns3d_update_halo(ns3d_ctx* ctx, double* const* fields, const int* sx, const int* sy, const int* sz,
                     int nfields, int nz);
/* MPI.Allreduce(x, MPI.MAX, comm) of one host double (NaN-propagating). */
NS3D_API int ns3d_allreduce_max(ns3d_ctx* ctx, double* h_inout);

/* ---- output path: interior extraction, save, heat-map planes ------------------
 * Replaces `Array(A)[2:end-1,2:end-1,2:end-1]` ahead of gather!/save_array (M:399-412, 481-523,
 * 528-532; G:169) and the planes `A_v[:,:,k]`, `A_v[:,j,:]` of the visualisation block
 * (M:422-431).  The box [x0,x1) x [y0,y1) x [z0,z1) (0-based) of the device array A (sx,sy,sz) is
 * packed on the device and copied densely, column-major, to h_out: Float64, or Float32 when
 * f32 != 0 (`convert.(Float32, A_v)`, M:408; round to nearest even like Julia).  The interior is
 * the box [1,sx-1) x [1,sy-1) x [1,sz-1); a plane is a box one point thick.  Synchronises.    */
NS3D_API int ns3d_box_d2h(ns3d_ctx* ctx, const double* A, int sx, int sy, int sz, int x0, int x1, int y0, int y1,
                          int z0, int z1, void* h_out, int f32);
/* gather!(A_inn, A_v) (M:399-403, 481-485, 528-532) on z-slabs: every rank passes the box of its local array
 * that belongs to the global array (same x-y extent on every rank; nplanes_all[r] = z1 - z0 of rank r,
 * the same list on every rank); the boxes are packed on the devices, sent to rank 0 over NCCL and
 * concatenated along z in h_out (rank 0 only; ignored elsewhere).  COLLECTIVE.  One rank: ns3d_box_d2h. */
NS3D_API int ns3d_gather_box(ns3d_ctx* ctx, const double* A, int sx, int sy, int sz, int x0, int x1, int y0, int y1,
                             int z0, int z1, const int* nplanes_all, void* h_out, int f32);

/* ---- level 2: fused fast path ------------------------------------------------- */
typedef struct ns3d_pt_params {
    int nx, ny, nz;         /* local grid */
    int variant;            /* NS3D_VARIANT_M | NS3D_VARIANT_G : pressure BC set */
    double rho, dt, dtau, damp, dx, dy, dz;
    double eps_it;          /* leave the loop when err < eps_it || !isfinite(err)  (M:469) */
    double err_num, err_den; /* err = max|Rp| * err_num / err_den   (ly^2, psc; M:466) */
    int niter, nchk;        /* M:328-329 / G:48-49 */
    int outlet_guard;       /* variant M: `xve_g == lx/2` (M:179) */
    double outlet_val;      /* variant M: 0.0 (M:463) */
    double g;               /* variant G: hydrostatic planes (G:257-261) use rho, g, dz, nz */
    int zchunk;             /* tuning: z-planes marched per CTA (0 = auto) */
    int reserved;
} ns3d_pt_params;

/* The whole pseudo-transient loop M:458-471 / G:126-137 in fused kernels, on pitched copies of
 * the three arrays owned by the context (packed on entry, unpacked on exit): K x (K5+K6+set_bc_Pr!
 * (+update_halo!(Pr) over peer memory on z-slabs)) per launch, K = 2 by default;
 * compute_res!+abs+maximum(+allreduce) one launch per check.
 * h_err_hist (capacity err_cap) receives err at every check (M:467).  On return Pr
 * and dPrdtau hold exactly the reference's iterates after *h_iters iterations. */
NS3D_API int ns3d_pt_solve(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV,
                  const ns3d_pt_params* p, int* h_iters, double* h_err_hist, int err_cap,
                  int* h_nchecks);
/* Exactly n fused iterations, no residual check (Poisson-only benchmark, config C). */
NS3D_API int ns3d_pt_iterate(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV,
                    const ns3d_pt_params* p, int n);

/* Which kernel the fused loop launches for these parameters on this context (mode, options, communicator):
 * a NUL-terminated description into buf (capacity cap) and the number of PT iterations one launch of
 * it performs -- what bench.py divides launch times by.  No reference counterpart (diagnostics).      */
NS3D_API int ns3d_pt_describe(ns3d_ctx* ctx, const ns3d_pt_params* p, char* buf, int cap, int* iters_per_launch);

typedef struct ns3d_fields {
    double *Pr, *dPrdtau, *C, *C_o, *txx, *tyy, *tzz, *txy, *txz, *tyz;
    double *Vx, *Vy, *Vz, *Vx_o, *Vy_o, *Vz_o, *divV, *Rp;
} ns3d_fields;

typedef struct ns3d_step_params {
    ns3d_pt_params pt;
    double mu, vin;
    double a2, b2, ox, oy, sinb, cosb; /* cylinder */
    double xco_g, yco_g;               /* variant M cylinder origin (M:363-364) */
    double lx, ly;                     /* variant G cylinder coordinates (G:337) */
    int inlet_guard;                   /* variant M: `xvo_g == -lx/2` (M:164) */
    int reserved;
} ns3d_step_params;

/* The three once-per-step groups around the PT loop, each with the effect of the level-1 sequence of its lines:
 *   predictor   M:449-455 / G:121-124  update_τ!, predict_V!, set_cylinder! (ONE kernel, the stresses never reach memory:
 *                                      the six stress arrays are not touched and may be NULL; Vx_o, Vy_o, Vz_o are used as
 *                                      scratch -- the reference overwrites them at M:475 before it reads them again),
 *                                      update_halo!(C,V), update_∇V!, update_halo!(∇V)
 *   corrector   M:472-474 / G:138-140  correct_V! + set_cylinder! (one kernel), set_bc_Vel! (halo inside)
 *   advect_swap M:475-477 / G:141-142  `A_o .= A` x4, advect!, update_halo!(Vx,Vy,Vz)
 * Asynchronous on the context's stream.                                                  */
NS3D_API int ns3d_predictor(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* p);
NS3D_API int ns3d_corrector(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* p);
NS3D_API int ns3d_advect_swap(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* p);

/* One whole time step M:449-477 / G:121-142 (everything between the `for it` line and the visualisation block) in
 * fused kernels: the velocity makes one round trip through the `_o` arrays (predictor V -> V_o, corrector and boundary
 * conditions in place on V_o, advection V_o -> V writing every entry), so the four copies of M:475 never happen as
 * copies.  On return EVERY array holds what the reference's step leaves in it -- the snapshots Vx_o, Vy_o, Vz_o, C_o
 * included, same PT iteration count and err history -- except the six stress arrays, which are not touched (NULL
 * allowed).                                                                                                          */
NS3D_API int ns3d_step(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* p, int* h_iters,
              double* h_err_hist, int err_cap, int* h_nchecks);

#ifdef __cplusplus
}
#endif
#endif /* NS3D_H */
