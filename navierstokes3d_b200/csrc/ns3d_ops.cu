// libns3d.so -- level-1 operators: one hand-written sm_100a kernel per reference
// ParallelStencil kernel, same arguments, same arithmetic.
//
// The whole library is compiled with --fmad=false: the reference's CPU (Threads) backend
// never contracts a*b+c, and double-precision `/` is IEEE-correct in CUDA, so every kernel
// in this file is bit-equal to the oracle by construction.  Indices are 0-based here; the
// formulas are the 1-based ones of SURVEY.md Appendix A shifted uniformly by one.
//
// All of these run once per time step (< 1 % of the bytes of a step, the PT loop in
// ns3d_pt.cu is the hot loop), so they are plain one-thread-per-point kernels with x on
// threadIdx.x: every global access is a coalesced row segment.
#include <algorithm>

#include "ns3d_internal.cuh"

#define VXI(i, j, k) Vx[idx3(i, j, k, nx + 1, ny)]
#define VYI(i, j, k) Vy[idx3(i, j, k, nx, ny + 1)]
#define VZI(i, j, k) Vz[idx3(i, j, k, nx, ny)]

static inline dim3 block3() { return dim3(32, 4, 2); }
static inline dim3 grid3(int sx, int sy, int sz)
{
    dim3 b = block3();
    return dim3(cdiv(sx, b.x), cdiv(sy, b.y), cdiv(sz, b.z));
}

#define THREAD_IJK()                                         \
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     \
    const int j = blockIdx.y * blockDim.y + threadIdx.y;     \
    const int k = blockIdx.z * blockDim.z + threadIdx.z;

// ---------------------------------------------------------------------------------------------
// K1 update_τ!  M:36-44
// ---------------------------------------------------------------------------------------------
__global__ void update_tau_kernel(double* __restrict__ txx, double* __restrict__ tyy, double* __restrict__ tzz,
                                  double* __restrict__ txy, double* __restrict__ txz, double* __restrict__ tyz,
                                  const double* __restrict__ Vx, const double* __restrict__ Vy,
                                  const double* __restrict__ Vz, double mu, double dx, double dy, double dz,
                                  int nx, int ny, int nz)
{
    THREAD_IJK();
    if (i >= nx || j >= ny || k >= nz) return;
    const double twomu = 2 * mu;
    {
        const double dxa = VXI(i + 1, j, k) - VXI(i, j, k);
        const double dya = VYI(i, j + 1, k) - VYI(i, j, k);
        const double dza = VZI(i, j, k + 1) - VZI(i, j, k);
        const double divv = (dxa / dx + dya / dy) + dza / dz;
        const size_t c = idx3(i, j, k, nx, ny);
        txx[c] = twomu * (dxa / dx - divv / 3.0);
        tyy[c] = twomu * (dya / dy - divv / 3.0);
        tzz[c] = twomu * (dza / dz - divv / 3.0);
    }
    if (i < nx - 1 && j < ny - 1 && k < nz - 1) {
        const double vx111 = VXI(i + 1, j + 1, k + 1), vy111 = VYI(i + 1, j + 1, k + 1), vz111 = VZI(i + 1, j + 1, k + 1);
        const double dyiVx = vx111 - VXI(i + 1, j, k + 1);
        const double dxiVy = vy111 - VYI(i, j + 1, k + 1);
        const double dziVx = vx111 - VXI(i + 1, j + 1, k);
        const double dxiVz = vz111 - VZI(i, j + 1, k + 1);
        const double dziVy = vy111 - VYI(i + 1, j + 1, k);
        const double dyiVz = vz111 - VZI(i + 1, j, k + 1);
        const size_t e = idx3(i, j, k, nx - 1, ny - 1);
        txy[e] = mu * (dyiVx / dy + dxiVy / dx);
        txz[e] = mu * (dziVx / dz + dxiVz / dx);
        tyz[e] = mu * (dziVy / dz + dyiVz / dy);
    }
}

extern "C" int ns3d_update_tau(ns3d_ctx* ctx, double* txx, double* tyy, double* tzz, double* txy, double* txz,
                               double* tyz, const double* Vx, const double* Vy, const double* Vz, double mu,
                               double dx, double dy, double dz, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    if (nx < 3 || ny < 3 || nz < 3) return ns3d_fail(ctx, NS3D_EINVAL, "grid must be at least 3^3");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    update_tau_kernel<<<grid3(nx, ny, nz), block3(), 0, ctx->stream>>>(txx, tyy, tzz, txy, txz, tyz, Vx, Vy, Vz, mu,
                                                                      dx, dy, dz, nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// ---------------------------------------------------------------------------------------------
// K2 predict_V!  M:50-55
// ---------------------------------------------------------------------------------------------
#define TXX(i, j, k) txx[idx3(i, j, k, nx, ny)]
#define TYY(i, j, k) tyy[idx3(i, j, k, nx, ny)]
#define TZZ(i, j, k) tzz[idx3(i, j, k, nx, ny)]
#define TXY(i, j, k) txy[idx3(i, j, k, nx - 1, ny - 1)]
#define TXZ(i, j, k) txz[idx3(i, j, k, nx - 1, ny - 1)]
#define TYZ(i, j, k) tyz[idx3(i, j, k, nx - 1, ny - 1)]

__global__ void predict_V_kernel(double* __restrict__ Vx, double* __restrict__ Vy, double* __restrict__ Vz,
                                 const double* __restrict__ txx, const double* __restrict__ tyy,
                                 const double* __restrict__ tzz, const double* __restrict__ txy,
                                 const double* __restrict__ txz, const double* __restrict__ tyz, double rho,
                                 double g, double dt, double dx, double dy, double dz, int nx, int ny, int nz)
{
    THREAD_IJK();
    const double dtr = dt / rho;
    if (i < nx - 1 && j < ny - 2 && k < nz - 2) {
        const double a = TXX(i + 1, j + 1, k + 1) - TXX(i, j + 1, k + 1);
        const double b = TXY(i, j + 1, k) - TXY(i, j, k);
        const double c = TXZ(i, j, k + 1) - TXZ(i, j, k);
        double* v = &VXI(i + 1, j + 1, k + 1);
        *v = *v + dtr * ((a / dx + b / dy) + c / dz);
    }
    if (i < nx - 2 && j < ny - 1 && k < nz - 2) {
        const double a = TYY(i + 1, j + 1, k + 1) - TYY(i + 1, j, k + 1);
        const double b = TXY(i + 1, j, k) - TXY(i, j, k);
        const double c = TYZ(i, j, k + 1) - TYZ(i, j, k);
        double* v = &VYI(i + 1, j + 1, k + 1);
        *v = *v + dtr * ((a / dy + b / dx) + c / dz);
    }
    if (i < nx - 2 && j < ny - 2 && k < nz - 1) {
        const double a = TZZ(i + 1, j + 1, k + 1) - TZZ(i + 1, j + 1, k);
        const double b = TXZ(i + 1, j, k) - TXZ(i, j, k);
        const double c = TYZ(i, j + 1, k) - TYZ(i, j, k);
        double* v = &VZI(i + 1, j + 1, k + 1);
        *v = *v + dtr * (((a / dz + b / dx) + c / dy) - rho * g);
    }
}

extern "C" int ns3d_predict_V(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, const double* txx,
                              const double* tyy, const double* tzz, const double* txy, const double* txz,
                              const double* tyz, double rho, double g, double dt, double dx, double dy, double dz,
                              int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    predict_V_kernel<<<grid3(nx - 1, ny - 1, nz - 1), block3(), 0, ctx->stream>>>(Vx, Vy, Vz, txx, tyy, tzz, txy, txz,
                                                                                 tyz, rho, g, dt, dx, dy, dz, nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// ---------------------------------------------------------------------------------------------
// K4 update_∇V!  M:61-64
// ---------------------------------------------------------------------------------------------
__global__ void update_divV_kernel(double* __restrict__ divV, const double* __restrict__ Vx,
                                   const double* __restrict__ Vy, const double* __restrict__ Vz, double dx,
                                   double dy, double dz, int nx, int ny, int nz)
{
    THREAD_IJK();
    if (i >= nx || j >= ny || k >= nz) return;
    const double dxa = VXI(i + 1, j, k) - VXI(i, j, k);
    const double dya = VYI(i, j + 1, k) - VYI(i, j, k);
    const double dza = VZI(i, j, k + 1) - VZI(i, j, k);
    divV[idx3(i, j, k, nx, ny)] = (dxa / dx + dya / dy) + dza / dz;
}

extern "C" int ns3d_update_divV(ns3d_ctx* ctx, double* divV, const double* Vx, const double* Vy, const double* Vz,
                                double dx, double dy, double dz, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    update_divV_kernel<<<grid3(nx, ny, nz), block3(), 0, ctx->stream>>>(divV, Vx, Vy, Vz, dx, dy, dz, nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// ---------------------------------------------------------------------------------------------
// K5 update_dPrdτ!  M:70-73 ; K6 update_Pr!  M:79-82 ; K8 compute_res!  M:88-91
// (unfused level-1 forms; the fused hot-loop kernel lives in ns3d_pt.cu)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pt_bracket_l1(const double* __restrict__ Pr, const double* __restrict__ divV,
                                                double rdt, double dx, double dy, double dz, int nx, int ny, int i,
                                                int j, int k)
{
    const double c = Pr[idx3(i + 1, j + 1, k + 1, nx, ny)];
    const double d2x = (Pr[idx3(i + 2, j + 1, k + 1, nx, ny)] - c) - (c - Pr[idx3(i, j + 1, k + 1, nx, ny)]);
    const double d2y = (Pr[idx3(i + 1, j + 2, k + 1, nx, ny)] - c) - (c - Pr[idx3(i + 1, j, k + 1, nx, ny)]);
    const double d2z = (Pr[idx3(i + 1, j + 1, k + 2, nx, ny)] - c) - (c - Pr[idx3(i + 1, j + 1, k, nx, ny)]);
    return ((d2x / dx / dx + d2y / dy / dy) + d2z / dz / dz) - rdt * divV[idx3(i + 1, j + 1, k + 1, nx, ny)];
}

__global__ void update_dPrdtau_kernel(const double* __restrict__ Pr, double* __restrict__ dPrdtau,
                                      const double* __restrict__ divV, double rho, double dt, double dtau,
                                      double damp, double dx, double dy, double dz, int nx, int ny, int nz)
{
    THREAD_IJK();
    if (i >= nx - 2 || j >= ny - 2 || k >= nz - 2) return;
    double* d = &dPrdtau[idx3(i, j, k, nx - 2, ny - 2)];
    *d = *d * (1.0 - damp) + dtau * pt_bracket_l1(Pr, divV, rho / dt, dx, dy, dz, nx, ny, i, j, k);
}

extern "C" int ns3d_update_dPrdtau(ns3d_ctx* ctx, const double* Pr, double* dPrdtau, const double* divV, double rho,
                                   double dt, double dtau, double damp, double dx, double dy, double dz, int nx,
                                   int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    update_dPrdtau_kernel<<<grid3(nx - 2, ny - 2, nz - 2), block3(), 0, ctx->stream>>>(Pr, dPrdtau, divV, rho, dt, dtau,
                                                                                      damp, dx, dy, dz, nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

__global__ void update_Pr_kernel(double* __restrict__ Pr, const double* __restrict__ dPrdtau, double dtau, int nx,
                                 int ny, int nz)
{
    THREAD_IJK();
    if (i >= nx - 2 || j >= ny - 2 || k >= nz - 2) return;
    double* p = &Pr[idx3(i + 1, j + 1, k + 1, nx, ny)];
    *p = *p + dtau * dPrdtau[idx3(i, j, k, nx - 2, ny - 2)];
}

extern "C" int ns3d_update_Pr(ns3d_ctx* ctx, double* Pr, const double* dPrdtau, double dtau, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    update_Pr_kernel<<<grid3(nx - 2, ny - 2, nz - 2), block3(), 0, ctx->stream>>>(Pr, dPrdtau, dtau, nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

__global__ void compute_res_kernel(double* __restrict__ Rp, const double* __restrict__ Pr,
                                   const double* __restrict__ divV, double rho, double dt, double dx, double dy,
                                   double dz, int nx, int ny, int nz)
{
    THREAD_IJK();
    if (i >= nx - 2 || j >= ny - 2 || k >= nz - 2) return;
    Rp[idx3(i, j, k, nx - 2, ny - 2)] = pt_bracket_l1(Pr, divV, rho / dt, dx, dy, dz, nx, ny, i, j, k);
}

extern "C" int ns3d_compute_res(ns3d_ctx* ctx, double* Rp, const double* Pr, const double* divV, double rho,
                                double dt, double dx, double dy, double dz, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    compute_res_kernel<<<grid3(nx - 2, ny - 2, nz - 2), block3(), 0, ctx->stream>>>(Rp, Pr, divV, rho, dt, dx, dy, dz,
                                                                                   nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// ---------------------------------------------------------------------------------------------
// K9 correct_V!  M:97-102
// ---------------------------------------------------------------------------------------------
__global__ void correct_V_kernel(double* __restrict__ Vx, double* __restrict__ Vy, double* __restrict__ Vz,
                                 const double* __restrict__ Pr, double dt, double rho, double dx, double dy,
                                 double dz, int nx, int ny, int nz)
{
    THREAD_IJK();
    const double dtr = dt / rho;
    if (i >= nx - 1 || j >= ny - 1 || k >= nz - 1) return;
    const double p111 = Pr[idx3(i + 1, j + 1, k + 1, nx, ny)];  // in bounds for every thread that got here
    if (j < ny - 2 && k < nz - 2) {
        double* v = &VXI(i + 1, j + 1, k + 1);
        *v = *v - dtr * (p111 - Pr[idx3(i, j + 1, k + 1, nx, ny)]) / dx;
    }
    if (i < nx - 2 && k < nz - 2) {
        double* v = &VYI(i + 1, j + 1, k + 1);
        *v = *v - dtr * (p111 - Pr[idx3(i + 1, j, k + 1, nx, ny)]) / dy;
    }
    if (i < nx - 2 && j < ny - 2) {
        double* v = &VZI(i + 1, j + 1, k + 1);
        *v = *v - dtr * (p111 - Pr[idx3(i + 1, j + 1, k, nx, ny)]) / dz;
    }
}

extern "C" int ns3d_correct_V(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, const double* Pr, double dt,
                              double rho, double dx, double dy, double dz, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    correct_V_kernel<<<grid3(nx - 1, ny - 1, nz - 1), block3(), 0, ctx->stream>>>(Vx, Vy, Vz, Pr, dt, rho, dx, dy, dz, nx,
                                                                                 ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// ---------------------------------------------------------------------------------------------
// K7 boundary-condition kernels  M:108-150, G:239-261.  One launch covers a whole face pair.
// op: 0 bc_x!  1 bc_y!  2 bc_z!  3 bc_x_Vx! (lo=v0)  4 bc_x_Pr! (hi=v0)  5 bc_zV!
//     6 bc_xhydstatic! (v0 = rho*g, v1 = dz, n0 = nz argument)
// ---------------------------------------------------------------------------------------------
__global__ void bc_kernel(double* __restrict__ A, int sx, int sy, int sz, int op, double v0, double v1, int n0)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y * blockDim.y + threadIdx.y;
    switch (op) {
        case 0:  // (a,b) = (iy,iz)
            if (a < sy && b < sz) {
                A[idx3(0, a, b, sx, sy)] = A[idx3(1, a, b, sx, sy)];
                A[idx3(sx - 1, a, b, sx, sy)] = A[idx3(sx - 2, a, b, sx, sy)];
            }
            break;
        case 1:  // (a,b) = (ix,iz)
            if (a < sx && b < sz) {
                A[idx3(a, 0, b, sx, sy)] = A[idx3(a, 1, b, sx, sy)];
                A[idx3(a, sy - 1, b, sx, sy)] = A[idx3(a, sy - 2, b, sx, sy)];
            }
            break;
        case 2:  // (a,b) = (ix,iy)
            if (a < sx && b < sy) {
                A[idx3(a, b, 0, sx, sy)] = A[idx3(a, b, 1, sx, sy)];
                A[idx3(a, b, sz - 1, sx, sy)] = A[idx3(a, b, sz - 2, sx, sy)];
            }
            break;
        case 3:
            if (a < sy && b < sz) A[idx3(0, a, b, sx, sy)] = v0;
            break;
        case 4:
            if (a < sy && b < sz) A[idx3(sx - 1, a, b, sx, sy)] = v0;
            break;
        case 5:
            if (a < sx && b < sy) {
                A[idx3(a, b, 0, sx, sy)] = 0.0;
                A[idx3(a, b, sz - 1, sx, sy)] = A[idx3(a, b, sz - 2, sx, sy)];
            }
            break;
        case 6:
            if (a < sy && b < sz) {
                // ρ*g*(nz-iz + 0.5)*dz with 1-based iz = b+1
                const double h = v0 * ((double)(n0 - (b + 1)) + 0.5) * v1;
                A[idx3(0, a, b, sx, sy)] = h + 100;
                A[idx3(sx - 1, a, b, sx, sy)] = h;
            }
            break;
    }
}

static int launch_bc(ns3d_ctx* ctx, double* A, int sx, int sy, int sz, int op, double v0 = 0.0, double v1 = 0.0,
                     int n0 = 0)
{
    if (!A || sx < 3 || sy < 3 || sz < 3) return ns3d_fail(ctx, NS3D_EINVAL, "bc: bad array (%d,%d,%d)", sx, sy, sz);
    int na, nb;
    if (op == 0 || op == 3 || op == 4 || op == 6) { na = sy; nb = sz; }
    else if (op == 1) { na = sx; nb = sz; }
    else { na = sx; nb = sy; }
    // x-face ops walk a strided plane (thread a = iy): keep a on threadIdx.x anyway, the planes are tiny.
    dim3 blk(32, 8);
    bc_kernel<<<dim3(cdiv(na, 32), cdiv(nb, 8)), blk, 0, ctx->stream>>>(A, sx, sy, sz, op, v0, v1, n0);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

#define BC_ENTRY(name, op)                                                         \
    extern "C" int name(ns3d_ctx* ctx, double* A, int sx, int sy, int sz)          \
    {                                                                              \
        NS3D_CHECK_CTX(ctx);                                                       \
        NS3D_CUDA(ctx, cudaSetDevice(ctx->device));                                \
        return launch_bc(ctx, A, sx, sy, sz, op);                                  \
    }
BC_ENTRY(ns3d_bc_x, 0)
BC_ENTRY(ns3d_bc_y, 1)
BC_ENTRY(ns3d_bc_z, 2)
BC_ENTRY(ns3d_bc_zV, 5)

extern "C" int ns3d_bc_x_Vx(ns3d_ctx* ctx, double* A, double V, int sx, int sy, int sz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_bc(ctx, A, sx, sy, sz, 3, V);
}
extern "C" int ns3d_bc_x_Pr(ns3d_ctx* ctx, double* A, double val, int sx, int sy, int sz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_bc(ctx, A, sx, sy, sz, 4, val);
}
extern "C" int ns3d_bc_xhydstatic(ns3d_ctx* ctx, double* A, double dz, int nz_arg, double g, double rho, int sx,
                                  int sy, int sz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_bc(ctx, A, sx, sy, sz, 6, rho * g, dz, nz_arg);
}

// set_bc_Vel!  M:156-169
extern "C" int ns3d_set_bc_Vel_M(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, int inlet_guard, double vin,
                                 int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_TRY(launch_bc(ctx, Vx, nx + 1, ny, nz, 0));
    NS3D_TRY(launch_bc(ctx, Vx, nx + 1, ny, nz, 1));
    NS3D_TRY(launch_bc(ctx, Vx, nx + 1, ny, nz, 2));
    NS3D_TRY(launch_bc(ctx, Vy, nx, ny + 1, nz, 0));
    NS3D_TRY(launch_bc(ctx, Vy, nx, ny + 1, nz, 2));
    NS3D_TRY(launch_bc(ctx, Vz, nx, ny, nz + 1, 0));
    NS3D_TRY(launch_bc(ctx, Vz, nx, ny, nz + 1, 1));
    if (inlet_guard) NS3D_TRY(launch_bc(ctx, Vx, nx + 1, ny, nz, 3, vin));
    double* fields[3] = {Vx, Vy, Vz};  // update_halo!(Vx,Vy,Vz) M:167
    const int sx[3] = {nx + 1, nx, nx}, sy[3] = {ny, ny + 1, ny}, sz[3] = {nz, nz, nz + 1};
    return ns3d_internal_halo_z(ctx, ctx->stream, fields, sx, sy, sz, 3, nz);
}

// set_bc_Vel!  G:264-279
extern "C" int ns3d_set_bc_Vel_G(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    double* f[3] = {Vx, Vy, Vz};
    const int sx[3] = {nx + 1, nx, nx}, sy[3] = {ny, ny + 1, ny}, sz[3] = {nz, nz, nz + 1};
    for (int q = 0; q < 3; ++q) {
        NS3D_TRY(launch_bc(ctx, f[q], sx[q], sy[q], sz[q], 0));
        NS3D_TRY(launch_bc(ctx, f[q], sx[q], sy[q], sz[q], 1));
        NS3D_TRY(launch_bc(ctx, f[q], sx[q], sy[q], sz[q], 5));
    }
    return NS3D_OK;
}

// set_bc_Pr!  M:175-184
extern "C" int ns3d_set_bc_Pr_M(ns3d_ctx* ctx, double* Pr, int outlet_guard, double val, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_TRY(launch_bc(ctx, Pr, nx, ny, nz, 0));
    NS3D_TRY(launch_bc(ctx, Pr, nx, ny, nz, 1));
    NS3D_TRY(launch_bc(ctx, Pr, nx, ny, nz, 2));
    if (outlet_guard) NS3D_TRY(launch_bc(ctx, Pr, nx, ny, nz, 4, val));
    double* fields[1] = {Pr};  // update_halo!(Pr) M:182
    return ns3d_internal_halo_z(ctx, ctx->stream, fields, &nx, &ny, &nz, 1, nz);
}

// set_bc_Pr!  G:281-286
extern "C" int ns3d_set_bc_Pr_G(ns3d_ctx* ctx, double* Pr, double dz, int nz_arg, double g, double rho, int nx,
                                int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_TRY(launch_bc(ctx, Pr, nx, ny, nz, 1));
    NS3D_TRY(launch_bc(ctx, Pr, nx, ny, nz, 2));
    return launch_bc(ctx, Pr, nx, ny, nz, 6, rho * g, dz, nz_arg);
}

// ---------------------------------------------------------------------------------------------
// K11 advect! + backtrack! + lerp  M:190-243
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double lerp_(double a, double b, double t) { return b * t + a * (1 - t); }  // M:211

__device__ __forceinline__ long long clampll(long long v, long long lo, long long hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

// backtrack!  M:190-205.  (ix,iy,iz) are the reference's 1-based indices, (sx,sy,sz)=size(A).
__device__ __forceinline__ double backtrack(const double* __restrict__ Ao, double vxc, double vyc, double vzc,
                                            double dt, double dx, double dy, double dz, int ix, int iy, int iz,
                                            int sx, int sy, int sz)
{
    double ddx = dt * vxc / dx, ddy = dt * vyc / dy, ddz = dt * vzc / dz;
    const long long ix1 = clampll(__double2ll_rd((double)ix - ddx), 1, sx);  // floor(Int, ix-δx)
    const long long iy1 = clampll(__double2ll_rd((double)iy - ddy), 1, sy);
    const long long iz1 = clampll(__double2ll_rd((double)iz - ddz), 1, sz);
    const long long ix2 = clampll(ix1 + 1, 1, sx), iy2 = clampll(iy1 + 1, 1, sy), iz2 = clampll(iz1 + 1, 1, sz);
    ddx = (ddx > 0 ? 1.0 : 0.0) - fmod(ddx, 1.0);
    ddy = (ddy > 0 ? 1.0 : 0.0) - fmod(ddy, 1.0);
    ddz = (ddz > 0 ? 1.0 : 0.0) - fmod(ddz, 1.0);
#define AO(i, j, k) Ao[idx3((int)(i)-1, (int)(j)-1, (int)(k)-1, sx, sy)]
    const double fy1z1 = lerp_(AO(ix1, iy1, iz1), AO(ix2, iy1, iz1), ddx);
    const double fy1z2 = lerp_(AO(ix1, iy1, iz2), AO(ix2, iy1, iz2), ddx);
    const double fy2z1 = lerp_(AO(ix1, iy2, iz1), AO(ix2, iy2, iz1), ddx);
    const double fy2z2 = lerp_(AO(ix1, iy2, iz2), AO(ix2, iy2, iz2), ddx);
#undef AO
    const double fz1 = lerp_(fy1z1, fy2z1, ddy);
    const double fz2 = lerp_(fy1z2, fy2z2, ddy);
    return lerp_(fz1, fz2, ddz);
}

#define VXO(i, j, k) Vx_o[idx3((i)-1, (j)-1, (k)-1, nx + 1, ny)]
#define VYO(i, j, k) Vy_o[idx3((i)-1, (j)-1, (k)-1, nx, ny + 1)]
#define VZO(i, j, k) Vz_o[idx3((i)-1, (j)-1, (k)-1, nx, ny)]

// One thread per (ix,iy,iz) of the (nx+1,ny+1,nz+1) launch range, like the reference.  The
// third branch targets Vy/Vy_o (M:234, sic): Vz is never advected, and where branches 2 and 3
// both fire the thread's second store to Vy[ix,iy,iz] wins -- here only the winner is stored.
// ALL (ns3d_step): the kernel writes EVERY entry of Vx, Vy, Vz -- where advect! leaves an array alone
// (the x-face planes of Vx, the rows of Vy no branch reaches, all of Vz) it copies the `_o` value, so
// that a step whose corrector worked on the `_o` arrays needs no `A_o .= A` copies (M:475).
template <bool ALL>
__global__ void advect_kernel(double* __restrict__ Vx, const double* __restrict__ Vx_o, double* __restrict__ Vy,
                              const double* __restrict__ Vy_o, double* __restrict__ Vz, const double* __restrict__ Vz_o,
                              double* __restrict__ C, const double* __restrict__ C_o, double dt, double dx,
                              double dy, double dz, int nx, int ny, int nz)
{
    const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int iz = blockIdx.z * blockDim.z + threadIdx.z + 1;
    if (ix > nx + 1 || iy > ny + 1 || iz > nz + 1) return;
    if (ix > 1 && ix < nx + 1 && iy <= ny && iz <= nz) {
        const double vxc = VXO(ix, iy, iz);
        const double vyc = 0.25 * (((VYO(ix - 1, iy, iz) + VYO(ix - 1, iy + 1, iz)) + VYO(ix, iy, iz)) + VYO(ix, iy + 1, iz));
        const double vzc = 0.25 * (((VZO(ix - 1, iy, iz) + VZO(ix - 1, iy, iz + 1)) + VZO(ix, iy, iz)) + VZO(ix, iy, iz + 1));
        Vx[idx3(ix - 1, iy - 1, iz - 1, nx + 1, ny)] = backtrack(Vx_o, vxc, vyc, vzc, dt, dx, dy, dz, ix, iy, iz, nx + 1, ny, nz);
    } else if (ALL && iy <= ny && iz <= nz) {
        Vx[idx3(ix - 1, iy - 1, iz - 1, nx + 1, ny)] = VXO(ix, iy, iz);
    }
    const bool b3 = iz > 1 && iz < nz + 1 && ix <= nx && iy <= ny;
    const bool b2 = iy > 1 && iy < ny + 1 && ix <= nx && iz <= nz;
    if (b3) {
        const double vxc = 0.25 * (((VXO(ix, iy, iz - 1) + VXO(ix + 1, iy, iz - 1)) + VXO(ix, iy, iz)) + VXO(ix + 1, iy, iz));
        const double vyc = 0.25 * (((VYO(ix, iy, iz - 1) + VYO(ix, iy + 1, iz - 1)) + VYO(ix, iy, iz)) + VYO(ix, iy + 1, iz));
        const double vzc = VZO(ix, iy, iz);
        Vy[idx3(ix - 1, iy - 1, iz - 1, nx, ny + 1)] = backtrack(Vy_o, vxc, vyc, vzc, dt, dx, dy, dz, ix, iy, iz, nx, ny + 1, nz);
    } else if (b2) {
        const double vxc = 0.25 * (((VXO(ix, iy - 1, iz) + VXO(ix + 1, iy - 1, iz)) + VXO(ix, iy, iz)) + VXO(ix + 1, iy, iz));
        const double vyc = VYO(ix, iy, iz);
        const double vzc = 0.25 * (((VZO(ix, iy - 1, iz) + VZO(ix, iy - 1, iz + 1)) + VZO(ix, iy, iz)) + VZO(ix, iy, iz + 1));
        Vy[idx3(ix - 1, iy - 1, iz - 1, nx, ny + 1)] = backtrack(Vy_o, vxc, vyc, vzc, dt, dx, dy, dz, ix, iy, iz, nx, ny + 1, nz);
    } else if (ALL && ix <= nx && iz <= nz) {
        Vy[idx3(ix - 1, iy - 1, iz - 1, nx, ny + 1)] = VYO(ix, iy, iz);
    }
    if (ALL && ix <= nx && iy <= ny) Vz[idx3(ix - 1, iy - 1, iz - 1, nx, ny)] = VZO(ix, iy, iz);
    if (ix <= nx && iy <= ny && iz <= nz) {
        const double vxc = 0.5 * (VXO(ix, iy, iz) + VXO(ix + 1, iy, iz));
        const double vyc = 0.5 * (VYO(ix, iy, iz) + VYO(ix, iy + 1, iz));
        const double vzc = 0.5 * (VZO(ix, iy, iz) + VZO(ix, iy, iz + 1));
        C[idx3(ix - 1, iy - 1, iz - 1, nx, ny)] = backtrack(C_o, vxc, vyc, vzc, dt, dx, dy, dz, ix, iy, iz, nx, ny, nz);
    }
}

extern "C" int ns3d_advect(ns3d_ctx* ctx, double* Vx, const double* Vx_o, double* Vy, const double* Vy_o,
                           double* Vz, const double* Vz_o, double* C, const double* C_o, double dt, double dx,
                           double dy, double dz, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    // Vz is never written by the reference (M:234)
    advect_kernel<false><<<grid3(nx + 1, ny + 1, nz + 1), block3(), 0, ctx->stream>>>(Vx, Vx_o, Vy, Vy_o, Vz, Vz_o, C, C_o, dt, dx,
                                                                                     dy, dz, nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// advect! that also carries the untouched entries over from the `_o` arrays (see advect_kernel<ALL>): ns3d_step.
int ns3d_internal_advect_all(ns3d_ctx* ctx, double* Vx, const double* Vx_o, double* Vy, const double* Vy_o, double* Vz,
                             const double* Vz_o, double* C, const double* C_o, double dt, double dx, double dy, double dz, int nx,
                             int ny, int nz)
{
    advect_kernel<true><<<grid3(nx + 1, ny + 1, nz + 1), block3(), 0, ctx->stream>>>(Vx, Vx_o, Vy, Vy_o, Vz, Vz_o, C, C_o, dt, dx, dy,
                                                                                    dz, nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// ---------------------------------------------------------------------------------------------
// K3 set_cylinder!  M:249-281 / G:336-368
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_ellipse(double X, double Y, double ox, double oy, double sinb, double cosb,
                                           double a2, double b2, double thr)
{
    const double xr = (X - ox) * cosb - (Y - oy) * sinb;
    const double yr = (X - ox) * sinb + (Y - oy) * cosb;
    return xr * xr / a2 + yr * yr / b2 < thr;
}

__global__ void set_cylinder_kernel(int variant, double* __restrict__ C, double* __restrict__ Vx,
                                    double* __restrict__ Vy, double* __restrict__ Vz, double a2, double b2,
                                    double ox, double oy, double sinb, double cosb, double xco_g, double yco_g,
                                    double lx, double ly, double dx, double dy, int nx, int ny, int nz)
{
    const int ix = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int iz = blockIdx.z * blockDim.z + threadIdx.z + 1;
    if (ix > nx + 1 || iy > ny + 1 || iz > nz + 1) return;
    double xc, yc, xv, yv;
    if (variant == NS3D_VARIANT_M) {  // M:250-251
        xc = xco_g + (ix - 1) * dx;
        yc = yco_g + (iy - 1) * dy;
        xv = xc - dx / 2;
        yv = yc - dy / 2;
    } else {  // G:337-338 (yc uses dx, sic)
        xv = (ix - 1) * dx - lx / 2;
        yv = (iy - 1) * dy - ly / 2;
        xc = xv + dx / 2;
        yc = yv + dx / 2;
    }
    if (ix <= nx && iy <= ny && iz <= nz)
        if (in_ellipse(xc, yc, ox, oy, sinb, cosb, a2, b2, 1.05)) C[idx3(ix - 1, iy - 1, iz - 1, nx, ny)] = 1.0;
    if (iy <= ny && iz <= nz)
        if (in_ellipse(xv, yc, ox, oy, sinb, cosb, a2, b2, 1.0)) Vx[idx3(ix - 1, iy - 1, iz - 1, nx + 1, ny)] = 0.0;
    if (ix <= nx && iz <= nz)
        if (in_ellipse(xc, yv, ox, oy, sinb, cosb, a2, b2, 1.0)) Vy[idx3(ix - 1, iy - 1, iz - 1, nx, ny + 1)] = 0.0;
    if (ix <= nx && iy <= ny)
        if (in_ellipse(xc, yc, ox, oy, sinb, cosb, a2, b2, 1.0)) Vz[idx3(ix - 1, iy - 1, iz - 1, nx, ny)] = 0.0;
}

extern "C" int ns3d_set_cylinder_M(ns3d_ctx* ctx, double* C, double* Vx, double* Vy, double* Vz, double a2,
                                   double b2, double ox, double oy, double sinb, double cosb, double xco_g,
                                   double yco_g, double dx, double dy, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    set_cylinder_kernel<<<grid3(nx + 1, ny + 1, nz + 1), block3(), 0, ctx->stream>>>(
        NS3D_VARIANT_M, C, Vx, Vy, Vz, a2, b2, ox, oy, sinb, cosb, xco_g, yco_g, 0.0, 0.0, dx, dy, nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

extern "C" int ns3d_set_cylinder_G(ns3d_ctx* ctx, double* C, double* Vx, double* Vy, double* Vz, double a2,
                                   double b2, double ox, double oy, double sinb, double cosb, double lx,
                                   double ly, double dx, double dy, int nx, int ny, int nz)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    set_cylinder_kernel<<<grid3(nx + 1, ny + 1, nz + 1), block3(), 0, ctx->stream>>>(
        NS3D_VARIANT_G, C, Vx, Vy, Vz, a2, b2, ox, oy, sinb, cosb, 0.0, 0.0, lx, ly, dx, dy, nx, ny, nz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}
