// libns3d.so -- level 2: the once-per-step part of a time step (everything around the PT loop) in fused kernels.
//
// Reference, per step (M:449-455, 472-477 / G:121-124, 138-142): update_τ! (9 field passes), predict_V! (12),
// set_cylinder!, update_∇V! (4), ... correct_V! (7), set_cylinder!, set_bc_Vel! (7-9 face kernels), four full-field
// copies `A_o .= A` (8), advect! (7): 47 field passes and six stress arrays that exist only between two kernels.
//
// Here (ns3d_step; 27 passes, no stress arrays, no copies):
//   predictor_kernel   Vx,Vy,Vz -> Vx_o,Vy_o,Vz_o   update_τ! + predict_V! + set_cylinder! in one z-march: the stresses of a
//                                                   plane live in a shared-memory ring (in-plane neighbours) and registers
//                                                   (z neighbours), each computed once per tile -- 19 IEEE divisions per cell
//                                                   like the reference, not the 43 a per-point recomputation costs.  The
//                                                   predicted velocity goes to the `_o` arrays, which hold nothing at this
//                                                   point of a step (M:475 overwrites them before their next use): a fused
//                                                   predictor cannot update V in place, other tiles still read the old one
//   update_divV_kernel Vx_o,Vy_o,Vz_o -> ∇V         (level 1, ns3d_ops.cu)
//   [PT loop]
//   corrector_kernel   Pr, V_o, C -> V_o, C, C_o    correct_V! + set_cylinder! in place (point-wise in V) + the snapshot C_o .= C
//   bc_kernel x 7-9    V_o faces                    set_bc_Vel! (level 1)
//   advect_all_kernel  V_o, C_o -> Vx,Vy,Vz,C       advect! writing EVERY entry: the advected ones, and copies of V_o where
//                                                   advect! leaves the array alone (its face planes; all of Vz, M:234) -- so
//                                                   `V_o .= V` has happened by construction and no copy is left
// On return every array of the reference holds what the reference's step leaves in it (V, C, Pr, dPrdτ, ∇V, V_o = the
// pre-advection snapshot, C_o) except the six stress arrays, which are not touched (they may be NULL).
//
// All arithmetic is the level-1 kernels' (IEEE division, no FMA: the library is compiled with --fmad=false), expression for
// expression, so the results are bit-equal to the level-1 sequence and to the oracle (tests/test_gpu_solver.py,
// tests/test_gpu_zz_output.py::test_step_groups_equal_the_fused_step).
#include <algorithm>

#include "ns3d_internal.cuh"

namespace {

#define VXI(i, j, k) Vx[idx3(i, j, k, nx + 1, ny)]
#define VYI(i, j, k) Vy[idx3(i, j, k, nx, ny + 1)]
#define VZI(i, j, k) Vz[idx3(i, j, k, nx, ny)]

__device__ __forceinline__ bool in_ellipse_s(double X, double Y, double ox, double oy, double sinb, double cosb, double a2, double b2,
                                             double thr)
{
    const double xr = (X - ox) * cosb - (Y - oy) * sinb;
    const double yr = (X - ox) * sinb + (Y - oy) * cosb;
    return xr * xr / a2 + yr * yr / b2 < thr;
}

struct CylArgs {
    int variant;
    double a2, b2, ox, oy, sinb, cosb, xco_g, yco_g, lx, ly, dx, dy;
};

// set_cylinder! (M:249-281 / G:336-368) for the point (ix,iy) (1-based, z-independent): which of C, Vx, Vy, Vz it sets.
struct CylMask {
    bool c, vx, vy, vz;
};
__device__ __forceinline__ CylMask cyl_mask(const CylArgs& a, int ix, int iy)
{
    double xc, yc, xv, yv;
    if (a.variant == NS3D_VARIANT_M) {  // M:250-251
        xc = a.xco_g + (ix - 1) * a.dx;
        yc = a.yco_g + (iy - 1) * a.dy;
        xv = xc - a.dx / 2;
        yv = yc - a.dy / 2;
    } else {  // G:337-338 (yc uses dx, sic)
        xv = (ix - 1) * a.dx - a.lx / 2;
        yv = (iy - 1) * a.dy - a.ly / 2;
        xc = xv + a.dx / 2;
        yc = yv + a.dx / 2;
    }
    CylMask m;
    m.c = in_ellipse_s(xc, yc, a.ox, a.oy, a.sinb, a.cosb, a.a2, a.b2, 1.05);
    m.vx = in_ellipse_s(xv, yc, a.ox, a.oy, a.sinb, a.cosb, a.a2, a.b2, 1.0);
    m.vy = in_ellipse_s(xc, yv, a.ox, a.oy, a.sinb, a.cosb, a.a2, a.b2, 1.0);
    m.vz = in_ellipse_s(xc, yc, a.ox, a.oy, a.sinb, a.cosb, a.a2, a.b2, 1.0);
    return m;
}

// ---- predictor_kernel ---------------------------------------------------------------------------------------------------------
// Thread (tx,ty) of a PT_X x PT_Y tile owns the column (i,j) = (X0+tx, Y0+ty) of the (nx+1, ny+1) index range and marches
// along z.  At step k it computes the stresses with index (i,j,k) -- normal ones at the cell, shear ones at the edge
// (M:37-43) -- publishes them in slot k mod 3 of the ring, and after the barrier updates the velocity points with index
// (i,j,k): M:51-53 read the stresses of the points (i-1,j,k), (i,j-1,k), (i-1,j-1,k-1), (i,j-1,k-1), (i-1,j,k-1) and
// (i-1,j-1,k) -- the low-side neighbours only, so a tile puts out every column but its first row and first column (which
// the neighbouring tile puts out; at the domain's low faces they are boundary points that predict_V! does not touch).
constexpr int PT_X = 32, PT_Y = 8;
enum { R_TXX = 0, R_TYY, R_TXY, R_TXZ, R_TYZ, R_N };

__global__ void __launch_bounds__(PT_X* PT_Y) predictor_kernel(double* __restrict__ Vxn, double* __restrict__ Vyn, double* __restrict__ Vzn,
                                                               double* __restrict__ C, const double* __restrict__ Vx,
                                                               const double* __restrict__ Vy, const double* __restrict__ Vz, double mu,
                                                               double rho, double g, double dt, double dx, double dy, double dz, int nx,
                                                               int ny, int nz, int zchunk, const CylArgs cyl)
{
    __shared__ double ring[3][R_N][PT_Y][PT_X];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int X0 = blockIdx.x * (PT_X - 1), Y0 = blockIdx.y * (PT_Y - 1);
    const int i = X0 + tx, j = Y0 + ty;
    const int kb = blockIdx.z * zchunk, ke = min(kb + zchunk, nz + 1);   // output planes [kb, ke) of the range 0..nz
    const bool inx = i <= nx, iny = j <= ny;
    const bool out = inx && iny && (tx > 0 || X0 == 0) && (ty > 0 || Y0 == 0);   // this column's points are put out here
    const CylMask cm = (inx && iny) ? cyl_mask(cyl, i + 1, j + 1) : CylMask{false, false, false, false};
    const double twomu = 2 * mu, dtr = dt / rho;
    const bool cell = i < nx && j < ny;              // (i,j) is a cell column
    const bool edge = i < nx - 1 && j < ny - 1;      // ... a shear-stress column
    double tzz_prev = 0.0;                           // τzz(i,j,k-1)
    for (int k = max(kb - 1, 0); k < ke; ++k) {
        const int s = k % 3;
        // ---- the stresses with index (i,j,k): update_τ!  M:37-43 ----
        double txx = 0.0, tyy = 0.0, tzz = 0.0, txy = 0.0, txz = 0.0, tyz = 0.0;
        if (cell && k < nz) {
            const double dxa = VXI(i + 1, j, k) - VXI(i, j, k);
            const double dya = VYI(i, j + 1, k) - VYI(i, j, k);
            const double dza = VZI(i, j, k + 1) - VZI(i, j, k);
            const double divv = (dxa / dx + dya / dy) + dza / dz;
            txx = twomu * (dxa / dx - divv / 3.0);
            tyy = twomu * (dya / dy - divv / 3.0);
            tzz = twomu * (dza / dz - divv / 3.0);
        }
        if (edge && k < nz - 1) {
            const double vx111 = VXI(i + 1, j + 1, k + 1), vy111 = VYI(i + 1, j + 1, k + 1), vz111 = VZI(i + 1, j + 1, k + 1);
            const double dyiVx = vx111 - VXI(i + 1, j, k + 1);
            const double dxiVy = vy111 - VYI(i, j + 1, k + 1);
            const double dziVx = vx111 - VXI(i + 1, j + 1, k);
            const double dxiVz = vz111 - VZI(i, j + 1, k + 1);
            const double dziVy = vy111 - VYI(i + 1, j + 1, k);
            const double dyiVz = vz111 - VZI(i + 1, j, k + 1);
            txy = mu * (dyiVx / dy + dxiVy / dx);
            txz = mu * (dziVx / dz + dxiVz / dx);
            tyz = mu * (dziVy / dz + dyiVz / dy);
        }
        ring[s][R_TXX][ty][tx] = txx;
        ring[s][R_TYY][ty][tx] = tyy;
        ring[s][R_TXY][ty][tx] = txy;
        ring[s][R_TXZ][ty][tx] = txz;
        ring[s][R_TYZ][ty][tx] = tyz;
        __syncthreads();   // slot s is complete; slot (k+1) mod 3, written next, was last read two steps ago
        if (k >= kb && out) {
            const int sp = (k + 2) % 3;   // slot of plane k-1
            // ---- Vx(i,j,k): predict_V! M:51 on faces 1..nx-1 x rows 1..ny-2 x planes 1..nz-2, else unchanged ----
            if (j < ny && k < nz) {
                double v = VXI(i, j, k);
                if (i >= 1 && i <= nx - 1 && j >= 1 && j <= ny - 2 && k >= 1 && k <= nz - 2) {
                    const double a = txx - ring[s][R_TXX][ty][tx - 1];
                    const double b = ring[sp][R_TXY][ty][tx - 1] - ring[sp][R_TXY][ty - 1][tx - 1];
                    const double c = ring[s][R_TXZ][ty - 1][tx - 1] - ring[sp][R_TXZ][ty - 1][tx - 1];
                    v = v + dtr * ((a / dx + b / dy) + c / dz);
                }
                if (cm.vx) v = 0.0;   // set_cylinder! M:452
                Vxn[idx3(i, j, k, nx + 1, ny)] = v;
            }
            if (i < nx && k < nz) {   // Vy(i,j,k): M:52
                double v = VYI(i, j, k);
                if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 1 && k >= 1 && k <= nz - 2) {
                    const double a = tyy - ring[s][R_TYY][ty - 1][tx];
                    const double b = ring[sp][R_TXY][ty - 1][tx] - ring[sp][R_TXY][ty - 1][tx - 1];
                    const double c = ring[s][R_TYZ][ty - 1][tx - 1] - ring[sp][R_TYZ][ty - 1][tx - 1];
                    v = v + dtr * ((a / dy + b / dx) + c / dz);
                }
                if (cm.vy) v = 0.0;
                Vyn[idx3(i, j, k, nx, ny + 1)] = v;
            }
            if (i < nx && j < ny) {   // Vz(i,j,k): M:53
                double v = VZI(i, j, k);
                if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2 && k >= 1 && k <= nz - 1) {
                    const double a = tzz - tzz_prev;
                    const double b = ring[sp][R_TXZ][ty - 1][tx] - ring[sp][R_TXZ][ty - 1][tx - 1];
                    const double c = ring[sp][R_TYZ][ty][tx - 1] - ring[sp][R_TYZ][ty - 1][tx - 1];
                    v = v + dtr * (((a / dz + b / dx) + c / dy) - rho * g);
                }
                if (cm.vz) v = 0.0;
                Vzn[idx3(i, j, k, nx, ny)] = v;
                if (k < nz && cm.c) C[idx3(i, j, k, nx, ny)] = 1.0;
            }
        }
        tzz_prev = tzz;
    }
}

// ---- corrector_kernel: correct_V! (M:97-102) + set_cylinder! (M:473), in place on V (every operation is point-wise in V),
// and optionally the snapshot C_o .= C of M:475.  One thread per point of the (nx+1, ny+1, nz+1) range. ----
__global__ void __launch_bounds__(256) corrector_kernel(double* __restrict__ Vx, double* __restrict__ Vy, double* __restrict__ Vz,
                                                        double* __restrict__ C, double* __restrict__ C_o, const double* __restrict__ Pr,
                                                        double dt, double rho, double dx, double dy, double dz, int nx, int ny, int nz,
                                                        const CylArgs cyl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i > nx || j > ny || k > nz) return;
    const CylMask cm = cyl_mask(cyl, i + 1, j + 1);
    const double dtr = dt / rho;
    const bool cell = i < nx && j < ny && k < nz;
    const double p = cell ? Pr[idx3(i, j, k, nx, ny)] : 0.0;
    if (j < ny && k < nz) {   // Vx(i,j,k)
        double* v = &VXI(i, j, k);
        if (cm.vx) {
            *v = 0.0;
        } else if (i >= 1 && i <= nx - 1 && j >= 1 && j <= ny - 2 && k >= 1 && k <= nz - 2) {
            *v = *v - dtr * (p - Pr[idx3(i - 1, j, k, nx, ny)]) / dx;
        }
    }
    if (i < nx && k < nz) {   // Vy(i,j,k)
        double* v = &VYI(i, j, k);
        if (cm.vy) {
            *v = 0.0;
        } else if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 1 && k >= 1 && k <= nz - 2) {
            *v = *v - dtr * (p - Pr[idx3(i, j - 1, k, nx, ny)]) / dy;
        }
    }
    if (i < nx && j < ny) {   // Vz(i,j,k)
        double* v = &VZI(i, j, k);
        if (cm.vz) {
            *v = 0.0;
        } else if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2 && k >= 1 && k <= nz - 1) {
            *v = *v - dtr * (p - Pr[idx3(i, j, k - 1, nx, ny)]) / dz;
        }
    }
    if (cell) {
        const size_t c = idx3(i, j, k, nx, ny);
        const double cv = cm.c ? 1.0 : C[c];
        if (cm.c) C[c] = 1.0;
        if (C_o) C_o[c] = cv;
    }
}

}  // namespace

static CylArgs cyl_args(const ns3d_step_params* sp)
{
    CylArgs a;
    a.variant = sp->pt.variant;
    a.a2 = sp->a2; a.b2 = sp->b2; a.ox = sp->ox; a.oy = sp->oy; a.sinb = sp->sinb; a.cosb = sp->cosb;
    a.xco_g = sp->xco_g; a.yco_g = sp->yco_g; a.lx = sp->lx; a.ly = sp->ly; a.dx = sp->pt.dx; a.dy = sp->pt.dy;
    return a;
}

// update_τ! + predict_V! + set_cylinder! (M:449-452 / G:121-123): (Vx,Vy,Vz) -> (Vxn,Vyn,Vzn), C masked in place.
int ns3d_internal_predict_fused(ns3d_ctx* ctx, double* Vxn, double* Vyn, double* Vzn, double* C, const double* Vx, const double* Vy,
                                const double* Vz, const ns3d_step_params* sp)
{
    const ns3d_pt_params& p = sp->pt;
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    // z-chunks: enough CTAs for a few waves, long enough that the extra plane every chunk starts with stays cheap
    const long long tiles = (long long)cdiv(nx + 1, PT_X - 1) * cdiv(ny + 1, PT_Y - 1);
    long long nch = (6LL * ctx->num_sms * 4 + tiles - 1) / tiles;
    int zchunk = (int)std::max<long long>(8, (nz + 1 + nch - 1) / std::max<long long>(nch, 1));
    zchunk = std::min(zchunk, nz + 1);
    const dim3 grid(cdiv(nx + 1, PT_X - 1), cdiv(ny + 1, PT_Y - 1), cdiv(nz + 1, zchunk));
    predictor_kernel<<<grid, dim3(PT_X, PT_Y, 1), 0, ctx->stream>>>(Vxn, Vyn, Vzn, C, Vx, Vy, Vz, sp->mu, p.rho, p.g, p.dt, p.dx, p.dy,
                                                                    p.dz, nx, ny, nz, zchunk, cyl_args(sp));
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// correct_V! + set_cylinder! (M:472-473 / G:138-139) in place on (Vx,Vy,Vz), C; C_o != NULL: also C_o .= C (M:475).
int ns3d_internal_correct_fused(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, double* C, double* C_o, const double* Pr,
                                const ns3d_step_params* sp)
{
    const ns3d_pt_params& p = sp->pt;
    const dim3 block(32, 4, 2);
    const dim3 grid(cdiv(p.nx + 1, block.x), cdiv(p.ny + 1, block.y), cdiv(p.nz + 1, block.z));
    corrector_kernel<<<grid, block, 0, ctx->stream>>>(Vx, Vy, Vz, C, C_o, Pr, p.dt, p.rho, p.dx, p.dy, p.dz, p.nx, p.ny, p.nz,
                                                      cyl_args(sp));
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}
