// Device code of the fused pseudo-transient (PT) pressure iteration: the kernels of ns3d_pt.cu.
//
// Kept in a header free of CUDA runtime includes so that the SAME source is compiled twice:
// by nvcc into libns3d.so (ns3d_pt.cu), and by g++ behind tests/emu/cuda_host_shim.h, where every
// CUDA thread of a block runs as a host thread and __syncthreads() is a barrier -- the CPU test
// suite executes these kernels bit-for-bit against the oracle without a GPU
// (tests/test_kernel_emu.py).  The includer provides the CUDA qualifiers and ns3d_shared.cuh.
//
// What every kernel here computes per cell is the reference's K5 + K6 + K7 (M:70-82, 175-184):
//     dPrdτ' = dPrdτ*(1-damp) + dτ*(d2x/dx/dx + d2y/dy/dy + d2z/dz/dz - ρ/dt*∇V)     M:71
//     Pr'    = Pr + dτ*dPrdτ'                                                       M:80
// on the interior, then set_bc_Pr!: after the zero-gradient copies in the order x, y, z every
// boundary point equals the new value at its index clamped into the interior (SURVEY.md Appendix
// A), then the outlet plane is overwritten (variant M, bc_x_Pr!) or both x planes get the
// hydrostatic profile (variant G, bc_xhydstatic!) -- so the thread that owns an interior point next
// to a face also stores its mirror images, and no boundary kernel exists.  Pr (and, in the
// two-iteration kernels, dPrdτ) ping-pong between two buffers so that every thread reads a
// consistent old iterate.  Thread mapping: x on threadIdx.x (coalesced rows), one column per
// thread marched along z with the z neighbours in registers (2.5-D blocking).
#pragma once

#include "ns3d_pt_common.cuh"

namespace {

struct PtK {
    int nx, ny, nz;
    double omd;   // 1.0 - damp
    double dtau;
    double rdt;   // rho / dt
    double dx, dy, dz;
    double rdx, rdy, rdz;     // RN(1/dx) ...      (FAST)
    double rdx2, rdy2, rdz2;  // RN(1/(dx*dx)) ... (FASTEST)
    int xlo_kind, xhi_kind;   // X_*
    double xlo_val, xhi_val;  // Dirichlet value / hydrostatic offset (+100 at the inlet, G:258)
    double rho_g, hyd_dz;     // hydrostatic: ((rho*g)*((hyd_nz-iz)+0.5))*dz, iz 1-based (G:258-259)
    int hyd_nz;
    int zlo_halo, zhi_halo;   // z faces that are slab interfaces: left to the halo exchange
    int zchunk;
    // planes this launch updates: [kbeg, kend) in chunks of zchunk, or -- when `faces` is set --
    // only the two outermost interior planes 1 and nz-2 (the ones a slab sends to its neighbours)
    int kbeg, kend, faces;
    // serpentine sweep: odd iterations walk the z-chunks downwards, so an iteration starts on the
    // planes the previous one touched last and finds them in the 126 MB L2
    int reverse;
    int serpentine;  // host-side policy flag (not read by the kernel)
    // peer-memory halo exchange (pt_iter_kernel<.,.,true>): where the planes this slab sends go
    // in the neighbours' new iterate, the mailboxes, and the number of CTAs per face
    double* peer_lo_plane;             // lower neighbour's halo plane nz-1 of its Pr'
    double* peer_hi_plane;             // upper neighbour's halo plane 0 of its Pr'
    unsigned long long* mbox;          // this rank's mailbox (NS3D_MB_*)
    unsigned long long* peer_lo_flag;  // lower neighbour's NS3D_MB_FLAG_HI
    unsigned long long* peer_hi_flag;  // upper neighbour's NS3D_MB_FLAG_LO
    // two-iterations-per-launch on slabs: the first iteration of the halo planes is recomputed
    // locally, which needs one more plane of the neighbour's CURRENT iterate and its dPrdτ plane
    const double* peer_lo_cur;  // lower neighbour's Pr plane nz-3   (= local plane -1)
    const double* peer_hi_cur;  // upper neighbour's Pr plane 2      (= local plane nz)
    const double* peer_lo_dp;   // lower neighbour's dPrdτ of its plane nz-2 (= local plane 0)
    const double* peer_hi_dp;   // upper neighbour's dPrdτ of its plane 1    (= local plane nz-1)
    // byte strides, precomputed on the host so that the kernel takes them from the constant bank
    // instead of re-deriving 64-bit products under register pressure
    long long rowB, planeB, dplaneB;
    int zchunk_tb;  // chunk length of the two-iterations-per-launch kernels
    int tb_ty;      // ... and their tile height (8, 16 or 32 rows of 32 columns)
    // pt_tb2s_kernel: byte displacements between the arrays of THIS launch (tb2s_set_offsets)
    long long oDV;    // ∇V - Pr
    long long oDVn;   // ∇V - Pr + one plane
    long long oZP2;   // two Pr planes
    long long oPr;    // PrN - Pr - one plane         : c + oPr is the stage-2 store of plane s-1
    long long oDP;    // dPN - dP - one dPrdτ plane   : d + oDP is the stage-2 store of plane s-1
    long long oDPt;   // dPN - dP                     : top-face store of plane s
};

template <int MODE>
__device__ __forceinline__ double bracket(const PtK& p, double pc, double xm, double xp, double ym, double yp,
                                          double zm, double zp, double divv)
{
    const double d2x = (xp - pc) - (pc - xm);
    const double d2y = (yp - pc) - (pc - ym);
    const double d2z = (zp - pc) - (pc - zm);
    if (MODE == NS3D_PARITY) {
        return ((d2x / p.dx / p.dx + d2y / p.dy / p.dy) + d2z / p.dz / p.dz) - p.rdt * divv;
    } else if (MODE == NS3D_FAST) {
        const double tx = div3(div3(d2x, p.dx, p.rdx), p.dx, p.rdx);
        const double ty = div3(div3(d2y, p.dy, p.rdy), p.dy, p.rdy);
        const double tz = div3(div3(d2z, p.dz, p.rdz), p.dz, p.rdz);
        return ((tx + ty) + tz) - p.rdt * divv;
    } else {
        return fma(-p.rdt, divv, fma(d2z, p.rdz2, fma(d2y, p.rdy2, d2x * p.rdx2)));
    }
}

// Value stored at x-face point (i in {0, nx-1}) of plane k (0-based) given the mirrored
// interior value u.  Neumann: u.  M outlet: val (bc_x_Pr!, M:147-150).  G: bc_xhydstatic!.
__device__ __forceinline__ double xface(const PtK& p, bool hi, int k, double u)
{
    const int kind = hi ? p.xhi_kind : p.xlo_kind;
    if (kind == X_NEUMANN) return u;
    if (kind == X_DIRICHLET) return hi ? p.xhi_val : p.xlo_val;
    const double h = p.rho_g * ((double)(p.hyd_nz - (k + 1)) + 0.5) * p.hyd_dz;
    return hi ? h : h + p.xlo_val;
}

// Stores the new value u of the interior point at x index i into row pointers of one plane:
// `row` is the row j itself, mirrors go to x faces (xl/xh) and to the y-face rows (ylo/yhi,
// NULL when the point is not next to that face).  k selects the hydrostatic value (variant G).
__device__ __forceinline__ void store_row(const PtK& p, double* __restrict__ row, int i, int k, double u, bool xl,
                                          bool xh)
{
    row[i] = u;
    if (xl) row[0] = xface(p, false, k, u);
    if (xh) row[p.nx - 1] = xface(p, true, k, u);
}

__device__ __forceinline__ void store_plane(const PtK& p, double* __restrict__ plane, int i, int j, int k, double u,
                                            bool xl, bool xh, bool yl, bool yh)
{
    store_row(p, plane + (size_t)j * p.nx, i, k, u, xl, xh);
    if (yl) store_row(p, plane, i, k, u, xl, xh);
    if (yh) store_row(p, plane + (size_t)(p.ny - 1) * p.nx, i, k, u, xl, xh);
}

// One fused PT iteration: K5 + K6 + set_bc_Pr!.
//
// Each thread owns one interior column (i,j) and marches planes [kb,ke).  The three streams
// that come from DRAM/L2 (Pr plane k+1, dPrdτ, ∇V) are software-pipelined: the loop is unrolled
// by two with two named register sets (A, B), the loads of plane k+1 are issued before the
// arithmetic of plane k and land in the other set, so no register move waits on them.  The
// four in-plane neighbours were brought into L1 one step earlier by this and the adjacent
// warps (as their "plane k+1" loads) and are read just in time.  The z neighbours stay in
// registers, and the face bookkeeping is hoisted into one per-thread flag.
struct StreamRegs {
    double zp, dq, dv;
};

// P2P = true: the CTAs that update plane 1 / nz-2 of a slab also store the new values -- with
// their x/y mirror images -- straight into the neighbour's halo plane over NVLink (mapped peer
// memory) and the last of them releases a flag in the neighbour's mailbox; the same CTAs of the
// next launch acquire the neighbour's flag before touching the halos.  One kernel does the
// update and the halo exchange; there is no separate communication step to overlap.
template <int MODE, int MINB, bool P2P>
__global__ void __launch_bounds__(256, MINB) pt_iter_kernel(const double* __restrict__ Pr, double* __restrict__ PrN,
                                                      double* __restrict__ dP, const double* __restrict__ divV,
                                                      const PtK p)
{
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const bool active = (i <= nx - 2) && (j <= ny - 2);
    if (!P2P && !active) return;
    int bz = blockIdx.z;
    if (P2P) {  // the two face chunks go first (lowest CTA indices are scheduled first), then the rest
        const int nc = gridDim.z;
        if (bz == 1) bz = nc - 1;
        else if (bz >= 2) bz = p.reverse ? nc - bz : bz - 1;
    } else if (p.reverse) {
        bz = gridDim.z - 1 - bz;
    }
    const int kb = p.faces ? (bz == 0 ? 1 : nz - 2) : p.kbeg + bz * p.zchunk;
    const int ke = p.faces ? kb + 1 : min(kb + p.zchunk, p.kend);  // interior planes [kb, ke)
    const bool lo_face = P2P && p.zlo_halo && kb == 1;
    const bool hi_face = P2P && p.zhi_halo && ke == nz - 1;
    if (P2P && (lo_face | hi_face)) {
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            if (lo_face) wait_neighbour(p.mbox, 0);
            if (hi_face) wait_neighbour(p.mbox, 1);
        }
        __syncthreads();
    }
    if (active) {
    const bool xl = (i == 1), xh = (i == nx - 2), yl = (j == 1), yh = (j == ny - 2);
    // One predicate keeps every boundary store off the hot path: threads next to an x/y face, and
    // whole CTAs whose chunk holds plane 1 or nz-2 (z faces / slab interfaces).
    const bool slow = xl | xh | yl | yh | (kb == 1) | (ke == nz - 1);
    const long long rowB = p.rowB, planeB = p.planeB, dplaneB = p.dplaneB;
    // byte pointers to this thread's column at plane kb; all other addresses are this + constants
    const char* c0 = (const char*)(Pr + idx3(i, j, kb, nx, ny));
    const long long dDV = (const char*)divV - (const char*)Pr;    // same shape: constant displacements
    const long long dOUT = (const char*)PrN - (const char*)Pr;
    char* d0 = (char*)(dP + idx3(i - 1, j - 1, kb - 1, nx - 2, ny - 2));
#define LD(ptr) (*(const double*)(ptr))
    auto load = [&](StreamRegs& r, const char* c, const char* d) {
        r.zp = LD(c + planeB);
        r.dq = LD(d);
        r.dv = LD(c + dDV);
    };
    double pm = LD(c0 - planeB), pc = LD(c0);
    auto compute = [&](const StreamRegs& r, int k, const char* c, char* d) {
        const double L = bracket<MODE>(p, pc, LD(c - 8), LD(c + 8), LD(c - rowB), LD(c + rowB), pm, r.zp, r.dv);
        double dn, u;
        if (MODE == NS3D_FASTEST) {
            dn = fma(p.dtau, L, r.dq * p.omd);
            u = fma(p.dtau, dn, pc);
        } else {
            dn = r.dq * p.omd + p.dtau * L;  // M:71
            u = pc + p.dtau * dn;            // M:80
        }
        *(double*)d = dn;
        *(double*)(const_cast<char*>(c) + dOUT) = u;
        if (slow) {
            const ptrdiff_t sxy = (ptrdiff_t)nx * ny;
            double* plane = PrN + (ptrdiff_t)k * sxy;
            // x/y mirror images in this plane (bc_x!, bc_y!; outlet / hydrostatic x faces)
            if (xl) plane[(ptrdiff_t)j * nx] = xface(p, false, k, u);
            if (xh) plane[(ptrdiff_t)j * nx + nx - 1] = xface(p, true, k, u);
            if (yl) store_row(p, plane, i, k, u, xl, xh);
            if (yh) store_row(p, plane + (ptrdiff_t)(ny - 1) * nx, i, k, u, xl, xh);
            if (k == 1) {
                if (!p.zlo_halo) store_plane(p, PrN, i, j, 0, u, xl, xh, yl, yh);  // bc_z! M:129
                else if (P2P) store_plane(p, p.peer_lo_plane, i, j, k, u, xl, xh, yl, yh);  // update_halo!(Pr)
            }
            if (k == nz - 2) {
                if (!p.zhi_halo) store_plane(p, PrN + (ptrdiff_t)(nz - 1) * sxy, i, j, nz - 1, u, xl, xh, yl, yh);  // M:130
                else if (P2P) store_plane(p, p.peer_hi_plane, i, j, k, u, xl, xh, yl, yh);
            }
        }
        pm = pc;
        pc = r.zp;
    };

    StreamRegs A, B;
    load(A, c0, d0);
    for (int k = kb; k < ke; k += 2) {
        // Prefetches are unconditional: past the chunk they touch planes other CTAs own, past the
        // array the allocator's padding (ns3d_zeros); such values are loaded and never used.
        load(B, c0 + planeB, d0 + dplaneB);
        compute(A, k, c0, d0);
        load(A, c0 + 2 * planeB, d0 + 2 * dplaneB);
        if (k + 1 < ke) compute(B, k + 1, c0 + planeB, d0 + dplaneB);
        c0 += 2 * planeB;
        d0 += 2 * dplaneB;
    }
#undef LD
    }  // active
    if (P2P && (lo_face | hi_face)) {
        __threadfence_system();  // this thread's peer stores are performed before the flag can be seen
        __syncthreads();
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            const unsigned nface = gridDim.x * gridDim.y;
            if (lo_face) signal_neighbour(p.mbox, 0, p.peer_lo_flag, nface);
            if (hi_face) signal_neighbour(p.mbox, 1, p.peer_hi_flag, nface);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Two PT iterations per launch (temporal blocking), first version.  Plain launches now run the
// re-write below (pt_tb2s_kernel, bit-identical results); this kernel stays for its P2P = true
// instantiation, which also performs update_halo!(Pr) over peer memory on slab interfaces.
//
// A CTA owns a 32x16 tile of columns and marches along z with a two-stage pipeline: stage 1
// computes the first iteration's pressure q = Pr^(1) of plane s for every tile column (from
// global memory, exactly like pt_iter_kernel), publishes it in a three-deep shared-memory ring,
// and stage 2 computes the second iteration of plane s-1 for the 30x14 inner columns from the
// ring (in-plane neighbours) and registers (z neighbours, dPrdτ^(1), ∇V) -- no global loads.
// Pr^(1) and dPrdτ^(1) never touch DRAM: 5 field passes per TWO iterations.  Tile rims and the
// two extra planes per z-chunk are recomputed by the neighbouring CTAs; columns on a domain face
// take the value of their index clamped into the interior (the folded bc_x!/bc_y!), so stage 1
// needs no extra synchronisation for the boundary conditions, and z faces are handled in
// registers (q[0] = q[1], q[nz-1] = q[nz-2]).  Same per-cell arithmetic as pt_iter_kernel, so
// PARITY mode stays bit-equal to the oracle.  dPrdτ ping-pongs with a context-owned shadow
// (rim columns of other CTAs read the old value while the owner writes the new one).
// ---------------------------------------------------------------------------------------------
constexpr int TB_X = 32;

template <int MODE, int TB_Y, bool P2P>
__global__ void __launch_bounds__(TB_X* TB_Y, 1024 / (TB_X * TB_Y)) pt_tb2_kernel(const double* __restrict__ Pr, double* __restrict__ PrN,
                                                              const double* __restrict__ dP, double* __restrict__ dPN,
                                                              const double* __restrict__ divV, const PtK p)
{
    __shared__ double ring[3][TB_Y][TB_X];
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * (TB_X - 2) + tx;  // tile columns include the rim (and the domain faces)
    const int j = blockIdx.y * (TB_Y - 2) + ty;
    const bool in_dom = (i <= nx - 1) && (j <= ny - 1);
    const int ci = min(max(i, 1), nx - 2), cj = min(max(j, 1), ny - 2);  // clamped into the interior
    const bool interior = in_dom && ci == i && cj == j;
    const bool owner = interior && tx >= 1 && tx <= TB_X - 2 && ty >= 1 && ty <= TB_Y - 2;  // stage-2 output
    int bz = blockIdx.z;
    if (P2P && !p.faces) {  // the chunks next to a slab interface go first
        const int nc = gridDim.z;
        if (bz == 1) bz = nc - 1;
        else if (bz >= 2) bz = p.reverse ? nc - bz : bz - 1;
    } else if (p.reverse && !p.faces) {
        bz = gridDim.z - 1 - bz;
    }
    // faces launch: only the two chunks of p.zchunk planes next to the z faces
    const int kb = p.faces ? (bz == 0 ? 1 : nz - 1 - p.zchunk) : p.kbeg + bz * p.zchunk;
    const int ke = p.faces ? kb + p.zchunk : min(kb + p.zchunk, p.kend);  // stage-2 planes [kb, ke)
    // On a slab interface the halo plane's first iteration is RECOMPUTED here (it is the
    // neighbour's plane nz-2 / 1) from the local halo plane plus one peer plane.
    const bool lo_face = P2P && p.zlo_halo && kb == 1;
    const bool hi_face = P2P && p.zhi_halo && ke == nz - 1;
    if (P2P && (lo_face | hi_face)) {
        if (tx == 0 && ty == 0) {
            if (lo_face) wait_neighbour(p.mbox, 0);
            if (hi_face) wait_neighbour(p.mbox, 1);
        }
        __syncthreads();
    }
    const int s0 = lo_face ? 0 : max(kb - 1, 1);
    const int s1 = hi_face ? nz - 1 : min(ke, nz - 2);  // stage-1 planes [s0, s1]
    const bool xl = (i == 1), xh = (i == nx - 2), yl = (j == 1), yh = (j == ny - 2);
    const long long rowB = p.rowB, planeB = p.planeB, dplaneB = p.dplaneB;
    const long long dDV = (const char*)divV - (const char*)Pr;
    const ptrdiff_t tcol = (ptrdiff_t)cj * nx + ci;                    // column offset in a Pr plane
    const ptrdiff_t dcol = (ptrdiff_t)(cj - 1) * (nx - 2) + (ci - 1);  // ... in a dPrdτ plane
    const char* c = (const char*)(Pr + (ptrdiff_t)s0 * nx * ny + tcol);
    const char* d = (const char*)(dP + ((ptrdiff_t)s0 - 1) * (nx - 2) * (ny - 2) + dcol);
#define LD(ptr) (*(const double*)(ptr))
    double pm = 0, pc = 0, zp = 0, dq = 0, dv = 0;
    if (in_dom) {
        if (P2P && lo_face) {  // plane -1 and the dPrdτ of plane 0 live on the lower neighbour
            pm = __ldcv(p.peer_lo_cur + tcol);
            dq = __ldcv(p.peer_lo_dp + dcol);
        } else {
            pm = LD(c - planeB);
            dq = LD(d);
        }
        pc = LD(c);
        zp = LD(c + planeB);
        dv = LD(c + dDV);
    }
    double q_m = 0, q_c = 0, d1_c = 0, dv_c = 0;  // stage-2 state of plane s-1 (and q of s-2)
    for (int s = s0; s <= s1; ++s) {
        // prefetch the three streamed values of plane s+1 (the allocator pads the arrays)
        double n_zp = 0, n_dq = 0, n_dv = 0;
        if (in_dom) {
            if (P2P && hi_face && s == nz - 2) {  // plane nz and the dPrdτ of plane nz-1: upper neighbour
                n_zp = __ldcv(p.peer_hi_cur + tcol);
                n_dq = __ldcv(p.peer_hi_dp + dcol);
            } else {
                n_zp = LD(c + 2 * planeB);
                n_dq = LD(d + dplaneB);
            }
            n_dv = LD(c + planeB + dDV);
        }
        // ---- stage 1: first iteration at the clamped column, plane s ---------------------------
        double q = 0, d1 = 0;
        if (in_dom) {
            const double L = bracket<MODE>(p, pc, LD(c - 8), LD(c + 8), LD(c - rowB), LD(c + rowB), pm, zp, dv);
            if (MODE == NS3D_FASTEST) {
                d1 = fma(p.dtau, L, dq * p.omd);
                q = fma(p.dtau, d1, pc);
            } else {
                d1 = dq * p.omd + p.dtau * L;
                q = pc + p.dtau * d1;
            }
            if (i == 0) q = xface(p, false, s, q);        // bc_x_Pr! / bc_xhydstatic! images
            if (i == nx - 1) q = xface(p, true, s, q);
        }
        ring[s % 3][ty][tx] = q;
        __syncthreads();
        // ---- stage 2: second iteration of plane s-1 ---------------------------------------------
        const int k2 = s - 1;
        if (owner && k2 >= kb) {
            const double(*rp)[TB_X] = ring[k2 % 3];
            const double L = bracket<MODE>(p, q_c, rp[ty][tx - 1], rp[ty][tx + 1], rp[ty - 1][tx], rp[ty + 1][tx], q_m, q, dv_c);
            double d2, u;
            if (MODE == NS3D_FASTEST) {
                d2 = fma(p.dtau, L, d1_c * p.omd);
                u = fma(p.dtau, d2, q_c);
            } else {
                d2 = d1_c * p.omd + p.dtau * L;
                u = q_c + p.dtau * d2;
            }
            dPN[idx3(i - 1, j - 1, k2 - 1, nx - 2, ny - 2)] = d2;
            store_plane(p, PrN + (ptrdiff_t)k2 * (ptrdiff_t)nx * ny, i, j, k2, u, xl, xh, yl, yh);
            if (k2 == 1) {
                if (!p.zlo_halo) store_plane(p, PrN, i, j, 0, u, xl, xh, yl, yh);  // bc_z! M:129
                else if (P2P) store_plane(p, p.peer_lo_plane, i, j, k2, u, xl, xh, yl, yh);  // update_halo!(Pr)
            }
            if (P2P && k2 == nz - 2 && p.zhi_halo) store_plane(p, p.peer_hi_plane, i, j, k2, u, xl, xh, yl, yh);
        }
        // rotate: plane s becomes "s-1"
        q_m = q_c;
        q_c = q;
        if (s == 1 && !p.zlo_halo) {  // bc_z!: q[0] is the image of q[1] (hydrostatic x faces depend on the plane)
            q_m = q;
            if (i == 0) q_m = xface(p, false, 0, q);
            if (i == nx - 1) q_m = xface(p, true, 0, q);
        }
        d1_c = d1;
        dv_c = dv;
        pm = pc; pc = zp; zp = n_zp; dq = n_dq; dv = n_dv;
        c += planeB;
        d += dplaneB;
    }
    // physical top face: plane nz-2 needs q[nz-1], the image of q[nz-2]
    if (!p.zhi_halo && s1 == nz - 2 && ke == nz - 1 && owner) {
        const int k2 = nz - 2;
        double q_p = q_c;
        const double(*rp)[TB_X] = ring[k2 % 3];
        const double L = bracket<MODE>(p, q_c, rp[ty][tx - 1], rp[ty][tx + 1], rp[ty - 1][tx], rp[ty + 1][tx], q_m, q_p, dv_c);
        double d2, u;
        if (MODE == NS3D_FASTEST) {
            d2 = fma(p.dtau, L, d1_c * p.omd);
            u = fma(p.dtau, d2, q_c);
        } else {
            d2 = d1_c * p.omd + p.dtau * L;
            u = q_c + p.dtau * d2;
        }
        dPN[idx3(i - 1, j - 1, k2 - 1, nx - 2, ny - 2)] = d2;
        store_plane(p, PrN + (ptrdiff_t)k2 * (ptrdiff_t)nx * ny, i, j, k2, u, xl, xh, yl, yh);
        if (k2 == 1) {
            if (!p.zlo_halo) store_plane(p, PrN, i, j, 0, u, xl, xh, yl, yh);
            else if (P2P) store_plane(p, p.peer_lo_plane, i, j, k2, u, xl, xh, yl, yh);
        }
        store_plane(p, PrN + (ptrdiff_t)(nz - 1) * (ptrdiff_t)nx * ny, i, j, nz - 1, u, xl, xh, yl, yh);  // bc_z! M:130
    }
#undef LD
    if (P2P && (lo_face | hi_face)) {
        __threadfence_system();
        __syncthreads();
        if (tx == 0 && ty == 0) {
            const unsigned nface = gridDim.x * gridDim.y;
            if (lo_face) signal_neighbour(p.mbox, 0, p.peer_lo_flag, nface);
            if (hi_face) signal_neighbour(p.mbox, 1, p.peer_hi_flag, nface);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pt_tb2s_kernel: the same two-iterations-per-launch scheme as pt_tb2_kernel (same tiles, same
// ring, same per-cell arithmetic, bit-identical results), re-written for instruction count.
// ncu showed pt_tb2_kernel issue-bound (63-65 % issue slots busy, DRAM 34 % busy, 128 instructions
// per cell and iteration against ~35 of arithmetic and memory operations): its loop body re-derived
// 64-bit addresses from (i,j,k) for every store, took the ring slot modulo 3, rotated nine doubles
// through register moves, and walked through the predicates of nine mirror-image stores per
// plane.  Here
//   * the loop is unrolled by three with the register roles rotated by NAME (ring slots are
//     compile-time constants, the Pr / ∇V / q chains need no moves);
//   * every address is a running byte pointer plus a launch constant (two 64-bit adds per plane);
//   * threads outside the domain work on a clamped duplicate column instead of being predicated;
//   * all boundary work (mirror images, z faces, non-Neumann x faces) sits behind ONE per-thread
//     flag OR-ed with a CTA-uniform plane test, in a compact loop instead of nine inlined stores.
// Plain (non-peer) launches only: slab-interface chunks keep pt_tb2_kernel<.,.,true>, which is
// interchangeable plane range by plane range.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void pt_update(const PtK& p, double L, double dq, double pc, double& dn, double& u)
{
    if (MODE == NS3D_FASTEST) {
        dn = fma(p.dtau, L, dq * p.omd);
        u = fma(p.dtau, dn, pc);
    } else {
        dn = dq * p.omd + p.dtau * L;  // M:71
        u = pc + p.dtau * dn;          // M:80
    }
}

// The mirror images of interior point (i,j) with new value u in plane k of PrN (bc_x!, bc_y!,
// bc_x_Pr! / bc_xhydstatic! folded in), and -- when k_face >= 0 -- the whole set once more in the
// z-face plane k_face (bc_z!).  Cold path: kept as a loop so that it costs little code.
// One plane's worth of the same: `plane` is the base of the target x-y plane (this rank's, or a
// neighbour's halo plane over peer memory), kk the plane index the x-face values are computed for.
__device__ __forceinline__ void tb2s_images_into(const PtK& p, double* __restrict__ plane, int kk, int i, int j, double u,
                                                 bool xl, bool xh, bool yl, bool yh)
{
    const double vlo = xface(p, false, kk, u), vhi = xface(p, true, kk, u);
#pragma unroll 1
    for (int r = 0; r < 3; ++r) {
        if ((r == 1 && !yl) || (r == 2 && !yh)) continue;
        double* row = plane + (ptrdiff_t)(r == 0 ? j : (r == 1 ? 0 : p.ny - 1)) * p.nx;
        row[i] = u;
        if (xl) row[0] = vlo;
        if (xh) row[p.nx - 1] = vhi;
    }
}

__device__ __forceinline__ void tb2s_images(const PtK& p, double* __restrict__ PrN, int i, int j, int k, int k_face,
                                            double u, bool xl, bool xh, bool yl, bool yh)
{
    const ptrdiff_t sxy = (ptrdiff_t)p.nx * p.ny;
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
        const int kk = t == 0 ? k : k_face;
        if (kk < 0) break;
        double* plane = PrN + (ptrdiff_t)kk * sxy;
        const double vlo = xface(p, false, kk, u), vhi = xface(p, true, kk, u);
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
            if ((r == 1 && !yl) || (r == 2 && !yh)) continue;
            double* row = plane + (ptrdiff_t)(r == 0 ? j : (r == 1 ? 0 : p.ny - 1)) * p.nx;
            row[i] = u;
            if (xl) row[0] = vlo;
            if (xh) row[p.nx - 1] = vhi;
        }
    }
}

struct Tb2sInv {
    double* PrN;
    int kb_own;    // first stage-2 plane of this thread (INT_MAX for threads that own no column)
    int top_own;   // nz-2 when this thread closes the physical top face in this chunk, else -1
    int edge;      // owner next to an x/y face: mirror images to store
    int xfix;      // x-face column with a non-Neumann condition: stage-1 value is replaced
    int bx, by;    // tile origin (the slow paths re-derive i, j from it)
    // P2P instantiation only: this CTA's chunk touches a slab interface; column offsets into the
    // neighbours' planes (Pr layout / dPrdτ layout)
    int lo_face, hi_face;
    ptrdiff_t tcol, dcol;
};

// Keeps a per-thread value in a register: without it ptxas re-derives loop invariants from
// %tid / %ctaid in every iteration to stay under the register cap (measured in SASS).
#ifdef NS3D_HOST_EMU
#define NS3D_KEEP(x) ((void)0)
#else
#define NS3D_KEEP(x) asm volatile("" : "+r"(x))
#endif

// One plane of the two-stage pipeline.  On entry PM, PC, ZP = Pr of planes s-1, s, s+1 at the
// (clamped) column, DQ = dPrdτ of plane s, DVC / DV = ∇V of planes s-1 / s, QM, QC = first-iteration
// pressure of planes s-2, s-1, D1C = first-iteration dPrdτ of plane s-1.  On exit PM holds Pr of
// plane s+2 and DVN ∇V of plane s+1, QN the first-iteration pressure of plane s, so the caller
// continues with the roles rotated: (PC,ZP,PM), (DV,DVN,DVC), (QC,QN,QM).
// Synchronisation of one tile row (= one warp, TB_X = 32) with the rows above and below it instead
// of the whole CTA: a row of the ring is read only by the warps of the adjacent rows, so warp w
// meets warp w-1 on named barrier w and warp w+1 on named barrier w+1 (64 threads each; ids 1..15,
// 0 stays __syncthreads).  Lower pair first on every warp: no cycle.  RAW: a neighbour's slot is
// read after the pair barrier of the step it was written in; WAR: a warp overwrites a slot two
// pair barriers after its neighbours read it.
__device__ __forceinline__ void pair_barrier(int id)
{
#ifdef NS3D_HOST_EMU
    emu::named_barrier(id, 64);
#else
    asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
#endif
}
template <int TB_Y>
__device__ __forceinline__ void row_barrier(int ty)
{
    if (ty > 0) pair_barrier(ty);
    if (ty < TB_Y - 1) pair_barrier(ty + 1);
}

// DRAM -> L2 prefetch of one line (no destination register; a no-op on the host emulation).
__device__ __forceinline__ void prefetch_l2(const void* ptr)
{
#ifdef NS3D_HOST_EMU
    (void)ptr;
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#endif
}

//
// NP (neighbour prefetch): the four in-plane neighbours of the NEXT plane are loaded one step
// ahead as well (NB[0..3] = x-, x+, y-, y+), so that no global-load latency sits on the per-plane
// dependency chain load -> stage 1 -> barrier -> stage 2 (ncu: the kernel is latency-bound, not
// issue-bound -- long-scoreboard stalls on exactly these loads; the rim columns always miss L1).
// PF = planes of additional DRAM -> L2 software prefetch ahead of the register prefetches.
// NXC, NYC > 0: the x-y extent of the grid is a compile-time constant (the sizes the reference's
// scripts and BASELINE.json's configurations run), so row and plane strides become immediate
// offsets of the load/store instructions instead of 64-bit additions (SASS: about 36 of the 110
// instructions per plane were address arithmetic); 0 = strides from the kernel parameters.
template <int NXC, int NYC>
struct Tb2sStride {
    static __device__ __forceinline__ long long row(const PtK& p) { return NXC > 0 ? 8LL * NXC : p.rowB; }
    static __device__ __forceinline__ long long plane(const PtK& p) { return NXC > 0 ? 8LL * NXC * NYC : p.planeB; }
    static __device__ __forceinline__ long long dplane(const PtK& p)
    {
        return NXC > 0 ? 8LL * (NXC - 2) * (NYC - 2) : p.dplaneB;
    }
};

// PB: pairwise row barriers (row_barrier) instead of __syncthreads -- ROUND-2 CANDIDATE, unmeasured.
// P2P: the chunk may touch a slab interface (see pt_tb2sp_kernel) -- ROUND-2 CANDIDATE, unmeasured.
template <int MODE, int TB_Y, int SLOT, int PF, bool NP, int NXC, int NYC, bool PB, bool P2P = false>
__device__ __forceinline__ void tb2s_step(const PtK& p, const Tb2sInv& v, const int s, const char*& c, const char*& d,
                                          double* __restrict__ sm, double& PM, double& PC, double& ZP, double& DQ,
                                          double& DVC, double& DV, double& DVN, double& QM, double& QC, double& QN,
                                          double& D1C, double (&NB)[4])
{
#define LD(ptr) (*(const double*)(ptr))
    constexpr int SLOTSZ = TB_Y * TB_X;
    typedef Tb2sStride<NXC, NYC> G;
    // (evaluated at the point of use: ptxas spills when these sit in named locals across the step)
#define rowB (G::row(p))
#define planeB (G::plane(p))
#define dplaneB (G::dplane(p))
    // ∇V of the next plane / Pr two planes ahead: with compile-time strides one base per array and
    // immediates, otherwise host-precomputed displacements from c (one 64-bit add each)
#define a_dvn (NXC > 0 ? (c + p.oDV) + planeB : c + p.oDVn)
#define a_zp2 (NXC > 0 ? c + 2 * planeB : c + p.oZP2)
    // ---- stage 1: first iteration at the clamped column, plane s ------------------------------
    double xm, xp, ym, yp;
    if (NP) {
        xm = NB[0]; xp = NB[1]; ym = NB[2]; yp = NB[3];
    } else {
        xm = LD(c - 8); xp = LD(c + 8); ym = LD(c - rowB); yp = LD(c + rowB);
    }
    DVN = LD(a_dvn);  // streams of the next plane (the allocator pads the arrays)
    const double L1 = bracket<MODE>(p, PC, xm, xp, ym, yp, PM, ZP, DV);
    double D1N;
    pt_update<MODE>(p, L1, DQ, PC, D1N, QN);
    if (P2P && v.hi_face && s == p.nz - 2) {
        // plane nz and the dPrdτ of plane nz-1 live on the upper neighbour (its plane 2 / its first dPrdτ plane)
        PM = __ldcv(p.peer_hi_cur + v.tcol);
        DQ = __ldcv(p.peer_hi_dp + v.dcol);
    } else {
        PM = LD(a_zp2);           // PM and DQ are dead: reuse them for planes s+2 / s+1
        DQ = LD(d + dplaneB);
    }
    if (NP) {
        const char* cn = c + planeB;
        NB[0] = LD(cn - 8); NB[1] = LD(cn + 8); NB[2] = LD(cn - rowB); NB[3] = LD(cn + rowB);
    }
    if (PF > 0) {  // further ahead into L2, so that the register prefetches above hit there
        prefetch_l2(a_zp2 + PF * planeB);
        prefetch_l2(d + (1 + PF) * dplaneB);
        prefetch_l2(a_dvn + PF * planeB);
    }
    if (v.xfix) QN = xface(p, v.xfix > 1, s, QN);  // bc_x_Pr! / bc_xhydstatic! images (x-face columns only)
    sm[SLOT * SLOTSZ] = QN;
    // ---- stage 2: second iteration of plane s-1 from the ring and registers ---------------------
    // It reads the ring slot of plane s-1, which the barrier of the PREVIOUS step published, so it
    // runs before this step's barrier: one synchronisation point per plane, and the two
    // arithmetic chains of a step are independent up to the z term (ncu: barrier and
    // fixed-latency waits were the top stalls once the loads were off the critical path).
    const int k2 = s - 1;
    if (k2 >= v.kb_own) {
        const double* r = sm + ((SLOT + 2) % 3) * SLOTSZ;
        const double L2 = bracket<MODE>(p, QC, r[-1], r[1], r[-TB_X], r[TB_X], QM, QN, DVC);
        double d2, u;
        pt_update<MODE>(p, L2, D1C, QC, d2, u);
        *(double*)(const_cast<char*>(d) + p.oDP) = d2;
        *(double*)(const_cast<char*>(c) + p.oPr) = u;
        if (v.edge | (k2 == 1) | (P2P && k2 == p.nz - 2)) {  // mirror images; bc_z! M:129 when plane 1 is next to a physical face
            const int i = v.bx + (int)threadIdx.x, j = v.by + (int)threadIdx.y;
            tb2s_images(p, v.PrN, i, j, k2, (k2 == 1 && !p.zlo_halo) ? 0 : -1, u, i == 1, i == p.nx - 2, j == 1,
                        j == p.ny - 2);
            if (P2P) {  // update_halo!(Pr): the planes a slab sends go straight into the neighbours' halo planes
                if (k2 == 1 && p.zlo_halo)
                    tb2s_images_into(p, p.peer_lo_plane, k2, i, j, u, i == 1, i == p.nx - 2, j == 1, j == p.ny - 2);
                if (k2 == p.nz - 2 && p.zhi_halo)
                    tb2s_images_into(p, p.peer_hi_plane, k2, i, j, u, i == 1, i == p.nx - 2, j == 1, j == p.ny - 2);
            }
        }
    }
    if (s == 1 && !p.zlo_halo) QC = QN;  // bc_z!: q[0] is the image of q[1] (QC becomes QM of the next plane)
    // slot SLOT is complete; slot SLOT+1 may be overwritten by the next step
    if (PB) row_barrier<TB_Y>((int)threadIdx.y);
    else __syncthreads();
    if (s == v.top_own) {
        // physical top face: plane nz-2 needs q[nz-1], the image of q[nz-2]; its second iteration
        // follows here because there is no further stage-1 plane to trigger it
        const double* r = sm + SLOT * SLOTSZ;
        const double L2 = bracket<MODE>(p, QN, r[-1], r[1], r[-TB_X], r[TB_X], QC, QN, DV);
        double d2, u;
        pt_update<MODE>(p, L2, D1N, QN, d2, u);
        *(double*)(const_cast<char*>(d) + p.oDPt) = d2;
        const int i = v.bx + (int)threadIdx.x, j = v.by + (int)threadIdx.y;
        const bool xl = i == 1, xh = i == p.nx - 2, yl = j == 1, yh = j == p.ny - 2;
        tb2s_images(p, v.PrN, i, j, s, p.nz - 1, u, xl, xh, yl, yh);                      // + bc_z! M:130
        if (s == 1 && !p.zlo_halo) tb2s_images(p, v.PrN, i, j, 0, -1, u, xl, xh, yl, yh);  // nz = 3
    }
    D1C = D1N;
    c += planeB;
    d += dplaneB;
#undef LD
#undef rowB
#undef planeB
#undef dplaneB
#undef a_dvn
#undef a_zp2
}

template <int MODE, int TB_Y, int PF, bool NP, int NXC, int NYC, bool PB = false>
__global__ void __launch_bounds__(TB_X* TB_Y, 1024 / (TB_X * TB_Y)) pt_tb2s_kernel(const double* Pr, double* PrN, const double* dP, double* dPN,
                                                               const double* divV, const PtK p)
{
    __shared__ double ring[3 * TB_Y * TB_X];
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * (TB_X - 2) + tx;  // tile columns include the rim (and the domain faces)
    const int j = blockIdx.y * (TB_Y - 2) + ty;
    const int ci = min(max(i, 1), nx - 2), cj = min(max(j, 1), ny - 2);  // clamped into the interior
    int bz = blockIdx.z;
    if (p.reverse) bz = gridDim.z - 1 - bz;
    const int kb = p.kbeg + bz * p.zchunk;
    const int ke = min(kb + p.zchunk, p.kend);  // stage-2 planes [kb, ke)
    const int s0 = max(kb - 1, 1);
    const int s1 = min(ke, nz - 2);             // stage-1 planes [s0, s1]
    Tb2sInv v;
    v.PrN = PrN;
    v.bx = blockIdx.x * (TB_X - 2); v.by = blockIdx.y * (TB_Y - 2);
    // stage-2 output: interior columns that are not on the tile rim
    const bool owner = (ci == i) && (cj == j) && tx >= 1 && tx <= TB_X - 2 && ty >= 1 && ty <= TB_Y - 2;
    v.kb_own = owner ? kb : 0x7fffffff;
    v.top_own = (owner && !p.zhi_halo && ke == nz - 1) ? nz - 2 : -1;
    v.edge = owner && ((i == 1) | (i == nx - 2) | (j == 1) | (j == ny - 2));
    v.xfix = (i == 0 && p.xlo_kind != X_NEUMANN) ? 1 : ((i == nx - 1 && p.xhi_kind != X_NEUMANN) ? 2 : 0);
    NS3D_KEEP(v.kb_own); NS3D_KEEP(v.top_own); NS3D_KEEP(v.edge); NS3D_KEEP(v.xfix);
    // Threads whose column lies outside the domain (last tiles) run the clamped column like a rim
    // thread: their ring entries are never read by an owner and they never store.
    const char* c = (const char*)(Pr + (ptrdiff_t)s0 * nx * ny + (ptrdiff_t)cj * nx + ci);
    const char* d = (const char*)(dP + ((ptrdiff_t)s0 - 1) * (nx - 2) * (ny - 2) + (ptrdiff_t)(cj - 1) * (nx - 2) + (ci - 1));
    int tslot = ty * TB_X + tx;
    NS3D_KEEP(tslot);
    double* sm = ring + tslot;
#define LD(ptr) (*(const double*)(ptr))
    const long long rowB = Tb2sStride<NXC, NYC>::row(p), planeB = Tb2sStride<NXC, NYC>::plane(p);
    double A = LD(c - planeB), B = LD(c), C = LD(c + planeB);  // Pr of planes s0-1, s0, s0+1
    double DQ = LD(d);
    double VA = 0, VB = LD(c + p.oDV), VC = 0;                         // ∇V of planes s-1, s, s+1
    double QA = 0, QB = 0, QC = 0, D1 = 0;
    double NB[4] = {0, 0, 0, 0};
    if (NP) {
        NB[0] = LD(c - 8); NB[1] = LD(c + 8); NB[2] = LD(c - rowB); NB[3] = LD(c + rowB);
    }
#undef LD
    int s = s0;
    while (true) {
        tb2s_step<MODE, TB_Y, 0, PF, NP, NXC, NYC, PB>(p, v, s, c, d, sm, A, B, C, DQ, VA, VB, VC, QA, QB, QC, D1, NB);
        if (s == s1) break;
        ++s;
        tb2s_step<MODE, TB_Y, 1, PF, NP, NXC, NYC, PB>(p, v, s, c, d, sm, B, C, A, DQ, VB, VC, VA, QB, QC, QA, D1, NB);
        if (s == s1) break;
        ++s;
        tb2s_step<MODE, TB_Y, 2, PF, NP, NXC, NYC, PB>(p, v, s, c, d, sm, C, A, B, DQ, VC, VA, VB, QC, QA, QB, D1, NB);
        if (s == s1) break;
        ++s;
    }
}

// pt_tb2sp_kernel: pt_tb2s_kernel for chunks on a slab interface -- what pt_tb2_kernel<.,.,true> does
// (DESIGN.md 3.4), on the slim pipeline.  ROUND-2 CANDIDATE (option "tb2_slim_faces", off by default):
// bit-exact in the single-process and multi-process emulations, not yet run on a device.  On an
// interface the halo plane's first iteration is recomputed here (it is the neighbour's plane nz-2 / 1)
// from the local halo plane plus one peer plane of the neighbour's current iterate and its dPrdτ
// plane; the second-iteration planes 1 / nz-2 are also stored into the neighbour's halo plane;
// hand-over through the mailbox exactly as in pt_tb2_kernel (wait at CTA start, signal at its end).
template <int MODE, int TB_Y, int PF>
__global__ void __launch_bounds__(TB_X* TB_Y, 1024 / (TB_X * TB_Y)) pt_tb2sp_kernel(const double* Pr, double* PrN, const double* dP,
                                                                double* dPN, const double* divV, const PtK p)
{
    __shared__ double ring[3 * TB_Y * TB_X];
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * (TB_X - 2) + tx;
    const int j = blockIdx.y * (TB_Y - 2) + ty;
    const int ci = min(max(i, 1), nx - 2), cj = min(max(j, 1), ny - 2);
    int bz = blockIdx.z;
    if (!p.faces) {  // unsplit launch: the chunks next to a slab interface go first
        const int nc = gridDim.z;
        if (bz == 1) bz = nc - 1;
        else if (bz >= 2) bz = p.reverse ? nc - bz : bz - 1;
    }
    // faces launch: only the two chunks of p.zchunk planes next to the z faces
    const int kb = p.faces ? (bz == 0 ? 1 : nz - 1 - p.zchunk) : p.kbeg + bz * p.zchunk;
    const int ke = p.faces ? kb + p.zchunk : min(kb + p.zchunk, p.kend);  // stage-2 planes [kb, ke)
    const bool lo_face = p.zlo_halo && kb == 1;
    const bool hi_face = p.zhi_halo && ke == nz - 1;
    if (lo_face | hi_face) {
        if (tx == 0 && ty == 0) {
            if (lo_face) wait_neighbour(p.mbox, 0);
            if (hi_face) wait_neighbour(p.mbox, 1);
        }
        __syncthreads();
    }
    const int s0 = lo_face ? 0 : max(kb - 1, 1);
    const int s1 = hi_face ? nz - 1 : min(ke, nz - 2);  // stage-1 planes [s0, s1]
    Tb2sInv v;
    v.PrN = PrN;
    v.bx = blockIdx.x * (TB_X - 2); v.by = blockIdx.y * (TB_Y - 2);
    const bool owner = (ci == i) && (cj == j) && tx >= 1 && tx <= TB_X - 2 && ty >= 1 && ty <= TB_Y - 2;
    v.kb_own = owner ? kb : 0x7fffffff;
    v.top_own = (owner && !p.zhi_halo && ke == nz - 1) ? nz - 2 : -1;
    v.edge = owner && ((i == 1) | (i == nx - 2) | (j == 1) | (j == ny - 2));
    v.xfix = (i == 0 && p.xlo_kind != X_NEUMANN) ? 1 : ((i == nx - 1 && p.xhi_kind != X_NEUMANN) ? 2 : 0);
    v.lo_face = lo_face; v.hi_face = hi_face;
    v.tcol = (ptrdiff_t)cj * nx + ci;
    v.dcol = (ptrdiff_t)(cj - 1) * (nx - 2) + (ci - 1);
    NS3D_KEEP(v.kb_own); NS3D_KEEP(v.top_own); NS3D_KEEP(v.edge); NS3D_KEEP(v.xfix);
    const char* c = (const char*)(Pr + (ptrdiff_t)s0 * nx * ny + v.tcol);
    const char* d = (const char*)(dP + ((ptrdiff_t)s0 - 1) * (nx - 2) * (ny - 2) + v.dcol);
    int tslot = ty * TB_X + tx;
    NS3D_KEEP(tslot);
    double* sm = ring + tslot;
#define LD(ptr) (*(const double*)(ptr))
    const long long rowB = p.rowB, planeB = p.planeB;
    double A, DQ;
    if (lo_face) {  // plane -1 and the dPrdτ of plane 0 live on the lower neighbour (its plane nz-3 / its last dPrdτ plane)
        A = __ldcv(p.peer_lo_cur + v.tcol);
        DQ = __ldcv(p.peer_lo_dp + v.dcol);
    } else {
        A = LD(c - planeB);
        DQ = LD(d);
    }
    double B = LD(c), C = LD(c + planeB);
    double VA = 0, VB = LD(c + p.oDV), VC = 0;
    double QA = 0, QB = 0, QC = 0, D1 = 0;
    double NB[4] = {LD(c - 8), LD(c + 8), LD(c - rowB), LD(c + rowB)};
#undef LD
    int s = s0;
    while (true) {
        tb2s_step<MODE, TB_Y, 0, PF, true, 0, 0, false, true>(p, v, s, c, d, sm, A, B, C, DQ, VA, VB, VC, QA, QB, QC, D1, NB);
        if (s == s1) break;
        ++s;
        tb2s_step<MODE, TB_Y, 1, PF, true, 0, 0, false, true>(p, v, s, c, d, sm, B, C, A, DQ, VB, VC, VA, QB, QC, QA, D1, NB);
        if (s == s1) break;
        ++s;
        tb2s_step<MODE, TB_Y, 2, PF, true, 0, 0, false, true>(p, v, s, c, d, sm, C, A, B, DQ, VC, VA, VB, QC, QA, QB, D1, NB);
        if (s == s1) break;
        ++s;
    }
    if (lo_face | hi_face) {
        __threadfence_system();  // this thread's peer stores are performed before the flag can be seen
        __syncthreads();
        if (tx == 0 && ty == 0) {
            const unsigned nface = gridDim.x * gridDim.y;
            if (lo_face) signal_neighbour(p.mbox, 0, p.peer_lo_flag, nface);
            if (hi_face) signal_neighbour(p.mbox, 1, p.peer_hi_flag, nface);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pt_tb2d_kernel: pt_tb2s_kernel with TWO tile rows per thread (rows 2*ty and 2*ty+1 of a 32 x TB_Y
// tile, TB_Y/2 thread rows).  CANDIDATE for round 2 -- bit-exact in the emulation and on the parity
// tests, NOT yet measured against pt_tb2s_kernel (option "tb2_dual", off by default).  Why it
// might pay (ncu of pt_tb2s_kernel at 255x153x153: issue slots 60 % busy; per issue 3.0 warps at
// the barrier, 2.1 on fixed-latency dependencies):
//   * 8-warp CTAs (cheap barriers, 4 per SM at 32x8 threads) with the rim ratio of 32x16 tiles
//     (82 % of the columns produce output instead of 70 %);
//   * two independent dependency chains per thread; half as many barriers per cell;
//   * the rows' mutual y neighbours are the thread's own registers: 6 instead of 8 neighbour
//     loads (stage 1) and shared-memory reads (stage 2) per two cells.
// It costs registers (two sets of per-column state): MINB = CTAs per SM the kernel is compiled for.
// ---------------------------------------------------------------------------------------------
struct Tb2dInv {
    double* PrN;
    int kb_own[2];   // first stage-2 plane of each row (INT_MAX for rows that own no column)
    int top_own[2];  // nz-2 when the row closes the physical top face in this chunk, else -1
    int edge[2];     // owner row next to an x/y face
    int xfix;        // x-face column with a non-Neumann condition (same for both rows)
    int adj;         // the rows are neighbours in MEMORY too (no clamping between them): stage 1 of
                     // row 0 takes its y+ neighbour from row 1's registers and vice versa
    int bx, by;
};

template <int MODE, int TB_Y, int SLOT, int PF, int NXC, int NYC>
__device__ __forceinline__ void tb2d_step(const PtK& p, const Tb2dInv& v, const int s, const char* (&c)[2],
                                          const char* (&d)[2], double* __restrict__ sm, double (&PM)[2], double (&PC)[2],
                                          double (&ZP)[2], double (&DQ)[2], double (&DVC)[2], double (&DV)[2],
                                          double (&DVN)[2], double (&QM)[2], double (&QC)[2], double (&QN)[2],
                                          double (&D1C)[2], double (&NX)[2][2], double (&NY)[2], double (&NM)[2])
{
#define LD(ptr) (*(const double*)(ptr))
    constexpr int SLOTSZ = TB_Y * TB_X;
    typedef Tb2sStride<NXC, NYC> G;
#define rowB (G::row(p))
#define planeB (G::plane(p))
#define dplaneB (G::dplane(p))
    // ---- stage 1: first iteration of plane s for both rows (operands were loaded one step ahead) ----
    double D1N[2];
    {
        const double yp0 = v.adj ? PC[1] : NM[0];  // y+ of row 0
        const double ym1 = v.adj ? PC[0] : NM[1];  // y- of row 1
        const double L0 = bracket<MODE>(p, PC[0], NX[0][0], NX[0][1], NY[0], yp0, PM[0], ZP[0], DV[0]);
        const double L1 = bracket<MODE>(p, PC[1], NX[1][0], NX[1][1], ym1, NY[1], PM[1], ZP[1], DV[1]);
        pt_update<MODE>(p, L0, DQ[0], PC[0], D1N[0], QN[0]);
        pt_update<MODE>(p, L1, DQ[1], PC[1], D1N[1], QN[1]);
    }
    // ---- operands of the next plane (PM, DQ, NX, NY, NM are dead) ------------------------------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const char* a_dvn = NXC > 0 ? (c[r] + p.oDV) + planeB : c[r] + p.oDVn;
        const char* a_zp2 = NXC > 0 ? c[r] + 2 * planeB : c[r] + p.oZP2;
        const char* cn = c[r] + planeB;
        DVN[r] = LD(a_dvn);
        PM[r] = LD(a_zp2);
        DQ[r] = LD(d[r] + dplaneB);
        NX[r][0] = LD(cn - 8);
        NX[r][1] = LD(cn + 8);
        if (PF > 0) {
            prefetch_l2(a_zp2 + PF * planeB);
            prefetch_l2(d[r] + (1 + PF) * dplaneB);
            prefetch_l2(a_dvn + PF * planeB);
        }
    }
    NY[0] = LD(c[0] + planeB - rowB);
    NY[1] = LD(c[1] + planeB + rowB);
    if (!v.adj) {
        NM[0] = LD(c[0] + planeB + rowB);
        NM[1] = LD(c[1] + planeB - rowB);
    }
    if (v.xfix) {  // bc_x_Pr! / bc_xhydstatic! images (x-face columns only)
        QN[0] = xface(p, v.xfix > 1, s, QN[0]);
        QN[1] = xface(p, v.xfix > 1, s, QN[1]);
    }
    sm[SLOT * SLOTSZ] = QN[0];
    sm[SLOT * SLOTSZ + TB_X] = QN[1];
    // ---- stage 2: second iteration of plane s-1 (ring slot published by the previous barrier) ---------
    const int k2 = s - 1;
    {
        const double* r0 = sm + ((SLOT + 2) % 3) * SLOTSZ;
        const double* r1 = r0 + TB_X;
        if (k2 >= v.kb_own[0]) {
            const double L2 = bracket<MODE>(p, QC[0], r0[-1], r0[1], r0[-TB_X], QC[1], QM[0], QN[0], DVC[0]);
            double d2, u;
            pt_update<MODE>(p, L2, D1C[0], QC[0], d2, u);
            *(double*)(const_cast<char*>(d[0]) + p.oDP) = d2;
            *(double*)(const_cast<char*>(c[0]) + p.oPr) = u;
            if (v.edge[0] | (k2 == 1)) {
                const int i = v.bx + (int)threadIdx.x, j = v.by + 2 * (int)threadIdx.y;
                tb2s_images(p, v.PrN, i, j, k2, (k2 == 1 && !p.zlo_halo) ? 0 : -1, u, i == 1, i == p.nx - 2, j == 1,
                            j == p.ny - 2);
            }
        }
        if (k2 >= v.kb_own[1]) {
            const double L2 = bracket<MODE>(p, QC[1], r1[-1], r1[1], QC[0], r1[TB_X], QM[1], QN[1], DVC[1]);
            double d2, u;
            pt_update<MODE>(p, L2, D1C[1], QC[1], d2, u);
            *(double*)(const_cast<char*>(d[1]) + p.oDP) = d2;
            *(double*)(const_cast<char*>(c[1]) + p.oPr) = u;
            if (v.edge[1] | (k2 == 1)) {
                const int i = v.bx + (int)threadIdx.x, j = v.by + 2 * (int)threadIdx.y + 1;
                tb2s_images(p, v.PrN, i, j, k2, (k2 == 1 && !p.zlo_halo) ? 0 : -1, u, i == 1, i == p.nx - 2, j == 1,
                            j == p.ny - 2);
            }
        }
    }
    if (s == 1 && !p.zlo_halo) {  // bc_z!: q[0] is the image of q[1]
        QC[0] = QN[0];
        QC[1] = QN[1];
    }
    __syncthreads();
    if (s == v.top_own[0] || s == v.top_own[1]) {  // physical top face, see tb2s_step
        const double* r0 = sm + SLOT * SLOTSZ;
        const double* r1 = r0 + TB_X;
        const int i = v.bx + (int)threadIdx.x;
        const bool xl = i == 1, xh = i == p.nx - 2;
        if (s == v.top_own[0]) {
            const double L2 = bracket<MODE>(p, QN[0], r0[-1], r0[1], r0[-TB_X], QN[1], QC[0], QN[0], DV[0]);
            double d2, u;
            pt_update<MODE>(p, L2, D1N[0], QN[0], d2, u);
            *(double*)(const_cast<char*>(d[0]) + p.oDPt) = d2;
            const int j = v.by + 2 * (int)threadIdx.y;
            tb2s_images(p, v.PrN, i, j, s, p.nz - 1, u, xl, xh, j == 1, j == p.ny - 2);
            if (s == 1 && !p.zlo_halo) tb2s_images(p, v.PrN, i, j, 0, -1, u, xl, xh, j == 1, j == p.ny - 2);
        }
        if (s == v.top_own[1]) {
            const double L2 = bracket<MODE>(p, QN[1], r1[-1], r1[1], QN[0], r1[TB_X], QC[1], QN[1], DV[1]);
            double d2, u;
            pt_update<MODE>(p, L2, D1N[1], QN[1], d2, u);
            *(double*)(const_cast<char*>(d[1]) + p.oDPt) = d2;
            const int j = v.by + 2 * (int)threadIdx.y + 1;
            tb2s_images(p, v.PrN, i, j, s, p.nz - 1, u, xl, xh, j == 1, j == p.ny - 2);
            if (s == 1 && !p.zlo_halo) tb2s_images(p, v.PrN, i, j, 0, -1, u, xl, xh, j == 1, j == p.ny - 2);
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        D1C[r] = D1N[r];
        c[r] += planeB;
        d[r] += dplaneB;
    }
#undef LD
#undef rowB
#undef planeB
#undef dplaneB
}

template <int MODE, int TB_Y, int PF, int MINB, int NXC, int NYC>
__global__ void __launch_bounds__(TB_X* TB_Y / 2, MINB) pt_tb2d_kernel(const double* Pr, double* PrN, const double* dP,
                                                                       double* dPN, const double* divV, const PtK p)
{
    __shared__ double ring[3 * TB_Y * TB_X];
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * (TB_X - 2) + tx;
    const int ci = min(max(i, 1), nx - 2);
    int bz = blockIdx.z;
    if (p.reverse) bz = gridDim.z - 1 - bz;
    const int kb = p.kbeg + bz * p.zchunk;
    const int ke = min(kb + p.zchunk, p.kend);  // stage-2 planes [kb, ke)
    const int s0 = max(kb - 1, 1);
    const int s1 = min(ke, nz - 2);             // stage-1 planes [s0, s1]
    Tb2dInv v;
    v.PrN = PrN;
    v.bx = blockIdx.x * (TB_X - 2);
    v.by = blockIdx.y * (TB_Y - 2);
    const char* c[2];
    const char* d[2];
    int cj[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int tr = 2 * ty + r;                 // tile row
        const int j = v.by + tr;
        cj[r] = min(max(j, 1), ny - 2);
        const bool owner = (ci == i) && (cj[r] == j) && tx >= 1 && tx <= TB_X - 2 && tr >= 1 && tr <= TB_Y - 2;
        v.kb_own[r] = owner ? kb : 0x7fffffff;
        v.top_own[r] = (owner && !p.zhi_halo && ke == nz - 1) ? nz - 2 : -1;
        v.edge[r] = owner && ((i == 1) | (i == nx - 2) | (j == 1) | (j == ny - 2));
        NS3D_KEEP(v.kb_own[r]); NS3D_KEEP(v.top_own[r]); NS3D_KEEP(v.edge[r]);
        c[r] = (const char*)(Pr + (ptrdiff_t)s0 * nx * ny + (ptrdiff_t)cj[r] * nx + ci);
        d[r] = (const char*)(dP + ((ptrdiff_t)s0 - 1) * (nx - 2) * (ny - 2) + (ptrdiff_t)(cj[r] - 1) * (nx - 2) + (ci - 1));
    }
    v.adj = cj[1] == cj[0] + 1;
    v.xfix = (i == 0 && p.xlo_kind != X_NEUMANN) ? 1 : ((i == nx - 1 && p.xhi_kind != X_NEUMANN) ? 2 : 0);
    NS3D_KEEP(v.adj); NS3D_KEEP(v.xfix);
    int tslot = 2 * ty * TB_X + tx;
    NS3D_KEEP(tslot);
    double* sm = ring + tslot;
#define LD(ptr) (*(const double*)(ptr))
    const long long rowB = Tb2sStride<NXC, NYC>::row(p), planeB = Tb2sStride<NXC, NYC>::plane(p);
    double A[2], B[2], C[2], DQ[2], VA[2], VB[2], VC[2], QA[2], QB[2], QC[2], D1[2], NX[2][2], NY[2], NM[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        A[r] = LD(c[r] - planeB);  // Pr of planes s0-1, s0, s0+1
        B[r] = LD(c[r]);
        C[r] = LD(c[r] + planeB);
        DQ[r] = LD(d[r]);
        VA[r] = 0; VB[r] = LD(c[r] + p.oDV); VC[r] = 0;
        QA[r] = QB[r] = QC[r] = D1[r] = 0;
        NX[r][0] = LD(c[r] - 8);
        NX[r][1] = LD(c[r] + 8);
    }
    NY[0] = LD(c[0] - rowB);
    NY[1] = LD(c[1] + rowB);
    NM[0] = LD(c[0] + rowB);
    NM[1] = LD(c[1] - rowB);
#undef LD
    int s = s0;
    while (true) {
        tb2d_step<MODE, TB_Y, 0, PF, NXC, NYC>(p, v, s, c, d, sm, A, B, C, DQ, VA, VB, VC, QA, QB, QC, D1, NX, NY, NM);
        if (s == s1) break;
        ++s;
        tb2d_step<MODE, TB_Y, 1, PF, NXC, NYC>(p, v, s, c, d, sm, B, C, A, DQ, VB, VC, VA, QB, QC, QA, D1, NX, NY, NM);
        if (s == s1) break;
        ++s;
        tb2d_step<MODE, TB_Y, 2, PF, NXC, NYC>(p, v, s, c, d, sm, C, A, B, DQ, VC, VA, VB, QC, QA, QB, D1, NX, NY, NM);
        if (s == s1) break;
        ++s;
    }
}

// Byte displacements between the arrays of one pt_tb2s_kernel launch: launch constants, so the
// kernel adds them from the constant bank instead of carrying 64-bit differences in registers.
inline void tb2s_set_offsets(PtK& k, const double* Pr, const double* PrN, const double* dP, const double* dPN,
                             const double* divV)
{
    k.oDV = (const char*)divV - (const char*)Pr;
    k.oDVn = k.oDV + k.planeB;
    k.oZP2 = 2 * k.planeB;
    k.oPr = ((const char*)PrN - (const char*)Pr) - k.planeB;
    k.oDPt = (const char*)dPN - (const char*)dP;
    k.oDP = k.oDPt - k.dplaneB;
}

// ---- host-side geometry shared by ns3d_pt.cu and the host emulation (tests/emu/) -----------------

// The arithmetic part of the kernel parameters (everything that does not depend on the context).
inline void ptk_fill(const ns3d_pt_params* p, PtK* k)
{
    k->nx = p->nx; k->ny = p->ny; k->nz = p->nz;
    k->omd = 1.0 - p->damp;
    k->dtau = p->dtau;
    k->rdt = p->rho / p->dt;
    k->dx = p->dx; k->dy = p->dy; k->dz = p->dz;
    k->rdx = 1.0 / p->dx; k->rdy = 1.0 / p->dy; k->rdz = 1.0 / p->dz;
    k->rdx2 = 1.0 / (p->dx * p->dx); k->rdy2 = 1.0 / (p->dy * p->dy); k->rdz2 = 1.0 / (p->dz * p->dz);
    if (p->variant == NS3D_VARIANT_M) {
        k->xlo_kind = X_NEUMANN;
        k->xhi_kind = p->outlet_guard ? X_DIRICHLET : X_NEUMANN;
        k->xhi_val = p->outlet_val;
    } else {
        k->xlo_kind = k->xhi_kind = X_HYDRO;
        k->xlo_val = 100;
        k->rho_g = p->rho * p->g;
        k->hyd_dz = p->dz;
        k->hyd_nz = p->nz;
    }
    k->kbeg = 1;
    k->kend = p->nz - 1;
    k->faces = 0;
    k->reverse = 0;
    k->rowB = 8LL * p->nx;
    k->planeB = 8LL * p->nx * p->ny;
    k->dplaneB = 8LL * (p->nx - 2) * (p->ny - 2);
}

// The neighbours' buffers as this rank sees them (CUDA IPC mappings on the device, shared memory in
// the host emulation): {Pr, Pr shadow, dPrdτ, dPrdτ shadow} of the lower / upper neighbour (NULL
// without one), this rank's mailbox and the neighbours' mailboxes.
struct PeerPtrs {
    double* lo[4];
    double* hi[4];
    unsigned long long* mbox;
    unsigned long long* lo_mbox;
    unsigned long long* hi_mbox;
};

// Where a launch on a slab interface reads and writes in its neighbours' memory.  All ranks
// ping-pong in lockstep, so a neighbour's buffers play the role of this rank's: `w_new` is the
// index (0 user array, 1 shadow) of the Pr buffer this launch WRITES, `w_dp` (2 user, 3 shadow) of
// the dPrdτ buffer it READS; w_dp < 0 for the one-iteration kernel, which reads nothing remote.
inline void ptk_set_peers(PtK& k, const PeerPtrs& pp, int w_new, int w_dp)
{
    const ptrdiff_t sxy = (ptrdiff_t)k.nx * k.ny, dxy = (ptrdiff_t)(k.nx - 2) * (k.ny - 2);
    k.mbox = pp.mbox;
    k.peer_lo_plane = pp.lo[w_new] ? pp.lo[w_new] + (ptrdiff_t)(k.nz - 1) * sxy : nullptr;  // its halo plane nz-1
    k.peer_hi_plane = pp.hi[w_new];                                                        // its halo plane 0
    k.peer_lo_flag = pp.lo_mbox ? pp.lo_mbox + NS3D_MB_FLAG_HI : nullptr;
    k.peer_hi_flag = pp.hi_mbox ? pp.hi_mbox + NS3D_MB_FLAG_LO : nullptr;
    if (w_dp >= 0) {
        const int w_cur = 1 - w_new;  // the neighbours' CURRENT iterate
        k.peer_lo_cur = pp.lo[w_cur] ? pp.lo[w_cur] + (ptrdiff_t)(k.nz - 3) * sxy : nullptr;
        k.peer_hi_cur = pp.hi[w_cur] ? pp.hi[w_cur] + 2 * sxy : nullptr;
        k.peer_lo_dp = pp.lo[w_dp] ? pp.lo[w_dp] + (ptrdiff_t)(k.nz - 3) * dxy : nullptr;
        k.peer_hi_dp = pp.hi[w_dp];
    }
}

// Balanced z-chunks whose last one keeps at least two planes: on slabs a neighbour reads plane
// nz-3 (resp. 2) of this rank, and the CTAs that own it are the ones holding the hand-over flag.
inline void balance_chunks(PtK& k)
{
    const int n = k.kend - k.kbeg;
    int nch = (n + k.zchunk - 1) / k.zchunk;
    int len = (n + nch - 1) / nch;
    if (nch > 1 && n - (nch - 1) * len == 1) {
        nch -= 1;
        len = (n + nch - 1) / nch;
    }
    k.zchunk = len;
}

}  // namespace
