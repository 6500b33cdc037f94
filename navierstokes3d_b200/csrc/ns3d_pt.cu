// libns3d.so -- the hot loop: fused pseudo-transient (PT) pressure iteration.
//
// Reference (per PT iteration, M:459-463 / G:127-129): update_dPrdτ! (K5), update_Pr! (K6),
// set_bc_Pr! = 3-4 face kernels (K7) and up to three update_halo! calls: >= 5 synchronous
// launches and 7+ full-field passes.  Here ONE launch per iteration does all of it in
// 5 passes (read Pr, dPrdτ, ∇V; write Pr', dPrdτ): Pr ping-pongs between the caller's array
// and a context-owned shadow so that every thread reads a consistent old iterate, and the
// boundary conditions are folded in: after x,y,z zero-gradient copies every boundary point
// equals the new value at its index clamped into the interior, so the
// thread that owns an interior point next to a face also stores its mirror images.
//
// Thread mapping: x on threadIdx.x (coalesced rows), a (32 x BY) tile of interior columns
// per CTA, each thread marches `zchunk` planes along z keeping Pr[k-1], Pr[k], Pr[k+1] of
// its column in registers (2.5-D blocking); the x/y neighbours come through L1.
//
// Arithmetic (template MODE): see NS3D_PARITY / NS3D_FAST / NS3D_FASTEST in ns3d.h.  The file
// is compiled with --fmad=false; FMA appears only where fma() is written explicitly.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ns3d_internal.cuh"

namespace {

enum { X_NEUMANN = 0, X_DIRICHLET = 1, X_HYDRO = 2 };

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
    int zchunk_tb;  // chunk length of the two-iterations-per-launch kernel
};

// a / b with y = RN(1/b): one multiply + two FMAs (Markstein's correction step).
__device__ __forceinline__ double div3(double a, double b, double y)
{
    const double q = a * y;
    const double r = fma(-b, q, a);
    return fma(r, y, q);
}

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

// ---- peer-memory halo protocol (device side) --------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long atom_add_acq_rel_gpu(unsigned long long* p, unsigned long long v)
{
    unsigned long long old;
    asm volatile("atom.add.acq_rel.gpu.global.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}

// Spins until the neighbour on `side` (0 lower, 1 upper) has finished the face work of every
// launch this rank has finished: then its stores into our halo plane have landed and it no
// longer reads the halo plane of its own that we are about to overwrite.  Bounded: a
// neighbour that never answers raises NS3D_MB_ERROR instead of hanging the GPU.
__device__ __forceinline__ void wait_neighbour(unsigned long long* mbox, int side)
{
    const unsigned long long need = ld_acquire_sys(mbox + NS3D_MB_EPOCH_LO + side);
    const long long t0 = clock64();
    while (ld_acquire_sys(mbox + NS3D_MB_FLAG_LO + side) < need) {
        if (clock64() - t0 > 8000000000LL) {  // ~4 s
            mbox[NS3D_MB_ERROR] = 1ULL + side;
            break;
        }
    }
}

// Last face CTA of this launch on `side`: close the epoch and tell the neighbour.
__device__ __forceinline__ void signal_neighbour(unsigned long long* mbox, int side, unsigned long long* peer_flag,
                                                 unsigned nface)
{
    const unsigned long long old = atom_add_acq_rel_gpu(mbox + NS3D_MB_ARRIVE_LO + side, 1ULL);
    if (old + 1 == nface) {
        mbox[NS3D_MB_ARRIVE_LO + side] = 0ULL;
        const unsigned long long e = mbox[NS3D_MB_EPOCH_LO + side] + 1ULL;
        mbox[NS3D_MB_EPOCH_LO + side] = e;
        __threadfence_system();
        st_release_sys(peer_flag, e);
    }
}

// After a chunk of launches: the halo planes of the current iterate are complete once both
// neighbours have signalled the epoch this rank has reached.
__global__ void pt_halo_wait_kernel(unsigned long long* mbox, int has_lo, int has_hi)
{
    if (threadIdx.x == 0) {
        if (has_lo) wait_neighbour(mbox, 0);
        if (has_hi) wait_neighbour(mbox, 1);
    }
}

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
// Two PT iterations per launch (temporal blocking, opt-in: ns3d_set_option("tb2", 1)).
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

// compute_res! + abs + maximum (K8 + K8') in one pass, no Rp array: max over the interior of
// the bit pattern of |bracket| (NaN-propagating, see absbits()).
template <int MODE>
__global__ void __launch_bounds__(256) pt_residual_kernel(const double* __restrict__ Pr,
                                                          const double* __restrict__ divV, const PtK p,
                                                          unsigned long long* __restrict__ out)
{
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    unsigned long long m = 0ULL;
    if (i <= nx - 2 && j <= ny - 2) {
        const int kb = p.kbeg + blockIdx.z * p.zchunk;
        const int ke = min(kb + p.zchunk, p.kend);
        const size_t sxy = (size_t)nx * ny;
        const double* c = Pr + idx3(i, j, kb, nx, ny);
        const double* dv = divV + idx3(i, j, kb, nx, ny);
        double pm = c[-(ptrdiff_t)sxy];
        double pc = c[0];
        for (int k = kb; k < ke; ++k) {
            const double pp = c[sxy];
            const double L = bracket<MODE>(p, pc, c[-1], c[1], c[-nx], c[nx], pm, pp, dv[0]);
            const unsigned long long b = absbits(L);
            m = b > m ? b : m;
            pm = pc;
            pc = pp;
            c += sxy;
            dv += sxy;
        }
    }
    block_max_to_global(m, out);
}

int make_ptk(ns3d_ctx* ctx, const ns3d_pt_params* p, PtK* k)
{
    if (!p) return ns3d_fail(ctx, NS3D_EINVAL, "pt: params is NULL");
    if (p->nx < 3 || p->ny < 3 || p->nz < 3) return ns3d_fail(ctx, NS3D_EINVAL, "pt: grid must be at least 3^3");
    if (p->variant != NS3D_VARIANT_M && p->variant != NS3D_VARIANT_G)
        return ns3d_fail(ctx, NS3D_EINVAL, "pt: unknown variant %d", p->variant);
    memset(k, 0, sizeof *k);
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
    // z-slab interfaces (variant M only: the G script is single-GPU)
    k->zlo_halo = ctx->nranks > 1 && ctx->rank > 0;
    k->zhi_halo = ctx->nranks > 1 && ctx->rank < ctx->nranks - 1;
    int zc = p->zchunk;
    if (zc <= 0) {
        // enough CTAs for several waves on 148 SMs, but chunks long enough to amortise the
        // two extra plane loads at the start of every chunk
        const long long xy = (long long)cdiv(p->nx - 2, 32) * cdiv(p->ny - 2, 8);
        zc = 16;
        while (zc > 4 && xy * cdiv(p->nz - 2, zc) < 2LL * 8 * ctx->num_sms) zc /= 2;  // measured: profiles/r01_*sweep*
    }
    k->zchunk = zc;
    k->zchunk_tb = p->zchunk > 0 ? p->zchunk : std::max(zc, 16);  // two extra stage-1 planes per chunk: keep them long
    k->kbeg = 1;
    k->kend = p->nz - 1;
    k->faces = 0;
    k->reverse = 0;
    k->rowB = 8LL * p->nx;
    k->planeB = 8LL * p->nx * p->ny;
    k->dplaneB = 8LL * (p->nx - 2) * (p->ny - 2);
    // serpentine pays while a good part of the 4-field working set can stay in L2 (measured:
    // +5..8 % at 255x153x153, neutral to -1 % at 511^3; profiles/r01_v4_sweep_serpentine.jsonl)
    k->serpentine = ctx->opt_serpentine < 0 ? (4.0 * 8.0 * p->nx * p->ny * p->nz < 6.0 * ctx->l2_bytes)
                                            : ctx->opt_serpentine;
    return NS3D_OK;
}

// The hot kernel prefetches past the end of its arrays into the allocator's padding, so the
// fields of the fused loop must be blocks handed out by ns3d_zeros of this context.
int check_owned(ns3d_ctx* ctx, const double* Pr, const double* dP, const double* divV)
{
    for (const double* a : {Pr, dP, divV})
        if (!ctx->allocs.count((void*)a))
            return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_pt_*: field %p was not allocated by ns3d_zeros of this context", (const void*)a);
    return NS3D_OK;
}

int ensure_shadow(ns3d_ctx* ctx, size_t count, size_t pad_bytes)
{
    count += pad_bytes / sizeof(double);
    if (ctx->pr_shadow_count >= count) return NS3D_OK;
    if (ctx->pr_shadow) {
        NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        auto it = ctx->p2p_map.find(ctx->pr_shadow);  // the neighbours' old shadows are going away too
        if (it != ctx->p2p_map.end()) {
            if (it->second.first) cudaIpcCloseMemHandle(it->second.first);
            if (it->second.second) cudaIpcCloseMemHandle(it->second.second);
            ctx->p2p_map.erase(it);
        }
        NS3D_CUDA(ctx, cudaFree(ctx->pr_shadow));
        ctx->pr_shadow = nullptr;
        ctx->pr_shadow_count = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->pr_shadow, count * sizeof(double));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ns3d_fail(ctx, NS3D_ENOMEM, "pt: cannot allocate the Pr shadow (%zu B)", count * sizeof(double));
    }
    ctx->pr_shadow_count = count;
    return NS3D_OK;
}

inline dim3 pt_block() { return dim3(32, 8, 1); }
inline dim3 pt_grid(const PtK& k)
{
    const unsigned gz = k.faces ? (k.nz > 3 ? 2u : 1u) : cdiv(k.kend - k.kbeg, k.zchunk);
    return dim3(cdiv(k.nx - 2, 32), cdiv(k.ny - 2, 8), gz);
}

int launch_iter(ns3d_ctx* ctx, cudaStream_t st, const PtK& k, const double* cur, double* nxt, double* dP,
                const double* divV)
{
#define PT_LAUNCH(MODE, MINB)                                                                      \
    do {                                                                                           \
        if (k.mbox) pt_iter_kernel<MODE, MINB, true><<<pt_grid(k), pt_block(), 0, st>>>(cur, nxt, dP, divV, k);  \
        else pt_iter_kernel<MODE, MINB, false><<<pt_grid(k), pt_block(), 0, st>>>(cur, nxt, dP, divV, k);        \
    } while (0)
#define PT_LAUNCH_MODE(MINB)                                  \
    switch (ctx->mode) {                                      \
        case NS3D_PARITY: PT_LAUNCH(NS3D_PARITY, MINB); break; \
        case NS3D_FAST: PT_LAUNCH(NS3D_FAST, MINB); break;     \
        default: PT_LAUNCH(NS3D_FASTEST, MINB); break;         \
    }
    // CTAs per SM the kernel is compiled for (register cap 65536 / (256 * MINB)); the defaults are
    // the measured optimum per mode (profiles/r01_v4_sweep_minb.jsonl): the largest occupancy
    // that does not spill.
    int minb = ctx->opt_pt_minb;
    if (minb == 0) minb = ctx->mode == NS3D_PARITY ? 3 : 5;
    if (k.mbox && minb > 4) minb = 4;  // the peer-store variant spills at 48 registers
    switch (minb) {
        case 3: PT_LAUNCH_MODE(3); break;
        case 5: PT_LAUNCH_MODE(5); break;
        case 6: PT_LAUNCH_MODE(6); break;
        default: PT_LAUNCH_MODE(4); break;
    }
#undef PT_LAUNCH_MODE
#undef PT_LAUNCH
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// Neighbours' buffers mapped through CUDA IPC (see peer_prepare).
struct PeerBufs {
    bool on = false;
    bool tb2 = false;                                       // dPrdτ buffers mapped too
    double* lo[4] = {nullptr, nullptr, nullptr, nullptr};  // lower neighbour's {Pr, Pr shadow, dPrdτ, dPrdτ shadow}
    double* hi[4] = {nullptr, nullptr, nullptr, nullptr};  // upper neighbour's
    const double* dP_user = nullptr;
};

// Balanced z-chunks whose last one keeps at least two planes: on slabs a neighbour reads plane
// nz-3 (resp. 2) of this rank, and the CTAs that own it are the ones holding the hand-over flag.
void balance_chunks(PtK& k)
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

int launch_tb2(ns3d_ctx* ctx, cudaStream_t st, const PtK& k_in, const double* cur, double* nxt, const double* dpc,
               double* dpn, const double* divV, const PeerBufs& pb, const double* Pr_user, bool peer)
{
    PtK k = k_in;
    if (!k.faces) balance_chunks(k);
    if (peer) {
        const int wp = (nxt == Pr_user) ? 0 : 1;        // neighbours' NEW iterate: same role as ours
        const int wc = 1 - wp;                          // ... CURRENT iterate
        const int wd = (dpc == pb.dP_user) ? 2 : 3;     // ... CURRENT dPrdτ
        const ptrdiff_t sxy = (ptrdiff_t)k.nx * k.ny, dxy = (ptrdiff_t)(k.nx - 2) * (k.ny - 2);
        k.mbox = ctx->mbox;
        k.peer_lo_plane = pb.lo[wp] ? pb.lo[wp] + (ptrdiff_t)(k.nz - 1) * sxy : nullptr;
        k.peer_hi_plane = pb.hi[wp];
        k.peer_lo_cur = pb.lo[wc] ? pb.lo[wc] + (ptrdiff_t)(k.nz - 3) * sxy : nullptr;
        k.peer_hi_cur = pb.hi[wc] ? pb.hi[wc] + 2 * sxy : nullptr;
        k.peer_lo_dp = pb.lo[wd] ? pb.lo[wd] + (ptrdiff_t)(k.nz - 3) * dxy : nullptr;
        k.peer_hi_dp = pb.hi[wd];
        k.peer_lo_flag = ctx->peer_mbox[0] ? ctx->peer_mbox[0] + NS3D_MB_FLAG_HI : nullptr;
        k.peer_hi_flag = ctx->peer_mbox[1] ? ctx->peer_mbox[1] + NS3D_MB_FLAG_LO : nullptr;
    }
    const int ty = k.mbox ? 16 : (ctx->opt_tb2_ty == 8 ? 8 : (ctx->opt_tb2_ty == 32 ? 32 : 16));
    const dim3 blk(TB_X, ty, 1);
    const dim3 grd(cdiv(k.nx - 2, TB_X - 2), cdiv(k.ny - 2, ty - 2), k.faces ? 2u : cdiv(k.kend - k.kbeg, k.zchunk));
#define TB_LAUNCH(MODE)                                                                                        \
    do {                                                                                                       \
        if (k.mbox) pt_tb2_kernel<MODE, 16, true><<<grd, blk, 0, st>>>(cur, nxt, dpc, dpn, divV, k);          \
        else if (ty == 8) pt_tb2_kernel<MODE, 8, false><<<grd, blk, 0, st>>>(cur, nxt, dpc, dpn, divV, k);    \
        else if (ty == 32) pt_tb2_kernel<MODE, 32, false><<<grd, blk, 0, st>>>(cur, nxt, dpc, dpn, divV, k);  \
        else pt_tb2_kernel<MODE, 16, false><<<grd, blk, 0, st>>>(cur, nxt, dpc, dpn, divV, k);                \
    } while (0)
    switch (ctx->mode) {
        case NS3D_PARITY: TB_LAUNCH(NS3D_PARITY); break;
        case NS3D_FAST: TB_LAUNCH(NS3D_FAST); break;
        default: TB_LAUNCH(NS3D_FASTEST); break;
    }
#undef TB_LAUNCH
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// Two iterations per launch is the default path wherever it applies (single rank, or slabs with
// the peer-memory halo path): measured sustained gains +11 % at 255x153x153 (T_eff 6.36 TB/s),
// +27 % at 511^3 (7.9-8.1 TB/s, above the HBM copy peak), +9.5 % per GPU on two slabs
// (profiles/r01_tb2_*.jsonl).  ns3d_set_option("tb2", 0) selects the one-iteration kernel.
bool use_tb2(const ns3d_ctx* ctx, const ns3d_pt_params* p, bool peer_on)
{
    (void)p;
    if (ctx->opt_tb2 == 0) return false;
    if (ctx->nranks > 1 && !peer_on) return false;  // slabs: needs the peer-memory path
    return true;
}

int ensure_dp_shadow(ns3d_ctx* ctx, size_t count)
{
    if (ctx->dp_shadow_count >= count) return NS3D_OK;
    if (ctx->dp_shadow) {
        NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        auto it = ctx->p2p_map.find(ctx->dp_shadow);
        if (it != ctx->p2p_map.end()) {
            if (it->second.first) cudaIpcCloseMemHandle(it->second.first);
            if (it->second.second) cudaIpcCloseMemHandle(it->second.second);
            ctx->p2p_map.erase(it);
        }
        NS3D_CUDA(ctx, cudaFree(ctx->dp_shadow));
        ctx->dp_shadow = nullptr;
        ctx->dp_shadow_count = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->dp_shadow, count * sizeof(double));
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ns3d_fail(ctx, NS3D_ENOMEM, "pt: cannot allocate the dPrdtau shadow (%zu B)", count * sizeof(double));
    }
    ctx->dp_shadow_count = count;
    return NS3D_OK;
}

int launch_residual(ns3d_ctx* ctx, const PtK& k, const double* cur, const double* divV)
{
    NS3D_CUDA(ctx, cudaMemsetAsync(ctx->d_maxbits, 0, sizeof(unsigned long long), ctx->stream));
    switch (ctx->mode) {
        case NS3D_PARITY: pt_residual_kernel<NS3D_PARITY><<<pt_grid(k), pt_block(), 0, ctx->stream>>>(cur, divV, k, ctx->d_maxbits); break;
        case NS3D_FAST: pt_residual_kernel<NS3D_FAST><<<pt_grid(k), pt_block(), 0, ctx->stream>>>(cur, divV, k, ctx->d_maxbits); break;
        default: pt_residual_kernel<NS3D_FASTEST><<<pt_grid(k), pt_block(), 0, ctx->stream>>>(cur, divV, k, ctx->d_maxbits); break;
    }
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// One PT iteration cur -> nxt including update_halo!(Pr) (replaces M:462 and M:182; the third
// call, update_halo!(∇V) M:460, is redundant: ∇V does not change inside the loop).
//
// Single rank: one launch.  z-slabs: the two planes a slab sends (1 and nz-2) are updated by a
// small launch on the high-priority communication stream, followed there by their exchange,
// while the remaining planes are updated on the main stream; the main stream joins before
// anything reads the new halos.  Event protocol (ev_a = "main stream finished reading the
// iterate that is about to be overwritten", ev_b = "faces + halos of the new iterate are in
// place"):
//     comm:  wait ev_a(n-1)  faces(n)  exchange(n)  record ev_b(n)
//     main:  interior(n)     record ev_a(n)         wait ev_b(n)
// The reference does the same work with three blocking update_halo! calls per iteration and
// no overlap.
int pt_begin(ns3d_ctx* ctx)
{
    if (ctx->nranks > 1) NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    return NS3D_OK;
}

// Peer-memory path usable for this solve?  Maps the neighbours' copies of both ping-pong buffers
// on first use (collective: every rank reaches this point with its own Pr / shadow).
int peer_prepare(ns3d_ctx* ctx, const PtK& k, double* Pr, double* shadow, PeerBufs* pb)
{
    pb->on = false;
    if (ctx->nranks == 1 || !ctx->opt_p2p || !ctx->p2p_ready || k.nz < 6) return NS3D_OK;
    void *l0, *h0, *l1, *h1;
    NS3D_TRY(ns3d_internal_p2p_map(ctx, Pr, &l0, &h0));
    NS3D_TRY(ns3d_internal_p2p_map(ctx, shadow, &l1, &h1));
    pb->lo[0] = (double*)l0; pb->hi[0] = (double*)h0;
    pb->lo[1] = (double*)l1; pb->hi[1] = (double*)h1;
    pb->on = true;
    return NS3D_OK;
}

// The two-iterations-per-launch kernel also reads the neighbours' dPrdτ (both ping-pong buffers).
int peer_prepare_tb2(ns3d_ctx* ctx, double* dP, double* dp_shadow, PeerBufs* pb)
{
    if (!pb->on) return NS3D_OK;
    void *l2, *h2, *l3, *h3;
    NS3D_TRY(ns3d_internal_p2p_map(ctx, dP, &l2, &h2));
    NS3D_TRY(ns3d_internal_p2p_map(ctx, dp_shadow, &l3, &h3));
    pb->lo[2] = (double*)l2; pb->hi[2] = (double*)h2;
    pb->lo[3] = (double*)l3; pb->hi[3] = (double*)h3;
    pb->dP_user = dP;
    pb->tb2 = true;
    return NS3D_OK;
}

int pt_iteration(ns3d_ctx* ctx, const PtK& k, const double* cur, double* nxt, double* dP, const double* divV,
                 const PeerBufs& pb, const double* Pr_user)
{
    if (ctx->nranks == 1) return launch_iter(ctx, ctx->stream, k, cur, nxt, dP, divV);
    if (pb.on && pb.tb2) {
        // Next to two-iterations-per-launch kernels (whose face CTAs read the neighbour's planes 2
        // and nz-3) a single iteration is one unsplit launch: the CTAs that own those planes are
        // then the ones that signal.
        PtK q = k;
        const int which = (nxt == Pr_user) ? 0 : 1;
        const ptrdiff_t sxy = (ptrdiff_t)k.nx * k.ny;
        balance_chunks(q);
        q.mbox = ctx->mbox;
        q.peer_lo_plane = pb.lo[which] ? pb.lo[which] + (ptrdiff_t)(k.nz - 1) * sxy : nullptr;
        q.peer_hi_plane = pb.hi[which];
        q.peer_lo_flag = ctx->peer_mbox[0] ? ctx->peer_mbox[0] + NS3D_MB_FLAG_HI : nullptr;
        q.peer_hi_flag = ctx->peer_mbox[1] ? ctx->peer_mbox[1] + NS3D_MB_FLAG_LO : nullptr;
        return launch_iter(ctx, ctx->stream, q, cur, nxt, dP, divV);
    }
    if (pb.on) {
        // The update of the two planes a slab sends and their delivery are ONE kernel: the face
        // CTAs store the new values into the neighbour's halo plane (peer memory over NVLink) and
        // hand over with mailbox flags.  It runs on the high-priority stream next to the launch
        // that updates the other planes; kernel-only, so whole chunks replay as a CUDA graph.
        PtK q = k, inner = k;
        const int which = (nxt == Pr_user) ? 0 : 1;  // all ranks ping-pong in lockstep
        const ptrdiff_t sxy = (ptrdiff_t)k.nx * k.ny;
        q.faces = 1;
        q.mbox = ctx->mbox;
        q.peer_lo_plane = pb.lo[which] ? pb.lo[which] + (ptrdiff_t)(k.nz - 1) * sxy : nullptr;
        q.peer_hi_plane = pb.hi[which];
        q.peer_lo_flag = ctx->peer_mbox[0] ? ctx->peer_mbox[0] + NS3D_MB_FLAG_HI : nullptr;
        q.peer_hi_flag = ctx->peer_mbox[1] ? ctx->peer_mbox[1] + NS3D_MB_FLAG_LO : nullptr;
        inner.kbeg = 2;
        inner.kend = k.nz - 2;
        NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
        NS3D_TRY(launch_iter(ctx, ctx->comm_stream, q, cur, nxt, dP, divV));
        NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->comm_stream));
        NS3D_TRY(launch_iter(ctx, ctx->stream, inner, cur, nxt, dP, divV));
        NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
        NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
        return NS3D_OK;
    }
    double* f[1] = {nxt};
    if (k.nz < 6) {  // too thin to split: update, then exchange, on one stream
        NS3D_TRY(launch_iter(ctx, ctx->stream, k, cur, nxt, dP, divV));
        return ns3d_internal_halo_z(ctx, ctx->stream, f, &k.nx, &k.ny, &k.nz, 1, k.nz);
    }
    PtK faces = k, inner = k;
    faces.faces = 1;
    inner.kbeg = 2;
    inner.kend = k.nz - 2;
    NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
    NS3D_TRY(launch_iter(ctx, ctx->comm_stream, faces, cur, nxt, dP, divV));
    NS3D_TRY(ns3d_internal_halo_z(ctx, ctx->comm_stream, f, &k.nx, &k.ny, &k.nz, 1, k.nz));
    NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->comm_stream));
    NS3D_TRY(launch_iter(ctx, ctx->stream, inner, cur, nxt, dP, divV));
    NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
    return NS3D_OK;
}

// The halos of the current iterate are complete once both neighbours have caught up.
int peer_join(ns3d_ctx* ctx, const PtK& k, const PeerBufs& pb)
{
    if (!pb.on) return NS3D_OK;
    pt_halo_wait_kernel<<<1, 32, 0, ctx->stream>>>(ctx->mbox, k.zlo_halo, k.zhi_halo);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// ---------------------------------------------------------------------------------------------
// n PT iterations, replayed as a CUDA graph when possible.
//
// Per iteration the host would otherwise issue 1 launch (single rank) or 2 launches, 2 event
// records, 2 stream waits and an NCCL group (slabs): ~5-30 us of CPU time against a 40-50 us
// kernel at 255x153x153 -- with 8 ranks on one host that made the loop launch-bound (measured:
// 52 us/iteration on 8 GPUs vs 45 on 4).  A chunk of nchk iterations, including the forked
// communication stream and the NCCL send/recv, is captured once and replayed with one call.
// ---------------------------------------------------------------------------------------------
struct PtGraph {
    cudaGraphExec_t exec = nullptr;
    PtK k;
    const double* cur = nullptr;
    double* nxt = nullptr;
    double* dP = nullptr;
    double* dPn = nullptr;
    const double* divV = nullptr;
    int n = 0, parity = 0, mode = 0, minb = 0;
    bool p2p = false;
    long long kernels = 0;
};
struct PtGraphCache {
    PtGraph slot[4];
    int next = 0;
};

int run_direct(ns3d_ctx* ctx, PtK& k, double*& cur, double*& nxt, double*& dP, double*& dPn, const double* divV, int n,
               int iter0, const PeerBufs& pb, const double* Pr_user)
{
    NS3D_TRY(pt_begin(ctx));
    int q = 0;
    if (dPn) {  // two iterations per launch; Pr and dPrdτ both ping-pong
        PtK k2 = k;
        k2.zchunk = k.zchunk_tb;
        const bool peer = pb.on && pb.tb2;
        const int zf = 8;  // planes per face chunk of the split slab launch
        const bool split = peer && (k.nz - 2) >= 2 * zf + 4;
        for (; q + 2 <= n; q += 2) {
            k2.reverse = k.serpentine && (((iter0 + q) >> 1) & 1);
            if (split) {
                // slabs: the two chunks next to the interfaces (peer loads/stores, mailbox flags) run
                // on the high-priority stream beside the launch that updates the other planes with
                // the plain variant; same event protocol as the single-iteration path
                PtK f = k2, in = k2;
                f.faces = 1;
                f.zchunk = zf;
                in.kbeg = 1 + zf;
                in.kend = k.nz - 1 - zf;
                NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
                NS3D_TRY(launch_tb2(ctx, ctx->comm_stream, f, cur, nxt, dP, dPn, divV, pb, Pr_user, true));
                NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->comm_stream));
                NS3D_TRY(launch_tb2(ctx, ctx->stream, in, cur, nxt, dP, dPn, divV, pb, Pr_user, false));
                NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
                NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
            } else {
                NS3D_TRY(launch_tb2(ctx, ctx->stream, k2, cur, nxt, dP, dPn, divV, pb, Pr_user, peer));
            }
            double* t = cur; cur = nxt; nxt = t;
            t = dP; dP = dPn; dPn = t;
        }
    }
    if (q > 0 && q < n) NS3D_TRY(pt_begin(ctx));  // the split path forks from HERE, not from the chunk start
    for (; q < n; ++q) {
        k.reverse = k.serpentine && ((iter0 + q) & 1);
        NS3D_TRY(pt_iteration(ctx, k, cur, nxt, dP, divV, pb, Pr_user));
        double* t = cur; cur = nxt; nxt = t;
    }
    return peer_join(ctx, k, pb);
}

// Runs iterations iter0 .. iter0+n-1 (0-based count since the start of the solve).
int run_iterations(ns3d_ctx* ctx, PtK& k, double*& cur, double*& nxt, double*& dP, double*& dPn, const double* divV,
                   int n, int iter0, const PeerBufs& pb, const double* Pr_user)
{
    // NCCL send/recv captured in a graph drags host-callback nodes along (proxy progress) and
    // replays slower than the stream version (measured 55.9 vs 44.9 us/iteration on 2 GPUs), so
    // only kernel-only iterations are replayed as graphs.
    const bool graphable = ctx->opt_graphs && n >= 8 && (ctx->nranks == 1 || pb.on);
    if (!graphable) return run_direct(ctx, k, cur, nxt, dP, dPn, divV, n, iter0, pb, Pr_user);
    if (!ctx->pt_graphs) ctx->pt_graphs = new PtGraphCache();
    PtGraphCache* cache = (PtGraphCache*)ctx->pt_graphs;
    PtK key;
    memcpy(&key, &k, sizeof key);  // byte copy: the cache compares with memcmp (padding included)
    key.reverse = 0;
    PtGraph* g = nullptr;
    for (PtGraph& c : cache->slot)
        if (c.exec && c.cur == cur && c.nxt == nxt && c.dP == dP && c.dPn == dPn && c.divV == divV && c.n == n &&
            c.parity == (iter0 & 3) && c.mode == ctx->mode && c.minb == ctx->opt_pt_minb && c.p2p == pb.on &&
            !memcmp(&c.k, &key, sizeof key))
            g = &c;
    if (!g) {
        g = &cache->slot[cache->next];
        cache->next = (cache->next + 1) % 4;
        if (g->exec) {
            cudaGraphExecDestroy(g->exec);
            g->exec = nullptr;
        }
        double *ccur = cur, *cnxt = nxt, *cdp = dP, *cdpn = dPn;
        const long long l0 = ctx->launches;
        NS3D_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = run_direct(ctx, k, ccur, cnxt, cdp, cdpn, divV, n, iter0, pb, Pr_user);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        const long long captured = ctx->launches - l0;
        ctx->launches = l0;  // nothing ran yet
        if (rc != NS3D_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (rc != NS3D_OK) return rc;
            return ns3d_fail(ctx, NS3D_ECUDA, "PT graph capture failed: %s", cudaGetErrorString(e));
        }
        const cudaError_t e2 = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) {
            g->exec = nullptr;
            return ns3d_fail(ctx, NS3D_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
        }
        memcpy(&g->k, &key, sizeof key);
        g->cur = cur; g->nxt = nxt; g->dP = dP; g->dPn = dPn; g->divV = divV; g->n = n;
        g->parity = iter0 & 3; g->p2p = pb.on; g->mode = ctx->mode; g->minb = ctx->opt_pt_minb; g->kernels = captured;
    }
    NS3D_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->kernels;
    if (dPn) {  // n/2 double launches swap both pairs, a trailing single launch swaps Pr only
        if ((n >> 1) & 1) {
            double* t = cur; cur = nxt; nxt = t;
            t = dP; dP = dPn; dPn = t;
        }
        if (n & 1) {
            double* t = cur; cur = nxt; nxt = t;
        }
    } else if (n & 1) {
        double* t = cur; cur = nxt; nxt = t;
    }
    return NS3D_OK;
}

}  // namespace

void ns3d_internal_pt_free_graphs(ns3d_ctx* ctx)
{
    PtGraphCache* cache = (PtGraphCache*)ctx->pt_graphs;
    if (!cache) return;
    for (PtGraph& c : cache->slot)
        if (c.exec) cudaGraphExecDestroy(c.exec);
    delete cache;
    ctx->pt_graphs = nullptr;
}

extern "C" int ns3d_pt_solve(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV,
                             const ns3d_pt_params* p, int* h_iters, double* h_err_hist, int err_cap,
                             int* h_nchecks)
{
    NS3D_CHECK_CTX(ctx);
    if (!Pr || !dPrdtau || !divV) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_pt_solve: NULL field");
    PtK k;
    NS3D_TRY(make_ptk(ctx, p, &k));
    if (p->nchk <= 0 || p->niter < 0) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_pt_solve: bad niter/nchk");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)p->nx * p->ny * p->nz;
    NS3D_TRY(check_owned(ctx, Pr, dPrdtau, divV));
    NS3D_TRY(ensure_shadow(ctx, n, 3 * (size_t)p->nx * p->ny * sizeof(double) + 256));
    double* cur = Pr;
    double* nxt = ctx->pr_shadow;
    int iters = 0, nc = 0;
    PeerBufs pb;
    NS3D_TRY(peer_prepare(ctx, k, Pr, ctx->pr_shadow, &pb));
    double *dpc = dPrdtau, *dpn = nullptr;
    if (use_tb2(ctx, p, pb.on)) {
        const size_t nd = (size_t)(p->nx - 2) * (p->ny - 2) * (p->nz - 2) + 3 * (size_t)(p->nx - 2) * (p->ny - 2) + 32;
        NS3D_TRY(ensure_dp_shadow(ctx, nd));
        dpn = ctx->dp_shadow;
        NS3D_TRY(peer_prepare_tb2(ctx, dPrdtau, ctx->dp_shadow, &pb));
    }
    while (iters < p->niter) {
        const int chunk = std::min(p->nchk - iters % p->nchk, p->niter - iters);  // up to the next check
        NS3D_TRY(run_iterations(ctx, k, cur, nxt, dpc, dpn, divV, chunk, iters, pb, Pr));
        iters += chunk;
        if (iters % p->nchk == 0) {
            NS3D_TRY(launch_residual(ctx, k, cur, divV));
            double m = 0.0;
            NS3D_TRY(ns3d_internal_read_max(ctx, &m));
            const double err = m * p->err_num / p->err_den;  // max*ly^2/psc  M:466
            if (h_err_hist && nc < err_cap) h_err_hist[nc] = err;
            ++nc;
            if (err < p->eps_it || !std::isfinite(err)) break;  // M:469
        }
    }
    if (cur != Pr) NS3D_CUDA(ctx, cudaMemcpyAsync(Pr, cur, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    if (dpc != dPrdtau)
        NS3D_CUDA(ctx, cudaMemcpyAsync(dPrdtau, dpc, (size_t)(p->nx - 2) * (p->ny - 2) * (p->nz - 2) * sizeof(double),
                                       cudaMemcpyDeviceToDevice, ctx->stream));
    if (pb.on)
        NS3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_maxbits + 3, ctx->mbox + NS3D_MB_ERROR, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (pb.on && ctx->h_maxbits[3] != 0ULL)
        return ns3d_fail(ctx, NS3D_ECOMM, "peer-memory halo exchange: neighbour %s did not answer within the spin limit",
                         ctx->h_maxbits[3] == 1ULL ? "below" : "above");
    if (h_iters) *h_iters = iters;
    if (h_nchecks) *h_nchecks = nc;
    return NS3D_OK;
}

extern "C" int ns3d_pt_iterate(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV,
                               const ns3d_pt_params* p, int n_iter)
{
    NS3D_CHECK_CTX(ctx);
    if (!Pr || !dPrdtau || !divV || n_iter < 0) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_pt_iterate: bad argument");
    PtK k;
    NS3D_TRY(make_ptk(ctx, p, &k));
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)p->nx * p->ny * p->nz;
    NS3D_TRY(check_owned(ctx, Pr, dPrdtau, divV));
    NS3D_TRY(ensure_shadow(ctx, n, 3 * (size_t)p->nx * p->ny * sizeof(double) + 256));
    double* cur = Pr;
    double* nxt = ctx->pr_shadow;
    PeerBufs pb;
    NS3D_TRY(peer_prepare(ctx, k, Pr, ctx->pr_shadow, &pb));
    double *dpc = dPrdtau, *dpn = nullptr;
    if (use_tb2(ctx, p, pb.on)) {
        const size_t nd = (size_t)(p->nx - 2) * (p->ny - 2) * (p->nz - 2) + 3 * (size_t)(p->nx - 2) * (p->ny - 2) + 32;
        NS3D_TRY(ensure_dp_shadow(ctx, nd));
        dpn = ctx->dp_shadow;
        NS3D_TRY(peer_prepare_tb2(ctx, dPrdtau, ctx->dp_shadow, &pb));
    }
    NS3D_TRY(run_iterations(ctx, k, cur, nxt, dpc, dpn, divV, n_iter, 0, pb, Pr));
    if (dpc != dPrdtau)
        NS3D_CUDA(ctx, cudaMemcpyAsync(dPrdtau, dpc, (size_t)(p->nx - 2) * (p->ny - 2) * (p->nz - 2) * sizeof(double),
                                       cudaMemcpyDeviceToDevice, ctx->stream));
    if (cur != Pr) NS3D_CUDA(ctx, cudaMemcpyAsync(Pr, cur, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return NS3D_OK;
}

// One time step, M:449-477 / G:121-142.
extern "C" int ns3d_step(ns3d_ctx* ctx, const ns3d_fields* f, const ns3d_step_params* sp, int* h_iters,
                         double* h_err_hist, int err_cap, int* h_nchecks)
{
    NS3D_CHECK_CTX(ctx);
    if (!f || !sp) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_step: NULL argument");
    const ns3d_pt_params& p = sp->pt;
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const bool M = p.variant == NS3D_VARIANT_M;
    auto cyl = [&]() {
        return M ? ns3d_set_cylinder_M(ctx, f->C, f->Vx, f->Vy, f->Vz, sp->a2, sp->b2, sp->ox, sp->oy, sp->sinb, sp->cosb,
                                       sp->xco_g, sp->yco_g, p.dx, p.dy, nx, ny, nz)
                 : ns3d_set_cylinder_G(ctx, f->C, f->Vx, f->Vy, f->Vz, sp->a2, sp->b2, sp->ox, sp->oy, sp->sinb, sp->cosb,
                                       sp->lx, sp->ly, p.dx, p.dy, nx, ny, nz);
    };
    NS3D_TRY(ns3d_update_tau(ctx, f->txx, f->tyy, f->tzz, f->txy, f->txz, f->tyz, f->Vx, f->Vy, f->Vz, sp->mu, p.dx, p.dy,
                             p.dz, nx, ny, nz));                                                       // M:449
    // update_halo!(τxx,τyy,τzz) M:450 is redundant: τ is computed on the halo cells too.
    NS3D_TRY(ns3d_predict_V(ctx, f->Vx, f->Vy, f->Vz, f->txx, f->tyy, f->tzz, f->txy, f->txz, f->tyz, p.rho, p.g, p.dt,
                            p.dx, p.dy, p.dz, nx, ny, nz));                                            // M:451
    NS3D_TRY(cyl());                                                                                   // M:452
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
    NS3D_TRY(ns3d_pt_solve(ctx, f->Pr, f->dPrdtau, f->divV, &p, h_iters, h_err_hist, err_cap, h_nchecks));  // M:458-471
    NS3D_TRY(ns3d_correct_V(ctx, f->Vx, f->Vy, f->Vz, f->Pr, p.dt, p.rho, p.dx, p.dy, p.dz, nx, ny, nz));   // M:472
    NS3D_TRY(cyl());                                                                                   // M:473
    if (M)
        NS3D_TRY(ns3d_set_bc_Vel_M(ctx, f->Vx, f->Vy, f->Vz, sp->inlet_guard, sp->vin, nx, ny, nz));   // M:474
    else
        NS3D_TRY(ns3d_set_bc_Vel_G(ctx, f->Vx, f->Vy, f->Vz, nx, ny, nz));                             // G:140
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
