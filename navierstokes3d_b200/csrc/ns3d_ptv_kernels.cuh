// ptv_kernel: K fused pseudo-transient (PT) iterations per launch on the library's INTERNAL, pitched
// copies of Pr, dPrdτ and ∇V -- the hot kernel of ns3d_pt_solve / ns3d_pt_iterate.
//
// Per cell and iteration it computes the reference's K5 + K6 + K7 (M:70-82, 175-184), bit for bit in
// PARITY mode:
//     dPrdτ' = dPrdτ*(1-damp) + dτ*(d2x/dx/dx + d2y/dy/dy + d2z/dz/dz - ρ/dt*∇V)     M:71
//     Pr'    = Pr + dτ*dPrdτ'                                                       M:80
// then set_bc_Pr!: after the zero-gradient copies in the order x, y, z every boundary point equals the
// new value at its index clamped into the interior, then the outlet plane is
// overwritten (variant M, bc_x_Pr!) or both x planes get the hydrostatic profile (variant G).
//
// Layout (ns3d_pt.cu packs / unpacks around a solve): the five arrays Pr, Pr', dPrdτ, dPrdτ', ∇V share
// ONE shape -- rows of `px` doubles (px = nx rounded up to 16, so every row starts on a 128-byte line),
// element (i,j,k) at ((k*ny + j)*px + i), dPrdτ indexed like Pr (its rim is unused).  One byte offset
// addresses the same cell in all five, and pairs of x-neighbours are 16-byte aligned.
//
// Thread mapping: a thread owns TWO x-adjacent columns of one row (128-bit shared-memory accesses and stores) and
// marches along z with the z neighbours in registers (2.5-D blocking).  A CTA covers a tile of W = 2*pxt columns by
// H = bty rows; the tile shape is a launch parameter, chosen by the host so that tiles cover the grid without a
// mostly empty last tile (nx = 2^n - 1).
//
// Staging (TMA): the x-y tile of one z-plane of Pr (with a halo of two columns / one row), of dPrdτ and of ∇V is
// brought into a ring of NS shared-memory slots by three cp.async.bulk.tensor copies that ONE thread issues, NS-1
// planes ahead of their use, completion counted on the slot's mbarrier.  Loads in flight hold no registers, so the
// threads stay light (several CTAs per SM) and no load latency sits on a thread's dependency chain; boxes that hang
// over the domain are zero-filled by the unit.  (ncu of the first version, which prefetched into registers: ptxas
// sank the loads to their uses under the register cap and 16 warps per SM could not hide them.)
//
// Temporal blocking: a K-stage pipeline over the planes of a z-chunk.  At step t, stage 1 computes iterate 1 of
// plane t-1 from the staged tiles, stage m computes iterate m of plane t-m from iterate m-1: z neighbours from
// registers, in-plane neighbours from a three-slot shared-memory ring per level; stage K stores Pr' and dPrdτ'.
// The intermediate iterates never touch DRAM: 5 field passes per K iterations.  Redundancy instead of
// synchronisation between CTAs: K-1 rim rows / columns per tile side and K-1 extra planes per chunk end are
// recomputed by the neighbouring CTAs; at domain faces nothing is recomputed -- the face values are images of the
// adjacent interior values (folded bc_x!/bc_y!/bc_z!), produced by the thread that owns those.
//
// P2P = true (z-slabs, K <= 2): the chunks next to a slab interface also perform update_halo!(Pr) over
// peer memory (NVLink): the first iteration of the halo plane is recomputed locally from the local
// halo plane plus one plane of the neighbour's current iterate and its dPrdτ plane (peer loads), the
// planes a slab sends are stored straight into the neighbour's halo plane, hand-over through the
// mailbox protocol of ns3d_pt_common.cuh.  The four slots of such a chunk that need a neighbour's memory (planes -1 / nz
// of Pr, dPrdτ of the halo planes) are staged with plain L1-bypassing loads (ptv_coop_stage); all others by TMA.
#pragma once

#include "ns3d_pt_common.cuh"

struct PtV {
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
    // pitched layout shared by the five arrays
    int px;                   // row pitch in doubles
    long long rowB, planeB;   // bytes
    const double* P;          // Pr, current iterate
    double* PN;               // Pr after K iterations
    const double* D;          // dPrdτ, current
    double* DN;
    const double* V;          // ∇V
    // tiling
    int pxt, bty;             // thread columns (pairs) / thread rows of a tile: W = 2*pxt, H = RY*bty
    int sx, sy;               // tile stride = output columns / rows of an inner tile
    int ex, ey;               // rim of a tile that is not output (x: even, >= K-1; y: K-1)
    int ntx, nty;
    // planes this launch updates: [kbeg, kend) in chunks of zchunk, or -- when `faces` is set -- only the two
    // chunks of zchunk planes next to the z faces (the slab interfaces)
    int kbeg, kend, zchunk, faces;
    int reverse;              // serpentine sweep: walk the z-chunks downwards
    int rw;                   // row pitch (doubles) of a plane box with halo in shared memory: W + 4
    // shared memory (byte offsets / sizes, multiples of 128): NS staging slots of {Pr box with halo, dPrdτ box, ∇V box}
    // behind their mbarriers, then the ring of the intermediate iterates ((K-1) x 3 boxes with halo)
    int ns;
    unsigned sm_bars, sm_stage, sm_pbox, sm_dbox, sm_slot, sm_qring, sm_total;
    unsigned tx_bytes;        // bytes the three TMA copies of one plane deliver
    // peer-memory halo exchange
    double* peer_lo_plane;             // lower neighbour's halo plane nz-1 of its Pr'
    double* peer_hi_plane;             // upper neighbour's halo plane 0 of its Pr'
    unsigned long long* mbox;          // this rank's mailbox (NS3D_MB_*)
    unsigned long long* peer_lo_flag;  // lower neighbour's NS3D_MB_FLAG_HI
    unsigned long long* peer_hi_flag;  // upper neighbour's NS3D_MB_FLAG_LO
    const double* peer_lo_cur;         // lower neighbour's Pr plane nz-3   (= local plane -1)
    const double* peer_hi_cur;         // upper neighbour's Pr plane 2      (= local plane nz)
    const double* peer_lo_dp;          // lower neighbour's dPrdτ plane nz-2 (= local plane 0)
    const double* peer_hi_dp;          // upper neighbour's dPrdτ plane 1    (= local plane nz-1)
    // persistent launch (ptv_flow_kernel): `nlaunch` launches of K iterations each in ONE kernel.  Work item w is tile
    // w % (ntx*nty) of z-chunk (w / (ntx*nty)) % nbz of launch w / (ntx*nty*nbz); launch l reads the buffers launch l-1
    // wrote (P/PN and D/DN swap roles every launch).  work[0] = the queue (next unclaimed item), work[PTV_WORK_DONE + l*nbz + c]
    // = finished items of chunk c of launch l.
    unsigned* work;
    unsigned* work_err;   // sticky: set when a dependency wait ran into the spin limit
    int nlaunch, nbz;
};
enum { PTV_WORK_DONE = 32 };   // the counters start one 128-byte line after the queue

// iterations per launch and launch-bounds variant of ptv_kernel (threads per CTA / CTAs per SM the kernel is
// compiled for: 0 = 256 / 2 (128 registers), 1 = 256 / 3 (80), 3 = 512 / 1 (128), 4 = 512 / 2 (64))
struct PtvPlan {
    int K = 2, lb = 1;
    int ns = 4;  // staging slots (planes in flight + in use)
    int zf = 8;  // planes per slab-interface chunk of the split launch
};
inline int ptv_lb_threads(int lb) { return lb >= 3 ? 512 : 256; }
inline int ptv_lb_ctas(int lb) { return lb == 0 ? 2 : (lb == 1 ? 3 : (lb == 3 ? 1 : 2)); }

// The TMA descriptors of one launch: Pr (current iterate), dPrdτ (current), ∇V -- 3-D tensor maps over the pitched
// arrays (dims px x ny x nz), box = the CTA's tile (with halo for Pr).  128 opaque bytes each (CUtensorMap).
// m[3], m[4]: Pr and dPrdτ of the OTHER pair of ping-pong buffers (odd launches of a persistent launch).
struct alignas(64) PtvMaps {
    struct alignas(64) Map {
        unsigned long long opaque[16];
    } m[5];
};
// Keeps a per-thread value in a register: without it ptxas re-derives loop invariants from %tid / %ctaid in every
// iteration to stay under the register cap (measured in SASS).
#ifdef NS3D_HOST_EMU
#define NS3D_KEEP(x) ((void)0)
#else
#define NS3D_KEEP(x) asm volatile("" : "+r"(x))
#endif
#ifdef NS3D_HOST_EMU
#define __grid_constant__
#define NS3D_NOINLINE __attribute__((noinline))
#else
#define NS3D_NOINLINE __noinline__
#endif

namespace {

struct D2 {
    double x, y;
};

#ifdef NS3D_HOST_EMU
__device__ __forceinline__ D2 ptv_ld2(const char* a) { return *(const D2*)a; }
__device__ __forceinline__ D2 ptv_ld2_l2(const char* a)
{
    const volatile double* q = (const volatile double*)a;
    return D2{q[0], q[1]};
}
__device__ __forceinline__ double ptv_ld1(const char* a) { return *(const double*)a; }
__device__ __forceinline__ double ptv_ld1_l2(const char* a) { return *(const volatile double*)a; }
__device__ __forceinline__ D2 ptv_ld2_peer(const double* a)
{
    const volatile double* q = (const volatile double*)a;
    return D2{q[0], q[1]};
}
__device__ __forceinline__ void ptv_st2(char* a, double x, double y) { ((double*)a)[0] = x; ((double*)a)[1] = y; }
__device__ __forceinline__ D2 ptv_lds2(const char* a) { return *(const D2*)a; }
__device__ __forceinline__ void ptv_sts2(char* a, double x, double y) { ((double*)a)[0] = x; ((double*)a)[1] = y; }
#else
__device__ __forceinline__ D2 ptv_ld2(const char* a)
{
    const double2 v = *(const double2*)a;
    return D2{v.x, v.y};
}
// L1-bypassing loads (ld.global.cg): what a neighbour GPU stores into this rank's halo planes while this
// kernel is resident must not be served from a stale L1 line
__device__ __forceinline__ D2 ptv_ld2_l2(const char* a)
{
    const double2 v = __ldcg((const double2*)a);
    return D2{v.x, v.y};
}
__device__ __forceinline__ double ptv_ld1(const char* a) { return *(const double*)a; }
__device__ __forceinline__ double ptv_ld1_l2(const char* a) { return __ldcg((const double*)a); }
__device__ __forceinline__ D2 ptv_ld2_peer(const double* a)
{
    const double2 v = __ldcv((const double2*)a);
    return D2{v.x, v.y};
}
__device__ __forceinline__ void ptv_st2(char* a, double x, double y) { *(double2*)a = make_double2(x, y); }
// 128-bit shared-memory accesses (the boxes keep pairs 16-byte aligned)
__device__ __forceinline__ D2 ptv_lds2(const char* a)
{
    const double2 v = *(const double2*)a;
    return D2{v.x, v.y};
}
__device__ __forceinline__ void ptv_sts2(char* a, double x, double y) { *(double2*)a = make_double2(x, y); }
#endif

template <int MODE>
__device__ __forceinline__ double ptv_bracket(const PtV& p, double pc, double xm, double xp, double ym, double yp,
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

template <int MODE>
__device__ __forceinline__ void ptv_update(const PtV& p, double L, double dq, double pc, double& dn, double& u)
{
    if (MODE == NS3D_FASTEST) {
        dn = fma(p.dtau, L, dq * p.omd);
        u = fma(p.dtau, dn, pc);
    } else {
        dn = dq * p.omd + p.dtau * L;  // M:71
        u = pc + p.dtau * dn;          // M:80
    }
}

// Value of an x-face point (i in {0, nx-1}) of plane k (0-based) given the mirrored interior value u.
// Neumann: u.  M outlet: val (bc_x_Pr!, M:147-150).  G: bc_xhydstatic! (G:257-261).
__device__ __forceinline__ double ptv_xface(const PtV& p, bool hi, int k, double u)
{
    const int kind = hi ? p.xhi_kind : p.xlo_kind;
    if (kind == X_NEUMANN) return u;
    if (kind == X_DIRICHLET) return hi ? p.xhi_val : p.xlo_val;
    const double h = p.rho_g * ((double)(p.hyd_nz - (k + 1)) + 0.5) * p.hyd_dz;
    return hi ? h : h + p.xlo_val;
}

// Per-thread constants of a launch.  Everything that concerns a domain face sits behind ONE flag per kind of
// store (`rs` for the ring of the intermediate iterates, `own == 2` for the final stores), so that warps without
// such threads -- almost all -- execute a single branch for it.  (Separate ints on purpose: ptxas keeps the ones the
// hot loop tests in predicate registers; packed into one word they cost general registers and the kernel spilled.)
struct PtvThread {
    int own;       // 0: stores nothing; 1: stores the results of its pair, no face nearby; 2: ... with face images
    int rs;        // the pair needs special treatment when it is published in the ring (a face column / row nearby)
    int lo_pair;   // the pair is columns (0, 1): column 0 is the image of column 1
    int hi_in;     // the pair is columns (nx-2, nx-1) (nx even): column nx-1 is the image of column nx-2
    int hi_src;    // the pair ends at column nx-2 and nx-1 lies in the NEXT pair (nx odd): this thread also stores the image
    int noring;    // the pair is columns (nx-1, nx): its values are never published
    int j0;        // row of the thread
    int peer_ok;   // P2P: both neighbours answered
};

// CTA-uniform constants of a work item: the steps at which stage m works (plane t - m), copies the upper z-face image,
// sets the lower one; strides in bytes.  (Kept in registers: read from shared memory in every step -- tried, to relieve the
// register allocation -- they put five dependent loads in front of each step and cost 7 %.)
template <int K>
struct PtvUni {
    int ts_lo, ts_hi;             // steady steps: every stage works and none of the z-face cases applies
    int last_load, t_last;        // last plane of iterate 0 the item reads; its last step
    int tlo[K + 1], thi[K + 1];   // stage m computes at steps tlo[m] <= t <= thi[m]
    int timg[K + 1];              // step at which stage m < K takes plane nz-1 as the image of plane nz-2 (-1: never)
    int tz1[K + 1];               // step at which stage m < K also sets plane 0 as the image of plane 1 (-1: never)
    int t_out1, t_outn;           // steps at which stage K puts out plane 1 / plane nz-2
    long long oDN;                // from a cell of the Pr buffer the item writes to the same cell of its dPrdτ buffer
    char* pn;                     // element (0,0,0) of the Pr buffer the item writes
};

// ---- shared memory by 32-bit address ------------------------------------------------------------------------------
// The kernel addresses its dynamic shared memory through 32-bit shared-window addresses computed once (ld.shared /
// st.shared with immediate offsets), not through generic pointers: with those the compiler re-derives the window base
// (S2UR SR_CgaCtaId, ULEA) at every use -- ncu: three times per step.  The host emulation adds a base pointer instead.
typedef unsigned sa_t;
#ifdef NS3D_HOST_EMU
inline thread_local char* ptv_emu_base = nullptr;
__device__ __forceinline__ sa_t sa_base(char* smem) { ptv_emu_base = smem; return 0u; }
__device__ __forceinline__ D2 sa_ld2(sa_t a) { return *(const D2*)(ptv_emu_base + a); }
__device__ __forceinline__ double sa_ld1(sa_t a) { return *(const double*)(ptv_emu_base + a); }
__device__ __forceinline__ void sa_st2(sa_t a, double x, double y) { double* q = (double*)(ptv_emu_base + a); q[0] = x; q[1] = y; }
__device__ __forceinline__ void sa_st1(sa_t a, double x) { *(double*)(ptv_emu_base + a) = x; }
__device__ __forceinline__ char* sa_ptr(sa_t a) { return ptv_emu_base + a; }
#else
__device__ __forceinline__ sa_t sa_base(char* smem) { return (sa_t)__cvta_generic_to_shared(smem); }
__device__ __forceinline__ D2 sa_ld2(sa_t a)
{
    D2 r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ double sa_ld1(sa_t a)
{
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ void sa_st2(sa_t a, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory"); }
__device__ __forceinline__ void sa_st1(sa_t a, double x) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(x) : "memory"); }
#endif

// Publishing iterate m next to a domain face (slow path): folded boundary conditions -- the face values are images of
// the adjacent interior values; a y-face row is written by the thread that owns the adjacent interior row.
__device__ __forceinline__ void ptv_ring_special(const PtV& p, const PtvThread& v, sa_t w, int rwB, int pm, double& u0, double& u1)
{
    if (v.lo_pair) u0 = ptv_xface(p, false, pm, u1);
    if (v.hi_in) u1 = ptv_xface(p, true, pm, u0);
    const bool skip = v.j0 == 0 || v.j0 == p.ny - 1;
    if (!v.noring && !skip) sa_st2(w, u0, u1);
    double uh = 0;
    if (v.hi_src) {
        uh = ptv_xface(p, true, pm, u1);
        sa_st1(w + 16, uh);
    }
    if (v.j0 == 1) {
        if (!v.noring) sa_st2(w - rwB, u0, u1);
        if (v.hi_src) sa_st1(w - rwB + 16, uh);
    }
    if (v.j0 == p.ny - 2) {
        if (!v.noring) sa_st2(w + rwB, u0, u1);
        if (v.hi_src) sa_st1(w + rwB + 16, uh);
    }
}

// A final store next to a domain face (slow path): the new values (u0, u1) of plane `kk` (x-face values are computed
// for that plane index) into the plane at `row`, with the images of the x faces and of the y faces.
__device__ __forceinline__ void ptv_store_special(const PtV& p, const PtvThread& v, char* row, int kk, double u0, double u1)
{
    if (v.lo_pair) u0 = ptv_xface(p, false, kk, u1);
    if (v.hi_in) u1 = ptv_xface(p, true, kk, u0);
    ptv_st2(row, u0, u1);
    double uh = 0;
    if (v.hi_src) {
        uh = ptv_xface(p, true, kk, u1);
        *(double*)(row + 16) = uh;
    }
    if (v.j0 == 1) {  // bc_y!: row 0 is the image of row 1
        ptv_st2(row - p.rowB, u0, u1);
        if (v.hi_src) *(double*)(row - p.rowB + 16) = uh;
    }
    if (v.j0 == p.ny - 2) {
        ptv_st2(row + p.rowB, u0, u1);
        if (v.hi_src) *(double*)(row + p.rowB + 16) = uh;
    }
}

__device__ __forceinline__ void ptv_store_plane(const PtV& p, const PtvThread& v, char* row, int kk, double u0, double u1)
{
    if (v.own == 1) ptv_st2(row, u0, u1);
    else if (v.own == 2) ptv_store_special(p, v, row, kk, u0, u1);
}

// Registers of the pipeline.  h[m] = iterate m (m = 0: the input) of three consecutive planes, d[m] the dPrdτ of
// iterate m >= 1, dv = ∇V; the slot of a plane is (plane - first step) mod 3, a compile-time constant in the
// three-fold unrolled loop, so nothing is ever moved.
template <int K>
struct PtvRegs {
    double h[K][3][2];
    double d[K][3][2];
    double dv[3][2];
};

// ---- staging of iterate 0: the x-y tile (with halo) of one z-plane of Pr, and the tiles of dPrdτ and ∇V -------------
// TMA: one elected thread issues three cp.async.bulk.tensor copies per plane into a ring of NS shared-memory slots,
// NS-1 planes ahead of their use; completion is a transaction count on the slot's mbarrier.  Loads in flight hold no
// registers, and out-of-range box coordinates are zero-filled by the unit (tiles may hang over the domain).
#if !defined(NS3D_HOST_EMU)
__device__ __forceinline__ void ptv_mbar_init(sa_t bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ptv_mbar_expect(sa_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded: a copy that never completes (a bad descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ bool ptv_mbar_try(sa_t bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void ptv_mbar_wait(sa_t bar, unsigned parity)
{
    for (unsigned n = 0;; ++n) {
        if (ptv_mbar_try(bar, parity)) return;
        if (n > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void ptv_tma_3d(sa_t dst, const void* tmap, int x, int y, int z, sa_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(tmap), "r"(bar), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
#endif

// Tile geometry: compile-time (PXT x BTY threads: every stride below is an immediate of the load / store that uses it)
// or, PXT = 0, from the kernel parameters.
template <int PXT, int BTY>
struct PtvGeo {
    static constexpr bool CT = PXT > 0;
    static constexpr unsigned up(unsigned b) { return (b + 127u) / 128u * 128u; }
    static constexpr int Wc = 2 * PXT, Hc = BTY;
    static constexpr unsigned rwBc = (unsigned)(Wc + 4) * 8u;
    static constexpr unsigned pboxBc = up((unsigned)((Hc + 2) * (Wc + 4) * 8));
    static constexpr unsigned dboxBc = up((unsigned)(Hc * Wc * 8));
    static constexpr unsigned slotBc = pboxBc + 2u * dboxBc;
    static __device__ __forceinline__ int pxt(const PtV& p) { return CT ? PXT : p.pxt; }
    static __device__ __forceinline__ int bty(const PtV& p) { return CT ? BTY : p.bty; }
    static __device__ __forceinline__ unsigned rwB(const PtV& p) { return CT ? rwBc : (unsigned)p.rw * 8u; }
    static __device__ __forceinline__ unsigned pboxB(const PtV& p) { return CT ? pboxBc : p.sm_pbox; }
    static __device__ __forceinline__ unsigned dboxB(const PtV& p) { return CT ? dboxBc : p.sm_dbox; }
    static __device__ __forceinline__ unsigned slotB(const PtV& p) { return CT ? slotBc : p.sm_slot; }
};

// One step of the pipeline.  PH = (t - first step) mod 3.  sc / sp: the thread's pair in the staged Pr box of planes
// t / t-1; dp: the thread's pair in the staged dPrdτ box of plane t-1 (∇V follows dboxB bytes later); qr[l]: the
// thread's pair in slot 0 of the ring of iterate l+1; o: the thread's pair in plane t-K of PN.
// STEADY: every stage works at this step and none of the z-face special cases applies (the common case, selected by one
// CTA-uniform branch in the caller): straight-line code whose shared-memory loads are all issued before the first
// dependent FP64 operation, so that the K independent dependency chains of a step overlap (ncu of the version with a
// branch per stage: short-scoreboard and fixed-latency waits were the top stalls at 5 warps per scheduler).
template <int MODE, int K, bool P2P, int PH, class G, bool STEADY>
__device__ __forceinline__ void ptv_step(const PtV& p, const PtvThread& v, const PtvUni<K>& u, PtvRegs<K>& s, const int t, char* o,
                                         const sa_t sc, const sa_t sp, const sa_t dp, const sa_t (&qr)[K > 1 ? K - 1 : 1])
{
    const unsigned rwB = G::rwB(p), pboxB = G::pboxB(p), dboxB = G::dboxB(p);
    bool act[K + 1];
    double xm[K + 1], xp[K + 1];
    D2 ym[K + 1], yp[K + 1];
    D2 w1 = {0, 0}, d1 = {0, 0}, g1 = {0, 0};
    // ---- shared-memory loads: the staged tiles of stage 1 and the in-plane neighbours of every stage.  In the steady
    // state they are all issued here, ahead of the first dependent FP64 operation; otherwise stage by stage below. ----
#pragma unroll
    for (int m = 1; m <= K; ++m) {
        const int S = ((PH - m) % 3 + 3) % 3;           // register slot of plane t-m (compile-time after unrolling)
        act[m] = STEADY || (t >= u.tlo[m] && t <= u.thi[m]);
        if (STEADY) {
            const sa_t q = (m == 1) ? sp : qr[m >= 2 ? m - 2 : 0] + (unsigned)S * pboxB;
            xm[m] = sa_ld1(q - 8); xp[m] = sa_ld1(q + 16);
            ym[m] = sa_ld2(q - rwB); yp[m] = sa_ld2(q + rwB);
            if (m == 1) {
                w1 = sa_ld2(sc);             // Pr of plane t: the z+1 neighbour
                d1 = sa_ld2(dp);
                g1 = sa_ld2(dp + dboxB);
            }
        }
    }
#pragma unroll
    for (int m = 1; m <= K; ++m) {
        const int S = ((PH - m) % 3 + 3) % 3;
        const int Sm = (S + 2) % 3, Sp = (S + 1) % 3;   // slots of planes t-m-1, t-m+1
        if (act[m]) {
            if (!STEADY) {
                const sa_t q = (m == 1) ? sp : qr[m >= 2 ? m - 2 : 0] + (unsigned)S * pboxB;
                xm[m] = sa_ld1(q - 8); xp[m] = sa_ld1(q + 16);
                ym[m] = sa_ld2(q - rwB); yp[m] = sa_ld2(q + rwB);
                if (m == 1) {
                    w1 = sa_ld2(sc);
                    d1 = sa_ld2(dp);
                    g1 = sa_ld2(dp + dboxB);
                }
            }
            double dq0, dq1;
            if (m == 1) {
                s.h[0][PH][0] = w1.x; s.h[0][PH][1] = w1.y;
                dq0 = d1.x; dq1 = d1.y;
                s.dv[S][0] = g1.x; s.dv[S][1] = g1.y;
            } else {
                dq0 = s.d[m - 1][S][0]; dq1 = s.d[m - 1][S][1];
            }
            // ---- the update of the thread's two cells ----------------------------------------------------------
            const double c0 = s.h[m - 1][S][0], c1 = s.h[m - 1][S][1];
            const double L0 = ptv_bracket<MODE>(p, c0, xm[m], c1, ym[m].x, yp[m].x, s.h[m - 1][Sm][0], s.h[m - 1][Sp][0], s.dv[S][0]);
            const double L1 = ptv_bracket<MODE>(p, c1, c0, xp[m], ym[m].y, yp[m].y, s.h[m - 1][Sm][1], s.h[m - 1][Sp][1], s.dv[S][1]);
            double u0, u1, dn0, dn1;
            ptv_update<MODE>(p, L0, dq0, c0, dn0, u0);
            ptv_update<MODE>(p, L1, dq1, c1, dn1, u1);
            if (m < K) {
                // ---- iterate m stays on chip: registers (z neighbours) and the ring (in-plane neighbours) ------
                const sa_t w = qr[m - 1] + (unsigned)S * pboxB;
                if (v.rs) ptv_ring_special(p, v, w, (int)rwB, t - m, u0, u1);
                else sa_st2(w, u0, u1);
                s.h[m][S][0] = u0; s.h[m][S][1] = u1;
                s.d[m][S][0] = dn0; s.d[m][S][1] = dn1;
                if (!STEADY && t == u.tz1[m]) {  // bc_z!: plane 0 is the image of plane 1 (M:129)
                    s.h[m][Sm][0] = u0; s.h[m][Sm][1] = u1;
                }
            } else {
                // ---- stage K: Pr' and dPrdτ' of plane t-K leave the chip --------------------------------------------
                if (v.own) ptv_st2(o + u.oDN, dn0, dn1);
                ptv_store_plane(p, v, o, t - K, u0, u1);
                if (!STEADY && t == u.t_out1) {
                    if (!p.zlo_halo) ptv_store_plane(p, v, o - p.planeB, 0, u0, u1);  // bc_z! M:129
                    else if (P2P && v.peer_ok)  // update_halo!(Pr): our plane 1 is the lower neighbour's halo plane
                        ptv_store_plane(p, v, (char*)p.peer_lo_plane + (o - (u.pn + p.planeB)), 1, u0, u1);
                }
                if (!STEADY && t == u.t_outn) {
                    if (!p.zhi_halo) ptv_store_plane(p, v, o + p.planeB, p.nz - 1, u0, u1);  // bc_z! M:130
                    else if (P2P && v.peer_ok)
                        ptv_store_plane(p, v, (char*)p.peer_hi_plane + (o - (u.pn + (long long)(p.nz - 2) * p.planeB)), p.nz - 2, u0, u1);
                }
            }
        } else if (!STEADY && m < K && t == u.timg[m]) {
            // bc_z!: plane nz-1 is the image of plane nz-2 (M:130), which this stage computed one step ago
            s.h[m][S][0] = s.h[m][Sm][0]; s.h[m][S][1] = s.h[m][Sm][1];
        }
    }
}

// Plain-load staging of one plane (z-slab interface chunks, whose first / last planes live in a neighbour's memory, and
// the host emulation): every thread copies box elements, zero-filling what lies outside the arrays like the TMA unit.
template <bool P2P>
__device__ __forceinline__ void ptv_coop_stage(const PtV& p, const double* P, const double* D, char* slot, int q, int X0, int Y0,
                                               bool lo_face, bool hi_face)
{
    const int W = 2 * p.pxt, H = p.bty;
    const int nthreads = blockDim.x, tid = threadIdx.x;
    const long long plane = p.planeB / 8;
    const double* Psrc = P + (long long)q * plane;
    const double* Dsrc = D + (long long)q * plane;
    const double* Vsrc = p.V + (long long)q * plane;
    if (P2P) {
        if (q == -1) Psrc = p.peer_lo_cur;          // the lower neighbour's plane nz-3
        if (q == p.nz) Psrc = p.peer_hi_cur;        // the upper neighbour's plane 2
        if (q == 0 && lo_face) Dsrc = p.peer_lo_dp;          // its dPrdτ plane nz-2
        if (q == p.nz - 1 && hi_face) Dsrc = p.peer_hi_dp;   // its dPrdτ plane 1
    }
    const bool zin = q >= 0 && q <= p.nz - 1;
    double* dP = (double*)slot;
    for (int e = tid; e < (H + 2) * (W + 4); e += nthreads) {
        const int r = e / (W + 4), c = e - r * (W + 4);
        const int gx = X0 - 2 + c, gy = Y0 - 1 + r;
        double val = 0.0;
        if (gx >= 0 && gx < p.px && gy >= 0 && gy < p.ny) {
            const double* a = Psrc + (long long)gy * p.px + gx;
            val = P2P ? ptv_ld1_l2((const char*)a) : *a;
        }
        dP[e] = val;
    }
    double* dD = (double*)(slot + p.sm_pbox);
    double* dV = (double*)(slot + p.sm_pbox + p.sm_dbox);
    for (int e = tid; e < H * W; e += nthreads) {
        const int r = e / W, c = e - r * W;
        const int gx = X0 + c, gy = Y0 + r;
        double d = 0.0, g = 0.0;
        if (zin && gx < p.px && gy < p.ny) {
            const long long a = (long long)gy * p.px + gx;
            d = P2P ? ptv_ld1_l2((const char*)(Dsrc + a)) : Dsrc[a];
            g = Vsrc[a];
        }
        dD[e] = d;
        dV[e] = g;
    }
}

// One work item: K iterations of the planes of z-chunk `bz` (of `nbz`) in the columns of tile `tile`.  `flip`: the launch
// reads PN / DN and writes P / D (odd launches of a persistent launch); `init_bars`: the CTA's first item initialises the
// slots' mbarriers; CARRY (persistent launch): later items of the CTA carry on with slots and mbarrier parities where the
// last one stopped (s_loads counts the CTA's copies) -- the mbarriers are never initialised twice (an mbarrier.inval +
// mbarrier.init pair per item, the first version of this, hung a launch-bounds variant of the kernel on hardware: the
// invalidation is not ordered with the store that initialises) -- and the item's first copies wait for `gate()`.
template <int MODE, int K, bool P2P, bool TMA, bool CARRY, int PXT, int BTY, class Gate>
__device__ __forceinline__ void ptv_item(const PtV& p, const PtvMaps& maps, char* smem, const int tile, int bz, const int nbz,
                                         const bool flip, const bool init_bars, Gate gate)
{
    typedef PtvGeo<PXT, BTY> G;
    const sa_t sb = sa_base(smem);
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const int NS = G::CT ? 4 : p.ns;
    const int pxt = G::pxt(p), bty = G::bty(p);
    const double* const P = flip ? p.PN : p.P;
    double* const PN = flip ? const_cast<double*>(p.P) : p.PN;
    const double* const D = flip ? p.DN : p.D;
    double* const DN = flip ? const_cast<double*>(p.D) : p.DN;
    int ty = (int)threadIdx.x / pxt;
    const int tx = (int)threadIdx.x - ty * pxt;
    const bool dup = ty >= bty;   // surplus threads of the last warp shadow a thread of the last row (same values, no stores)
    if (dup) ty = bty - 1;
    const int bxi = tile / p.nty, byi = tile - bxi * p.nty;   // y tiles fastest
    if (p.reverse && !p.faces) bz = nbz - 1 - bz;
    const int kb = p.faces ? (bz == 0 ? 1 : nz - 1 - p.zchunk) : p.kbeg + bz * p.zchunk;
    const int ke = p.faces ? kb + p.zchunk : min(kb + p.zchunk, p.kend);   // stage-K planes [kb, ke)
    const bool lo_face = P2P && p.zlo_halo && kb == 1;
    const bool hi_face = P2P && p.zhi_halo && ke == nz - 1;
    __shared__ int s_peer_ok;
    __shared__ unsigned s_loads;   // CARRY: TMA copies the CTA has issued in its earlier items
    PtvThread v;
    bool peer_ok = true;
    const unsigned pboxB = G::pboxB(p), dboxB = G::dboxB(p), slotB = G::slotB(p);
    const sa_t bars = sb + p.sm_bars, stage0 = sb + p.sm_stage;
#if !defined(NS3D_HOST_EMU)
    if (TMA && init_bars && threadIdx.x == 0) {
        for (int q = 0; q < NS; ++q) ptv_mbar_init(bars + 8 * q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    if (CARRY && init_bars && threadIdx.x == 0) s_loads = 0u;
    if (P2P && (lo_face | hi_face)) {
        if (threadIdx.x == 0) {
            bool ok = true;
            if (lo_face) ok = wait_neighbour(p.mbox, 0) && ok;
            if (hi_face) ok = wait_neighbour(p.mbox, 1) && ok;
            s_peer_ok = ok;
        }
        __syncthreads();
        peer_ok = s_peer_ok != 0;
    } else {
        __syncthreads();   // the mbarriers are initialised
    }
    const int W = 2 * pxt, H = bty;
    const int X0 = bxi * p.sx, Y0 = byi * p.sy;
    const int i0 = X0 + 2 * tx;
    v.peer_ok = peer_ok;
    v.j0 = Y0 + ty;
    v.lo_pair = i0 == 0;
    v.hi_in = i0 == nx - 2;
    v.hi_src = i0 + 1 == nx - 2;
    v.noring = i0 == nx - 1;
    const bool yface = v.j0 <= 1 || v.j0 >= ny - 2;   // a y-face row or the interior row next to one
    v.rs = v.lo_pair | v.hi_in | v.hi_src | v.noring | (int)yface;
    {
        const int oxl = bxi == 0 ? 0 : X0 + p.ex, oxh = bxi == p.ntx - 1 ? nx : X0 + W - p.ex;
        const int oyl = byi == 0 ? 1 : Y0 + p.ey, oyh = byi == p.nty - 1 ? ny - 1 : Y0 + H - p.ey;
        const bool own = !dup && i0 >= oxl && i0 < oxh && i0 <= nx - 2 && v.j0 >= oyl && v.j0 < oyh && v.j0 >= 1 && v.j0 <= ny - 2;
        v.own = own ? (v.rs ? 2 : 1) : 0;
    }
    // stage 1 range: interior planes, on a slab interface of the K = 2 pipeline also the halo plane
    const int lo1 = (K > 1 && lo_face) ? 0 : 1, hi1 = (K > 1 && hi_face) ? nz - 1 : nz - 2;
    const int a1 = max(kb - (K - 1), lo1);               // first plane stage 1 computes
    const int b1 = min(ke - 1 + (K - 1), hi1);           // last one
    const int q0 = a1 - 1;                               // first plane of iterate 0 the chunk reads
    const int t0 = a1 + 1;
    PtvUni<K> u;
    u.last_load = b1 + 1;                                // last plane of iterate 0 it reads
    u.t_last = ke - 1 + K;
#pragma unroll
    for (int m = 1; m <= K; ++m) {
        const int need_lo = kb - (K - m), need_hi = ke - 1 + (K - m);   // planes stage m has to produce for this chunk
        const int clo = m == 1 ? lo1 : 1, chi = m == 1 ? hi1 : nz - 2;  // planes it can compute
        u.tlo[m] = max(need_lo, clo) + m;
        u.thi[m] = min(need_hi, chi) + m;
        u.timg[m] = (m < K && !p.zhi_halo && nz - 1 <= need_hi) ? nz - 1 + m : -1;
        u.tz1[m] = (m < K && !p.zlo_halo) ? 1 + m : -1;
    }
    u.tlo[0] = u.thi[0] = u.timg[0] = u.tz1[0] = 0;
    u.t_out1 = 1 + K;
    u.t_outn = nz - 2 + K;
    u.ts_lo = u.tlo[1];
    u.ts_hi = u.thi[1];
#pragma unroll
    for (int m = 2; m <= K; ++m) {
        u.ts_lo = max(u.ts_lo, u.tlo[m]);
        u.ts_hi = min(u.ts_hi, u.thi[m]);
    }
    u.ts_lo = max(u.ts_lo, K + 2);            // past the steps that set the lower z-face images (t <= 1 + K)
    u.ts_hi = min(u.ts_hi, nz - 3 + K);       // before the one that stores the upper z-face image / the peer plane
    u.oDN = (const char*)DN - (const char*)PN;
    u.pn = (char*)PN;
    // the thread's pair inside a box with halo / without
    unsigned ringo = (unsigned)(ty + 1) * G::rwB(p) + (unsigned)(2 * tx + 2) * 8u;
    unsigned cello = (unsigned)(ty * W + 2 * tx) * 8u;
    NS3D_KEEP(ringo); NS3D_KEEP(cello);
    sa_t qr[K > 1 ? K - 1 : 1];
#pragma unroll
    for (int l = 0; l < (K > 1 ? K - 1 : 1); ++l) {
        qr[l] = sb + p.sm_qring + (unsigned)(3 * l) * pboxB + ringo;
        NS3D_KEEP(qr[l]);   // ptxas otherwise re-derives these from %tid and the shared-window base at every use (ncu)
    }
    // ---- staging: the item's copy number j (plane q0 + j) is the CTA's copy number g0 + j; copy number g lives in slot
    // g mod NS and completes the slot's mbarrier with parity (g / NS) & 1.  One-item kernel: g0 = 0.  CARRY (persistent
    // launch): g0 = what the CTA's earlier items have issued -- an item carries on with slots and parities where the last
    // one stopped, so the mbarriers are never initialised twice and nothing per step distinguishes it from a first item.
    const unsigned g0 = CARRY ? s_loads : 0u;
    const unsigned s0 = CARRY ? g0 % (unsigned)NS : 0u;             // slot of plane q0
    auto slot_of = [&](unsigned j) { const unsigned c = s0 + j; return c >= (unsigned)NS ? c - (unsigned)NS : c; };   // j < NS
    auto stage_plane = [&](int q, sa_t slot, sa_t bar) {
        if (q > u.last_load) return;
#if !defined(NS3D_HOST_EMU)
        if (TMA) {
            // slab interfaces: the planes beyond the halo (-1, nz) and the dPrdτ of the halo planes (0, nz-1) live in a
            // neighbour's memory -- those few slots are filled by plain loads of all threads, and the elected thread
            // completes the slot's mbarrier by hand; every other plane goes through the TMA unit
            const bool remote = P2P && ((lo_face && q <= 0) || (hi_face && q >= nz - 1));
            if (remote) {
                ptv_coop_stage<P2P>(p, P, D, smem + (slot - sb), q, X0, Y0, lo_face, hi_face);
                __syncthreads();
                if (threadIdx.x == 0) ptv_mbar_expect(bar, 0);
                return;
            }
            if (threadIdx.x == 0) {
                ptv_mbar_expect(bar, p.tx_bytes);
                ptv_tma_3d(slot, flip ? &maps.m[3] : &maps.m[0], X0 - 2, Y0 - 1, q, bar);
                ptv_tma_3d(slot + pboxB, flip ? &maps.m[4] : &maps.m[1], X0, Y0, q, bar);
                ptv_tma_3d(slot + pboxB + dboxB, &maps.m[2], X0, Y0, q, bar);
            }
            return;
        }
#endif
        (void)bar;
        ptv_coop_stage<P2P>(p, P, D, smem + (slot - sb), q, X0, Y0, lo_face, hi_face);
    };
    // CARRY (persistent launch): the item's first copies wait for gate() -- the issuing thread's look at the item's
    // predecessors.  With TMA staging that wait happens INSIDE the loop in which every thread waits for the first plane
    // anyway (below): as a loop of its own anywhere in the kernel, it cost the hot loop its spill-free register allocation.
    bool issued = !(TMA && CARRY);
    if (!TMA && CARRY) {
        if (threadIdx.x == 0)
            while (!gate()) spin_pause();
        __syncthreads();
    }
    if (issued)
        for (int j = 0; j < NS; ++j) stage_plane(q0 + j, stage0 + slot_of((unsigned)j) * slotB, bars + 8u * slot_of((unsigned)j));
    if (!TMA) __syncthreads();
    PtvRegs<K> s;
    // ---- prologue: the thread's own Pr of planes a1-1 and a1 (register slots 1 and 2) -------------------------------
    {
        const unsigned c0 = s0, c1 = slot_of(1u);
#if !defined(NS3D_HOST_EMU)
        if (TMA) {
            for (unsigned n = 0;; ++n) {
                if (CARRY && !issued && threadIdx.x == 0 && gate()) {
                    for (int j = 0; j < NS; ++j)
                        stage_plane(q0 + j, stage0 + slot_of((unsigned)j) * slotB, bars + 8u * slot_of((unsigned)j));
                    issued = true;
                }
                if (ptv_mbar_try(bars + 8u * c0, (g0 / (unsigned)NS) & 1u)) break;
                if (n > (1u << 24)) __trap();   // bounded: a copy (or a predecessor) that never arrives traps instead of hanging the GPU
            }
        }
#endif
        D2 w = sa_ld2(stage0 + c0 * slotB + ringo);
        s.h[0][1][0] = w.x; s.h[0][1][1] = w.y;
#if !defined(NS3D_HOST_EMU)
        if (TMA) ptv_mbar_wait(bars + 8u * c1, ((g0 + 1u) / (unsigned)NS) & 1u);
#endif
        w = sa_ld2(stage0 + c1 * slotB + ringo);
        s.h[0][2][0] = w.x; s.h[0][2][1] = w.y;
        __syncthreads();            // every thread has read its pair of plane q0 (and s_loads): its slot is free
        if (CARRY && threadIdx.x == 0) s_loads = g0 + (unsigned)(u.last_load - q0 + 1);   // where the CTA's next item carries on
        stage_plane(q0 + NS, stage0 + c0 * slotB, bars + 8u * c0);
    }
    // the thread's pair in plane t0 - K of PN
    char* o = (char*)PN + (long long)(t0 - K) * p.planeB + (long long)v.j0 * p.rowB + (long long)i0 * 8;
    int t = t0;
    sa_t prv = stage0 + slot_of(1u) * slotB;       // staging slot of plane t0 - 1
    sa_t pbar = bars + 8u * slot_of(1u);
    sa_t cur = stage0 + slot_of(2u) * slotB;       // ... of plane t0 (= q0 + 2; NS >= 3)
    sa_t bar = bars + 8u * slot_of(2u);            // mbarrier of the current slot
    unsigned par = ((g0 + 2u) / (unsigned)NS) & 1u;   // parity it completes with
#ifdef NS3D_HOST_EMU
#define PTV_MBAR_WAIT(b, parity) ((void)0)
#else
#define PTV_MBAR_WAIT(b, parity) ptv_mbar_wait((b), (parity))
#endif
#define PTV_STEP(PH)                                                                                                       \
    {                                                                                                                      \
        if (TMA && t <= u.last_load) PTV_MBAR_WAIT(bar, par);                                                              \
        if (t >= u.ts_lo && t <= u.ts_hi)                                                                                  \
            ptv_step<MODE, K, P2P, PH, G, true>(p, v, u, s, t, o, cur + ringo, prv + ringo, prv + pboxB + cello, qr);      \
        else                                                                                                               \
            ptv_step<MODE, K, P2P, PH, G, false>(p, v, u, s, t, o, cur + ringo, prv + ringo, prv + pboxB + cello, qr);     \
        __syncthreads(); /* ring slots written in this step are read in the next; the staging slot of plane t-1 is free */ \
        stage_plane(t - 1 + NS, prv, pbar);                                                                                \
    }
#define PTV_NEXT()                                                                          \
    {                                                                                       \
        ++t;                                                                                \
        o += p.planeB;                                                                      \
        prv = cur; pbar = bar;                                                              \
        cur += slotB; bar += 8;                                                             \
        if (cur == stage0 + (unsigned)NS * slotB) { cur = stage0; bar = bars; par ^= 1u; }  \
    }
    while (true) {
        PTV_STEP(0)
        if (t == u.t_last) break;
        PTV_NEXT()
        PTV_STEP(1)
        if (t == u.t_last) break;
        PTV_NEXT()
        PTV_STEP(2)
        if (t == u.t_last) break;
        PTV_NEXT()
    }
#undef PTV_STEP
#undef PTV_NEXT
#undef PTV_MBAR_WAIT
    if (P2P && (lo_face | hi_face)) {
        __threadfence_system();  // this thread's peer stores are performed before the flag can be seen
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned nface = (unsigned)(p.ntx * p.nty);
            if (lo_face) signal_neighbour(p.mbox, 0, p.peer_lo_flag, nface);
            if (hi_face) signal_neighbour(p.mbox, 1, p.peer_hi_flag, nface);
        }
    }
}

#ifdef NS3D_HOST_EMU
#define PTV_SMEM(name) char* name = (char*)emu::dyn_smem()
#else
#define PTV_SMEM(name)                                   \
    extern __shared__ __align__(128) char ptv_smem_[];   \
    char* name = ptv_smem_
#endif

// One launch = K iterations; one CTA per (tile, z-chunk): grid (ntx*nty, chunks).
template <int MODE, int K, bool P2P, bool TMA, int PXT, int BTY, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ptv_kernel(const PtV p, const __grid_constant__ PtvMaps maps)
{
    PTV_SMEM(smem);
    ptv_item<MODE, K, P2P, TMA, false, PXT, BTY>(p, maps, smem, (int)blockIdx.x, (int)blockIdx.y, (int)gridDim.y, false, true, [] { return true; });
}

// Persistent launch: p.nlaunch launches of K iterations in ONE kernel, without the idle ramp and tail between dependent
// launches (ncu of ptv_kernel at 255x153x153: the SMs are idle 17 % of a launch) and without their launch gaps.  The
// CTAs -- as many as are resident at a time -- claim work items from a queue in launch-major, chunk-major order.  An item
// of launch l reads planes of the iterate that launch l-1 wrote and overwrites planes that launch l-1 read, both within
// K planes of its chunk: it waits until every tile of the chunks within that range has finished launch l-1 (a counter
// per launch and chunk; release / acquire at gpu scope, then a proxy fence because the readers are TMA copies).  Items
// are claimed in order, so everything an item waits for has been claimed by a CTA that is running: no deadlock, whatever
// the number of resident CTAs.  In steady state the wait is over before it starts -- the chunks an item needs were
// claimed a whole launch earlier.
//
// The predecessors of work item w: launch l-1 must have finished every tile of the chunks whose planes [kb - K, ke - 1 + K]
// the item reads or overwrites -- the item's own chunk and its two neighbours (chunks are at least K planes long, the host
// sees to that), clamped at the ends.  dep[0..2]: indices of their counters in p.work; dep[3]: the count they must reach
// (0: nothing to wait for).
__device__ __forceinline__ void ptv_flow_deps(const PtV& p, unsigned w, unsigned ntiles, unsigned per_launch, unsigned* dep)
{
    const unsigned l = w / per_launch, bz = (w - l * per_launch) / ntiles;
    const unsigned base = PTV_WORK_DONE + (l > 0 ? l - 1 : 0) * (unsigned)p.nbz;
    dep[0] = base + (bz > 0 ? bz - 1 : 0);
    dep[1] = base + bz;
    dep[2] = base + min(bz + 1, (unsigned)p.nbz - 1);
    dep[3] = (l > 0 && l < (unsigned)p.nlaunch) ? ntiles : 0u;   // (no such item: the CTA is about to leave)
}
// One look at them: three relaxed polls, served from L2, then ONE fence (an acquire load is followed by an invalidation of
// the SM's whole L1, which a waiting thread would repeat with every look).
__device__ __forceinline__ bool ptv_flow_ready(const unsigned* work, const unsigned* dep)
{
    const bool ok = ld_min3_relaxed(work + dep[0], work + dep[1], work + dep[2]) >= dep[3];
    if (ok) {
        fence_acq_rel_gpu();
        fence_proxy_async();   // ... and what was acquired is visible to the TMA copies this thread issues next
    }
    return ok;
}

// Register allocation of this kernel is fragile: its hot loop (ptv_item) sits at the 80-register limit of three CTAs per SM,
// and ptxas loses the spill-free allocation (ptxas -v: 0 -> 112 bytes of stack, reloads inside the hot loop that the L1
// cannot hold next to three CTAs' shared memory: ncu, 45 % of the stall samples) as soon as the kernel contains another loop
// in which one thread waits, a NANOSLEEP or a non-inlined call -- wherever it stands and however it is written.  Hence the
// look at an item's predecessors sits inside ptv_item's own wait for the item's first plane (`gate`), reads what it needs
// from shared memory, and nothing else in this kernel waits.
template <int MODE, int K, bool TMA, int PXT, int BTY, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ptv_flow_kernel(const PtV p, const __grid_constant__ PtvMaps maps)
{
    PTV_SMEM(smem);
    __shared__ unsigned s_w;        // the CTA's work item
    __shared__ unsigned s_dep[4];   // ... and its predecessors (ptv_flow_deps)
    __shared__ unsigned s_ready;    // ... which were complete at the first look
    if (threadIdx.x == 0) {
        const unsigned ntiles = (unsigned)(p.ntx * p.nty), per_launch = ntiles * (unsigned)p.nbz;
        const unsigned w = atom_add_relaxed_gpu(p.work, 1u);
        s_w = w;
        ptv_flow_deps(p, w, ntiles, per_launch, s_dep);
    }
    bool first = true;
    for (;;) {   // (re-)entry: the item's predecessors are waited for
        if (threadIdx.x == 0) {
            unsigned looks = 0;
            while (!ptv_flow_ready(p.work, s_dep)) {
                if (++looks > NS3D_FLOW_WAIT_LOOKS) {   // bounded: a bug must not hang the GPU; the host reports the error word
                    *p.work_err = 1u;
                    break;
                }
            }
        }
        __syncthreads();
        for (;;) {   // items whose predecessors are complete on arrival
            {
                const unsigned ntiles = (unsigned)(p.ntx * p.nty), per_launch = ntiles * (unsigned)p.nbz;
                const unsigned w = s_w;
                if (w >= per_launch * (unsigned)p.nlaunch) return;
                const int l = (int)(w / per_launch);
                const unsigned r = w - (unsigned)l * per_launch;
                const int bz = (int)(r / ntiles), tile = (int)(r - (unsigned)bz * ntiles);
                ptv_item<MODE, K, false, TMA, true, PXT, BTY>(p, maps, smem, tile, bz, p.nbz, (l & 1) != 0, first, [] { return true; });
                first = false;
            }
            __syncthreads();   // the item's stores happen before the counter moves (and s_w, s_dep have been read)
            if (threadIdx.x == 0) {
                const unsigned ntiles = (unsigned)(p.ntx * p.nty), per_launch = ntiles * (unsigned)p.nbz;
                atom_add_release_gpu(p.work + s_dep[1] + (s_dep[3] ? (unsigned)p.nbz : 0u), 1u);   // = done[l][bz]
                const unsigned w = atom_add_relaxed_gpu(p.work, 1u);
                s_w = w;
                ptv_flow_deps(p, w, ntiles, per_launch, s_dep);
                s_ready = (w >= per_launch * (unsigned)p.nlaunch || ptv_flow_ready(p.work, s_dep)) ? 1u : 0u;
            }
            __syncthreads();
            if (!s_ready) break;
        }
    }
}

// ---- the pitched copies: pack (reference-shaped arrays -> internal) and unpack ------------------------------
// mode 0: src is (nx,ny,nz) dense; mode 1: src is dPrdτ (nx-2,ny-2,nz-2), placed at the Pr index of its cells,
// rim zero-filled.  One thread per internal element (pad columns included, zero-filled).
__global__ void __launch_bounds__(256) ptv_pack_kernel(double* __restrict__ dst, const double* __restrict__ src, int nx, int ny,
                                                       int nz, int px, int inner)
{
    const size_t n = (size_t)px * ny * nz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(q % px);
        const size_t jk = q / px;
        const int j = (int)(jk % ny), k = (int)(jk / ny);
        double val = 0.0;
        if (!inner) {
            if (i < nx) val = src[idx3(i, j, k, nx, ny)];
        } else if (i >= 1 && i <= nx - 2 && j >= 1 && j <= ny - 2 && k >= 1 && k <= nz - 2) {
            val = src[idx3(i - 1, j - 1, k - 1, nx - 2, ny - 2)];
        }
        dst[q] = val;
    }
}

// One thread per element of the reference-shaped destination.
__global__ void __launch_bounds__(256) ptv_unpack_kernel(double* __restrict__ dst, const double* __restrict__ src, int nx, int ny,
                                                         int nz, int px, int inner)
{
    const int sx = inner ? nx - 2 : nx, sy = inner ? ny - 2 : ny, sz = inner ? nz - 2 : nz, o = inner ? 1 : 0;
    const size_t n = (size_t)sx * sy * sz;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(q % sx);
        const size_t jk = q / sx;
        const int j = (int)(jk % sy), k = (int)(jk / sy);
        dst[q] = src[((size_t)(k + o) * ny + (j + o)) * px + (i + o)];
    }
}

// ---- host-side geometry (shared by ns3d_pt.cu and its host emulation) ---------------------------------------

inline int ptv_pitch(int nx) { return (nx + 15) / 16 * 16; }

// The arithmetic part of the kernel parameters (everything that does not depend on the context).
inline void ptv_fill(const ns3d_pt_params* p, PtV* k)
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
    k->px = ptv_pitch(p->nx);
    k->rowB = 8LL * k->px;
    k->planeB = k->rowB * p->ny;
    k->kbeg = 1;
    k->kend = p->nz - 1;
    k->faces = 0;
    k->reverse = 0;
}

// Tile shape for K iterations per launch and at most `nt` threads per CTA.
// A tile of W x H cells puts out (W - 2 ex) x (H - 2 ey) of them (more at the domain faces, where nothing is
// recomputed); W = 2*pxt, H = bty, pxt*bty <= nt.  x: all tiles of a row have the same width, the smallest
// that covers it with the chosen number of tiles -- so that nx = 2^n - 1 does not leave a nearly empty last
// tile; the number of x tiles (hence the tile height the threads allow) is the one that wastes the fewest
// cell updates.  want_pxt / want_bty > 0 fix the shape (tests, sweeps).
inline void ptv_tile_set(PtV& k, int K, int pxt, int bty)
{
    const int ex = ((K - 1) + 1) / 2 * 2, ey = K - 1;
    k.pxt = pxt; k.bty = bty;
    k.ex = ex; k.ey = ey;
    const int W = 2 * pxt, H = bty;
    k.sx = W - 2 * ex; k.sy = H - 2 * ey;
    k.ntx = 1;
    while (k.sx > 0 && (long long)(k.ntx - 1) * k.sx + W < k.nx - 1) ++k.ntx;
    k.nty = 1;
    while (k.sy > 0 && (long long)(k.nty - 1) * k.sy + H < k.ny - 1) ++k.nty;
    k.rw = W + 4;
    auto up = [](unsigned b) { return (b + 127u) / 128u * 128u; };
    if (k.ns < 3) k.ns = 4;
    k.sm_pbox = up((unsigned)((H + 2) * (W + 4) * 8));
    k.sm_dbox = up((unsigned)(H * W * 8));
    k.sm_slot = k.sm_pbox + 2 * k.sm_dbox;
    k.sm_bars = 0;
    k.sm_stage = 128;
    k.sm_qring = k.sm_stage + (unsigned)k.ns * k.sm_slot;
    k.sm_total = k.sm_qring + (unsigned)(K > 1 ? (K - 1) * 3 : 0) * k.sm_pbox;
    k.tx_bytes = (unsigned)((H + 2) * (W + 4) * 8 + 2 * H * W * 8);
}

inline bool ptv_tile(PtV& k, int K, int nt, int want_pxt, int want_bty)
{
    const int ex = ((K - 1) + 1) / 2 * 2, ey = K - 1;
    const int pairs = (k.nx + 1) / 2;                 // pairs that hold a column of a row
    const int rows = k.ny;                            // thread rows that hold a row of a plane
    double best = -1.0;
    int best_pxt = 0, best_bty = 0;
    for (int ntx = 1; ntx <= 64; ++ntx) {
        int pxt;
        if (want_pxt > 0) {
            pxt = want_pxt;
        } else if (ntx == 1) {
            pxt = pairs;
        } else {
            // ntx tiles of W columns, stride W - 2 ex, must reach column nx-2: (ntx-1)(W-2ex) + W >= nx-1
            const int W = (k.nx - 1 + 2 * ex * (ntx - 1) + ntx - 1) / ntx;
            pxt = (W + 1) / 2;
        }
        if (pxt <= ex) pxt = ex + 1;                 // a tile must put out at least one pair
        if (pxt > nt || pxt > 126) continue;         // TMA boxes hold at most 256 elements per dimension (W + 4)
        if (want_pxt <= 0 && pxt < 32 && pxt < pairs) break;   // rows of a tile span at least one warp (512 contiguous bytes)
        int bty = want_bty > 0 ? want_bty : nt / pxt;
        if (bty > rows) bty = rows;
        if (bty < 1) bty = 1;
        if (pxt * bty > nt) continue;
        if (bty <= 2 * ey && bty < k.ny) {  // a tile must put out at least one row (unless it holds every row)
            if (want_pxt > 0) break;
            continue;
        }
        PtV t = k;
        ptv_tile_set(t, K, pxt, bty);
        const double useful = (double)(k.nx - 2) * (k.ny - 2);
        const double done = (double)t.ntx * (2 * pxt) * (double)t.nty * bty;
        const double threads_used = (double)(pxt * bty) / ((pxt * bty + 31) / 32 * 32);
        const double eff = useful / done * threads_used;
        if (eff > best * 1.02) {   // prefer fewer, wider tiles unless narrower ones win clearly
            best = eff;
            best_pxt = pxt;
            best_bty = bty;
        }
        if (want_pxt > 0) break;
    }
    if (best < 0) return false;
    ptv_tile_set(k, K, best_pxt, best_bty);
    return true;
}

inline size_t ptv_smem_bytes(const PtV& k, int K) { (void)K; return k.sm_total; }
inline int ptv_threads(const PtV& k) { return (k.pxt * k.bty + 31) / 32 * 32; }

// The neighbours' buffers as this rank sees them: {Pr A, Pr B, dPrdτ A, dPrdτ B} of the lower / upper neighbour
// (NULL without one), this rank's mailbox and the neighbours' mailboxes.
struct PtvPeers {
    double* lo[4];
    double* hi[4];
    unsigned long long* mbox;
    unsigned long long* lo_mbox;
    unsigned long long* hi_mbox;
};

// Where a launch on a slab interface reads and writes in its neighbours' memory.  All ranks ping-pong in
// lockstep, so a neighbour's buffers play the role of this rank's: w_new (0/1) is the Pr buffer this launch
// WRITES, w_dp (2/3) the dPrdτ buffer it READS.
inline void ptv_set_peers(PtV& k, const PtvPeers& pp, int w_new, int w_dp)
{
    const long long plane = k.planeB / 8;
    k.mbox = pp.mbox;
    k.peer_lo_plane = pp.lo[w_new] ? pp.lo[w_new] + (long long)(k.nz - 1) * plane : nullptr;  // its halo plane nz-1
    k.peer_hi_plane = pp.hi[w_new];                                                          // its halo plane 0
    k.peer_lo_flag = pp.lo_mbox ? pp.lo_mbox + NS3D_MB_FLAG_HI : nullptr;
    k.peer_hi_flag = pp.hi_mbox ? pp.hi_mbox + NS3D_MB_FLAG_LO : nullptr;
    const int w_cur = 1 - w_new;  // the neighbours' CURRENT iterate
    k.peer_lo_cur = pp.lo[w_cur] ? pp.lo[w_cur] + (long long)(k.nz - 3) * plane : nullptr;
    k.peer_hi_cur = pp.hi[w_cur] ? pp.hi[w_cur] + 2 * plane : nullptr;
    k.peer_lo_dp = pp.lo[w_dp] ? pp.lo[w_dp] + (long long)(k.nz - 2) * plane : nullptr;
    k.peer_hi_dp = pp.hi[w_dp] ? pp.hi[w_dp] + plane : nullptr;
}

// Balanced z-chunks whose last one keeps at least two planes: on slabs a neighbour reads plane nz-3 (resp. 2)
// of this rank, and the CTAs that own it are the ones holding the hand-over flag.
inline void ptv_balance_chunks(PtV& k)
{
    const int n = k.kend - k.kbeg;
    int nch = (n + k.zchunk - 1) / k.zchunk;
    if (nch < 1) nch = 1;
    int len = (n + nch - 1) / nch;
    if (nch > 1 && n - (nch - 1) * len == 1) {
        nch -= 1;
        len = (n + nch - 1) / nch;
    }
    k.zchunk = len > 0 ? len : 1;
}

}  // namespace
