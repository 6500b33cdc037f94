// libns3d.so -- the hot loop: fused pseudo-transient (PT) pressure iteration, host side.
//
// Reference (per PT iteration, M:459-463 / G:127-129): update_dPrdτ! (K5), update_Pr! (K6),
// set_bc_Pr! = 3-4 face kernels (K7) and up to three update_halo! calls: >= 5 synchronous
// launches and 7+ full-field passes.  The device code that replaces it lives in
// ns3d_pt_kernels.cuh (a header without CUDA runtime includes, so that the CPU test suite can
// execute the same source on host threads, tests/emu/):
//   pt_tb2s_kernel  two iterations per launch, 5 field passes per TWO iterations  (default)
//   pt_tb2_kernel   its first version; <.,.,true> also exchanges the slab halos over peer memory
//   pt_iter_kernel  one iteration per launch (odd trailing iteration, option tb2=0)
// This file holds what needs the runtime: kernel parameters and launch geometry (make_ptk),
// the ping-pong buffers, stream/event protocol and peer mappings of the slab path, CUDA-graph
// replay of iteration chunks, the residual reduction, and the entry points ns3d_pt_solve,
// ns3d_pt_iterate and ns3d_step.
//
// Arithmetic (template MODE): see NS3D_PARITY / NS3D_FAST / NS3D_FASTEST in ns3d.h.  The library
// is compiled with --fmad=false; FMA appears only where fma() is written explicitly.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ns3d_internal.cuh"
#include "ns3d_pt_kernels.cuh"

namespace {

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
    ptk_fill(p, k);
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
    // Two-iteration kernels recompute two extra stage-1 planes per chunk.  Measured (profiles/
    // r01_tb2s_sweep_*.jsonl): 12-plane chunks of 32x8 tiles at 255x153x153 (3.9 waves of 8-warp
    // CTAs), 64-plane chunks of 32x16 tiles at 511^3.
    const bool small = (double)(p->nx - 2) * (p->ny - 2) * (p->nz - 2) < 3.0e7;
    k->zchunk_tb = p->zchunk > 0 ? p->zchunk : (small ? 12 : 64);
    k->tb_ty = ctx->opt_tb2_ty ? ctx->opt_tb2_ty : (small ? 8 : 16);
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

PeerPtrs peer_ptrs(const ns3d_ctx* ctx, const PeerBufs& pb)
{
    PeerPtrs pp;
    for (int q = 0; q < 4; ++q) {
        pp.lo[q] = pb.lo[q];
        pp.hi[q] = pb.hi[q];
    }
    pp.mbox = ctx->mbox;
    pp.lo_mbox = ctx->peer_mbox[0];
    pp.hi_mbox = ctx->peer_mbox[1];
    return pp;
}

int launch_tb2(ns3d_ctx* ctx, cudaStream_t st, const PtK& k_in, const double* cur, double* nxt, const double* dpc,
               double* dpn, const double* divV, const PeerBufs& pb, const double* Pr_user, bool peer)
{
    PtK k = k_in;
    if (!k.faces) balance_chunks(k);
    if (peer) {
        ptk_set_peers(k, peer_ptrs(ctx, pb), (nxt == Pr_user) ? 0 : 1, (dpc == pb.dP_user) ? 2 : 3);
    }
    // round-2 candidate (unmeasured, off by default): two tile rows per thread, 32x16 tiles of 8 warps
    const int dual = k.mbox ? 0 : ctx->opt_tb2_dual;
    const int ty = (k.mbox || dual) ? 16 : k.tb_ty;
    const dim3 blk(TB_X, dual ? ty / 2 : ty, 1);
    const dim3 grd(cdiv(k.nx - 2, TB_X - 2), cdiv(k.ny - 2, ty - 2), k.faces ? 2u : cdiv(k.kend - k.kbeg, k.zchunk));
    // plain launches run the slim re-write (pt_tb2s_kernel, same results); chunks on a slab
    // interface need the peer loads/stores of pt_tb2_kernel<.,.,true>
    const bool slim = !k.mbox && ctx->opt_tb2_slim;
    // round-2 candidate (off by default): the slim pipeline on the slab interfaces too
    const bool slim_faces = k.mbox && ctx->opt_tb2_slim_faces;
    if (slim || dual || slim_faces) tb2s_set_offsets(k, cur, nxt, dpc, dpn, divV);
    // Grids whose x-y extent has a compile-time instantiation (default variant only): the
    // reference scripts' nx = 255 and BASELINE.json's 511^2 / 1023x511 planes.
#define TBS_ARGS <<<grd, blk, 0, st>>>(cur, nxt, dpc, dpn, divV, k)
#define TBS_LAUNCH(MODE, TY)                                                                                      \
    do {                                                                                                          \
        const int pf = ctx->opt_tb2_pf, np = ctx->opt_tb2_np;                                                     \
        const bool spec = ctx->opt_tb2_spec && np && pf == 1;                                                     \
        const bool pb = ctx->opt_tb2_pb && np && pf == 1 && TY != 32;  /* round-2 candidate, off by default */    \
        if (pb && spec && TY == 8 && k.nx == 255 && k.ny == 153) pt_tb2s_kernel<MODE, 8, 1, true, 255, 153, true> TBS_ARGS; \
        else if (pb && spec && TY == 16 && k.nx == 511 && k.ny == 511) pt_tb2s_kernel<MODE, 16, 1, true, 511, 511, true> TBS_ARGS; \
        else if (pb && TY == 8) pt_tb2s_kernel<MODE, 8, 1, true, 0, 0, true> TBS_ARGS;                            \
        else if (pb) pt_tb2s_kernel<MODE, 16, 1, true, 0, 0, true> TBS_ARGS;                                      \
        else if (spec && TY == 8 && k.nx == 255 && k.ny == 153) pt_tb2s_kernel<MODE, 8, 1, true, 255, 153> TBS_ARGS; \
        else if (spec && TY == 16 && k.nx == 511 && k.ny == 511) pt_tb2s_kernel<MODE, 16, 1, true, 511, 511> TBS_ARGS; \
        else if (spec && TY == 16 && k.nx == 1023 && k.ny == 511) pt_tb2s_kernel<MODE, 16, 1, true, 1023, 511> TBS_ARGS; \
        else if (np && pf == 2) pt_tb2s_kernel<MODE, TY, 2, true, 0, 0> TBS_ARGS;                                \
        else if (np && pf == 1) pt_tb2s_kernel<MODE, TY, 1, true, 0, 0> TBS_ARGS;                                \
        else if (np) pt_tb2s_kernel<MODE, TY, 0, true, 0, 0> TBS_ARGS;                                           \
        else pt_tb2s_kernel<MODE, TY, 0, false, 0, 0> TBS_ARGS; /* tb2_np=0: no prefetch at all (the baseline) */ \
    } while (0)
#define TBD_LAUNCH(MODE, MINB)                                                                                 \
    do {                                                                                                       \
        if (k.nx == 255 && k.ny == 153) pt_tb2d_kernel<MODE, 16, 1, MINB, 255, 153> TBS_ARGS;                  \
        else if (k.nx == 511 && k.ny == 511) pt_tb2d_kernel<MODE, 16, 1, MINB, 511, 511> TBS_ARGS;             \
        else pt_tb2d_kernel<MODE, 16, 1, MINB, 0, 0> TBS_ARGS;                                                 \
    } while (0)
#define TB_LAUNCH(MODE)                                                                                        \
    do {                                                                                                       \
        if (slim_faces) pt_tb2sp_kernel<MODE, 16, 1><<<grd, blk, 0, st>>>(cur, nxt, dpc, dpn, divV, k);        \
        else if (k.mbox) pt_tb2_kernel<MODE, 16, true><<<grd, blk, 0, st>>>(cur, nxt, dpc, dpn, divV, k);     \
        else if (dual) TBD_LAUNCH(MODE, 2);                                                                    \
        else if (slim && ty == 8) TBS_LAUNCH(MODE, 8);                                                         \
        else if (slim && ty == 32) TBS_LAUNCH(MODE, 32);                                                       \
        else if (slim) TBS_LAUNCH(MODE, 16);                                                                   \
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
#undef TBS_LAUNCH
#undef TBD_LAUNCH
#undef TBS_ARGS
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// Two iterations per launch is the default path wherever it applies (single rank, or slabs with
// the peer-memory halo path).  Measured against the one-iteration kernel: 39.4 -> 31.3 us per
// iteration at 255x153x153 (T_eff 7.06 TB/s for whole time steps), 856 -> 548 us at 511^3
// (9.7 TB/s, above the HBM copy peak); DESIGN.md 3.4-3.5, profiles/README.md.
// ns3d_set_option("tb2", 0) selects the one-iteration kernel.
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
        balance_chunks(q);
        ptk_set_peers(q, peer_ptrs(ctx, pb), (nxt == Pr_user) ? 0 : 1, -1);
        return launch_iter(ctx, ctx->stream, q, cur, nxt, dP, divV);
    }
    if (pb.on) {
        // The update of the two planes a slab sends and their delivery are ONE kernel: the face
        // CTAs store the new values into the neighbour's halo plane (peer memory over NVLink) and
        // hand over with mailbox flags.  It runs on the high-priority stream next to the launch
        // that updates the other planes; kernel-only, so whole chunks replay as a CUDA graph.
        PtK q = k, inner = k;
        q.faces = 1;
        ptk_set_peers(q, peer_ptrs(ctx, pb), (nxt == Pr_user) ? 0 : 1, -1);
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
    int n = 0, parity = 0, mode = 0, minb = 0, opts = 0;
    bool p2p = false;
    long long kernels = 0;
};
struct PtGraphCache {
    PtGraph slot[4];
    int next = 0;
};

// ---- z-band pipelining (option "pt_bands" = NB >= 2, single rank; CANDIDATE, not yet run on a device) ----
// One launch per two iterations leaves the SMs idle while its last CTAs drain and the next launch
// ramps up (ncu: ~10 % of a launch at 255x153x153).  Here every launch is split into NB z-bands on
// NB streams, and band b of launch n+1 depends only on bands b-1, b, b+1 of launch n -- the planes
// it reads reach two planes into the adjacent bands, and those same launches are the last readers
// of what it overwrites (the WAR dependency is the RAW dependency one launch later).  Band kernels
// of consecutive launches then overlap like a wavefront; inside a captured chunk the dependencies
// become graph edges.  Plane ranges compose exactly (tests/test_kernel_emu.py, split launches).
int bands_prepare(ns3d_ctx* ctx, int nb)
{
    if (ctx->bands_ready >= nb) return NS3D_OK;
    if (!ctx->band_fork) NS3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->band_fork, cudaEventDisableTiming));
    for (int b = ctx->bands_ready; b < nb; ++b) {
        NS3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->band_stream[b], cudaStreamNonBlocking));
        for (int q = 0; q < 2; ++q) NS3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->band_ev[q][b], cudaEventDisableTiming));
    }
    ctx->bands_ready = nb;
    return NS3D_OK;
}

// n2 double launches (2*n2 iterations) as NB pipelined bands; the ping-pong pointers advance as in run_direct.
int run_bands(ns3d_ctx* ctx, const PtK& k2, int nb, double*& cur, double*& nxt, double*& dP, double*& dPn,
              const double* divV, int n2, int iter0, const PeerBufs& pb, const double* Pr_user)
{
    NS3D_TRY(bands_prepare(ctx, nb));
    const int planes = k2.kend - k2.kbeg;
    const int nchunks = (planes + k2.zchunk - 1) / k2.zchunk;
    const int per_band = (nchunks + nb - 1) / nb * k2.zchunk;  // bands are whole chunks (the last one may be shorter)
    NS3D_CUDA(ctx, cudaEventRecord(ctx->band_fork, ctx->stream));
    for (int b = 0; b < nb; ++b) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->band_stream[b], ctx->band_fork, 0));
    for (int q = 0; q < n2; ++q) {
        const int par = q & 1;
        for (int b = 0; b < nb; ++b) {
            PtK kb = k2;
            kb.kbeg = k2.kbeg + b * per_band;
            kb.kend = std::min(kb.kbeg + per_band, k2.kend);
            if (kb.kbeg >= kb.kend) continue;  // nb was clamped so that this cannot happen; belt and braces
            kb.reverse = k2.serpentine && (((iter0 >> 1) + q) & 1);
            if (q > 0) {  // launch q-1 of the adjacent bands (this band's own is ordered by its stream)
                if (b > 0) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->band_stream[b], ctx->band_ev[1 - par][b - 1], 0));
                if (b < nb - 1) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->band_stream[b], ctx->band_ev[1 - par][b + 1], 0));
            }
            NS3D_TRY(launch_tb2(ctx, ctx->band_stream[b], kb, cur, nxt, dP, dPn, divV, pb, Pr_user, false));
            NS3D_CUDA(ctx, cudaEventRecord(ctx->band_ev[par][b], ctx->band_stream[b]));
        }
        double* t = cur; cur = nxt; nxt = t;
        t = dP; dP = dPn; dPn = t;
    }
    const int last = (n2 - 1) & 1;
    for (int b = 0; b < nb; ++b) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->band_ev[last][b], 0));  // join
    return NS3D_OK;
}

int run_direct(ns3d_ctx* ctx, PtK& k, double*& cur, double*& nxt, double*& dP, double*& dPn, const double* divV, int n,
               int iter0, const PeerBufs& pb, const double* Pr_user)
{
    NS3D_TRY(pt_begin(ctx));
    int q = 0;
    if (dPn && ctx->opt_pt_bands >= 2 && ctx->nranks == 1 && n >= 2) {
        PtK k2 = k;
        k2.zchunk = k.zchunk_tb;
        balance_chunks(k2);
        const int nchunks = (k2.kend - k2.kbeg + k2.zchunk - 1) / k2.zchunk;
        const int nb = std::min(ctx->opt_pt_bands, nchunks);  // every band gets at least one chunk
        const int per_band_chunks = (nchunks + nb - 1) / nb;
        const int nb_eff = (nchunks + per_band_chunks - 1) / per_band_chunks;
        // a band must hold two planes at least: a launch reads two planes beyond its own range, and only
        // the ADJACENT bands of the previous launch are waited for
        if (nb_eff >= 2 && per_band_chunks * k2.zchunk >= 2) {
            NS3D_TRY(run_bands(ctx, k2, nb_eff, cur, nxt, dP, dPn, divV, n / 2, iter0, pb, Pr_user));
            q = n & ~1;
        }
    }
    if (dPn && q == 0) {  // two iterations per launch; Pr and dPrdτ both ping-pong
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
    // streams and events of the band pipeline are created outside any capture
    if (ctx->opt_pt_bands >= 2 && ctx->nranks == 1) NS3D_TRY(bands_prepare(ctx, ctx->opt_pt_bands));
    const bool graphable = ctx->opt_graphs && n >= 8 && (ctx->nranks == 1 || pb.on);
    if (!graphable) return run_direct(ctx, k, cur, nxt, dP, dPn, divV, n, iter0, pb, Pr_user);
    if (!ctx->pt_graphs) ctx->pt_graphs = new PtGraphCache();
    PtGraphCache* cache = (PtGraphCache*)ctx->pt_graphs;
    PtK key;
    memcpy(&key, &k, sizeof key);  // byte copy: the cache compares with memcmp (padding included)
    key.reverse = 0;
    // every tuning option that selects a kernel or its launch shape is part of the key
    const int opts = ctx->opt_tb2 | (ctx->opt_tb2_slim << 1) | (ctx->opt_tb2_np << 2) | (ctx->opt_tb2_pf << 3) |
                     (ctx->opt_tb2_spec << 5) | (ctx->opt_tb2_dual << 6) | (ctx->opt_tb2_ty << 8) | (ctx->opt_tb2_pb << 16) |
                     (ctx->opt_pt_bands << 20) | (ctx->opt_tb2_slim_faces << 24);
    PtGraph* g = nullptr;
    for (PtGraph& c : cache->slot)
        if (c.exec && c.cur == cur && c.nxt == nxt && c.dP == dP && c.dPn == dPn && c.divV == divV && c.n == n &&
            c.parity == (iter0 & 3) && c.mode == ctx->mode && c.minb == ctx->opt_pt_minb && c.opts == opts && c.p2p == pb.on &&
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
        g->parity = iter0 & 3; g->p2p = pb.on; g->mode = ctx->mode; g->minb = ctx->opt_pt_minb; g->opts = opts; g->kernels = captured;
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
    if (ctx->opt_ptv) return ns3d_internal_ptv_solve(ctx, Pr, dPrdtau, divV, p, h_iters, h_err_hist, err_cap, h_nchecks);
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
    if (ctx->opt_ptv) return ns3d_internal_ptv_iterate(ctx, Pr, dPrdtau, divV, p, n_iter);
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

extern "C" int ns3d_pt_describe(ns3d_ctx* ctx, const ns3d_pt_params* p, char* buf, int cap, int* iters_per_launch)
{
    NS3D_CHECK_CTX(ctx);
    PtK k;
    NS3D_TRY(make_ptk(ctx, p, &k));
    if (ctx->opt_ptv) return ns3d_internal_ptv_describe(ctx, p, buf, cap, iters_per_launch);
    const bool slabs = ctx->nranks > 1;
    const bool tb2 = use_tb2(ctx, p, !slabs || (ctx->opt_p2p && ctx->p2p_ready && k.nz >= 6));
    const char* mode = ctx->mode == NS3D_PARITY ? "PARITY" : (ctx->mode == NS3D_FAST ? "FAST" : "FASTEST");
    if (buf && cap > 0) {
        if (tb2)
            snprintf(buf, cap, "pt_tb2s_kernel<%s,TY=%d> (two fused PT iterations per launch: 2 x (K5+K6+set_bc_Pr!), %d-plane chunks%s)",
                     mode, k.tb_ty, k.zchunk_tb, slabs ? "; slab-interface chunks: pt_tb2_kernel<.,16,true> with update_halo!(Pr) over peer memory" : "");
        else
            snprintf(buf, cap, "pt_iter_kernel<%s> (fused K5+K6+set_bc_Pr!, one PT iteration per launch)", mode);
    }
    if (iters_per_launch) *iters_per_launch = tb2 ? 2 : 1;
    return NS3D_OK;
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
