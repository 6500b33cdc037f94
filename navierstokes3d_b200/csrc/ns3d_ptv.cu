// libns3d.so -- the hot loop on the library's internal pitched layout: host side of ptv_kernel
// (ns3d_ptv_kernels.cuh), K fused pseudo-transient iterations per launch.
//
// Reference (per PT iteration, M:459-463 / G:127-129): update_dPrdτ! (K5), update_Pr! (K6), set_bc_Pr! = 3-4
// face kernels (K7) and up to three update_halo! calls.  ns3d_pt_solve / ns3d_pt_iterate (ns3d_pt.cu) call
// in here: the caller's Pr, dPrdτ and ∇V are packed into pitched, 128-byte-row-aligned copies owned by the
// context (two Pr buffers and two dPrdτ buffers that ping-pong, one ∇V), the loop runs on those -- replayed
// as CUDA graphs chunk by chunk, on z-slabs with the halo exchange fused into the kernels over peer memory --
// and the result is unpacked into the caller's arrays: 10 field passes per solve against 5 per K iterations.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ns3d_internal.cuh"
#ifndef NS3D_HOST_EMU
#include <cuda.h>   // CUtensorMap and the enums of cuTensorMapEncodeTiled (the entry point itself comes from the runtime)
#endif
#include "ns3d_ptv_kernels.cuh"

namespace {

const int PTV_HMAX = 64;  // tallest tile (rows): the buffers are padded so that a tile may read past the last row

// compute_res! + abs + maximum (K8 + K8') in one pass over the pitched arrays, no Rp array: max over the
// interior of the bit pattern of |bracket| (NaN-propagating, see absbits()).
template <int MODE>
__global__ void __launch_bounds__(256) ptv_residual_kernel(const PtV p, unsigned long long* __restrict__ out)
{
    const int nx = p.nx, ny = p.ny, nz = p.nz;
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    unsigned long long m = 0ULL;
    if (i <= nx - 2 && j <= ny - 2) {
        const int kb = p.kbeg + blockIdx.z * p.zchunk;
        const int ke = min(kb + p.zchunk, p.kend);
        const size_t row = (size_t)p.px, sxy = (size_t)p.px * ny;
        const double* c = p.P + (size_t)kb * sxy + (size_t)j * row + i;
        const double* dv = p.V + (size_t)kb * sxy + (size_t)j * row + i;
        double pm = c[-(ptrdiff_t)sxy];
        double pc = c[0];
        for (int k = kb; k < ke; ++k) {
            const double pp = c[sxy];
            const double L = ptv_bracket<MODE>(p, pc, c[-1], c[1], c[-(ptrdiff_t)row], c[row], pm, pp, dv[0]);
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

// Iterations per launch / rows per thread / CTA size: the options, else the measured defaults.
PtvPlan make_plan(const ns3d_ctx* ctx, bool slabs)
{
    PtvPlan pl;
    pl.K = ctx->opt_ptv_k > 0 ? ctx->opt_ptv_k : 2;
    if (slabs && pl.K > 2) pl.K = 2;  // the peer-memory halo exchange provides one plane per side
    // 256-thread CTAs, three per SM (80 registers) for K <= 2, two per SM (128 registers) for K = 3 (measured)
    pl.lb = ctx->opt_ptv_lb >= 0 ? ctx->opt_ptv_lb : (pl.K >= 3 ? 0 : 1);
    pl.ns = ctx->opt_ptv_ns >= 3 ? ctx->opt_ptv_ns : 4;
    return pl;
}

// z-chunks: short enough that the SMs get several waves of CTAs (the hardware scheduler then balances them; with
// about one wave some SMs hold 3 CTAs and others 2 until the end), long enough to amortise the 2(K-1) planes every
// chunk recomputes and its prologue.  Measured at 255x153x153 (profiles/r02_ptv_sweep_B.jsonl): 10-13 planes beat 19
// and 38; at 511^3 the tiles alone give more than four waves and chunks are a hundred planes long.
int auto_zchunk(const ns3d_ctx* ctx, const PtV& k, int K, int planes, int lb, bool flow = false)
{
    const long long tiles = (long long)k.ntx * k.nty;
    const int ctas = std::max(1, std::min(ptv_lb_ctas(lb), (int)(227u * 1024u / std::max(1u, k.sm_total + 1024u))));
    const long long slots = (long long)ctx->num_sms * ctas;   // CTAs resident at a time: registers, shared memory
    // persistent launch: no waves to fill -- the chunks only have to be numerous enough that the items an item waits for
    // (the adjacent chunks of the previous launch) were claimed about two sets of resident CTAs earlier
    const long long want = flow ? 2 * slots + 2 * tiles : 4 * slots;   // CTAs for four waves
    long long nch = (want + tiles - 1) / tiles;
    if (nch < 1) nch = 1;
    int len = (int)((planes + nch - 1) / nch);
    const int min_len = 4 * K + 2;
    if (len < min_len) len = min_len;
    // ... and no longer than 32 planes: the CTAs of a wave then stay close enough in z for the tile halos their
    // neighbours re-read to be L2 hits (511^3: 16 planes 667 us/iteration, 32: 613, 64: 659, 170: 899)
    if (len > 32) len = 32;
    if (len > planes) len = planes;
    return len;
}

int make_ptv(ns3d_ctx* ctx, const ns3d_pt_params* p, const PtvPlan& pl, int K, PtV* k, bool flow = false)
{
    memset(k, 0, sizeof *k);
    ptv_fill(p, k);
    k->zlo_halo = ctx->nranks > 1 && ctx->rank > 0;
    k->zhi_halo = ctx->nranks > 1 && ctx->rank < ctx->nranks - 1;
    k->ns = pl.ns;
    // default tile: 32 x 16 cells (16 x 16 threads), the shape with a compile-time instantiation that measured best at
    // 255x153x153 and 511^3; small grids and explicit requests go through the general chooser
    int want_pxt = ctx->opt_ptv_pxt, want_bty = ctx->opt_ptv_bty;
    if (want_pxt == 0 && want_bty == 0 && p->nx >= 64 && p->ny >= 32 && ptv_lb_threads(pl.lb) == 256) {
        want_pxt = 16;
        want_bty = 16;
    }
    if (!ptv_tile(*k, K, ptv_lb_threads(pl.lb), want_pxt, want_bty))
        return ns3d_fail(ctx, NS3D_EINVAL, "pt: no tile shape for K = %d, %d threads (ptv_pxt = %d, ptv_bty = %d)", K,
                         ptv_lb_threads(pl.lb), ctx->opt_ptv_pxt, ctx->opt_ptv_bty);
    if (k->sm_total > 227u * 1024u)
        return ns3d_fail(ctx, NS3D_EINVAL, "pt: tile of %d x %d threads needs %u B of shared memory", k->pxt, k->bty, k->sm_total);
    if (k->bty > PTV_HMAX || ptv_threads(*k) > ptv_lb_threads(pl.lb))
        return ns3d_fail(ctx, NS3D_EINVAL, "pt: tile of %d x %d threads does not fit the kernel (at most %d threads, %d rows)", k->pxt,
                         k->bty, ptv_lb_threads(pl.lb), PTV_HMAX);
    k->zchunk = p->zchunk > 0 ? p->zchunk : auto_zchunk(ctx, *k, K, p->nz - 2, pl.lb, flow);
    return NS3D_OK;
}

// ---- the pitched buffers -------------------------------------------------------------------------------------
size_t ptv_front_pad(int px) { return 2 * (size_t)px + 16; }
size_t ptv_alloc_count(int px, int ny, int nz) { return ptv_front_pad(px) + (size_t)px * ny * nz + (size_t)(PTV_HMAX + 2) * px + 16; }

void ptv_unmap(ns3d_ctx* ctx, void* base)
{
    auto it = ctx->p2p_map.find(base);
    if (it == ctx->p2p_map.end()) return;
    if (it->second.first) cudaIpcCloseMemHandle(it->second.first);
    if (it->second.second) cudaIpcCloseMemHandle(it->second.second);
    ctx->p2p_map.erase(it);
}

int ptv_ensure(ns3d_ctx* ctx, int nx, int ny, int nz)
{
    if (ctx->ptv_raw[0] && ctx->ptv_nx == nx && ctx->ptv_ny == ny && ctx->ptv_nz == nz) return NS3D_OK;
    const int px = ptv_pitch(nx);
    const size_t count = ptv_alloc_count(px, ny, nz);
    if (ctx->ptv_raw[0]) {
        NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
        ns3d_internal_ptv_free(ctx);
        for (int q = 0; q < 5; ++q) {
            ptv_unmap(ctx, ctx->ptv_raw[q]);
            cudaFree(ctx->ptv_raw[q]);
            ctx->ptv_raw[q] = nullptr;
        }
        ctx->ptv_peers_mapped = false;
    }
    for (int q = 0; q < 5; ++q) {
        cudaError_t e = cudaMalloc(&ctx->ptv_raw[q], count * sizeof(double));
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int r = 0; r < q; ++r) {
                cudaFree(ctx->ptv_raw[r]);
                ctx->ptv_raw[r] = nullptr;
            }
            return ns3d_fail(ctx, NS3D_ENOMEM, "pt: cannot allocate the pitched working copies (5 x %zu B)", count * sizeof(double));
        }
        NS3D_CUDA(ctx, cudaMemsetAsync(ctx->ptv_raw[q], 0, count * sizeof(double), ctx->stream));
        ctx->ptv[q] = ctx->ptv_raw[q] + ptv_front_pad(px);
    }
    ctx->ptv_nx = nx; ctx->ptv_ny = ny; ctx->ptv_nz = nz;
    return NS3D_OK;
}

// Peer-memory path usable for this solve?  Maps the neighbours' four ping-pong buffers on first use
// (COLLECTIVE: every rank reaches this point with its own buffers).
int ptv_peer_prepare(ns3d_ctx* ctx, int nz, PtvPeers* pp, bool* on)
{
    *on = false;
    memset(pp, 0, sizeof *pp);
    if (ctx->nranks == 1 || !ctx->opt_p2p || !ctx->p2p_ready || nz < 6) return NS3D_OK;
    const size_t pad = ptv_front_pad(ptv_pitch(ctx->ptv_nx));
    for (int q = 0; q < 4; ++q) {
        void *lo = nullptr, *hi = nullptr;
        NS3D_TRY(ns3d_internal_p2p_map(ctx, ctx->ptv_raw[q], &lo, &hi));
        pp->lo[q] = lo ? (double*)lo + pad : nullptr;
        pp->hi[q] = hi ? (double*)hi + pad : nullptr;
    }
    pp->mbox = ctx->mbox;
    pp->lo_mbox = ctx->peer_mbox[0];
    pp->hi_mbox = ctx->peer_mbox[1];
    ctx->ptv_peers_mapped = true;
    *on = true;
    return NS3D_OK;
}

// Work queue and completion counters of the persistent launch, for up to `iters` iterations in one launch
// (at most one counter per launch and plane).  Allocated outside stream capture; growing it drops the cached graphs,
// which hold its address.
int ptv_work_ensure(ns3d_ctx* ctx, int iters, int nz)
{
    const size_t words = PTV_WORK_DONE + (size_t)std::max(iters, 1) * (size_t)nz + 32;   // last word: the sticky error flag
    if (words <= ctx->ptv_work_words) return NS3D_OK;
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ns3d_internal_ptv_free(ctx);
    if (ctx->ptv_work) cudaFree(ctx->ptv_work);
    ctx->ptv_work = nullptr;
    ctx->ptv_work_words = 0;
    NS3D_CUDA(ctx, cudaMalloc(&ctx->ptv_work, words * sizeof(unsigned)));
    NS3D_CUDA(ctx, cudaMemsetAsync(ctx->ptv_work, 0, words * sizeof(unsigned), ctx->stream));
    ctx->ptv_work_words = words;
    ctx->h_maxbits[5] = 0ULL;
    return NS3D_OK;
}

int ptv_pack(ns3d_ctx* ctx, double* dst, const double* src, int nx, int ny, int nz, int inner)
{
    const int px = ptv_pitch(nx);
    const size_t n = (size_t)px * ny * nz;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->num_sms * 32);
    ptv_pack_kernel<<<blocks, 256, 0, ctx->stream>>>(dst, src, nx, ny, nz, px, inner);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

int ptv_unpack(ns3d_ctx* ctx, double* dst, const double* src, int nx, int ny, int nz, int inner)
{
    const size_t n = inner ? (size_t)(nx - 2) * (ny - 2) * (nz - 2) : (size_t)nx * ny * nz;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->num_sms * 32);
    ptv_unpack_kernel<<<blocks ? blocks : 1, 256, 0, ctx->stream>>>(dst, src, nx, ny, nz, ptv_pitch(nx), inner);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

}  // namespace

// ---- launches ---------------------------------------------------------------------------------------------------
// ptv_kernel is instantiated per arithmetic mode in its own translation unit (ns3d_ptv_mode{0,1,2}.cu, compiled
// in parallel): K x rows per thread x launch bounds x {plain, peer-memory} instantiations each.
int ns3d_internal_ptv_launch_parity(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool p2p,
                                    bool tma, dim3 grid, size_t smem);
int ns3d_internal_ptv_launch_fast(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool p2p,
                                  bool tma, dim3 grid, size_t smem);
int ns3d_internal_ptv_launch_fastest(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool p2p,
                                     bool tma, dim3 grid, size_t smem);
int ns3d_internal_ptv_flow_launch_parity(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool tma,
                                         size_t smem);
int ns3d_internal_ptv_flow_launch_fast(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool tma,
                                       size_t smem);
int ns3d_internal_ptv_flow_launch_fastest(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool tma,
                                          size_t smem);

namespace {

// ---- TMA descriptors ---------------------------------------------------------------------------------------------
#ifndef NS3D_HOST_EMU
typedef CUresult (*PtvEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PtvEncodeTiled ptv_encode_fn()
{
    static PtvEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PtvEncodeTiled)f;
        else
            cudaGetLastError();
    }
    return fn;
}
#endif

// 3-D tensor map over one pitched array (px x ny x nz doubles), box = bx x by x 1: the tile of a CTA in one z-plane.
int ptv_make_map(ns3d_ctx* ctx, PtvMaps::Map* out, const double* base, int px, int ny, int nz, int bx, int by)
{
#ifndef NS3D_HOST_EMU
    static_assert(sizeof(CUtensorMap) == sizeof(PtvMaps::Map), "CUtensorMap is 128 bytes");
    PtvEncodeTiled enc = ptv_encode_fn();
    if (!enc) return ns3d_fail(ctx, NS3D_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)px, (cuuint64_t)ny, (cuuint64_t)nz};
    const cuuint64_t strides[2] = {(cuuint64_t)px * 8, (cuuint64_t)px * ny * 8};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ns3d_fail(ctx, NS3D_ECUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d box", (int)r, bx, by);
#else
    (void)ctx; (void)out; (void)base; (void)px; (void)ny; (void)nz; (void)bx; (void)by;
#endif
    return NS3D_OK;
}

// One launch = K iterations of the planes [k.kbeg, k.kend) (or of the two interface chunks when k.faces);
// nlaunch > 0: the persistent launch, nlaunch x K iterations.
int ptv_launch(ns3d_ctx* ctx, cudaStream_t st, PtV k, const PtvPlan& pl, int K, bool p2p, int nlaunch = 0)
{
    if (!k.faces) ptv_balance_chunks(k);
    const dim3 grid((unsigned)(k.ntx * k.nty), k.faces ? 2u : cdiv(k.kend - k.kbeg, k.zchunk), 1);
    const size_t smem = ptv_smem_bytes(k, K);
    PtvMaps maps;
    memset(&maps, 0, sizeof maps);
    bool tma = false;
#ifndef NS3D_HOST_EMU
    tma = ctx->opt_ptv_tma != 0;
    if (tma) {
        const int W = 2 * k.pxt, H = k.bty;
        NS3D_TRY(ptv_make_map(ctx, &maps.m[0], k.P, k.px, k.ny, k.nz, W + 4, H + 2));
        NS3D_TRY(ptv_make_map(ctx, &maps.m[1], k.D, k.px, k.ny, k.nz, W, H));
        NS3D_TRY(ptv_make_map(ctx, &maps.m[2], k.V, k.px, k.ny, k.nz, W, H));
        if (nlaunch > 0) {   // odd launches read what even ones write
            NS3D_TRY(ptv_make_map(ctx, &maps.m[3], k.PN, k.px, k.ny, k.nz, W + 4, H + 2));
            NS3D_TRY(ptv_make_map(ctx, &maps.m[4], k.DN, k.px, k.ny, k.nz, W, H));
        }
    }
#endif
    if (nlaunch > 0) {
        k.nlaunch = nlaunch;
        k.nbz = (int)grid.y;
        const size_t words = PTV_WORK_DONE + (size_t)nlaunch * k.nbz;
        if (words + 32 > ctx->ptv_work_words)   // ptv_work_ensure() sizes it ahead of any stream capture
            return ns3d_fail(ctx, NS3D_EINVAL, "pt: the work queue holds %zu words, %zu needed", ctx->ptv_work_words, words);
        k.work = ctx->ptv_work;
        k.work_err = ctx->ptv_work + ctx->ptv_work_words - 1;
        NS3D_CUDA(ctx, cudaMemsetAsync(ctx->ptv_work, 0, words * sizeof(unsigned), st));
        switch (ctx->mode) {
            case NS3D_PARITY: return ns3d_internal_ptv_flow_launch_parity(ctx, st, k, maps, pl, K, tma, smem);
            case NS3D_FAST: return ns3d_internal_ptv_flow_launch_fast(ctx, st, k, maps, pl, K, tma, smem);
            default: return ns3d_internal_ptv_flow_launch_fastest(ctx, st, k, maps, pl, K, tma, smem);
        }
    }
    switch (ctx->mode) {
        case NS3D_PARITY: return ns3d_internal_ptv_launch_parity(ctx, st, k, maps, pl, K, p2p, tma, grid, smem);
        case NS3D_FAST: return ns3d_internal_ptv_launch_fast(ctx, st, k, maps, pl, K, p2p, tma, grid, smem);
        default: return ns3d_internal_ptv_launch_fastest(ctx, st, k, maps, pl, K, p2p, tma, grid, smem);
    }
}

struct PtvRun {
    const ns3d_pt_params* p;
    PtvPlan pl;
    PtvPeers peers;
    bool peer_on = false;
    int cur = 0;  // index of the buffers that hold the current iterate: Pr = ptv[cur], dPrdτ = ptv[2 + cur]
};

void ptv_bind(const ns3d_ctx* ctx, PtV& k, int cur)
{
    k.P = ctx->ptv[cur];
    k.PN = ctx->ptv[1 - cur];
    k.D = ctx->ptv[2 + cur];
    k.DN = ctx->ptv[3 - cur];
    k.V = ctx->ptv[4];
}

// ---- z-bands (single rank) -------------------------------------------------------------------------------------------
// ncu of ptv_kernel at 255x153x153: the SMs are idle 17 % of a launch (ramp, tail, 3.1 waves of CTAs), and a pass cannot
// start before the last CTA of the previous one has finished.  Here a pass is `nb` launches, one per band of z-chunks, on
// `nb` streams; band b of pass n+1 depends on bands b-1, b, b+1 of pass n only -- its reads reach K planes into the
// adjacent bands, and those launches are also the last readers of what it overwrites -- so the first bands of the next
// pass run while the last bands of this one drain.  The dependencies are events (graph edges inside a captured chunk);
// no device-side waiting (cf. ptv_flow_kernel).
int ptv_bands_ensure(ns3d_ctx* ctx, int nb)
{
    if (!ctx->band_fork) NS3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->band_fork, cudaEventDisableTiming));
    for (int b = 0; b < nb; ++b) {
        if (!ctx->band_stream[b]) NS3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->band_stream[b], cudaStreamNonBlocking));
        for (int q = 0; q < 2; ++q)
            if (!ctx->band_ev[q][b]) NS3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->band_ev[q][b], cudaEventDisableTiming));
    }
    return NS3D_OK;
}

// Bands for L passes of K iterations: how many (0 = do not band); may lengthen k.zchunk (chunks need not fill waves any more).
// Measured at 255x153x153 (profiles/r02_bands_sweep_B.jsonl): 8 bands of one 16-19-plane chunk each, 25.5 us per iteration
// against 32.0 for one launch per pass; 511^3 (25 waves of CTAs per launch) loses 4-6 % with bands.
int ptv_bands_plan(const ns3d_ctx* ctx, PtV& k, const PtvPlan& pl, int K, int L, bool zchunk_given)
{
    int nb = ctx->opt_ptv_bands;
    const int planes = k.kend - k.kbeg;
    if (nb == 0 || nb == 1 || L < 2) return 0;
    if (nb < 0) {
        // by default only where a launch is a few waves of CTAs: there its ramp and tail are a good part of it
        const long long tiles = (long long)k.ntx * k.nty;
        const int ctas = std::max(1, std::min(ptv_lb_ctas(pl.lb), (int)(227u * 1024u / std::max(1u, k.sm_total + 1024u))));
        const long long nch = (planes + k.zchunk - 1) / k.zchunk;
        if (tiles * nch >= 6LL * ctx->num_sms * ctas) return 0;
        nb = 8;
    }
    if (nb > 16) nb = 16;
    if (!zchunk_given) {   // one chunk per band, at most 32 planes long (see auto_zchunk)
        int len = (planes + nb - 1) / nb;
        len = std::max(len, 4 * K + 2);
        len = std::min(std::min(len, 32), planes);
        k.zchunk = len;
        ptv_balance_chunks(k);
    }
    const int nch = (planes + k.zchunk - 1) / k.zchunk;
    if (nb > nch) nb = nch;
    if (nb < 3 || k.zchunk < K) return 0;   // fewer than three bands overlap nothing
    return nb;
}

// k: the planes the bands cover.  Slabs (faces != NULL): those are the planes between the interface chunks, and the two
// interface chunks -- one launch of the P2P instantiation on the high-priority stream, as without bands -- are one more
// "band" F next to band 0 and band nb-1: F(n+1) waits for F(n), band 0 (n) and band nb-1 (n); bands 0 and nb-1 wait for F.
int ptv_run_banded(ns3d_ctx* ctx, PtvRun& r, PtV k, int K, int L, int nb, const PtV* faces)
{
    NS3D_TRY(ptv_bands_ensure(ctx, nb));
    const int planes = k.kend - k.kbeg, zc = k.zchunk;
    const int nch = (planes + zc - 1) / zc;
    int first[17];   // band b = chunks [first[b], first[b+1])
    for (int b = 0; b <= nb; ++b) first[b] = (int)((long long)nch * b / nb);
    NS3D_CUDA(ctx, cudaEventRecord(ctx->band_fork, ctx->stream));
    for (int b = 0; b < nb; ++b) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->band_stream[b], ctx->band_fork, 0));
    if (faces) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->band_fork, 0));
    cudaEvent_t evF[2] = {ctx->ev_a, ctx->ev_b};   // the interface launches of even / odd passes
    const int kbeg0 = k.kbeg, kend0 = k.kend;
    for (int l = 0; l < L; ++l) {
        ptv_bind(ctx, k, r.cur);
        if (faces) {
            PtV f = *faces;
            ptv_bind(ctx, f, r.cur);
            ptv_set_peers(f, r.peers, 1 - r.cur, 2 + r.cur);
            if (l > 0) {
                NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->band_ev[(l - 1) & 1][0], 0));
                NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->band_ev[(l - 1) & 1][nb - 1], 0));
            }
            NS3D_TRY(ptv_launch(ctx, ctx->comm_stream, f, r.pl, K, true, 0));
            NS3D_CUDA(ctx, cudaEventRecord(evF[l & 1], ctx->comm_stream));
        }
        for (int b = 0; b < nb; ++b) {
            if (l > 0) {
                for (int q = std::max(b - 1, 0); q <= std::min(b + 1, nb - 1); ++q)
                    if (q != b) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->band_stream[b], ctx->band_ev[(l - 1) & 1][q], 0));
                if (faces && (b == 0 || b == nb - 1)) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->band_stream[b], evF[(l - 1) & 1], 0));
            }
            PtV kb = k;
            kb.kbeg = kbeg0 + first[b] * zc;
            kb.kend = std::min(kbeg0 + first[b + 1] * zc, kend0);
            kb.zchunk = zc;
            kb.reverse = 0;
            kb.mbox = nullptr;
            NS3D_TRY(ptv_launch(ctx, ctx->band_stream[b], kb, r.pl, K, false, 0));
            NS3D_CUDA(ctx, cudaEventRecord(ctx->band_ev[l & 1][b], ctx->band_stream[b]));
        }
        r.cur = 1 - r.cur;
    }
    for (int b = 0; b < nb; ++b) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->band_ev[(L - 1) & 1][b], 0));
    if (faces) NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, evF[(L - 1) & 1], 0));
    return NS3D_OK;
}

// n iterations as launches of K (then of fewer) iterations; on z-slabs the two chunks next to the interfaces
// (peer loads / stores, mailbox flags) run on the high-priority stream beside the launch that updates the
// other planes.  Event protocol (ev_a = "main stream finished reading the iterate that is about to be
// overwritten", ev_b = "interface chunks of the new iterate are done"):
//     comm:  wait ev_a(n-1)   faces(n)      record ev_b(n)
//     main:  interior(n)      record ev_a(n)  wait ev_b(n)
int ptv_run_direct(ns3d_ctx* ctx, PtvRun& r, int n, int iter0)
{
    const ns3d_pt_params* p = r.p;
    if (ctx->nranks > 1) NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    int done = 0, launch = 0;
    // single rank: all full launches of K iterations as ONE persistent launch (ptv_flow_kernel); an odd rest follows below
    if (ctx->nranks == 1 && ctx->opt_ptv_flow && ptv_lb_threads(r.pl.lb) == 256 && n / r.pl.K >= 2) {
        const int K = r.pl.K, L = n / K;
        PtV k;
        NS3D_TRY(make_ptv(ctx, p, r.pl, K, &k, true));
        PtV probe = k;
        ptv_balance_chunks(probe);
        // an item waits for its own chunk and the two next to it in the previous launch: its planes +- K must lie in them
        if (probe.zchunk >= K) {
            ptv_bind(ctx, k, r.cur);
            NS3D_TRY(ptv_launch(ctx, ctx->stream, k, r.pl, K, false, L));
            if (L & 1) r.cur = 1 - r.cur;
            done += L * K;
            launch += L;
        }
    }
    // single rank: the full passes as bands of launches that overlap consecutive passes
    if (ctx->nranks == 1 && done == 0 && n / r.pl.K >= 2) {
        const int K = r.pl.K, L = n / K;
        PtV k;
        NS3D_TRY(make_ptv(ctx, p, r.pl, K, &k));
        ptv_balance_chunks(k);
        const int nb = ptv_bands_plan(ctx, k, r.pl, K, L, p->zchunk > 0);
        if (nb > 0) {
            NS3D_TRY(ptv_run_banded(ctx, r, k, K, L, nb, nullptr));
            done += L * K;
            launch += L;
        }
    }
    // z-slabs over peer memory, split launches: the planes between the interface chunks in bands, the interface chunks as
    // one more band on the high-priority stream
    if (r.peer_on && ctx->opt_p2p_split && done == 0 && n / r.pl.K >= 2 && (p->nz - 2) >= 2 * r.pl.zf + 4) {
        const int K = r.pl.K, L = n / K, zf = r.pl.zf;
        PtV k;
        NS3D_TRY(make_ptv(ctx, p, r.pl, K, &k));
        PtV f = k, in = k;
        f.faces = 1;
        f.zchunk = zf;
        in.kbeg = 1 + zf;
        in.kend = p->nz - 1 - zf;
        if (p->zchunk <= 0) in.zchunk = auto_zchunk(ctx, in, K, in.kend - in.kbeg, r.pl.lb);
        ptv_balance_chunks(in);
        const int nb = ptv_bands_plan(ctx, in, r.pl, K, L, p->zchunk > 0);
        if (nb > 0) {
            NS3D_TRY(ptv_run_banded(ctx, r, in, K, L, nb, &f));
            done += L * K;
            launch += L;
            NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));   // the protocol of the loop below starts from here
        }
    }
    while (done < n) {
        const int K = std::min(r.pl.K, n - done);
        PtV k;
        NS3D_TRY(make_ptv(ctx, p, r.pl, K, &k));
        ptv_bind(ctx, k, r.cur);
        // serpentine sweep: every other launch walks the z-chunks downwards, so a launch starts on the planes the
        // previous one touched last and finds them in the 126 MB L2; on while a good part of the working set fits
        const bool serp = ctx->opt_serpentine < 0 ? (4.0 * 8.0 * p->nx * p->ny * p->nz < 6.0 * ctx->l2_bytes) : ctx->opt_serpentine != 0;
        k.reverse = serp && ((iter0 / r.pl.K + launch) & 1);
        if (r.peer_on) {
            ptv_set_peers(k, r.peers, 1 - r.cur, 2 + r.cur);
            const int zf = r.pl.zf;
            const bool split = ctx->opt_p2p_split && (p->nz - 2) >= 2 * zf + 4;
            if (split) {
                PtV f = k, in = k;
                f.faces = 1;
                f.zchunk = zf;
                in.kbeg = 1 + zf;
                in.kend = p->nz - 1 - zf;
                in.mbox = nullptr;
                if (p->zchunk <= 0) in.zchunk = auto_zchunk(ctx, in, K, in.kend - in.kbeg, r.pl.lb);
                NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_a, 0));
                NS3D_TRY(ptv_launch(ctx, ctx->comm_stream, f, r.pl, K, true));
                NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->comm_stream));
                NS3D_TRY(ptv_launch(ctx, ctx->stream, in, r.pl, K, false));
                NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
                NS3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0));
            } else {
                NS3D_TRY(ptv_launch(ctx, ctx->stream, k, r.pl, K, true));
            }
        } else {
            NS3D_TRY(ptv_launch(ctx, ctx->stream, k, r.pl, K, false));
            if (ctx->nranks > 1) {  // no peer memory: update_halo!(Pr) (M:462,182) over NCCL after every iteration (K = 1 here)
                double* f[1] = {ctx->ptv[1 - r.cur]};
                const int sx = ptv_pitch(p->nx);
                NS3D_TRY(ns3d_internal_halo_z(ctx, ctx->stream, f, &sx, &p->ny, &p->nz, 1, p->nz));
            }
        }
        r.cur = 1 - r.cur;
        done += K;
        ++launch;
    }
    if (r.peer_on) {  // the halos of the current iterate are complete once both neighbours have caught up
        PtV k;
        NS3D_TRY(make_ptv(ctx, p, r.pl, 1, &k));
        pt_halo_wait_kernel<<<1, 32, 0, ctx->stream>>>(ctx->mbox, k.zlo_halo, k.zhi_halo);
        NS3D_LAUNCH_CHECK(ctx);
    }
    return NS3D_OK;
}

// ---- CUDA-graph replay of chunks of iterations -------------------------------------------------------------------
// Per launch the host would otherwise issue 1 launch (single rank) or 2 launches, 2 event records and 2 stream
// waits (slabs): 5-30 us of CPU time against a 40-60 us kernel pair at 255x153x153, which with 8 ranks on one
// host made the loop launch-bound.  A chunk of nchk iterations, both streams, is captured once and replayed.
struct PtvGraph {
    cudaGraphExec_t exec = nullptr;
    ns3d_pt_params p;
    int n = 0, iter0 = 0, cur = 0, mode = 0, opts[8] = {};
    bool peer_on = false;
    long long kernels = 0;
    int cur_after = 0;
};
struct PtvGraphCache {
    PtvGraph slot[6];
    int next = 0;
};

void opts_key(const ns3d_ctx* ctx, int (&o)[8])
{
    o[0] = ctx->opt_ptv_k; o[1] = ctx->opt_ptv_ns; o[2] = ctx->opt_ptv_lb; o[3] = ctx->opt_ptv_pxt;
    o[4] = ctx->opt_ptv_bty; o[5] = ctx->opt_serpentine; o[6] = ctx->opt_p2p;
    o[7] = ctx->opt_ptv_tma | (ctx->opt_ptv_flow << 1) | (ctx->opt_p2p_split << 2) | ((ctx->opt_ptv_bands + 1) << 3);
}

int ptv_run(ns3d_ctx* ctx, PtvRun& r, int n, int iter0)
{
    const bool graphable = ctx->opt_graphs && n >= 8 && (ctx->nranks == 1 || r.peer_on);
    if (!graphable) return ptv_run_direct(ctx, r, n, iter0);
    if (!ctx->ptv_graphs) ctx->ptv_graphs = new PtvGraphCache();
    PtvGraphCache* cache = (PtvGraphCache*)ctx->ptv_graphs;
    int ok[8];
    opts_key(ctx, ok);
    const int phase = (iter0 / r.pl.K) & 1;  // the serpentine direction of the first launch
    PtvGraph* g = nullptr;
    for (PtvGraph& c : cache->slot)
        if (c.exec && c.n == n && c.iter0 == phase && c.cur == r.cur && c.mode == ctx->mode && c.peer_on == r.peer_on &&
            !memcmp(c.opts, ok, sizeof ok) && !memcmp(&c.p, r.p, sizeof c.p))
            g = &c;
    if (!g) {
        g = &cache->slot[cache->next];
        cache->next = (cache->next + 1) % 6;
        if (g->exec) {
            cudaGraphExecDestroy(g->exec);
            g->exec = nullptr;
        }
        PtvRun rr = r;
        const long long l0 = ctx->launches;
        NS3D_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = ptv_run_direct(ctx, rr, n, iter0);
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
        g->p = *r.p; g->n = n; g->iter0 = phase; g->cur = r.cur; g->mode = ctx->mode; g->peer_on = r.peer_on;
        memcpy(g->opts, ok, sizeof ok);
        g->kernels = captured;
        g->cur_after = rr.cur;
    }
    NS3D_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->kernels;
    r.cur = g->cur_after;
    return NS3D_OK;
}

// A chunk of iterations as several graphs launched back to back (option graph_pieces; off by default).  The chunk between
// two residual checks is one graph of hundreds of kernel nodes (255x153x153 with z-bands: 76 passes x 8 launches), launched
// after a host synchronisation; in pieces only the first piece's launch would be exposed.  Measured
// (profiles/r02_graph_pieces_bench.log): 8 428 GB/s in one piece, 8 334 in two, 8 197 in four, 8 010 in eight -- the launch
// is not what the step waits for, and every piece boundary drains the bands.  Pieces hold an even number of passes (the
// ping-pong buffers end where they started: one cached graph serves them all).
int ptv_run_pieces(ns3d_ctx* ctx, PtvRun& r, int n, int iter0)
{
    const int K = r.pl.K;
    int pieces = ctx->opt_graph_pieces;
    if (pieces < 0) pieces = 1;
    if (!ctx->opt_graphs || pieces <= 1 || n / K < 8 * pieces) return ptv_run(ctx, r, n, iter0);
    int per = (n / K + pieces - 1) / pieces;   // passes per piece ...
    per += per & 1;                            // ... an even number
    int done = 0;
    while (done < n) {
        const int m = std::min(per * K, n - done);
        NS3D_TRY(ptv_run(ctx, r, m, iter0 + done));
        done += m;
    }
    return NS3D_OK;
}

int ptv_residual(ns3d_ctx* ctx, const PtvRun& r)
{
    PtV k;
    NS3D_TRY(make_ptv(ctx, r.p, r.pl, 1, &k));
    ptv_bind(ctx, k, r.cur);
    k.zchunk = 16;
    NS3D_CUDA(ctx, cudaMemsetAsync(ctx->d_maxbits, 0, sizeof(unsigned long long), ctx->stream));
    const dim3 grid(cdiv(k.nx - 2, 32), cdiv(k.ny - 2, 8), cdiv(k.kend - k.kbeg, k.zchunk));
    switch (ctx->mode) {
        case NS3D_PARITY: ptv_residual_kernel<NS3D_PARITY><<<grid, dim3(32, 8, 1), 0, ctx->stream>>>(k, ctx->d_maxbits); break;
        case NS3D_FAST: ptv_residual_kernel<NS3D_FAST><<<grid, dim3(32, 8, 1), 0, ctx->stream>>>(k, ctx->d_maxbits); break;
        default: ptv_residual_kernel<NS3D_FASTEST><<<grid, dim3(32, 8, 1), 0, ctx->stream>>>(k, ctx->d_maxbits); break;
    }
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// Pack the caller's arrays, agree with the neighbours that the previous solve is over.
int ptv_begin(ns3d_ctx* ctx, PtvRun& r, const double* Pr, const double* dPrdtau, const double* divV, int max_chunk_iters)
{
    const ns3d_pt_params* p = r.p;
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_TRY(ptv_ensure(ctx, p->nx, p->ny, p->nz));
    if (ctx->nranks == 1 && ctx->opt_ptv_flow) NS3D_TRY(ptv_work_ensure(ctx, max_chunk_iters, p->nz));
    r.pl = make_plan(ctx, ctx->nranks > 1);
    NS3D_TRY(ptv_peer_prepare(ctx, p->nz, &r.peers, &r.peer_on));
    if (ctx->nranks > 1 && !r.peer_on) r.pl.K = 1;  // NCCL halo exchange after every iteration
    r.cur = 0;
    NS3D_TRY(ptv_pack(ctx, ctx->ptv[0], Pr, p->nx, p->ny, p->nz, 0));
    NS3D_TRY(ptv_pack(ctx, ctx->ptv[2], dPrdtau, p->nx, p->ny, p->nz, 1));
    NS3D_TRY(ptv_pack(ctx, ctx->ptv[4], divV, p->nx, p->ny, p->nz, 0));
    if (r.peer_on) {
        pt_halo_barrier_kernel<<<1, 32, 0, ctx->stream>>>(ctx->mbox, r.peers.lo_mbox ? r.peers.lo_mbox + NS3D_MB_FLAG_HI : nullptr,
                                                          r.peers.hi_mbox ? r.peers.hi_mbox + NS3D_MB_FLAG_LO : nullptr);
        NS3D_LAUNCH_CHECK(ctx);
    }
    return NS3D_OK;
}

int ptv_end(ns3d_ctx* ctx, PtvRun& r, double* Pr, double* dPrdtau, bool sync)
{
    const ns3d_pt_params* p = r.p;
    NS3D_TRY(ptv_unpack(ctx, Pr, ctx->ptv[r.cur], p->nx, p->ny, p->nz, 0));
    NS3D_TRY(ptv_unpack(ctx, dPrdtau, ctx->ptv[2 + r.cur], p->nx, p->ny, p->nz, 1));
    if (r.peer_on)
        NS3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_maxbits + 3, ctx->mbox + NS3D_MB_ERROR, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->ptv_work)
        NS3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_maxbits + 5, ctx->ptv_work + ctx->ptv_work_words - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (sync || r.peer_on) {
        NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->ptv_work && (ctx->h_maxbits[5] & 0xffffffffULL) != 0ULL) {
            cudaMemsetAsync(ctx->ptv_work + ctx->ptv_work_words - 1, 0, 4, ctx->stream);
            return ns3d_fail(ctx, NS3D_ECUDA, "persistent PT launch: a work item waited for its predecessors beyond the spin limit");
        }
        if (r.peer_on && ctx->h_maxbits[3] != 0ULL) {
            const unsigned long long who = ctx->h_maxbits[3];
            cudaMemsetAsync(ctx->mbox + NS3D_MB_ERROR, 0, 8, ctx->stream);  // reported: the next solve starts clean
            return ns3d_fail(ctx, NS3D_ECOMM, "peer-memory halo exchange: neighbour %s did not answer within the spin limit",
                             who == 1ULL ? "below" : "above");
        }
    }
    return NS3D_OK;
}

}  // namespace

void ns3d_internal_ptv_free(ns3d_ctx* ctx)
{
    PtvGraphCache* cache = (PtvGraphCache*)ctx->ptv_graphs;
    if (cache) {
        for (PtvGraph& c : cache->slot)
            if (c.exec) cudaGraphExecDestroy(c.exec);
        delete cache;
        ctx->ptv_graphs = nullptr;
    }
}

void ns3d_internal_ptv_release(ns3d_ctx* ctx)
{
    ns3d_internal_ptv_free(ctx);
    for (int b = 0; b < 16; ++b) {
        if (ctx->band_stream[b]) cudaStreamDestroy(ctx->band_stream[b]);
        ctx->band_stream[b] = nullptr;
        for (int q = 0; q < 2; ++q) {
            if (ctx->band_ev[q][b]) cudaEventDestroy(ctx->band_ev[q][b]);
            ctx->band_ev[q][b] = nullptr;
        }
    }
    if (ctx->band_fork) cudaEventDestroy(ctx->band_fork);
    ctx->band_fork = nullptr;
    if (ctx->ptv_work) {
        cudaFree(ctx->ptv_work);
        ctx->ptv_work = nullptr;
        ctx->ptv_work_words = 0;
    }
    for (int q = 0; q < 5; ++q) {
        if (!ctx->ptv_raw[q]) continue;
        ptv_unmap(ctx, ctx->ptv_raw[q]);
        cudaFree(ctx->ptv_raw[q]);
        ctx->ptv_raw[q] = nullptr;
    }
}

int ns3d_internal_ptv_solve(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV, const ns3d_pt_params* p, int* h_iters,
                            double* h_err_hist, int err_cap, int* h_nchecks)
{
    PtvRun r;
    r.p = p;
    NS3D_TRY(ptv_begin(ctx, r, Pr, dPrdtau, divV, std::min(p->nchk, p->niter)));
    int iters = 0, nc = 0;
    while (iters < p->niter) {
        const int chunk = std::min(p->nchk - iters % p->nchk, p->niter - iters);  // up to the next check
        NS3D_TRY(ptv_run_pieces(ctx, r, chunk, iters));
        iters += chunk;
        if (iters % p->nchk == 0) {
            NS3D_TRY(ptv_residual(ctx, r));
            double m = 0.0;
            NS3D_TRY(ns3d_internal_read_max(ctx, &m));
            const double err = m * p->err_num / p->err_den;  // max*ly^2/psc  M:466
            if (h_err_hist && nc < err_cap) h_err_hist[nc] = err;
            ++nc;
            if (err < p->eps_it || !std::isfinite(err)) break;  // M:469
        }
    }
    NS3D_TRY(ptv_end(ctx, r, Pr, dPrdtau, true));
    if (h_iters) *h_iters = iters;
    if (h_nchecks) *h_nchecks = nc;
    return NS3D_OK;
}

int ns3d_internal_ptv_iterate(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV, const ns3d_pt_params* p, int n_iter)
{
    PtvRun r;
    r.p = p;
    NS3D_TRY(ptv_begin(ctx, r, Pr, dPrdtau, divV, n_iter));
    NS3D_TRY(ptv_run_pieces(ctx, r, n_iter, 0));
    return ptv_end(ctx, r, Pr, dPrdtau, false);
}

int ns3d_internal_ptv_describe(ns3d_ctx* ctx, const ns3d_pt_params* p, char* buf, int cap, int* iters_per_launch)
{
    const bool slabs = ctx->nranks > 1;
    PtvPlan pl = make_plan(ctx, slabs);
    const bool peer = slabs && ctx->opt_p2p && ctx->p2p_ready && p->nz >= 6;
    if (slabs && !peer) pl.K = 1;
    const bool flow = !slabs && ctx->opt_ptv_flow && ptv_lb_threads(pl.lb) == 256;
    PtV k;
    NS3D_TRY(make_ptv(ctx, p, pl, pl.K, &k, flow));
    ptv_balance_chunks(k);
    // the band plan of a chunk of iterations (ptv_run_direct)
    int nb = 0;
    const bool split = peer && ctx->opt_p2p_split && (p->nz - 2) >= 2 * pl.zf + 4;
    if (!flow && !slabs) {
        nb = ptv_bands_plan(ctx, k, pl, pl.K, 2, p->zchunk > 0);
    } else if (split) {
        PtV in = k;
        in.kbeg = 1 + pl.zf;
        in.kend = p->nz - 1 - pl.zf;
        if (p->zchunk <= 0) in.zchunk = auto_zchunk(ctx, in, pl.K, in.kend - in.kbeg, pl.lb);
        ptv_balance_chunks(in);
        nb = ptv_bands_plan(ctx, in, pl, pl.K, 2, p->zchunk > 0);
        k.zchunk = in.zchunk;
    }
    const char* mode = ctx->mode == NS3D_PARITY ? "PARITY" : (ctx->mode == NS3D_FAST ? "FAST" : "FASTEST");
    char bands[160] = "";
    if (nb > 0)
        snprintf(bands, sizeof bands, "; a pass = %d launches (z-bands on %d streams, band b waits for bands b-1, b, b+1 of the previous pass)",
                 nb + (split ? 1 : 0), nb + (split ? 1 : 0));
    if (buf && cap > 0)
        snprintf(buf, cap,
                 "%s<%s,K=%d> (%s%d fused PT iterations per pass over the fields: %d x (K5+K6+set_bc_Pr!) on the pitched copies, z-plane "
                 "tiles staged by TMA; tiles of %d x %d cells, %d x %d tiles, %d-plane chunks, %d threads, %u B shared memory%s%s)",
                 flow ? "ptv_flow_kernel" : "ptv_kernel", mode, pl.K,
                 flow ? "persistent: all passes between two residual checks are ONE launch, work items ordered by pass and z-chunk with "
                        "per-chunk completion counters; "
                      : "",
                 pl.K, pl.K, 2 * k.pxt, k.bty, k.ntx, k.nty, k.zchunk, ptv_threads(k), k.sm_total, bands,
                 slabs ? "; slab-interface chunks: the P2P instantiation with update_halo!(Pr) over peer memory" : "");
    if (iters_per_launch) *iters_per_launch = pl.K;
    return NS3D_OK;
}
