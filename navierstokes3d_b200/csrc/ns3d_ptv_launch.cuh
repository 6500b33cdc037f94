// Instantiations of ptv_kernel for ONE arithmetic mode and their dispatch (included by ns3d_ptv_mode*.cu).
#pragma once

#include "ns3d_internal.cuh"
#include "ns3d_ptv_kernels.cuh"

namespace {

template <int MODE, int K, bool P2P, bool TMA, int PXT, int BTY, int NT, int MINB>
int ptv_launch_t(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, dim3 grid, size_t smem)
{
    auto ptv_kernel_fn = ptv_kernel<MODE, K, P2P, TMA, PXT, BTY, NT, MINB>;
    static size_t s_smem_set[64] = {};  // per instantiation and device: dynamic shared memory the function may use
    if (smem > 48 * 1024 && smem > s_smem_set[ctx->device & 63]) {
        NS3D_CUDA(ctx, cudaFuncSetAttribute(ptv_kernel_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s_smem_set[ctx->device & 63] = smem;
    }
    ptv_kernel_fn<<<grid, dim3(ptv_threads(k), 1, 1), smem, st>>>(k, maps);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// launch bounds: lb 0 = 256 threads / 2 CTAs per SM, 1 = 256 / 3, 3 = 512 / 1, 4 = 512 / 2
template <int MODE, int K, bool P2P, bool TMA, int PXT, int BTY>
int ptv_launch_lb(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, int lb, dim3 grid, size_t smem)
{
    if constexpr (PXT * BTY == 512) {
        if (lb == 4) return ptv_launch_t<MODE, K, P2P, TMA, PXT, BTY, 512, 2>(ctx, st, k, maps, grid, smem);
        return ptv_launch_t<MODE, K, P2P, TMA, PXT, BTY, 512, 1>(ctx, st, k, maps, grid, smem);
    } else if constexpr (PXT * BTY == 256) {
        if (lb == 0) return ptv_launch_t<MODE, K, P2P, TMA, PXT, BTY, 256, 2>(ctx, st, k, maps, grid, smem);
        return ptv_launch_t<MODE, K, P2P, TMA, PXT, BTY, 256, 3>(ctx, st, k, maps, grid, smem);
    } else {   // geometry from the kernel parameters
        if (lb == 4) return ptv_launch_t<MODE, K, P2P, TMA, PXT, BTY, 512, 2>(ctx, st, k, maps, grid, smem);
        if (lb == 3) return ptv_launch_t<MODE, K, P2P, TMA, PXT, BTY, 512, 1>(ctx, st, k, maps, grid, smem);
        if (lb == 0) return ptv_launch_t<MODE, K, P2P, TMA, PXT, BTY, 256, 2>(ctx, st, k, maps, grid, smem);
        return ptv_launch_t<MODE, K, P2P, TMA, PXT, BTY, 256, 3>(ctx, st, k, maps, grid, smem);
    }
}

// Tile shapes with a compile-time instantiation (every shared-memory stride an immediate); any other shape runs the
// instantiation that takes the geometry from the kernel parameters.
template <int MODE, int K>
int ptv_launch_shape(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, int lb, dim3 grid, size_t smem)
{
    if (k.ns == 4) {
        if (k.pxt == 16 && k.bty == 16) return ptv_launch_lb<MODE, K, false, true, 16, 16>(ctx, st, k, maps, lb, grid, smem);
        if (k.pxt == 32 && k.bty == 8) return ptv_launch_lb<MODE, K, false, true, 32, 8>(ctx, st, k, maps, lb, grid, smem);
    }
    return ptv_launch_lb<MODE, K, false, true, 0, 0>(ctx, st, k, maps, lb, grid, smem);
}

// ... on a slab interface (P2P): the default tile or the general geometry
template <int MODE, int K>
int ptv_launch_shape_p2p(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, int lb, dim3 grid, size_t smem)
{
    if (k.ns == 4 && k.pxt == 16 && k.bty == 16) return ptv_launch_lb<MODE, K, true, true, 16, 16>(ctx, st, k, maps, lb, grid, smem);
    return ptv_launch_lb<MODE, K, true, true, 0, 0>(ctx, st, k, maps, lb, grid, smem);
}

// ---- the persistent launch (ptv_flow_kernel): as many CTAs as are resident at a time, or one per work item if fewer ----
template <int MODE, int K, bool TMA, int PXT, int BTY, int NT, int MINB>
int ptv_flow_launch_t(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, size_t smem)
{
    auto ptv_flow_kernel_fn = ptv_flow_kernel<MODE, K, TMA, PXT, BTY, NT, MINB>;
    static size_t s_smem_set[64] = {};
    static int s_resident[64] = {};
    const int dev = ctx->device & 63;
    if (smem > 48 * 1024 && smem > s_smem_set[dev]) {
        NS3D_CUDA(ctx, cudaFuncSetAttribute(ptv_flow_kernel_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s_smem_set[dev] = smem;
        s_resident[dev] = 0;
    }
    int per_sm = MINB;
#ifndef NS3D_HOST_EMU
    if (s_resident[dev] == 0) {
        NS3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s_resident[dev], ptv_flow_kernel_fn, ptv_threads(k), smem));
        if (s_resident[dev] < 1) s_resident[dev] = 1;
    }
    per_sm = s_resident[dev];
#endif
    const long long items = (long long)k.ntx * k.nty * k.nbz * k.nlaunch;
    const long long resident = (long long)ctx->num_sms * per_sm;
    const dim3 grid((unsigned)std::min(items, resident), 1, 1);
    ptv_flow_kernel_fn<<<grid, dim3(ptv_threads(k), 1, 1), smem, st>>>(k, maps);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

template <int MODE, int K, bool TMA, int PXT, int BTY>
int ptv_flow_launch_lb(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, int lb, size_t smem)
{
    if (lb == 0) return ptv_flow_launch_t<MODE, K, TMA, PXT, BTY, 256, 2>(ctx, st, k, maps, smem);
    return ptv_flow_launch_t<MODE, K, TMA, PXT, BTY, 256, 3>(ctx, st, k, maps, smem);
}

template <int MODE, int K>
int ptv_flow_launch_k(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, int lb, bool tma, size_t smem)
{
#ifndef NS3D_HOST_EMU
    if (tma) {
        if (k.ns == 4 && k.pxt == 16 && k.bty == 16) return ptv_flow_launch_lb<MODE, K, true, 16, 16>(ctx, st, k, maps, lb, smem);
        return ptv_flow_launch_lb<MODE, K, true, 0, 0>(ctx, st, k, maps, lb, smem);
    }
#endif
    (void)tma;
    return ptv_flow_launch_lb<MODE, K, false, 0, 0>(ctx, st, k, maps, lb, smem);
}

// lb: 0 or 1 (256-thread CTAs)
template <int MODE>
int ptv_flow_launch_m(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool tma, size_t smem)
{
    if (K == 1) return ptv_flow_launch_k<MODE, 1>(ctx, st, k, maps, pl.lb, tma, smem);
    if (K == 2) return ptv_flow_launch_k<MODE, 2>(ctx, st, k, maps, pl.lb, tma, smem);
    return ptv_flow_launch_k<MODE, 3>(ctx, st, k, maps, pl.lb, tma, smem);
}

// tma: the staging ring is filled by the TMA unit (launches on a device); otherwise by plain loads of all threads (the
// host emulation of the library, option ptv_tma = 0)
template <int MODE>
int ptv_launch_m(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool p2p, bool tma,
                 dim3 grid, size_t smem)
{
#ifndef NS3D_HOST_EMU
    if (tma && p2p) {
        if (K == 1) return ptv_launch_shape_p2p<MODE, 1>(ctx, st, k, maps, pl.lb, grid, smem);
        return ptv_launch_shape_p2p<MODE, 2>(ctx, st, k, maps, pl.lb, grid, smem);
    }
    if (tma) {
        if (K == 1) return ptv_launch_shape<MODE, 1>(ctx, st, k, maps, pl.lb, grid, smem);
        if (K == 2) return ptv_launch_shape<MODE, 2>(ctx, st, k, maps, pl.lb, grid, smem);
        return ptv_launch_shape<MODE, 3>(ctx, st, k, maps, pl.lb, grid, smem);
    }
#endif
    if (p2p) {
        if (K == 1) return ptv_launch_lb<MODE, 1, true, false, 0, 0>(ctx, st, k, maps, pl.lb, grid, smem);
        return ptv_launch_lb<MODE, 2, true, false, 0, 0>(ctx, st, k, maps, pl.lb, grid, smem);
    }
    if (K == 1) return ptv_launch_lb<MODE, 1, false, false, 0, 0>(ctx, st, k, maps, pl.lb, grid, smem);
    if (K == 2) return ptv_launch_lb<MODE, 2, false, false, 0, 0>(ctx, st, k, maps, pl.lb, grid, smem);
    return ptv_launch_lb<MODE, 3, false, false, 0, 0>(ctx, st, k, maps, pl.lb, grid, smem);
}

}  // namespace
