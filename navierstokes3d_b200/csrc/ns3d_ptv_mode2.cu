// libns3d.so -- ptv_kernel instantiated for NS3D_FASTEST arithmetic (see ns3d_ptv_launch.cuh, ns3d_ptv.cu).
#include "ns3d_ptv_launch.cuh"

int ns3d_internal_ptv_launch_fastest(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool p2p,
                                   bool tma, dim3 grid, size_t smem)
{
    return ptv_launch_m<NS3D_FASTEST>(ctx, st, k, maps, pl, K, p2p, tma, grid, smem);
}

int ns3d_internal_ptv_flow_launch_fastest(ns3d_ctx* ctx, cudaStream_t st, const PtV& k, const PtvMaps& maps, const PtvPlan& pl, int K, bool tma,
                                         size_t smem)
{
    return ptv_flow_launch_m<NS3D_FASTEST>(ctx, st, k, maps, pl, K, tma, smem);
}
