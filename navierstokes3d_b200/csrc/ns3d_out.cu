// libns3d.so -- output path: what the reference does with
// `Array(A)[2:end-1,2:end-1,2:end-1]` before gather!/save_array (M:399-412, 481-523, 528-532) and with
// the heat-map planes `A_v[:,:,k]` / `A_v[:,j,:]` (M:422-431).  The reference copies the WHOLE
// field to the host and slices there; here a kernel packs just the requested box in device memory
// (optionally converting to Float32 like `convert.(Float32, A_v)`, M:408) and only the box
// crosses PCIe: a mid-plane of a 511^3 field is 2 MB instead of 1 GB.
#include <vector>

#include "ns3d_internal.cuh"

namespace {

// out[(i-x0) + bx*((j-y0) + by*(k-z0))] = A[i,j,k] over the box; x on threadIdx.x (coalesced rows).
template <class T>
__global__ void __launch_bounds__(256) pack_box_kernel(const double* __restrict__ A, int sx, int sy, int x0, int y0,
                                                       int z0, int bx, int by, int bz, T* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= bx || j >= by) return;
    for (int k = blockIdx.z; k < bz; k += gridDim.z)
        out[(size_t)i + (size_t)bx * ((size_t)j + (size_t)by * (size_t)k)] =
            static_cast<T>(A[idx3(x0 + i, y0 + j, z0 + k, sx, sy)]);  // double -> float: round to nearest even, as Julia
}

int ensure_stage(ns3d_ctx* ctx, size_t bytes)
{
    if (ctx->out_stage_bytes >= bytes) return NS3D_OK;
    if (ctx->out_stage) {
        NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        NS3D_CUDA(ctx, cudaFree(ctx->out_stage));
        ctx->out_stage = nullptr;
        ctx->out_stage_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->out_stage, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ns3d_fail(ctx, NS3D_ENOMEM, "ns3d_box_d2h: cannot allocate %zu B of staging memory", bytes);
    }
    ctx->out_stage_bytes = bytes;
    return NS3D_OK;
}

}  // namespace

void ns3d_internal_out_free(ns3d_ctx* ctx)
{
    if (ctx->out_stage) cudaFree(ctx->out_stage);
    if (ctx->gather_stage) cudaFree(ctx->gather_stage);
    ctx->out_stage = ctx->gather_stage = nullptr;
    ctx->out_stage_bytes = ctx->gather_stage_bytes = 0;
}

namespace {

// Validates the box and packs it into ctx->out_stage (Float64 or Float32); *bytes = its size.
int pack_box(ns3d_ctx* ctx, const char* who, const double* A, int sx, int sy, int sz, int x0, int x1, int y0, int y1,
             int z0, int z1, int f32, size_t* bytes)
{
    if (!A || sx <= 0 || sy <= 0 || sz <= 0) return ns3d_fail(ctx, NS3D_EINVAL, "%s: bad array (%d,%d,%d)", who, sx, sy, sz);
    if (x0 < 0 || y0 < 0 || z0 < 0 || x1 > sx || y1 > sy || z1 > sz || x1 < x0 || y1 < y0 || z1 < z0)
        return ns3d_fail(ctx, NS3D_EINVAL, "%s: box [%d,%d)x[%d,%d)x[%d,%d) outside (%d,%d,%d)", who, x0, x1, y0, y1, z0, z1,
                         sx, sy, sz);
    const int bx = x1 - x0, by = y1 - y0, bz = z1 - z0;
    const size_t count = (size_t)bx * by * bz;
    *bytes = count * (f32 ? sizeof(float) : sizeof(double));
    if (count == 0) return NS3D_OK;  // empty box (e.g. the interior of a 2-point-wide array)
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_TRY(ensure_stage(ctx, *bytes));
    const dim3 blk(32, 8, 1);
    const dim3 grd(cdiv(bx, 32), cdiv(by, 8), (unsigned)std::min(bz, 65535));
    if (f32) pack_box_kernel<float><<<grd, blk, 0, ctx->stream>>>(A, sx, sy, x0, y0, z0, bx, by, bz, (float*)ctx->out_stage);
    else pack_box_kernel<double><<<grd, blk, 0, ctx->stream>>>(A, sx, sy, x0, y0, z0, bx, by, bz, (double*)ctx->out_stage);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

}  // namespace

extern "C" int ns3d_box_d2h(ns3d_ctx* ctx, const double* A, int sx, int sy, int sz, int x0, int x1, int y0, int y1,
                            int z0, int z1, void* h_out, int f32)
{
    NS3D_CHECK_CTX(ctx);
    if (!h_out) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_box_d2h: NULL output");
    size_t bytes = 0;
    NS3D_TRY(pack_box(ctx, "ns3d_box_d2h", A, sx, sy, sz, x0, x1, y0, y1, z0, z1, f32, &bytes));
    if (bytes == 0) return NS3D_OK;
    NS3D_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->out_stage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NS3D_OK;
}

// gather!(A_inn, A_v) (M:399-403, 481-485, 528-532) for z-slabs: every rank packs its box on the device,
// the boxes travel to rank 0 over NCCL and are concatenated along z there (the x-y extent of the
// box is the same on every rank; nplanes_all[r] = z1 - z0 of rank r), then one device-to-host copy.
extern "C" int ns3d_gather_box(ns3d_ctx* ctx, const double* A, int sx, int sy, int sz, int x0, int x1, int y0, int y1,
                               int z0, int z1, const int* nplanes_all, void* h_out, int f32)
{
    NS3D_CHECK_CTX(ctx);
    if (ctx->nranks == 1) return ns3d_box_d2h(ctx, A, sx, sy, sz, x0, x1, y0, y1, z0, z1, h_out, f32);
    if (!nplanes_all) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_gather_box: nplanes_all is NULL");
    if (nplanes_all[ctx->rank] != z1 - z0)
        return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_gather_box: rank %d passes %d planes, nplanes_all says %d", ctx->rank, z1 - z0,
                         nplanes_all[ctx->rank]);
    if (ctx->rank == 0 && !h_out) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_gather_box: NULL output on rank 0");
    size_t bytes = 0;
    NS3D_TRY(pack_box(ctx, "ns3d_gather_box", A, sx, sy, sz, x0, x1, y0, y1, z0, z1, f32, &bytes));
    const size_t plane_bytes = (size_t)(x1 - x0) * (y1 - y0) * (f32 ? sizeof(float) : sizeof(double));
    std::vector<size_t> bytes_all(ctx->nranks, 0);
    size_t total = 0;
    for (int r = 0; r < ctx->nranks; ++r) {
        if (nplanes_all[r] < 0) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_gather_box: negative plane count for rank %d", r);
        bytes_all[r] = plane_bytes * (size_t)nplanes_all[r];
        total += bytes_all[r];
    }
    if (ctx->rank == 0 && ctx->gather_stage_bytes < total) {
        if (ctx->gather_stage) {
            NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            NS3D_CUDA(ctx, cudaFree(ctx->gather_stage));
            ctx->gather_stage = nullptr;
            ctx->gather_stage_bytes = 0;
        }
        if (cudaMalloc(&ctx->gather_stage, total) != cudaSuccess) {
            cudaGetLastError();
            return ns3d_fail(ctx, NS3D_ENOMEM, "ns3d_gather_box: cannot allocate %zu B on rank 0", total);
        }
        ctx->gather_stage_bytes = total;
    }
    NS3D_TRY(ns3d_internal_gather_bytes(ctx, ctx->out_stage, bytes, ctx->gather_stage, bytes_all.data()));
    if (ctx->rank == 0 && total)
        NS3D_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->gather_stage, total, cudaMemcpyDeviceToHost, ctx->stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NS3D_OK;
}
