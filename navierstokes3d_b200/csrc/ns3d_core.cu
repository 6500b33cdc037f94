// libns3d.so -- lifecycle, device field allocator, reductions, z-slab communicator.
//
// Replaces: @init_parallel_stencil / IGG device selection, the
// ParallelStencil allocator `@zeros` (M:343-360), `Data.Array(host)` (M:370), `Array(dev)`
// (M:399), ImplicitGlobalGrid's update_halo! and MPI.Allreduce(MAX) in max_g (M:21).
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>

#include "ns3d_internal.cuh"

static thread_local std::string g_create_error;

int ns3d_fail(ns3d_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_error = buf;
    return code;
}

extern "C" const char* ns3d_version(void) { return "ns3d-b200 0.1.0 (sm_100a)"; }

extern "C" const char* ns3d_last_error(const ns3d_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int ns3d_create(int device, ns3d_ctx** out)
{
    if (!out) return ns3d_fail(nullptr, NS3D_EINVAL, "ns3d_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return ns3d_fail(nullptr, NS3D_ENODEV, "ns3d_create: no CUDA device (%s); there is no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev)
        return ns3d_fail(nullptr, NS3D_EINVAL, "ns3d_create: device %d out of range [0,%d)", device, ndev);
    ns3d_ctx* ctx = new ns3d_ctx();
    ctx->device = device;
#define CREATE_CUDA(call)                                                                          \
    do {                                                                                           \
        cudaError_t e2 = (call);                                                                   \
        if (e2 != cudaSuccess) {                                                                   \
            int rc = ns3d_fail(nullptr, NS3D_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e2)); \
            delete ctx;                                                                            \
            return rc;                                                                             \
        }                                                                                          \
    } while (0)
    CREATE_CUDA(cudaSetDevice(device));
    CREATE_CUDA(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device));
    int l2 = 0;
    CREATE_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, device));
    ctx->l2_bytes = (size_t)l2;
    int lo = 0, hi = 0;
    CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CREATE_CUDA(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, lo));
    CREATE_CUDA(cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
    CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_xfer, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming));
    CREATE_CUDA(cudaMalloc(&ctx->d_maxbits, 64 * sizeof(unsigned long long)));
    CREATE_CUDA(cudaMemset(ctx->d_maxbits, 0, 64 * sizeof(unsigned long long)));
    CREATE_CUDA(cudaMallocHost(&ctx->h_maxbits, 64 * sizeof(unsigned long long)));
#undef CREATE_CUDA
    *out = ctx;
    return NS3D_OK;
}

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

// NCCL is bound lazily so that single-GPU users need no libnccl at all.  If the host
// process already loaded one (torch ships libnccl.so.2) the loader hands us that copy.
static int load_nccl(ns3d_ctx* ctx)
{
    if (g_nccl.handle) return NS3D_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return ns3d_fail(ctx, NS3D_ECOMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                              \
    do {                                                                              \
        *(void**)(&g_nccl.field) = dlsym(h, name);                                    \
        if (!g_nccl.field) return ns3d_fail(ctx, NS3D_ECOMM, "libnccl lacks %s", name); \
    } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.handle = h;
    return NS3D_OK;
}

#define NS3D_NCCL(ctx, call)                                                                    \
    do {                                                                                        \
        ncclResult_t r__ = (call);                                                              \
        if (r__ != ncclSuccess)                                                                 \
            return ns3d_fail((ctx), NS3D_ECOMM, "%s failed: %s", #call, g_nccl.GetErrorString(r__)); \
    } while (0)

extern "C" int ns3d_destroy(ns3d_ctx* ctx)
{
    NS3D_CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->comm_stream);
    ns3d_internal_ptv_release(ctx);
    ns3d_internal_out_free(ctx);
    for (auto& kv : ctx->p2p_map) {
        if (kv.second.first) cudaIpcCloseMemHandle(kv.second.first);
        if (kv.second.second) cudaIpcCloseMemHandle(kv.second.second);
    }
    for (int q = 0; q < 2; ++q)
        if (ctx->peer_mbox[q]) cudaIpcCloseMemHandle(ctx->peer_mbox[q]);
    if (ctx->mbox) cudaFree(ctx->mbox);
    if (ctx->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl);
    for (auto& kv : ctx->allocs) cudaFree(kv.first);
    cudaFree(ctx->d_maxbits);
    cudaFreeHost(ctx->h_maxbits);
    cudaEventDestroy(ctx->ev_a);
    cudaEventDestroy(ctx->ev_b);
    if (ctx->ev_xfer) cudaEventDestroy(ctx->ev_xfer);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->comm_stream);
    delete ctx;
    return NS3D_OK;
}

extern "C" int ns3d_set_mode(ns3d_ctx* ctx, int mode)
{
    NS3D_CHECK_CTX(ctx);
    if (mode != NS3D_PARITY && mode != NS3D_FAST && mode != NS3D_FASTEST)
        return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_set_mode: unknown mode %d", mode);
    ctx->mode = mode;
    return NS3D_OK;
}
extern "C" int ns3d_set_option(ns3d_ctx* ctx, const char* name, int value)
{
    NS3D_CHECK_CTX(ctx);
    if (!name) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_set_option: NULL name");
    // the peer-memory halo path of the fused loop on z-slabs (0: NCCL send/recv after every iteration)
    if (!strcmp(name, "p2p_halo")) { ctx->opt_p2p = value != 0; return NS3D_OK; }
    // ptv_kernel: PT iterations per launch (0 = default 2; z-slabs: at most 2)
    if (!strcmp(name, "ptv_k")) {
        if (value < 0 || value > 3) return ns3d_fail(ctx, NS3D_EINVAL, "ptv_k must be 0 (default) .. 3");
        ctx->opt_ptv_k = value;
        return NS3D_OK;
    }
    if (!strcmp(name, "ptv_ns")) {
        if (value != 0 && (value < 3 || value > 8)) return ns3d_fail(ctx, NS3D_EINVAL, "ptv_ns must be 0 (default) or 3 .. 8");
        ctx->opt_ptv_ns = value;
        return NS3D_OK;
    }
    if (!strcmp(name, "ptv_tma")) { ctx->opt_ptv_tma = value != 0; return NS3D_OK; }
    // persistent launch: every chunk of iterations between two residual checks is ONE kernel (single rank)
    if (!strcmp(name, "ptv_flow")) { ctx->opt_ptv_flow = value != 0; return NS3D_OK; }
    if (!strcmp(name, "p2p_split")) { ctx->opt_p2p_split = value != 0; return NS3D_OK; }
    if (!strcmp(name, "ptv_bands")) {
        if (value < -1 || value > 16) return ns3d_fail(ctx, NS3D_EINVAL, "ptv_bands must be -1 (default) .. 16");
        ctx->opt_ptv_bands = value;
        return NS3D_OK;
    }
    if (!strcmp(name, "ptv_lb")) {
        if (value < -1 || value > 4 || value == 2) return ns3d_fail(ctx, NS3D_EINVAL, "ptv_lb must be -1 (default), 0, 1, 3 or 4");
        ctx->opt_ptv_lb = value;
        return NS3D_OK;
    }
    if (!strcmp(name, "ptv_pxt")) { ctx->opt_ptv_pxt = value < 0 ? 0 : value; return NS3D_OK; }
    if (!strcmp(name, "ptv_bty")) { ctx->opt_ptv_bty = value < 0 ? 0 : value; return NS3D_OK; }
    if (!strcmp(name, "graphs")) { ctx->opt_graphs = value != 0; return NS3D_OK; }
    if (!strcmp(name, "graph_pieces")) { ctx->opt_graph_pieces = value < 0 ? -1 : value; return NS3D_OK; }
    if (!strcmp(name, "serpentine")) { ctx->opt_serpentine = value < 0 ? -1 : (value != 0); return NS3D_OK; }
    return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_set_option: unknown option '%s'", name);
}

extern "C" int ns3d_get_mode(const ns3d_ctx* ctx) { return ctx ? ctx->mode : NS3D_EINVAL; }

extern "C" int ns3d_sync(ns3d_ctx* ctx)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->h2d_stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->d2h_stream));
    return NS3D_OK;
}

extern "C" long long ns3d_launch_count(const ns3d_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void* ns3d_stream(ns3d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---------------------------------------------------------------------------------------------
// allocator
// ---------------------------------------------------------------------------------------------
extern "C" int ns3d_zeros(ns3d_ctx* ctx, int sx, int sy, int sz, double** dptr)
{
    NS3D_CHECK_CTX(ctx);
    if (!dptr || sx <= 0 || sy <= 0 || sz <= 0)
        return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_zeros: bad shape (%d,%d,%d)", sx, sy, sz);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    // The block is padded by three x-y planes behind the array: the software-pipelined hot kernel
    // prefetches up to two planes ahead unconditionally (values past the end are never used).
    size_t bytes = ((size_t)sx * sy * sz + 3 * (size_t)sx * sy) * sizeof(double) + 256;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);  // cudaMalloc returns >= 256-byte aligned blocks
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ns3d_fail(ctx, NS3D_ENOMEM, "ns3d_zeros: cudaMalloc(%zu B) failed: %s", bytes, cudaGetErrorString(e));
    }
    NS3D_CUDA(ctx, cudaMemsetAsync(p, 0, bytes, ctx->stream));
    ctx->allocs[p] = bytes;
    ctx->bytes += bytes;
    *dptr = (double*)p;
    return NS3D_OK;
}

extern "C" int ns3d_free(ns3d_ctx* ctx, double* dptr)
{
    NS3D_CHECK_CTX(ctx);
    auto it = ctx->allocs.find((void*)dptr);
    if (it == ctx->allocs.end()) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_free: pointer not owned by this context");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    // nothing may keep the pointer: peer mappings of this block go first (the fused loop's graphs only hold
    // context-owned buffers)
    auto pm = ctx->p2p_map.find((const void*)dptr);
    if (pm != ctx->p2p_map.end()) {
        if (pm->second.first) cudaIpcCloseMemHandle(pm->second.first);
        if (pm->second.second) cudaIpcCloseMemHandle(pm->second.second);
        ctx->p2p_map.erase(pm);
    }
    NS3D_CUDA(ctx, cudaFree(dptr));
    ctx->bytes -= it->second;
    ctx->allocs.erase(it);
    return NS3D_OK;
}

extern "C" size_t ns3d_bytes_allocated(const ns3d_ctx* ctx) { return ctx ? ctx->bytes : 0; }

extern "C" int ns3d_h2d(ns3d_ctx* ctx, double* dptr, const double* h_src, size_t count)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_CUDA(ctx, cudaMemcpyAsync(dptr, h_src, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host buffer may be pageable / reused
    return NS3D_OK;
}

extern "C" int ns3d_d2h(ns3d_ctx* ctx, double* h_dst, const double* dptr, size_t count)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_CUDA(ctx, cudaMemcpyAsync(h_dst, dptr, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NS3D_OK;
}

// ---- asynchronous host <-> device copies on their own streams (a driver that overlaps them with time steps) ----
static cudaStream_t xfer_stream(ns3d_ctx* ctx, int which)
{
    return which == NS3D_STREAM_COMPUTE ? ctx->stream : (which == NS3D_STREAM_H2D ? ctx->h2d_stream : (which == NS3D_STREAM_D2H ? ctx->d2h_stream : nullptr));
}

extern "C" int ns3d_h2d_async(ns3d_ctx* ctx, double* dptr, const double* h_pinned_src, size_t count)
{
    NS3D_CHECK_CTX(ctx);
    if (!dptr || !h_pinned_src) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_h2d_async: NULL pointer");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_CUDA(ctx, cudaMemcpyAsync(dptr, h_pinned_src, count * sizeof(double), cudaMemcpyHostToDevice, ctx->h2d_stream));
    return NS3D_OK;
}

extern "C" int ns3d_d2h_async(ns3d_ctx* ctx, double* h_pinned_dst, const double* dptr, size_t count)
{
    NS3D_CHECK_CTX(ctx);
    if (!dptr || !h_pinned_dst) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_d2h_async: NULL pointer");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_CUDA(ctx, cudaMemcpyAsync(h_pinned_dst, dptr, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    return NS3D_OK;
}

extern "C" int ns3d_stream_wait(ns3d_ctx* ctx, int waiter, int signaller)
{
    NS3D_CHECK_CTX(ctx);
    cudaStream_t w = xfer_stream(ctx, waiter), s = xfer_stream(ctx, signaller);
    if (!w || !s || w == s) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_stream_wait: bad stream ids %d, %d", waiter, signaller);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_CUDA(ctx, cudaEventRecord(ctx->ev_xfer, s));
    NS3D_CUDA(ctx, cudaStreamWaitEvent(w, ctx->ev_xfer, 0));
    return NS3D_OK;
}

extern "C" int ns3d_stream_sync(ns3d_ctx* ctx, int which)
{
    NS3D_CHECK_CTX(ctx);
    cudaStream_t s = xfer_stream(ctx, which);
    if (!s) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_stream_sync: bad stream id %d", which);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_CUDA(ctx, cudaStreamSynchronize(s));
    return NS3D_OK;
}

extern "C" int ns3d_copy(ns3d_ctx* ctx, double* dst, const double* src, size_t count)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_CUDA(ctx, cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return NS3D_OK;
}

__global__ void fill_kernel(double* __restrict__ a, double v, size_t n)
{
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) a[q] = v;
}

extern "C" int ns3d_fill(ns3d_ctx* ctx, double* dptr, double value, size_t count)
{
    NS3D_CHECK_CTX(ctx);
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, (size_t)ctx->num_sms * 16);
    fill_kernel<<<blocks ? blocks : 1, 256, 0, ctx->stream>>>(dptr, value, count);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// ---------------------------------------------------------------------------------------------
// device-side initialisers: the scripts build their initial 3-D arrays on the host
// with comprehensions (G:86-87, M:370) and upload them; every one of them depends on z alone, so the
// host computes the nz profile values (the power law of G:86 uses `^(1/6)`, which CUDA's pow does not
// round like the host's libm -- it stays on the host) and the device broadcasts them.
// ---------------------------------------------------------------------------------------------
__global__ void fill_profile_z_kernel(double* __restrict__ a, const double* __restrict__ prof, size_t sxy, size_t n)
{
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) a[q] = prof[q / sxy];
}

// A[ix,iy,iz] = (prof[iz] + add_y[iy]) + add_z[iz]: the comprehension of M:370, `-(z_g(iz,dz,C)-dz/2)*ρ*g + 0*yc[iy] + 0*zc[iz]`,
// term by term -- with g = 0 (M:316, Fr = Inf) every term is a SIGNED zero and the sum is -0.0 exactly where all three are
__global__ void fill_profile_zy_kernel(double* __restrict__ a, const double* __restrict__ prof, const double* __restrict__ add_y,
                                       const double* __restrict__ add_z, int sx, int sy, size_t n)
{
    const size_t sxy = (size_t)sx * sy;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const size_t k = q / sxy;
        const int j = (int)((q - k * sxy) / sx);
        a[q] = (prof[k] + add_y[j]) + add_z[k];
    }
}

__global__ void fill_plane_x_kernel(double* __restrict__ a, int sx, int ix, double v, size_t nyz)
{
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < nyz; q += (size_t)gridDim.x * blockDim.x)
        a[(size_t)ix + (size_t)sx * q] = v;
}

extern "C" int ns3d_fill_profile_z(ns3d_ctx* ctx, double* A, int sx, int sy, int sz, const double* h_profile)
{
    NS3D_CHECK_CTX(ctx);
    if (!A || !h_profile || sx <= 0 || sy <= 0 || sz <= 0) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_fill_profile_z: bad argument");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    double* d_prof = nullptr;
    NS3D_CUDA(ctx, cudaMalloc(&d_prof, (size_t)sz * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_prof, h_profile, (size_t)sz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const size_t n = (size_t)sx * sy * sz;
        const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->num_sms * 16);
        fill_profile_z_kernel<<<blocks, 256, 0, ctx->stream>>>(A, d_prof, (size_t)sx * sy, n);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // h_profile may go away; d_prof is freed below
    cudaFree(d_prof);
    if (e != cudaSuccess) return ns3d_fail(ctx, NS3D_ECUDA, "ns3d_fill_profile_z failed: %s", cudaGetErrorString(e));
    return NS3D_OK;
}

extern "C" int ns3d_fill_profile_zy(ns3d_ctx* ctx, double* A, int sx, int sy, int sz, const double* h_profile, const double* h_add_y,
                                    const double* h_add_z)
{
    NS3D_CHECK_CTX(ctx);
    if (!A || !h_profile || !h_add_y || !h_add_z || sx <= 0 || sy <= 0 || sz <= 0)
        return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_fill_profile_zy: bad argument");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    double* d = nullptr;   // prof (sz) | add_z (sz) | add_y (sy)
    NS3D_CUDA(ctx, cudaMalloc(&d, (size_t)(2 * sz + sy) * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, h_profile, (size_t)sz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + sz, h_add_z, (size_t)sz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + 2 * sz, h_add_y, (size_t)sy * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const size_t n = (size_t)sx * sy * sz;
        const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)ctx->num_sms * 16);
        fill_profile_zy_kernel<<<blocks, 256, 0, ctx->stream>>>(A, d, d + 2 * sz, d + sz, sx, sy, n);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // the host arrays may go away; d is freed below
    cudaFree(d);
    if (e != cudaSuccess) return ns3d_fail(ctx, NS3D_ECUDA, "ns3d_fill_profile_zy failed: %s", cudaGetErrorString(e));
    return NS3D_OK;
}

extern "C" int ns3d_fill_plane_x(ns3d_ctx* ctx, double* A, int sx, int sy, int sz, int ix, double value)
{
    NS3D_CHECK_CTX(ctx);
    if (!A || sx <= 0 || sy <= 0 || sz <= 0 || ix < 0 || ix >= sx) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_fill_plane_x: bad argument");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nyz = (size_t)sy * sz;
    const unsigned blocks = (unsigned)std::min<size_t>((nyz + 255) / 256, (size_t)ctx->num_sms * 16);
    fill_plane_x_kernel<<<blocks, 256, 0, ctx->stream>>>(A, sx, ix, value, nyz);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// ---------------------------------------------------------------------------------------------
// maximum(abs.(A))   (K8' of SURVEY.md: the reference allocates abs.(Rp) and reduces it;
// here one pass, no temporary)
// ---------------------------------------------------------------------------------------------
__global__ void max_abs_kernel(const double* __restrict__ a, size_t n, unsigned long long* out)
{
    unsigned long long m = 0ULL;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        unsigned long long b = absbits(a[q]);
        m = b > m ? b : m;
    }
    block_max_to_global(m, out);
}

int ns3d_internal_max_abs_async(ns3d_ctx* ctx, const double* A, size_t count)
{
    NS3D_CUDA(ctx, cudaMemsetAsync(ctx->d_maxbits, 0, sizeof(unsigned long long), ctx->stream));
    unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, (size_t)ctx->num_sms * 8);
    max_abs_kernel<<<blocks ? blocks : 1, 256, 0, ctx->stream>>>(A, count, ctx->d_maxbits);
    NS3D_LAUNCH_CHECK(ctx);
    return NS3D_OK;
}

// Reads ctx->d_maxbits[0] back (synchronising) and max-reduces it over the ranks.
int ns3d_internal_read_max(ns3d_ctx* ctx, double* h_out)
{
    if (ctx->nccl) {
        // Unsigned 64-bit max over bit patterns of |x| == NaN-propagating max over |x|.
        NS3D_NCCL(ctx, g_nccl.AllReduce(ctx->d_maxbits, ctx->d_maxbits, 1, ncclUint64, ncclMax,
                                        (ncclComm_t)ctx->nccl, ctx->stream));
    }
    NS3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_maxbits, ctx->d_maxbits, sizeof(unsigned long long),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long b = ctx->h_maxbits[0];
    double v;
    memcpy(&v, &b, sizeof v);
    *h_out = v;
    return NS3D_OK;
}

extern "C" int ns3d_max_abs(ns3d_ctx* ctx, const double* A, size_t count, double* h_out)
{
    NS3D_CHECK_CTX(ctx);
    if (!A || !h_out || count == 0) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_max_abs: bad argument");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    NS3D_TRY(ns3d_internal_max_abs_async(ctx, A, count));
    return ns3d_internal_read_max(ctx, h_out);
}

// ---------------------------------------------------------------------------------------------
// communicator: z-slabs, rank r <-> r+-1
// ---------------------------------------------------------------------------------------------
extern "C" int ns3d_comm_unique_id(char id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!id) return NS3D_EINVAL;
    NS3D_TRY(load_nccl(nullptr));
    ncclUniqueId uid;
    ncclResult_t r = g_nccl.GetUniqueId(&uid);
    if (r != ncclSuccess) return ns3d_fail(nullptr, NS3D_ECOMM, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(id, &uid, 128);
    return NS3D_OK;
}

extern "C" int ns3d_comm_init(ns3d_ctx* ctx, int rank, int nranks, const char id[128])
{
    NS3D_CHECK_CTX(ctx);
    if (nranks < 1 || rank < 0 || rank >= nranks || !id)
        return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_comm_init: bad rank %d / nranks %d", rank, nranks);
    if (ctx->nccl) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_comm_init: communicator already attached");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->rank = rank;
    ctx->nranks = nranks;
    if (nranks == 1) return NS3D_OK;  // singleton: every update_halo! is a no-op, like IGG on one rank
    NS3D_TRY(load_nccl(ctx));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclComm_t comm;
    NS3D_NCCL(ctx, g_nccl.CommInitRank(&comm, nranks, uid, rank));
    ctx->nccl = comm;
    // Peer-memory halo path: every rank owns a mailbox and maps its neighbours' through CUDA IPC.
    // If IPC is not available the NCCL send/recv path stays in charge (p2p_ready == false).
    NS3D_CUDA(ctx, cudaMalloc(&ctx->mbox, NS3D_MB_WORDS * sizeof(unsigned long long)));
    NS3D_CUDA(ctx, cudaMemset(ctx->mbox, 0, NS3D_MB_WORDS * sizeof(unsigned long long)));
    void *lo = nullptr, *hi = nullptr;
    const int rc = ns3d_internal_p2p_map(ctx, ctx->mbox, &lo, &hi);
    ctx->p2p_map.erase(ctx->mbox);
    ctx->peer_mbox[0] = (unsigned long long*)lo;
    ctx->peer_mbox[1] = (unsigned long long*)hi;
    ctx->p2p_ready = rc == NS3D_OK;   // agreed on by all ranks inside ns3d_internal_p2p_map
    ctx->err.clear();
    return NS3D_OK;
}

// Exchanges the CUDA IPC handle of `local_base` (the base of a cudaMalloc block) with both slab
// neighbours and maps theirs.  COLLECTIVE: every rank must call it at the same point with its
// corresponding buffer, and every rank takes part in the whole exchange even when something failed
// locally (it then sends a null handle): the outcome is agreed on with a min-all-reduce, so either
// every rank returns NS3D_OK or every rank returns an error -- nobody is left waiting in a send/recv
// group.  Cached per local pointer; entries are dropped when the block is freed.
int ns3d_internal_p2p_map(ns3d_ctx* ctx, const void* local_base, void** peer_lo, void** peer_hi)
{
    auto it = ctx->p2p_map.find(local_base);
    if (it != ctx->p2p_map.end()) {
        *peer_lo = it->second.first;
        *peer_hi = it->second.second;
        return NS3D_OK;
    }
    *peer_lo = *peer_hi = nullptr;
    if (!ctx->nccl) return ns3d_fail(ctx, NS3D_ECOMM, "p2p_map: no communicator");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    ncclComm_t comm = (ncclComm_t)ctx->nccl;
    const int lo = ctx->rank - 1, hi = ctx->rank + 1;
    int rc = NS3D_OK;            // first local failure; the collective steps below run regardless
    auto note_cuda = [&](cudaError_t e, const char* what) {
        if (e == cudaSuccess) return;
        cudaGetLastError();
        if (rc == NS3D_OK) rc = ns3d_fail(ctx, NS3D_ECOMM, "p2p_map: %s failed: %s", what, cudaGetErrorString(e));
    };
    auto note_nccl = [&](ncclResult_t r, const char* what) {
        if (r != ncclSuccess && rc == NS3D_OK) rc = ns3d_fail(ctx, NS3D_ECOMM, "p2p_map: %s failed: %s", what, g_nccl.GetErrorString(r));
    };
    // [0,64) mine, [64,128) from lower, [128,192) from upper, [192,200) outcome: scratch behind the reduction words
    unsigned char* stage = (unsigned char*)(ctx->d_maxbits + 8);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    note_cuda(cudaIpcGetMemHandle(&mine, const_cast<void*>(local_base)), "cudaIpcGetMemHandle");
    if (rc != NS3D_OK) memset(&mine, 0, sizeof mine);
    note_cuda(cudaMemcpyAsync(stage, &mine, 64, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync");
    note_nccl(g_nccl.GroupStart(), "ncclGroupStart");
    if (lo >= 0) {
        note_nccl(g_nccl.Send(stage, 64, ncclUint8, lo, comm, ctx->stream), "ncclSend");
        note_nccl(g_nccl.Recv(stage + 64, 64, ncclUint8, lo, comm, ctx->stream), "ncclRecv");
    }
    if (hi < ctx->nranks) {
        note_nccl(g_nccl.Send(stage, 64, ncclUint8, hi, comm, ctx->stream), "ncclSend");
        note_nccl(g_nccl.Recv(stage + 128, 64, ncclUint8, hi, comm, ctx->stream), "ncclRecv");
    }
    note_nccl(g_nccl.GroupEnd(), "ncclGroupEnd");   // the group is closed on every path
    cudaIpcMemHandle_t theirs[2];
    memset(theirs, 0, sizeof theirs);
    note_cuda(cudaMemcpyAsync(theirs, stage + 64, 128, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync");
    note_cuda(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    ctx->halo_calls++;
    void* mapped[2] = {nullptr, nullptr};
    for (int q = 0; q < 2 && rc == NS3D_OK; ++q) {
        const int nb = q == 0 ? lo : hi;
        if (nb < 0 || nb >= ctx->nranks) continue;
        cudaError_t e = cudaIpcOpenMemHandle(&mapped[q], theirs[q], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            mapped[q] = nullptr;
            rc = ns3d_fail(ctx, NS3D_ECOMM, "cudaIpcOpenMemHandle (rank %d -> %d) failed: %s", ctx->rank, nb, cudaGetErrorString(e));
        }
    }
    // agree on the outcome
    ctx->h_maxbits[4] = rc == NS3D_OK ? 1ULL : 0ULL;
    int rc2 = NS3D_OK;
    if (cudaMemcpyAsync(ctx->d_maxbits + 4, ctx->h_maxbits + 4, 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        g_nccl.AllReduce(ctx->d_maxbits + 4, ctx->d_maxbits + 4, 1, ncclUint64, ncclMin, comm, ctx->stream) != ncclSuccess ||
        cudaMemcpyAsync(ctx->h_maxbits + 4, ctx->d_maxbits + 4, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        cudaGetLastError();
        rc2 = NS3D_ECOMM;
    }
    if (rc == NS3D_OK && (rc2 != NS3D_OK || ctx->h_maxbits[4] != 1ULL))
        rc = ns3d_fail(ctx, NS3D_ECOMM, "p2p_map: a neighbour could not share or map its buffer");
    if (rc != NS3D_OK) {
        for (int q = 0; q < 2; ++q)
            if (mapped[q]) cudaIpcCloseMemHandle(mapped[q]);
        return rc;
    }
    ctx->p2p_map[local_base] = {mapped[0], mapped[1]};
    *peer_lo = mapped[0];
    *peer_hi = mapped[1];
    return NS3D_OK;
}

// Gathers one block of bytes per rank on rank 0 (device memory on both sides): rank r sends `bytes`
// bytes from d_send, rank 0 receives rank r's block at d_recv + offs[r] (its own block is copied).
// COLLECTIVE over the communicator; `bytes_all` (length nranks) is meaningful on rank 0 only.
int ns3d_internal_gather_bytes(ns3d_ctx* ctx, const void* d_send, size_t bytes, void* d_recv, const size_t* bytes_all)
{
    if (!ctx->nccl) return ns3d_fail(ctx, NS3D_ECOMM, "gather: no communicator attached");
    ncclComm_t comm = (ncclComm_t)ctx->nccl;
    if (ctx->rank != 0) {
        if (bytes) NS3D_NCCL(ctx, g_nccl.Send(d_send, bytes, ncclUint8, 0, comm, ctx->stream));
        return NS3D_OK;
    }
    if (bytes) NS3D_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    size_t off = bytes_all[0];
    NS3D_NCCL(ctx, g_nccl.GroupStart());
    for (int r = 1; r < ctx->nranks; ++r) {
        if (bytes_all[r]) NS3D_NCCL(ctx, g_nccl.Recv((char*)d_recv + off, bytes_all[r], ncclUint8, r, comm, ctx->stream));
        off += bytes_all[r];
    }
    NS3D_NCCL(ctx, g_nccl.GroupEnd());
    return NS3D_OK;
}

extern "C" int ns3d_comm_rank(const ns3d_ctx* ctx) { return ctx ? ctx->rank : NS3D_EINVAL; }
extern "C" int ns3d_comm_size(const ns3d_ctx* ctx) { return ctx ? ctx->nranks : NS3D_EINVAL; }

// IGG update_halo! for dims=(1,1,N), overlap 2, halo width 1: a field
// with sz planes has overlap ol = 2 + (sz - nz); 1-based plane `ol` goes to the lower
// neighbour's plane `sz`, plane `sz-ol+1` to the upper neighbour's plane 1.  z-planes are
// contiguous (x fastest, z slowest) so no packing is needed.
int ns3d_internal_halo_z(ns3d_ctx* ctx, cudaStream_t s, double* const* fields, const int* sx, const int* sy,
                         const int* sz, int nfields, int nz)
{
    if (ctx->nranks == 1) return NS3D_OK;
    if (!ctx->nccl) return ns3d_fail(ctx, NS3D_ECOMM, "update_halo: no communicator attached");
    ncclComm_t comm = (ncclComm_t)ctx->nccl;
    const int lo = ctx->rank - 1, hi = ctx->rank + 1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &cap);
    if (cap == cudaStreamCaptureStatusNone) ctx->halo_calls++;
    NS3D_NCCL(ctx, g_nccl.GroupStart());
    for (int f = 0; f < nfields; ++f) {
        const int ol = 2 + (sz[f] - nz);
        if (ol < 2) {
            g_nccl.GroupEnd();
            return ns3d_fail(ctx, NS3D_EINVAL, "update_halo: field %d has overlap %d < 2 and cannot be exchanged", f, ol);
        }
        const size_t plane = (size_t)sx[f] * sy[f];
        double* a = fields[f];
        if (lo >= 0) {
            NS3D_NCCL(ctx, g_nccl.Send(a + plane * (size_t)(ol - 1), plane, ncclDouble, lo, comm, s));
            NS3D_NCCL(ctx, g_nccl.Recv(a, plane, ncclDouble, lo, comm, s));
        }
        if (hi < ctx->nranks) {
            NS3D_NCCL(ctx, g_nccl.Send(a + plane * (size_t)(sz[f] - ol), plane, ncclDouble, hi, comm, s));
            NS3D_NCCL(ctx, g_nccl.Recv(a + plane * (size_t)(sz[f] - 1), plane, ncclDouble, hi, comm, s));
        }
    }
    NS3D_NCCL(ctx, g_nccl.GroupEnd());
    return NS3D_OK;
}

extern "C" int ns3d_update_halo(ns3d_ctx* ctx, double* const* fields, const int* sx, const int* sy,
                                const int* sz, int nfields, int nz)
{
    NS3D_CHECK_CTX(ctx);
    if (nfields < 0 || (nfields > 0 && (!fields || !sx || !sy || !sz)))
        return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_update_halo: bad argument");
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    return ns3d_internal_halo_z(ctx, ctx->stream, fields, sx, sy, sz, nfields, nz);
}

extern "C" int ns3d_allreduce_max(ns3d_ctx* ctx, double* h_inout)
{
    NS3D_CHECK_CTX(ctx);
    if (!h_inout) return ns3d_fail(ctx, NS3D_EINVAL, "ns3d_allreduce_max: NULL");
    if (ctx->nranks == 1) return NS3D_OK;
    NS3D_CUDA(ctx, cudaSetDevice(ctx->device));
    // MPI.MAX on doubles; NaN must win (Julia's maximum already produced NaN locally).
    double v = *h_inout;
    unsigned long long b;
    if (std::isnan(v)) {
        b = 0xffffffffffffffffULL;
    } else {
        // order-preserving map double -> uint64
        memcpy(&b, &v, sizeof b);
        b = (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
    }
    ctx->h_maxbits[1] = b;
    NS3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_maxbits + 1, ctx->h_maxbits + 1, sizeof b, cudaMemcpyHostToDevice, ctx->stream));
    NS3D_NCCL(ctx, g_nccl.AllReduce(ctx->d_maxbits + 1, ctx->d_maxbits + 1, 1, ncclUint64, ncclMax,
                                    (ncclComm_t)ctx->nccl, ctx->stream));
    NS3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_maxbits + 1, ctx->d_maxbits + 1, sizeof b, cudaMemcpyDeviceToHost, ctx->stream));
    NS3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    b = ctx->h_maxbits[1];
    if (b == 0xffffffffffffffffULL) {
        *h_inout = NAN;
    } else {
        b = (b & 0x8000000000000000ULL) ? (b & 0x7fffffffffffffffULL) : ~b;
        memcpy(h_inout, &b, sizeof b);
    }
    return NS3D_OK;
}
