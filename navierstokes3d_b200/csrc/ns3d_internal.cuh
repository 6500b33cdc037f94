// Internal declarations shared by the translation units of libns3d.so.
// Nothing in here crosses the C ABI (include/ns3d.h).
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>

#include "../../include/ns3d.h"
#include "ns3d_shared.cuh"

struct ns3d_ctx {
    int device = 0;
    int mode = NS3D_PARITY;
    int num_sms = 0;
    cudaStream_t stream = nullptr;       // all operators run here
    cudaStream_t comm_stream = nullptr;  // halo exchange, overlapped with interior compute
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    // copy engines beside the compute stream: asynchronous uploads / downloads of a driver that pipelines
    // host <-> device traffic with the time steps (ns3d_h2d_async, ns3d_d2h_async, ns3d_stream_wait)
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_xfer = nullptr;
    long long launches = 0;
    std::string err;
    std::unordered_map<void*, size_t> allocs;
    size_t bytes = 0;
    // scratch owned by the context
    void* out_stage = nullptr;    // device staging of the output path (ns3d_box_d2h)
    size_t out_stage_bytes = 0;
    void* gather_stage = nullptr;  // rank 0: where the ranks' boxes land (ns3d_gather_box)
    size_t gather_stage_bytes = 0;
    unsigned long long* d_maxbits = nullptr;  // device accumulator of max |x| bit patterns
    unsigned long long* h_maxbits = nullptr;  // pinned host mirror
    // communicator (z-slabs, one rank per GPU)
    void* nccl = nullptr;
    int rank = 0, nranks = 1;
    // peer-memory halo path: a mailbox of 64-bit words in device memory (see NS3D_MB_*), the
    // neighbours' mailboxes and field buffers mapped through CUDA IPC
    unsigned long long* mbox = nullptr;
    unsigned long long* peer_mbox[2] = {nullptr, nullptr};  // [0] lower, [1] upper neighbour
    bool p2p_ready = false;
    std::unordered_map<const void*, std::pair<void*, void*>> p2p_map;  // local base -> (lower, upper) peer base
    int opt_p2p = 1;
    int opt_p2p_split = 1;    // z-slabs: the interface chunks of a pass as a launch of their own on the high-priority stream (0: one launch per pass)
    // tuning knobs (ns3d_set_option)
    int opt_serpentine = -1;  // -1 = by working-set size
    int opt_graphs = 1;       // replay chunks of PT iterations as CUDA graphs
    int opt_graph_pieces = -1; // ... each chunk as this many graphs launched back to back (-1 = default 1; see ptv_run_pieces)
    long long halo_calls = 0; // uncaptured halo exchanges so far (NCCL peers connected)
    // the fused loop's pitched working copies (ns3d_ptv.cu): Pr x2, dPrdtau x2 (ping-pong), divV
    double* ptv_raw[5] = {};    // cudaMalloc blocks
    double* ptv[5] = {};        // element (0,0,0) inside them (front padding skipped)
    int ptv_nx = 0, ptv_ny = 0, ptv_nz = 0;
    bool ptv_peers_mapped = false;
    void* ptv_graphs = nullptr;  // graph cache owned by ns3d_ptv.cu
    int opt_ptv_k = 0;        // PT iterations per launch (0 = default)
    int opt_ptv_ns = 0;       // staging slots of the TMA ring (0 = default 4; >= 3)
    unsigned* ptv_work = nullptr;  // work queue + completion counters of the persistent launch (ptv_flow_kernel)
    size_t ptv_work_words = 0;
    int opt_ptv_flow = 0;     // persistent launch of whole chunks of iterations (single rank; off: see ptv_flow_kernel)
    // z-bands (single rank): every pass as `bands` launches on as many streams; band b of pass n+1 waits for bands b-1, b, b+1
    // of pass n only, so consecutive passes overlap like a wavefront instead of idling in every launch's ramp and tail
    int opt_ptv_bands = -1;   // -1 = by grid size, 0 / 1 = off
    cudaStream_t band_stream[16] = {};
    cudaEvent_t band_ev[2][16] = {};
    cudaEvent_t band_fork = nullptr;
    int opt_ptv_tma = 1;      // stage the z-plane tiles with the TMA unit (0 = plain loads by all threads)
    int opt_ptv_lb = -1;      // launch-bounds variant (threads / CTAs per SM): 0 = 256/2, 1 = 256/3, 2 = 256/4, 3 = 512/1, 4 = 512/2 (-1 = default)
    int opt_ptv_pxt = 0;      // thread columns per tile (0 = balanced automatically)
    int opt_ptv_bty = 0;      // thread rows per tile (0 = as many as the CTA has threads for)
    size_t l2_bytes = 0;
};

int ns3d_fail(ns3d_ctx* ctx, int code, const char* fmt, ...);

#define NS3D_CHECK_CTX(ctx)            \
    do {                               \
        if (!(ctx)) return NS3D_EINVAL; \
    } while (0)

#define NS3D_CUDA(ctx, call)                                                                  \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return ns3d_fail((ctx), NS3D_ECUDA, "%s failed: %s (%s:%d)", #call,               \
                             cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)

#define NS3D_LAUNCH_CHECK(ctx)                                                                \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess)                                                               \
            return ns3d_fail((ctx), NS3D_ECUDA, "kernel launch failed: %s (%s:%d)",           \
                             cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)

#define NS3D_TRY(expr)            \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != NS3D_OK) return rc__; \
    } while (0)

static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// Bit pattern of |x|: for non-negative doubles the unsigned order of the patterns is the
// numeric order, +Inf sorts above every finite value and every NaN above +Inf -- so an
// unsigned max over patterns is a NaN-propagating maximum(abs.(A)) like Julia's.
__device__ __forceinline__ unsigned long long absbits(double x)
{
    return (unsigned long long)__double_as_longlong(x) & 0x7fffffffffffffffULL;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// Block-wide max of bit patterns, then one atomicMax per CTA.
__device__ __forceinline__ void block_max_to_global(unsigned long long v, unsigned long long* out)
{
    __shared__ unsigned long long smax[32];
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    v = warp_max_u64(v);
    if ((tid & 31) == 0) smax[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        v = tid < (nthreads + 31) / 32 ? smax[tid] : 0ULL;
        v = warp_max_u64(v);
        if (tid == 0 && v != 0ULL) atomicMax(out, v);
    }
}

// internal cross-TU entry points
int ns3d_internal_p2p_map(ns3d_ctx* ctx, const void* local_base, void** peer_lo, void** peer_hi);
void ns3d_internal_ptv_free(ns3d_ctx* ctx);     // graph cache of ns3d_ptv.cu
void ns3d_internal_ptv_release(ns3d_ctx* ctx);  // ... and its buffers
int ns3d_internal_ptv_solve(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV, const ns3d_pt_params* p, int* h_iters,
                            double* h_err_hist, int err_cap, int* h_nchecks);
int ns3d_internal_ptv_iterate(ns3d_ctx* ctx, double* Pr, double* dPrdtau, const double* divV, const ns3d_pt_params* p, int n_iter);
int ns3d_internal_ptv_describe(ns3d_ctx* ctx, const ns3d_pt_params* p, char* buf, int cap, int* iters_per_launch);
// fused once-per-step kernels (ns3d_step.cu, ns3d_ops.cu)
int ns3d_internal_predict_fused(ns3d_ctx* ctx, double* Vxn, double* Vyn, double* Vzn, double* C, const double* Vx, const double* Vy,
                                const double* Vz, const ns3d_step_params* sp);
int ns3d_internal_correct_fused(ns3d_ctx* ctx, double* Vx, double* Vy, double* Vz, double* C, double* C_o, const double* Pr,
                                const ns3d_step_params* sp);
int ns3d_internal_advect_all(ns3d_ctx* ctx, double* Vx, const double* Vx_o, double* Vy, const double* Vy_o, double* Vz,
                             const double* Vz_o, double* C, const double* C_o, double dt, double dx, double dy, double dz, int nx,
                             int ny, int nz);
void ns3d_internal_out_free(ns3d_ctx* ctx);
int ns3d_internal_gather_bytes(ns3d_ctx* ctx, const void* d_send, size_t bytes, void* d_recv, const size_t* bytes_all);
int ns3d_internal_max_abs_async(ns3d_ctx* ctx, const double* A, size_t count);  // result -> ctx->d_maxbits
int ns3d_internal_read_max(ns3d_ctx* ctx, double* h_out);                       // sync + allreduce
int ns3d_internal_halo_z(ns3d_ctx* ctx, cudaStream_t s, double* const* fields, const int* sx,
                         const int* sy, const int* sz, int nfields, int nz);
