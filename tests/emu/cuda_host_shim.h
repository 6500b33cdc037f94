// CUDA-on-host shim (test infrastructure): lets g++ compile the device code of libns3d.so
// unchanged.  Every CUDA thread of a block runs as a host thread, __syncthreads() is a pthread
// barrier over the block, named barriers (bar.sync id, n) are pthread barriers too, __shared__
// arrays are function-local statics (blocks run one after the other), blocks of a grid run
// sequentially.  Kernels that never synchronise can be run without threads (launch_serial).
// Only what the library's kernels use is provided.  Not a product path: nothing outside tests/
// includes it.
#pragma once

#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define NS3D_HOST_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {
inline thread_local dim3 tls_threadIdx, tls_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline pthread_barrier_t g_barrier;
inline pthread_barrier_t g_named[16];      // bar.sync id, count: initialised on first use within a block
inline int g_named_count[16];
inline pthread_mutex_t g_named_lock = PTHREAD_MUTEX_INITIALIZER;
inline bool g_serial = false;              // the running launch has no host threads behind its CUDA threads
}  // namespace emu
#define threadIdx (emu::tls_threadIdx)
#define blockIdx (emu::tls_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

using std::max;
using std::min;

inline void __syncthreads()
{
    if (emu::g_serial) {
        std::fprintf(stderr, "emu: __syncthreads() inside a kernel that was launched without threads\n");
        std::abort();
    }
    pthread_barrier_wait(&emu::g_barrier);
}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline long long clock64()  // 10 ns ticks: the kernels' spin limits (8e9 "cycles") become 80 s, generous for a loaded CI box
{
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() / 10;
}
template <class T>
inline T __ldcv(const T* p)
{
    return *(const volatile T*)p;
}
inline long long __double_as_longlong(double x)
{
    long long v;
    std::memcpy(&v, &x, sizeof v);
    return v;
}
// cvt.rmi.s64.f64: round towards -inf, saturating; NaN -> 0 (PTX cvt of NaN to an integer type)
inline long long __double2ll_rd(double x)
{
    if (x != x) return 0;
    const double f = std::floor(x);
    if (f >= 9223372036854775808.0) return INT64_MAX;
    if (f < -9223372036854775808.0) return INT64_MIN;
    return (long long)f;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
    return old;
}

namespace emu {

inline void spin_pause() { sched_yield(); }

// `bar.sync id, count` (PTX named barrier): every participating thread passes the same count.
inline void named_barrier(int id, int count)
{
    pthread_mutex_lock(&g_named_lock);
    if (g_named_count[id] == 0) {
        pthread_barrier_init(&g_named[id], nullptr, (unsigned)count);
        g_named_count[id] = count;
    }
    pthread_mutex_unlock(&g_named_lock);
    pthread_barrier_wait(&g_named[id]);
}

// Warp shuffle among the 32 host threads of a warp: exchange through a per-warp slot array, two
// warp-wide barriers around the read (full-mask calls only, which is all the library uses).
struct WarpXchg {
    pthread_barrier_t bar;
    unsigned long long slot[32];
};
inline std::vector<WarpXchg>* g_warps = nullptr;
inline unsigned linear_tid()
{
    return tls_threadIdx.x + g_blockDim.x * (tls_threadIdx.y + g_blockDim.y * tls_threadIdx.z);
}

struct ThreadArg {
    void (*fn)(void*);
    void* ctx;
    dim3 tid;
    dim3 grid;
};

// One host thread per CUDA thread of a block, alive for the whole launch: it runs its CUDA thread of
// every block in turn, with a block-wide barrier between blocks (blocks execute one after the other,
// so function-local `static` really is per-block shared memory).
inline void* thread_main(void* a)
{
    ThreadArg* t = (ThreadArg*)a;
    tls_threadIdx = t->tid;
    for (unsigned bz = 0; bz < t->grid.z; ++bz)
        for (unsigned by = 0; by < t->grid.y; ++by)
            for (unsigned bx = 0; bx < t->grid.x; ++bx) {
                tls_blockIdx = dim3(bx, by, bz);
                t->fn(t->ctx);
                pthread_barrier_wait(&g_barrier);   // end of block: every thread is out of the kernel
                if (t->tid.x == 0 && t->tid.y == 0 && t->tid.z == 0) {
                    for (int id = 0; id < 16; ++id)
                        if (g_named_count[id]) {
                            pthread_barrier_destroy(&g_named[id]);
                            g_named_count[id] = 0;
                        }
                }
                pthread_barrier_wait(&g_barrier);
            }
    return nullptr;
}

// Runs `body()` once per CUDA thread of a grid x block launch, CUDA threads = host threads.
template <class F>
void launch(dim3 grid, dim3 block, F body)
{
    const unsigned nthreads = block.x * block.y * block.z;
    g_gridDim = grid;
    g_blockDim = block;
    g_serial = false;
    std::vector<WarpXchg> warps((nthreads + 31) / 32);
    for (size_t w = 0; w < warps.size(); ++w)
        pthread_barrier_init(&warps[w].bar, nullptr, std::min(32u, nthreads - 32u * (unsigned)w));
    g_warps = &warps;
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    std::vector<pthread_t> th(nthreads);
    std::vector<ThreadArg> args(nthreads);
    auto tramp = [](void* c) { (*(F*)c)(); };
    pthread_barrier_init(&g_barrier, nullptr, nthreads);
    unsigned t = 0;
    for (unsigned tz = 0; tz < block.z; ++tz)
        for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx, ++t) {
                args[t] = ThreadArg{tramp, &body, dim3(tx, ty, tz), grid};
                if (pthread_create(&th[t], &attr, thread_main, &args[t]) != 0) {
                    std::fprintf(stderr, "emu: pthread_create failed\n");
                    std::abort();
                }
            }
    for (unsigned q = 0; q < nthreads; ++q) pthread_join(th[q], nullptr);
    pthread_barrier_destroy(&g_barrier);
    for (auto& w : warps) pthread_barrier_destroy(&w.bar);
    g_warps = nullptr;
    pthread_attr_destroy(&attr);
}

// The same for kernels that never synchronise (no __syncthreads, shuffles or named barriers): the
// CUDA threads run one after the other on the calling thread.
template <class F>
void launch_serial(dim3 grid, dim3 block, F body)
{
    g_gridDim = grid;
    g_blockDim = block;
    g_serial = true;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                tls_blockIdx = dim3(bx, by, bz);
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx) {
                            tls_threadIdx = dim3(tx, ty, tz);
                            body();
                        }
            }
    g_serial = false;
}

}  // namespace emu

inline unsigned long long __shfl_xor_sync(unsigned, unsigned long long v, int lane_mask)
{
    const unsigned tid = emu::linear_tid();
    emu::WarpXchg& w = (*emu::g_warps)[tid >> 5];
    w.slot[tid & 31] = v;
    pthread_barrier_wait(&w.bar);
    const unsigned long long r = w.slot[(tid & 31) ^ (unsigned)lane_mask];
    pthread_barrier_wait(&w.bar);
    return r;
}
