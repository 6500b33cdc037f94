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

inline void __syncthreads();
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
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
inline unsigned long long wall_ns()
{
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// dynamic shared memory (`extern __shared__`): blocks run one after the other, so one buffer serves them all
inline void* dyn_smem()
{
    alignas(16) static char buf[1 << 20];
    return buf;
}

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

// One host thread per CUDA thread of a block.  The threads live in a POOL per block size and are
// reused by every launch of that size (creating 256-512 threads per launch dominated the run time
// of many-launch tests): a launch publishes its job, releases the pool through a start barrier and
// waits for it at a done barrier.  Inside a launch each pool thread runs its CUDA thread of every
// block in turn, with a block-wide barrier between blocks (blocks execute one after the other, so
// function-local `static` really is per-block shared memory).
struct Job {
    void (*fn)(void*) = nullptr;
    void* ctx = nullptr;
    dim3 grid, block;
};
struct Pool {
    unsigned nthreads = 0;
    pthread_barrier_t start, done, sync;   // start/done: pool + launcher; sync: the pool (= __syncthreads)
    Job job;
    std::vector<WarpXchg> warps;
};
inline Pool* g_pool = nullptr;             // the pool of the running launch

struct PoolThread {
    Pool* pool;
    unsigned tid;
};

inline void* pool_main(void* a)
{
    PoolThread* me = (PoolThread*)a;
    Pool* p = me->pool;
    for (;;) {
        pthread_barrier_wait(&p->start);
        const Job& j = p->job;
        const unsigned t = me->tid;
        tls_threadIdx = dim3(t % j.block.x, (t / j.block.x) % j.block.y, t / (j.block.x * j.block.y));
        for (unsigned bz = 0; bz < j.grid.z; ++bz)
            for (unsigned by = 0; by < j.grid.y; ++by)
                for (unsigned bx = 0; bx < j.grid.x; ++bx) {
                    tls_blockIdx = dim3(bx, by, bz);
                    j.fn(j.ctx);
                    pthread_barrier_wait(&p->sync);   // end of block: every thread is out of the kernel
                    if (t == 0) {
                        for (int id = 0; id < 16; ++id)
                            if (g_named_count[id]) {
                                pthread_barrier_destroy(&g_named[id]);
                                g_named_count[id] = 0;
                            }
                    }
                    pthread_barrier_wait(&p->sync);
                }
        pthread_barrier_wait(&p->done);
    }
    return nullptr;
}

inline Pool* get_pool(unsigned nthreads)
{
    static std::vector<Pool*> pools;
    for (Pool* p : pools)
        if (p->nthreads == nthreads) return p;
    Pool* p = new Pool();
    p->nthreads = nthreads;
    pthread_barrier_init(&p->start, nullptr, nthreads + 1);
    pthread_barrier_init(&p->done, nullptr, nthreads + 1);
    pthread_barrier_init(&p->sync, nullptr, nthreads);
    p->warps.resize((nthreads + 31) / 32);
    for (size_t w = 0; w < p->warps.size(); ++w)
        pthread_barrier_init(&p->warps[w].bar, nullptr, std::min(32u, nthreads - 32u * (unsigned)w));
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    pthread_attr_setdetachstate(&attr, PTHREAD_CREATE_DETACHED);
    for (unsigned t = 0; t < nthreads; ++t) {
        pthread_t th;
        if (pthread_create(&th, &attr, pool_main, new PoolThread{p, t}) != 0) {
            std::fprintf(stderr, "emu: pthread_create failed\n");
            std::abort();
        }
    }
    pthread_attr_destroy(&attr);
    pools.push_back(p);
    return p;
}

// Runs `body()` once per CUDA thread of a grid x block launch, CUDA threads = host threads.
template <class F>
void launch(dim3 grid, dim3 block, F body)
{
    Pool* p = get_pool(block.x * block.y * block.z);
    g_gridDim = grid;
    g_blockDim = block;
    g_serial = false;
    g_pool = p;
    g_warps = &p->warps;
    p->job.fn = [](void* c) { (*(F*)c)(); };
    p->job.ctx = &body;
    p->job.grid = grid;
    p->job.block = block;
    pthread_barrier_wait(&p->start);
    pthread_barrier_wait(&p->done);
    g_pool = nullptr;
    g_warps = nullptr;
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

inline void __syncthreads()
{
    if (emu::g_serial) {
        std::fprintf(stderr, "emu: __syncthreads() inside a kernel that was launched without threads\n");
        std::abort();
    }
    pthread_barrier_wait(&emu::g_pool->sync);
}

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
