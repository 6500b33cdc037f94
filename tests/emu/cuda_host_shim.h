// CUDA-on-host shim (test infrastructure): lets g++ compile the device code of
// navierstokes3d_b200/csrc/ns3d_pt_kernels.cuh unchanged.  Every CUDA thread of a block runs as a
// host thread, __syncthreads() is a pthread barrier over the block, __shared__ arrays are
// function-local statics (blocks run one after the other), blocks of a grid run sequentially.
// Only what those kernels use is provided.  Not a product path: nothing outside tests/ includes it.
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
inline thread_local dim3 tls_threadIdx;
inline dim3 g_blockIdx, g_blockDim, g_gridDim;
inline pthread_barrier_t g_barrier;
inline pthread_barrier_t g_named[16];      // bar.sync id, count: initialised on first use within a block
inline int g_named_count[16];
inline pthread_mutex_t g_named_lock = PTHREAD_MUTEX_INITIALIZER;
}  // namespace emu
#define threadIdx (emu::tls_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

using std::max;
using std::min;

inline void __syncthreads() { pthread_barrier_wait(&emu::g_barrier); }
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

struct ThreadArg {
    void (*fn)(void*);
    void* ctx;
    dim3 tid;
};

inline void* thread_main(void* a)
{
    ThreadArg* t = (ThreadArg*)a;
    tls_threadIdx = t->tid;
    t->fn(t->ctx);
    return nullptr;
}

// Runs `body()` once per CUDA thread of a grid x block launch.
template <class F>
void launch(dim3 grid, dim3 block, F body)
{
    const unsigned nthreads = block.x * block.y * block.z;
    g_gridDim = grid;
    g_blockDim = block;
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    std::vector<pthread_t> th(nthreads);
    std::vector<ThreadArg> args(nthreads);
    auto tramp = [](void* c) { (*(F*)c)(); };
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_blockIdx = dim3(bx, by, bz);
                pthread_barrier_init(&g_barrier, nullptr, nthreads);
                unsigned t = 0;
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx, ++t) {
                            args[t] = ThreadArg{tramp, &body, dim3(tx, ty, tz)};
                            if (pthread_create(&th[t], &attr, thread_main, &args[t]) != 0) {
                                std::fprintf(stderr, "emu: pthread_create failed\n");
                                std::abort();
                            }
                        }
                for (unsigned q = 0; q < nthreads; ++q) pthread_join(th[q], nullptr);
                pthread_barrier_destroy(&g_barrier);
                for (int id = 0; id < 16; ++id)
                    if (g_named_count[id]) {
                        pthread_barrier_destroy(&g_named[id]);
                        g_named_count[id] = 0;
                    }
            }
    pthread_attr_destroy(&attr);
}

}  // namespace emu
