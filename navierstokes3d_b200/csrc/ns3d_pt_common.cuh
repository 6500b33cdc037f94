// Device helpers shared by the kernels of the fused PT loop: the arithmetic of one cell update in
// the three modes, the x-face boundary values, and the mailbox protocol of the peer-memory halo
// exchange.  Free of CUDA runtime includes (also compiled by g++ for the host emulation, tests/emu/).
#pragma once

#include "ns3d_shared.cuh"

namespace {

enum { X_NEUMANN = 0, X_DIRICHLET = 1, X_HYDRO = 2 };

// a / b with y = RN(1/b): one multiply + two FMAs (Markstein's correction step).
__device__ __forceinline__ double div3(double a, double b, double y)
{
    const double q = a * y;
    const double r = fma(-b, q, a);
    return fma(r, y, q);
}

// ---- peer-memory halo protocol (device side) --------------------------------------------------
#ifdef NS3D_HOST_EMU  // host emulation of the kernels (tests/emu/): the same orderings with GCC atomics
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    emu::spin_pause();  // only ever polled in a spin loop: let the other ranks' threads run
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
}
__device__ __forceinline__ unsigned long long atom_add_acq_rel_gpu(unsigned long long* p, unsigned long long v)
{
    return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL);
}
__device__ __forceinline__ unsigned long long wall_ns() { return emu::wall_ns(); }
// work queue / completion counters of the persistent launch (ptv_flow_kernel)
__device__ __forceinline__ unsigned atom_add_relaxed_gpu(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
__device__ __forceinline__ void atom_add_release_gpu(unsigned* p, unsigned v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
__device__ __forceinline__ void fence_acq_rel_gpu() { __atomic_thread_fence(__ATOMIC_ACQ_REL); }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void spin_pause() { emu::spin_pause(); }
__device__ __forceinline__ unsigned ld_min3_relaxed(const unsigned* a, const unsigned* b, const unsigned* c)
{
    return min(min(ld_relaxed_gpu(a), ld_relaxed_gpu(b)), ld_relaxed_gpu(c));
}
#else
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long atom_add_acq_rel_gpu(unsigned long long* p, unsigned long long v)
{
    unsigned long long old;
    asm volatile("atom.add.acq_rel.gpu.global.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
    return old;
}
// work queue / completion counters of the persistent launch (ptv_flow_kernel)
__device__ __forceinline__ unsigned atom_add_relaxed_gpu(unsigned* p, unsigned v)
{
    unsigned old;
    asm volatile("atom.add.relaxed.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void atom_add_release_gpu(unsigned* p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Polling without side effects: an acquire load is followed by an invalidation of the SM's whole L1 (CCTL.IVALL), which a
// spinning thread repeats in every round -- ncu showed the other CTAs of the SM missing L1 on everything they had there.
// The counters are polled relaxed (served from L2), ONE fence follows the successful poll.
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// min(*a, *b, *c), polled relaxed (served from L2: no invalidation of the SM's L1 like after an acquire load)
__device__ __forceinline__ unsigned ld_min3_relaxed(const unsigned* a, const unsigned* b, const unsigned* c)
{
    return min(min(ld_relaxed_gpu(a), ld_relaxed_gpu(b)), ld_relaxed_gpu(c));
}
// what this thread has observed through the generic proxy is visible to the TMA copies (async proxy) it issues next
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void spin_pause() { __nanosleep(64); }
// nanoseconds of wall-clock time, independent of the SM clock (the spin limit must not depend on it)
__device__ __forceinline__ unsigned long long wall_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

#define NS3D_SPIN_LIMIT_NS 20000000000ULL  // 20 s of wall-clock time
#define NS3D_FLOW_WAIT_LOOKS 1000000u  // looks at the counters (~2 us each) before a work item of ptv_flow_kernel gives up: ~2 s

// Spins until the neighbour on `side` (0 lower, 1 upper) has finished the face work of every
// launch this rank has finished: then its stores into our halo plane have landed and it no
// longer reads the halo plane of its own that we are about to overwrite.  Bounded in wall-clock
// time: a neighbour that never answers raises NS3D_MB_ERROR (the solve then fails with NS3D_ECOMM)
// instead of hanging the GPU, and the caller gets `false` so that it can leave the neighbour's
// memory alone.
__device__ __forceinline__ bool wait_neighbour(unsigned long long* mbox, int side)
{
    const unsigned long long need = ld_acquire_sys(mbox + NS3D_MB_EPOCH_LO + side);
    if (ld_acquire_sys(mbox + NS3D_MB_FLAG_LO + side) >= need) return true;
    const unsigned long long t0 = wall_ns();
    while (ld_acquire_sys(mbox + NS3D_MB_FLAG_LO + side) < need) {
        if (wall_ns() - t0 > NS3D_SPIN_LIMIT_NS) {
            mbox[NS3D_MB_ERROR] = 1ULL + side;
            return false;
        }
    }
    return true;
}

// Last face CTA of this launch on `side`: close the epoch and tell the neighbour.
__device__ __forceinline__ void signal_neighbour(unsigned long long* mbox, int side, unsigned long long* peer_flag,
                                                 unsigned nface)
{
    const unsigned long long old = atom_add_acq_rel_gpu(mbox + NS3D_MB_ARRIVE_LO + side, 1ULL);
    if (old + 1 == nface) {
        mbox[NS3D_MB_ARRIVE_LO + side] = 0ULL;
        const unsigned long long e = mbox[NS3D_MB_EPOCH_LO + side] + 1ULL;
        mbox[NS3D_MB_EPOCH_LO + side] = e;
        __threadfence_system();
        st_release_sys(peer_flag, e);
    }
}

// After a chunk of launches: the halo planes of the current iterate are complete once both
// neighbours have signalled the epoch this rank has reached.
__global__ void pt_halo_wait_kernel(unsigned long long* mbox, int has_lo, int has_hi)
{
    if (threadIdx.x == 0) {
        if (has_lo) wait_neighbour(mbox, 0);
        if (has_hi) wait_neighbour(mbox, 1);
    }
}

// Hand-over between two solves (one thread): this rank has finished READING its ping-pong buffers
// (the copy-out of the previous solve precedes this kernel in stream order) and has WRITTEN the first
// iterate of the next one; it says so to both neighbours as one more epoch and waits for theirs.  Only
// then may a neighbour's first launch store into this rank's halo planes and read this rank's planes.
__global__ void pt_halo_barrier_kernel(unsigned long long* mbox, unsigned long long* peer_lo_flag,
                                       unsigned long long* peer_hi_flag)
{
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (peer_lo_flag) signal_neighbour(mbox, 0, peer_lo_flag, 1u);
        if (peer_hi_flag) signal_neighbour(mbox, 1, peer_hi_flag, 1u);
        if (peer_lo_flag) wait_neighbour(mbox, 0);
        if (peer_hi_flag) wait_neighbour(mbox, 1);
    }
}

}  // namespace
