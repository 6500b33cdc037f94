// A fake CUDA runtime (test infrastructure): just enough of the API for libns3d.so's host code to
// run on the CPU, on top of the CUDA-on-host shim for its device code.  "Device" memory is host
// memory, streams execute immediately and in order, events are no-ops, stream capture RECORDS the
// launches and copies and a graph launch replays them -- so the library's ping-pong bookkeeping
// across captured chunks is exercised for real.  One device, no peers (IPC fails cleanly).
// tests/emu/build_lib.py compiles the library's translation units against this header after
// rewriting `kernel<<<grid, block, smem, stream>>>(args)` into emu::launch_on(...).
#pragma once

#include "../cuda_host_shim.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <functional>
#include <map>
#include <string>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
struct FakeStream {
    int id;
};
typedef FakeStream* cudaStream_t;
struct FakeEvent {
    int id;
};
typedef FakeEvent* cudaEvent_t;
typedef std::vector<std::function<void()>>* cudaGraph_t;
typedef std::vector<std::function<void()>>* cudaGraphExec_t;
struct cudaIpcMemHandle_t {
    char reserved[64];
};
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1 };
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrL2CacheSize = 38 };

namespace emu {
inline std::vector<std::function<void()>>* g_capture = nullptr;  // non-null while a stream capture is open
template <class F>
inline void enqueue(F f)
{
    if (g_capture) g_capture->push_back(f);
    else f();
}
// What build_lib.py turns a <<<>>> launch into.  `threads`: the kernel synchronises (barriers, shuffles).
template <class F>
inline void launch_on(dim3 grid, dim3 block, cudaStream_t, bool threads, F body)
{
    enqueue([=]() {
        if (threads) launch(grid, block, body);
        else launch_serial(grid, block, body);
    });
}
}  // namespace emu

enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake CUDA runtime: unsupported"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// one device per process unless NS3D_EMU_DEVICES says otherwise (rank processes of a "node" on which every rank sees all
// GPUs and selects its own by ordinal, as ImplicitGlobalGrid does); every ordinal is this process's own fake device
inline cudaError_t cudaGetDeviceCount(int* n) { const char* e = getenv("NS3D_EMU_DEVICES"); *n = e ? atoi(e) : 1; if (*n < 1) *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int)
{
    *v = a == cudaDevAttrMultiProcessorCount ? 148 : 126 * 1024 * 1024;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -5; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = new FakeStream{1}; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new FakeStream{2}; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new FakeEvent{0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
// Multi-rank emulation (ranks = processes, tests/test_emulated_multi_rank.py): with
// NS3D_EMU_SHARED_ARENA=1 "device" memory comes from a per-process POSIX shared-memory arena, so
// that cudaIpcGetMemHandle / cudaIpcOpenMemHandle can hand a peer process a mapping of it -- what
// CUDA IPC does between GPUs of one box.  A bump allocator: cudaFree is a no-op there.
namespace emu {
struct Arena {
    char* base = nullptr;
    size_t size = 0, used = 0;
    std::string name;
};
inline Arena& arena()
{
    static Arena a;
    if (!a.base && std::getenv("NS3D_EMU_SHARED_ARENA")) {
        a.size = (size_t)1 << 30;  // sparse: pages exist once touched
        a.name = "/ns3d_emu_" + std::to_string((long)getpid());
        const int fd = shm_open(a.name.c_str(), O_CREAT | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)a.size) != 0) {
            std::perror("emu arena");
            std::abort();
        }
        a.base = (char*)mmap(nullptr, a.size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
        close(fd);
        if (a.base == MAP_FAILED) {
            std::perror("emu arena mmap");
            std::abort();
        }
        std::atexit([]() { shm_unlink(arena().name.c_str()); });
    }
    return a;
}
struct IpcHandle {  // the 64 bytes of a cudaIpcMemHandle_t
    unsigned long long magic;
    long pid;
    unsigned long long offset;
};
inline std::map<long, char*>& peer_arenas()
{
    static std::map<long, char*> m;
    return m;
}
}  // namespace emu

inline cudaError_t cudaMalloc(void** p, size_t n)
{
    emu::Arena& a = emu::arena();
    if (a.base) {
        const size_t off = (a.used + 255) / 256 * 256;
        if (off + n > a.size) return cudaErrorMemoryAllocation;
        *p = a.base + off;
        a.used = off + n;
    } else {
        *p = std::aligned_alloc(256, (n + 255) / 256 * 256);
        if (!*p) return cudaErrorMemoryAllocation;
    }
    std::memset(*p, 0xff, n);  // NaN patterns: uninitialised device memory must not look like zeros
    return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p)
{
    if (!emu::arena().base) std::free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T>
inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t)
{
    emu::enqueue([=]() { std::memmove(d, s, n); });
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t)
{
    emu::enqueue([=]() { std::memset(d, v, n); });
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
// stream capture: record, then replay
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode)
{
    if (emu::g_capture) return cudaErrorInvalidValue;
    emu::g_capture = new std::vector<std::function<void()>>();
    return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g)
{
    *g = emu::g_capture;
    emu::g_capture = nullptr;
    return *g ? cudaSuccess : cudaErrorInvalidValue;
}
inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* st)
{
    *st = emu::g_capture ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
    return cudaSuccess;
}
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long)
{
    *e = new std::vector<std::function<void()>>(*g);
    return cudaSuccess;
}
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t)
{
    for (auto& f : *e) emu::enqueue(f);
    return cudaSuccess;
}
// peers exist only in shared-arena mode (see emu::arena)
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p)
{
    emu::Arena& a = emu::arena();
    if (!a.base || (char*)p < a.base || (char*)p >= a.base + a.size) return cudaErrorNotSupported;
    emu::IpcHandle ih{0x4e533344454d55ULL, (long)getpid(), (unsigned long long)((char*)p - a.base)};
    std::memset(h, 0, sizeof *h);
    std::memcpy(h, &ih, sizeof ih);
    return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned)
{
    emu::IpcHandle ih;
    std::memcpy(&ih, &h, sizeof ih);
    if (ih.magic != 0x4e533344454d55ULL) return cudaErrorInvalidValue;
    char*& base = emu::peer_arenas()[ih.pid];
    if (!base) {
        const std::string name = "/ns3d_emu_" + std::to_string(ih.pid);
        const int fd = shm_open(name.c_str(), O_RDWR, 0600);
        if (fd < 0) return cudaErrorInvalidValue;
        base = (char*)mmap(nullptr, (size_t)1 << 30, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
        close(fd);
        if (base == MAP_FAILED) {
            base = nullptr;
            return cudaErrorInvalidValue;
        }
    }
    *p = base + ih.offset;
    return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
