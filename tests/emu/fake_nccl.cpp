// A fake libnccl.so.2 (test infrastructure): the nine NCCL entry points libns3d.so binds with
// dlopen, implemented over POSIX shared memory between the rank PROCESSES of an emulated multi-rank
// run (tests/test_emulated_multi_rank.py).  "Device" buffers are host memory there, so send/recv
// are copies through per-pair byte rings and the one-element all-reduce goes through shared slots
// and a barrier.  Semantics kept: operations of a group take effect at ncclGroupEnd, sends read
// their buffers before any receive of the same group writes (in-place halo exchange), messages
// between two ranks arrive in order.  Streams are ignored: the fake runtime executes in order.
//
//   g++ -O1 -std=c++17 -shared -fPIC fake_nccl.cpp -o _build/libnccl.so.2 -lrt -pthread
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;  // 1 uint8, 5 uint64, 8 float64 (fake_cuda/nccl.h)
typedef int ncclRedOp_t;     // 2 max, 3 min
typedef void* cudaStream_t;
}

namespace {

constexpr int MAX_RANKS = 8;
constexpr size_t RING = (size_t)8 << 20;  // bytes per ordered pair of ranks

struct Ring {
    volatile uint64_t head, tail;  // bytes written / read so far
    char data[RING];
};
struct Shared {
    volatile uint64_t arrived[2];  // two-phase barrier counters
    volatile uint64_t red[2][MAX_RANKS];
    Ring ring[MAX_RANKS][MAX_RANKS];  // [src][dst]
};

struct Op {
    bool send;
    void* buf;
    size_t bytes;
    int peer;
};
}  // namespace

struct ncclComm {
    int rank, nranks;
    Shared* sh;
    std::string name;
    uint64_t barrier_gen = 0, red_gen = 0;
    int group_depth = 0;
    std::vector<Op> ops;
};

namespace {

size_t type_size(ncclDataType_t t) { return t == 1 ? 1 : 8; }

void ring_write(Ring& r, const void* src, size_t n)
{
    const char* s = (const char*)src;
    while (n) {
        while (r.head - r.tail == RING) sched_yield();  // full: the receiver is behind
        const size_t pos = r.head % RING;
        const size_t chunk = std::min(std::min(n, RING - pos), (size_t)(RING - (r.head - r.tail)));
        std::memcpy(r.data + pos, s, chunk);
        __atomic_thread_fence(__ATOMIC_RELEASE);
        r.head += chunk;
        s += chunk;
        n -= chunk;
    }
}
void ring_read(Ring& r, void* dst, size_t n)
{
    char* d = (char*)dst;
    while (n) {
        while (r.head == r.tail) sched_yield();
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
        const size_t pos = r.tail % RING;
        const size_t chunk = std::min(std::min(n, RING - pos), (size_t)(r.head - r.tail));
        std::memcpy(d, r.data + pos, chunk);
        __atomic_thread_fence(__ATOMIC_RELEASE);
        r.tail += chunk;
        d += chunk;
        n -= chunk;
    }
}
void barrier(ncclComm* c)
{
    const int ph = (int)(c->barrier_gen & 1);
    const uint64_t target = (c->barrier_gen / 2 + 1) * (uint64_t)c->nranks;
    __atomic_fetch_add(&c->sh->arrived[ph], 1, __ATOMIC_ACQ_REL);
    while (__atomic_load_n(&c->sh->arrived[ph], __ATOMIC_ACQUIRE) < target) sched_yield();
    ++c->barrier_gen;
}
void flush(ncclComm* c)
{
    for (const Op& o : c->ops)  // all sends of the group first: they read before any receive writes
        if (o.send) {
            const uint64_t n = o.bytes;
            ring_write(c->sh->ring[c->rank][o.peer], &n, sizeof n);
            ring_write(c->sh->ring[c->rank][o.peer], o.buf, o.bytes);
        }
    for (const Op& o : c->ops)
        if (!o.send) {
            uint64_t n = 0;
            ring_read(c->sh->ring[o.peer][c->rank], &n, sizeof n);
            if (n != o.bytes) {
                std::fprintf(stderr, "fake nccl: rank %d expected %zu bytes from %d, got %llu\n", c->rank, o.bytes, o.peer,
                             (unsigned long long)n);
                std::abort();
            }
            ring_read(c->sh->ring[o.peer][c->rank], o.buf, o.bytes);
        }
    c->ops.clear();
}
thread_local ncclComm* g_group_comm = nullptr;
thread_local int g_group_depth = 0;

}  // namespace

extern "C" {

__attribute__((visibility("default"))) const char* ncclGetErrorString(ncclResult_t r) { return r == 0 ? "no error" : "fake nccl error"; }

__attribute__((visibility("default"))) ncclResult_t ncclGetUniqueId(ncclUniqueId* id)
{
    std::memset(id, 0, sizeof *id);
    std::snprintf(id->internal, sizeof id->internal, "/ns3d_nccl_%ld_%ld", (long)getpid(), (long)random());
    return 0;
}

__attribute__((visibility("default"))) ncclResult_t ncclCommInitRank(ncclComm_t* out, int nranks, ncclUniqueId id, int rank)
{
    if (nranks > MAX_RANKS || rank < 0 || rank >= nranks) return 3;
    ncclComm* c = new ncclComm();
    c->rank = rank;
    c->nranks = nranks;
    c->name = id.internal;
    // every rank may create it: O_CREAT without O_EXCL, a fresh segment is zero-filled
    const int fd = shm_open(c->name.c_str(), O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)sizeof(Shared)) != 0) return 3;
    c->sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
    close(fd);
    if (c->sh == MAP_FAILED) return 3;
    barrier(c);  // everybody is attached: the name can go
    if (rank == 0) shm_unlink(c->name.c_str());
    *out = c;
    return 0;
}

__attribute__((visibility("default"))) ncclResult_t ncclCommDestroy(ncclComm_t c)
{
    if (c) {
        munmap(c->sh, sizeof(Shared));
        delete c;
    }
    return 0;
}

__attribute__((visibility("default"))) ncclResult_t ncclGroupStart(void)
{
    ++g_group_depth;
    return 0;
}

__attribute__((visibility("default"))) ncclResult_t ncclGroupEnd(void)
{
    if (--g_group_depth == 0 && g_group_comm) {
        flush(g_group_comm);
        g_group_comm = nullptr;
    }
    return 0;
}

static ncclResult_t p2p(bool send, void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c)
{
    if (peer < 0 || peer >= c->nranks || peer == c->rank) return 3;
    c->ops.push_back(Op{send, buf, count * type_size(t), peer});
    if (g_group_depth == 0) flush(c);
    else g_group_comm = c;
    return 0;
}
__attribute__((visibility("default"))) ncclResult_t ncclSend(const void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t)
{
    return p2p(true, const_cast<void*>(buf), count, t, peer, c);
}
__attribute__((visibility("default"))) ncclResult_t ncclRecv(void* buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t)
{
    return p2p(false, buf, count, t, peer, c);
}

// one-element uint64 max / min is all the library reduces
__attribute__((visibility("default"))) ncclResult_t ncclAllReduce(const void* sendbuf, void* recvbuf, size_t count, ncclDataType_t t, ncclRedOp_t op,
                                                                 ncclComm_t c, cudaStream_t)
{
    if (count != 1 || t != 5 || (op != 2 && op != 3)) return 3;
    const int ph = (int)(c->red_gen & 1);
    c->sh->red[ph][c->rank] = *(const uint64_t*)sendbuf;
    barrier(c);
    uint64_t v = c->sh->red[ph][0];
    for (int r = 1; r < c->nranks; ++r) {
        const uint64_t w = c->sh->red[ph][r];
        v = op == 2 ? (w > v ? w : v) : (w < v ? w : v);
    }
    *(uint64_t*)recvbuf = v;
    barrier(c);  // nobody overwrites this phase's slots before everybody has read them
    ++c->red_gen;
    return 0;
}
}
