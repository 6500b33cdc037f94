// Host emulation of the fused PT kernels (test infrastructure, see cuda_host_shim.h): the device
// code of navierstokes3d_b200/csrc/ns3d_pt_kernels.cuh compiled by g++ and driven like
// run_direct() in ns3d_pt.cu drives it on one rank.  tests/test_kernel_emu.py compares the
// result bit for bit with the CPU oracle.
//
//   g++ -O1 -ffp-contract=off -std=c++17 -shared -fPIC -pthread pt_emu.cpp -o _build/libpt_emu.so
#include "cuda_host_shim.h"

#include <cstring>
#include <limits>

#include "../../navierstokes3d_b200/csrc/ns3d_pt_kernels.cuh"

namespace {

unsigned cdivu(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// Array copied into a block padded like ns3d_zeros pads (three x-y planes behind the array); the
// padding holds NaN so that a kernel USING a value it only prefetched is caught by the comparison.
struct Padded {
    std::vector<double> buf;
    size_t count;
    Padded(const double* src, size_t n, size_t plane) : buf(n + 3 * plane + 32, std::numeric_limits<double>::quiet_NaN()), count(n)
    {
        if (src) std::memcpy(buf.data(), src, n * sizeof(double));
    }
    double* p() { return buf.data(); }
};

template <int MODE>
void launch_iter(const PtK& k, const double* cur, double* nxt, double* dP, const double* divV)
{
    const dim3 grid(cdivu(k.nx - 2, 32), cdivu(k.ny - 2, 8), cdivu(k.kend - k.kbeg, k.zchunk));
    emu::launch(grid, dim3(32, 8, 1), [=]() { pt_iter_kernel<MODE, 4, false>(cur, nxt, dP, divV, k); });
}

template <int MODE, int TY>
void launch_tb2(int kernel, PtK k, const double* cur, double* nxt, const double* dpc, double* dpn, const double* divV)
{
    balance_chunks(k);
    const dim3 grid(cdivu(k.nx - 2, TB_X - 2), cdivu(k.ny - 2, TY - 2), cdivu(k.kend - k.kbeg, k.zchunk));
    if (kernel == 3) {  // two tile rows per thread: 32 x TY tiles of TY/2 thread rows
        tb2s_set_offsets(k, cur, nxt, dpc, dpn, divV);
        if (TY == 16 && k.nx == 255 && k.ny == 153)
            emu::launch(grid, dim3(TB_X, TY / 2, 1), [=]() { pt_tb2d_kernel<MODE, 16, 1, 2, 255, 153>(cur, nxt, dpc, dpn, divV, k); });
        else
            emu::launch(grid, dim3(TB_X, TY / 2, 1), [=]() { pt_tb2d_kernel<MODE, TY, 1, 2, 0, 0>(cur, nxt, dpc, dpn, divV, k); });
    } else if (kernel == 4) {  // pt_tb2s_kernel with pairwise row barriers
        tb2s_set_offsets(k, cur, nxt, dpc, dpn, divV);
        if (TY == 8 && k.nx == 255 && k.ny == 153)
            emu::launch(grid, dim3(TB_X, TY, 1), [=]() { pt_tb2s_kernel<MODE, 8, 1, true, 255, 153, true>(cur, nxt, dpc, dpn, divV, k); });
        else if (TY == 8)
            emu::launch(grid, dim3(TB_X, TY, 1), [=]() { pt_tb2s_kernel<MODE, 8, 1, true, 0, 0, true>(cur, nxt, dpc, dpn, divV, k); });
        else
            emu::launch(grid, dim3(TB_X, 16, 1), [=]() { pt_tb2s_kernel<MODE, 16, 1, true, 0, 0, true>(cur, nxt, dpc, dpn, divV, k); });
    } else if (kernel == 2) {
        tb2s_set_offsets(k, cur, nxt, dpc, dpn, divV);
        // like launch_tb2() in ns3d_pt.cu: grids with a compile-time-stride instantiation use it
        if (TY == 8 && k.nx == 255 && k.ny == 153)
            emu::launch(grid, dim3(TB_X, TY, 1), [=]() { pt_tb2s_kernel<MODE, 8, 1, true, 255, 153>(cur, nxt, dpc, dpn, divV, k); });
        else
            emu::launch(grid, dim3(TB_X, TY, 1), [=]() { pt_tb2s_kernel<MODE, TY, 1, true, 0, 0>(cur, nxt, dpc, dpn, divV, k); });
    } else {
        emu::launch(grid, dim3(TB_X, TY, 1), [=]() { pt_tb2_kernel<MODE, TY, false>(cur, nxt, dpc, dpn, divV, k); });
    }
}

template <int MODE>
int run(int kernel, int ty, PtK k, int serpentine, double* Pr, double* dP, const double* divV, int n_iter, int zchunk_iter,
        int zchunk_tb, long long* launches)
{
    const size_t sxy = (size_t)k.nx * k.ny, n = sxy * k.nz;
    const size_t dxy = (size_t)(k.nx - 2) * (k.ny - 2), nd = dxy * (k.nz - 2);
    Padded a(Pr, n, sxy), b(nullptr, n, sxy), da(dP, nd, dxy), db(nullptr, nd, dxy), dv(divV, n, sxy);
    double *cur = a.p(), *nxt = b.p(), *dpc = da.p(), *dpn = db.p();
    long long nl = 0;
    int q = 0;
    if (kernel != 0) {
        PtK k2 = k;
        k2.zchunk = zchunk_tb;
        for (; q + 2 <= n_iter; q += 2) {
            k2.reverse = serpentine && ((q >> 1) & 1);
            if (ty == 8) launch_tb2<MODE, 8>(kernel, k2, cur, nxt, dpc, dpn, dv.p());
            else if (ty == 32) launch_tb2<MODE, 32>(kernel, k2, cur, nxt, dpc, dpn, dv.p());
            else launch_tb2<MODE, 16>(kernel, k2, cur, nxt, dpc, dpn, dv.p());
            std::swap(cur, nxt);
            std::swap(dpc, dpn);
            ++nl;
        }
    }
    k.zchunk = zchunk_iter;
    for (; q < n_iter; ++q) {
        k.reverse = serpentine && (q & 1);
        launch_iter<MODE>(k, cur, nxt, dpc, dv.p());
        std::swap(cur, nxt);
        ++nl;
    }
    std::memcpy(Pr, cur, n * sizeof(double));
    std::memcpy(dP, dpc, nd * sizeof(double));
    if (launches) *launches = nl;
    return 0;
}

// Like the split launch of run_direct() on slabs: the chunks next to the z faces and the planes
// between them are updated by DIFFERENT launches (here: pt_tb2_kernel on [1,klo) and [khi,nz-1),
// `kernel_mid` on [klo,khi)), which must compose to the same iterate.
template <int MODE>
int run_split(int kernel_mid, int ty_mid, PtK k, double* Pr, double* dP, const double* divV, int n_pairs, int klo, int khi,
              int zchunk_mid)
{
    const size_t sxy = (size_t)k.nx * k.ny, n = sxy * k.nz;
    const size_t dxy = (size_t)(k.nx - 2) * (k.ny - 2), nd = dxy * (k.nz - 2);
    Padded a(Pr, n, sxy), b(nullptr, n, sxy), da(dP, nd, dxy), db(nullptr, nd, dxy), dv(divV, n, sxy);
    double *cur = a.p(), *nxt = b.p(), *dpc = da.p(), *dpn = db.p();
    for (int q = 0; q < n_pairs; ++q) {
        PtK lo = k, mid = k, hi = k;
        lo.kbeg = 1; lo.kend = klo; lo.zchunk = 8;
        hi.kbeg = khi; hi.kend = k.nz - 1; hi.zchunk = 8;
        mid.kbeg = klo; mid.kend = khi; mid.zchunk = zchunk_mid;
        mid.reverse = q & 1;
        launch_tb2<MODE, 16>(1, lo, cur, nxt, dpc, dpn, dv.p());
        launch_tb2<MODE, 16>(1, hi, cur, nxt, dpc, dpn, dv.p());
        if (ty_mid == 8) launch_tb2<MODE, 8>(kernel_mid, mid, cur, nxt, dpc, dpn, dv.p());
        else launch_tb2<MODE, 16>(kernel_mid, mid, cur, nxt, dpc, dpn, dv.p());
        std::swap(cur, nxt);
        std::swap(dpc, dpn);
    }
    std::memcpy(Pr, cur, n * sizeof(double));
    std::memcpy(dP, dpc, nd * sizeof(double));
    return 0;
}

// ---- one rank of a z-slab run: the peer-memory halo exchange fused into the kernels ----------------
// Every rank is a separate PROCESS; its fields, shadows and mailbox live in shared memory that the
// neighbours map too -- what CUDA IPC does for GPUs.  The launches below are the ones run_direct()
// and pt_iteration() issue on slabs (ns3d_pt.cu), minus streams and graphs: the emulation runs a
// rank's launches one after the other, while the ranks themselves run concurrently and meet only
// through the mailbox protocol.
struct EmuSlab {
    double* pr[2];            // this rank's Pr: user array, shadow
    double* dp[2];            // this rank's dPrdτ: user array, shadow
    const double* divV;
    PeerPtrs peers;           // neighbours' {Pr, Pr shadow, dPrdτ, dPrdτ shadow}, the three mailboxes
};

template <int MODE>
int run_slab(int kernel, PtK k, const EmuSlab& b, int n_iter, int kernel_mid, int ty_mid, int zchunk_tb, int zchunk_iter,
             int* which_pr, int* which_dp)
{
    int ip = 0, id = 0;  // index of the CURRENT Pr / dPrdτ buffer
    int q = 0;
    const int zf = 8;    // planes per face chunk of the split launch (run_direct)
    if (kernel != 0) {
        const bool split = (k.nz - 2) >= 2 * zf + 4;
        for (; q + 2 <= n_iter; q += 2) {
            PtK k2 = k;
            k2.zchunk = zchunk_tb;
            k2.reverse = (q >> 1) & 1;
            const double* cur = b.pr[ip];
            double* nxt = b.pr[1 - ip];
            const double* dpc = b.dp[id];
            double* dpn = b.dp[1 - id];
            if (split) {
                PtK f = k2, in = k2;
                f.faces = 1;
                f.zchunk = zf;
                ptk_set_peers(f, b.peers, 1 - ip, 2 + id);
                in.kbeg = 1 + zf;
                in.kend = k.nz - 1 - zf;
                emu::launch(dim3(cdivu(k.nx - 2, TB_X - 2), cdivu(k.ny - 2, 16 - 2), 2), dim3(TB_X, 16, 1),
                            [=]() { pt_tb2_kernel<MODE, 16, true>(cur, nxt, dpc, dpn, b.divV, f); });
                if (ty_mid == 8) launch_tb2<MODE, 8>(kernel_mid, in, cur, nxt, dpc, dpn, b.divV);
                else launch_tb2<MODE, 16>(kernel_mid, in, cur, nxt, dpc, dpn, b.divV);
            } else {
                PtK u = k2;
                balance_chunks(u);
                ptk_set_peers(u, b.peers, 1 - ip, 2 + id);
                emu::launch(dim3(cdivu(k.nx - 2, TB_X - 2), cdivu(k.ny - 2, 16 - 2), cdivu(u.kend - u.kbeg, u.zchunk)),
                            dim3(TB_X, 16, 1), [=]() { pt_tb2_kernel<MODE, 16, true>(cur, nxt, dpc, dpn, b.divV, u); });
            }
            ip = 1 - ip;
            id = 1 - id;
        }
    }
    for (; q < n_iter; ++q) {  // pt_iteration(): one unsplit launch of the peer-store instantiation
        PtK u = k;
        u.zchunk = zchunk_iter;
        u.reverse = q & 1;
        balance_chunks(u);
        ptk_set_peers(u, b.peers, 1 - ip, -1);
        const double* cur = b.pr[ip];
        double* nxt = b.pr[1 - ip];
        double* dP = b.dp[id];
        emu::launch(dim3(cdivu(k.nx - 2, 32), cdivu(k.ny - 2, 8), cdivu(u.kend - u.kbeg, u.zchunk)), dim3(32, 8, 1),
                    [=]() { pt_iter_kernel<MODE, 4, true>(cur, nxt, dP, b.divV, u); });
        ip = 1 - ip;
    }
    // peer_join(): the halos of the current iterate are complete once both neighbours have caught up
    if (k.zlo_halo) wait_neighbour(b.peers.mbox, 0);
    if (k.zhi_halo) wait_neighbour(b.peers.mbox, 1);
    *which_pr = ip;
    *which_dp = id;
    return b.peers.mbox[NS3D_MB_ERROR] ? -2 : 0;
}

}  // namespace

// One rank of an N-rank z-slab run (call it from N processes at once).  All buffers are padded like
// ns3d_zeros pads; *which_pr / *which_dp tell which of the two buffers hold the result.
extern "C" int emu_pt_slab_iterate(int kernel, int mode, const ns3d_pt_params* pp, int rank, int nranks, const EmuSlab* b,
                                   int n_iter, int kernel_mid, int ty_mid, int* which_pr, int* which_dp)
{
    if (!pp || !b || nranks < 2 || pp->nz < 6) return -1;
    PtK k;
    std::memset(&k, 0, sizeof k);
    ptk_fill(pp, &k);
    k.zlo_halo = rank > 0;
    k.zhi_halo = rank < nranks - 1;
    const int zc_iter = pp->zchunk > 0 ? pp->zchunk : 8;
    const int zc_tb = pp->zchunk > 0 ? pp->zchunk : 12;
    switch (mode) {
        case NS3D_PARITY: return run_slab<NS3D_PARITY>(kernel, k, *b, n_iter, kernel_mid, ty_mid, zc_tb, zc_iter, which_pr, which_dp);
        case NS3D_FAST: return run_slab<NS3D_FAST>(kernel, k, *b, n_iter, kernel_mid, ty_mid, zc_tb, zc_iter, which_pr, which_dp);
        case NS3D_FASTEST: return run_slab<NS3D_FASTEST>(kernel, k, *b, n_iter, kernel_mid, ty_mid, zc_tb, zc_iter, which_pr, which_dp);
    }
    return -1;
}

extern "C" int emu_pt_tb2_split(int kernel_mid, int mode, int ty_mid, const ns3d_pt_params* pp, double* Pr, double* dP,
                                const double* divV, int n_pairs, int klo, int khi, int zchunk_mid)
{
    if (!pp || pp->nx < 3 || pp->ny < 3 || pp->nz < 3 || klo < 2 || khi <= klo || khi > pp->nz - 2) return -1;
    PtK k;
    std::memset(&k, 0, sizeof k);
    ptk_fill(pp, &k);
    switch (mode) {
        case NS3D_PARITY: return run_split<NS3D_PARITY>(kernel_mid, ty_mid, k, Pr, dP, divV, n_pairs, klo, khi, zchunk_mid);
        case NS3D_FAST: return run_split<NS3D_FAST>(kernel_mid, ty_mid, k, Pr, dP, divV, n_pairs, klo, khi, zchunk_mid);
        case NS3D_FASTEST: return run_split<NS3D_FASTEST>(kernel_mid, ty_mid, k, Pr, dP, divV, n_pairs, klo, khi, zchunk_mid);
    }
    return -1;
}

// kernel: 0 = pt_iter_kernel, 1 = pt_tb2_kernel (+ pt_iter_kernel for an odd tail), 2 = pt_tb2s_kernel (+ tail),
// 3 = pt_tb2d_kernel (+ tail), 4 = pt_tb2s_kernel with pairwise row barriers (tile heights 8 and 16).
// zlo_halo / zhi_halo mark z faces that are slab interfaces (left to the halo exchange).
extern "C" int emu_pt_iterate(int kernel, int mode, int ty, const ns3d_pt_params* pp, int zlo_halo, int zhi_halo,
                              int serpentine, double* Pr, double* dP, const double* divV, int n_iter, long long* launches)
{
    if (!pp || pp->nx < 3 || pp->ny < 3 || pp->nz < 3) return -1;
    PtK k;
    std::memset(&k, 0, sizeof k);
    ptk_fill(pp, &k);
    k.zlo_halo = zlo_halo;
    k.zhi_halo = zhi_halo;
    const int zc_iter = pp->zchunk > 0 ? pp->zchunk : 8;
    const int zc_tb = pp->zchunk > 0 ? pp->zchunk : 16;
    switch (mode) {
        case NS3D_PARITY: return run<NS3D_PARITY>(kernel, ty, k, serpentine, Pr, dP, divV, n_iter, zc_iter, zc_tb, launches);
        case NS3D_FAST: return run<NS3D_FAST>(kernel, ty, k, serpentine, Pr, dP, divV, n_iter, zc_iter, zc_tb, launches);
        case NS3D_FASTEST: return run<NS3D_FASTEST>(kernel, ty, k, serpentine, Pr, dP, divV, n_iter, zc_iter, zc_tb, launches);
    }
    return -1;
}
