// Definitions shared by host code and device code of libns3d.so that need no CUDA runtime header
// (also compiled by g++ for the host emulation of the kernels, tests/emu/).
#pragma once

#include <cstddef>
#include <cstdint>

#include "../../include/ns3d.h"

// 0-based column-major index into an array with leading sizes (sx, sy).
__host__ __device__ __forceinline__ size_t idx3(int i, int j, int k, int sx, int sy)
{
    return (size_t)i + (size_t)sx * ((size_t)j + (size_t)sy * (size_t)k);
}

// mailbox words of the peer-memory halo protocol
enum {
    NS3D_MB_FLAG_LO = 0,    // written by the lower neighbour: epochs whose face work it has finished
    NS3D_MB_FLAG_HI = 1,    // same, upper neighbour
    NS3D_MB_ARRIVE_LO = 2,  // face CTAs of the running launch that are done (reset by the last one)
    NS3D_MB_ARRIVE_HI = 3,
    NS3D_MB_EPOCH_LO = 4,   // launches whose lower-face work is complete on this rank
    NS3D_MB_EPOCH_HI = 5,
    NS3D_MB_ERROR = 6,      // set when a spin-wait timed out
    NS3D_MB_WORDS = 16
};
