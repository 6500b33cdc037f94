// Types of <nccl.h> that ns3d_core.cu names (test infrastructure).  The library binds NCCL's
// functions at run time with dlopen and only when a communicator with more than one rank is
// attached, which the emulated single-"device" library never does.
#pragma once
#include <cuda_runtime.h>
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclInternalError = 3 } ncclResult_t;
typedef enum { ncclUint8 = 1, ncclUint64 = 5, ncclFloat64 = 8, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
