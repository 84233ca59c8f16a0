// gwz_nccl.h -- NCCL entry points resolved at run time (dlopen), so that libgutzwiller_b200.so
// loads on machines without NCCL and shares the process's already-loaded libnccl.so.2 (e.g. the
// one PyTorch ships) instead of pulling in a second copy.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace gwz {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { kNcclSuccess = 0, kNcclSum = 0, kNcclFloat64 = 8 };

struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId*);
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    int (*CommInitAll)(ncclComm_t*, int, const int*);
    int (*CommDestroy)(ncclComm_t);
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char* (*GetErrorString)(int);
    int (*GetVersion)(int*);
};

// nullptr if libnccl cannot be loaded; *why receives a message
const NcclApi* nccl_api(const char** why);

}  // namespace gwz
