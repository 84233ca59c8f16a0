// gwz_handles.h -- handle layouts and host-side helpers shared by the C-ABI translation units
// (gwz_api.cu: HubbardReal1D + GutzwillerAnsatz fast path; gwz_api_generic.cu: ansatz-generic engine).
#pragma once
#include <atomic>
#include <string>
#include <vector>

#include "gutzwiller_b200.h"
#include "gwz_internal.h"
#include "gwz_nccl.h"

namespace gwz {
int api_fail(int code, const std::string& msg);   // records the thread-local message, returns code
void api_count_launch();                           // every kernel launch of the library is counted
int api_check_err_flag(gwz_context* c);            // "Infinite residence time" (qmc.jl:34-36)
void binomial_table(std::vector<uint64_t>& tab);   // [65][65], saturating
// symmetric ranked sweeps (gwz_api.cu): the rank sub-range that can hold orbit representatives, and their stored basis
struct RepRange {
    uint64_t lo, count;      // ranks [lo, lo + count) are visited
    uint64_t extra_state;    // the uniform state at commensurate filling, evaluated with weight 1 ...
    int has_extra;           // ... if its rank lies in the caller's range
};
RepRange representative_range(const std::vector<uint64_t>& binom, int N, int M, int B, uint64_t rank_begin, uint64_t dim);
}  // namespace gwz
struct gwz_context;
namespace gwz {
int build_representative_basis(gwz_context* c, int N, int M, int B, const uint64_t* d_binom, const RepRange& r,
                               uint64_t** d_rep, uint32_t** d_mult, uint64_t* nrep, int* stored);
}  // namespace gwz

#define GWZ_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return gwz::api_fail(GWZ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
#define GWZ_NCCL(api, expr)                                                                         \
    do {                                                                                            \
        int _r = (expr);                                                                            \
        if (_r != gwz::kNcclSuccess)                                                                \
            return gwz::api_fail(GWZ_ERR_NCCL, std::string(#expr) + ": " + (api)->GetErrorString(_r)); \
    } while (0)
#define GWZ_REQUIRE(cond, msg)                                       \
    do {                                                             \
        if (!(cond)) return gwz::api_fail(GWZ_ERR_INVALID, (msg));   \
    } while (0)
#define GWZ_LAUNCH(expr)         \
    do {                         \
        gwz::api_count_launch(); \
        GWZ_CUDA(expr);          \
    } while (0)

constexpr int GROUP_ROWS = 64;

struct gwz_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    int clock_khz = 0;
    gwz::ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    double* d_acc = nullptr;    // NUM_ACC doubles: send/recv buffer of the all-reduce
    double* d_group = nullptr;  // GROUP_ROWS x NUM_ACC scratch
    double* h_acc = nullptr;    // pinned
    int* d_err = nullptr;
    int* h_err = nullptr;  // pinned
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

struct gwz_model {
    gwz_context* ctx;
    int N, M, B, W, nw;
    double u, t;
    double v = 0.0;  // ExtendedHubbardReal1D nearest-neighbour interaction (generic engine only)
    gwz::DevModel dm;
};

static inline int use_device(const gwz_context* ctx) {
    GWZ_CUDA(cudaSetDevice(ctx->device));
    return GWZ_OK;
}

// releases a half-built handle when a create function fails midway (the destroy functions accept partially
// initialised objects: cudaFree(nullptr) is a no-op)
template <typename H, int (*DESTROY)(H*)>
struct HandleGuard {
    H* h;
    explicit HandleGuard(H* p) : h(p) {}
    HandleGuard(const HandleGuard&) = delete;
    HandleGuard& operator=(const HandleGuard&) = delete;
    ~HandleGuard() {
        if (h) DESTROY(h);
    }
    H* release() {
        H* t = h;
        h = nullptr;
        return t;
    }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * (n ? n : 1)); }
};
