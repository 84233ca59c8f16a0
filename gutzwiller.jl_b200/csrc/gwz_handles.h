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
enum HandleTag : uint32_t {
    GWZ_TAG_CONTEXT = 0x477a0001u, GWZ_TAG_MODEL, GWZ_TAG_WALKERS, GWZ_TAG_SWEEP, GWZ_TAG_ANSATZ, GWZ_TAG_GWALKERS,
    GWZ_TAG_GSWEEP, GWZ_TAG_VANSATZ, GWZ_TAG_VWALKERS, GWZ_TAG_SPARSE, GWZ_TAG_SWALKERS
};
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
// every opaque handle starts with a type tag: passing a handle of another type (or a destroyed one) is an error
// message, not a read through the wrong struct layout
#define GWZ_HANDLE(h, TAG, msg)                                                         \
    do {                                                                                \
        if ((h) == nullptr || (h)->magic != (TAG)) return gwz::api_fail(GWZ_ERR_INVALID, (msg)); \
    } while (0)
#define GWZ_LAUNCH(expr)         \
    do {                         \
        gwz::api_count_launch(); \
        GWZ_CUDA(expr);          \
    } while (0)

constexpr int GROUP_ROWS = 64;
constexpr int NUM_VEC = gwz::NUM_ACC + gwz::NUM_BACC;  // sampling accumulators followed by the blocking accumulators

struct gwz_context {
    uint32_t magic = gwz::GWZ_TAG_CONTEXT;  // handle type tag, checked by every entry point
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    int clock_khz = 0;
    gwz::ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    double* d_acc = nullptr;    // NUM_VEC doubles: send/recv buffer of the all-reduce
    double* d_group = nullptr;  // GROUP_ROWS x NUM_VEC scratch
    double* h_acc = nullptr;    // pinned
    int* d_err = nullptr;
    int* h_err = nullptr;  // pinned
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // sampling kernel(s) of the last run
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;  // blocking kernel of the last analysis
};

struct gwz_model {
    uint32_t magic = gwz::GWZ_TAG_MODEL;  // handle type tag, checked by every entry point
    gwz_context* ctx;
    int N, M, B, W, nw;
    double u, t;
    double v = 0.0;  // ExtendedHubbardReal1D nearest-neighbour interaction (generic engine only)
    gwz::DevModel dm;
};

struct gwz_walkers {
    uint32_t magic = gwz::GWZ_TAG_WALKERS;  // handle type tag, checked by every entry point
    int device = 0;  // destroy must not dereference the parent handles: a host GC may release them in any order
    gwz_model* m;
    size_t n;
    uint64_t offset, seed;
    uint64_t steps_done;  // Philox step counter
    uint64_t acc_steps;   // steps currently held in the estimators (per walker)
    uint64_t* d_addr = nullptr;
    double* d_acc = nullptr;
    double* d_partials = nullptr;
    int partials_cap = 0;
    double* d_vec = nullptr;  // NUM_VEC doubles: this set's reduced vector (sampling + blocking accumulators)
    double* d_shift = nullptr;  // scratch for the reference-point kernel (first estimator only)
    double* h_shift = nullptr;  // pinned
    double shift_e = 0.0, shift_g = 0.0;  // (E_ref, G_ref): reference point of the current estimator's sums
    double next_e = 0.0, next_g = 0.0;    // pooled (E, G) of the last run: reference point of the next estimator
    bool have_next = false;
    // Where the walkers currently live: unary bitstrings (d_addr) or occupation bit-planes (d_planes).
    // The production kernel works on planes; everything that needs addresses converts back first.
    uint32_t* d_planes = nullptr;
    uint64_t* d_ref_addr = nullptr;  // W words: unary copy of walker 0 for the reference-point kernel
    bool in_planes = false;
    // (tau, E_loc) series of the accumulated steps, kept in HBM by runs with GWZ_RUN_SERIES (step-major
    // [step][walker]); ser_steps == acc_steps whenever the series covers exactly what the running sums hold
    double* d_ser_tau = nullptr;
    double* d_ser_eloc = nullptr;
    uint64_t ser_cap = 0, ser_steps = 0;
    double* d_ser_v = nullptr;  // scratch_cap doubles: scratch rows of the blocking kernel (one per resident thread)
    uint64_t scratch_cap = 0;
    double* d_blk = nullptr;       // [5][n] per-walker blocking results of the last analysis
    double* d_bpartials = nullptr;  // [bgrid_cap][NUM_BACC]
    int bpartials_cap = 0;
    bool blk_valid = false;
    uint64_t blk_len = 0;  // resample length of the last analysis
    int grid_key = 0, grid_cached = 0;  // persistent grid of the kernel used last (occupancy query cached)
    bool vec_in_ctx = false;            // the last run reduced its sampling sums straight into ctx->d_acc
    bool vec_mirrored = false;          // ... and the reduction kernel also wrote them to the pinned host mirror
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
