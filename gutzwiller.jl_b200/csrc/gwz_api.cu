// gwz_api.cu -- the C ABI of libgutzwiller_b200.so (see include/gutzwiller_b200.h).
//
// Host-side orchestration only: handles, HBM buffers, kernel launches on the context's stream,
// the NCCL all-reduce of the accumulator vector, and the (host, O(parameters)) AMSGrad update.
// There is no CPU implementation of any sampled or swept quantity in this file.
#include "gutzwiller_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "gwz_handles.h"
#include "gwz_blocking.cuh"

using namespace gwz;

// ------------------------------------------------------------------------------------------
// errors, launch counter (shared with gwz_api_generic.cu through gwz_handles.h)
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
int gwz::api_fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
static int fail(int code, const std::string& msg) { return gwz::api_fail(code, msg); }

// number of kernels this library has launched (every launch goes through GWZ_LAUNCH)
static std::atomic<uint64_t> g_launches{0};
void gwz::api_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int gwz_launch_count(uint64_t* n) {
    if (!n) return fail(GWZ_ERR_INVALID, "gwz_launch_count: NULL argument");
    *n = g_launches.load(std::memory_order_relaxed);
    return GWZ_OK;
}

extern "C" const char* gwz_last_error(void) { return g_last_error.c_str(); }
extern "C" int gwz_version(void) { return GWZ_VERSION; }

static bool planes_enabled() {
    const char* e = std::getenv("GWZ_PLANES");
    return !(e && e[0] == '0');
}
// GWZ_INCR=0 falls back from the incremental class-count kernel to the full re-count per step
// (same plane format, bitwise identical results; kept for A/B measurements and cross-checks)
static bool incr_enabled() {
    const char* e = std::getenv("GWZ_INCR");
    return !(e && e[0] == '0');
}

// GWZ_FAST=0 falls back from the round-2 production kernel (gwz_walker_fast.cu: integer hop counts per rate class)
// to the round-1 incremental kernel (gwz_walker_incr.cu); both sample the same distribution with different u -> hop maps
static bool fast_enabled() {
    const char* e = std::getenv("GWZ_FAST");
    return !(e && e[0] == '0');
}

// make d_addr (unary) the valid copy
static int to_unary(gwz_walkers* w) {
    if (!w->in_planes) return GWZ_OK;
    gwz_context* c = w->m->ctx;
    GWZ_LAUNCH(launch_planes_to_unary(w->d_planes, w->n, w->m->dm, 0, w->n, w->d_addr, c->stream));
    w->in_planes = false;
    return GWZ_OK;
}
// make d_planes the valid copy
static int to_planes(gwz_walkers* w) {
    if (w->in_planes) return GWZ_OK;
    gwz_context* c = w->m->ctx;
    if (!w->d_planes) {
        int nwm, np;
        planes_shape(w->m->N, w->m->M, &nwm, &np);
        GWZ_CUDA(cudaMalloc(&w->d_planes, sizeof(uint32_t) * (size_t)nwm * np * w->n));
        GWZ_CUDA(cudaMalloc(&w->d_ref_addr, sizeof(uint64_t) * w->m->W));
    }
    GWZ_LAUNCH(launch_unary_to_planes(w->d_addr, w->n, w->m->dm, w->d_planes, c->stream));
    w->in_planes = true;
    return GWZ_OK;
}

struct gwz_sweep {
    uint32_t magic = gwz::GWZ_TAG_SWEEP;  // handle type tag, checked by every entry point
    int device = 0;
    gwz_model* m;
    uint64_t* d_basis = nullptr;  // SoA [W][dim], null => ranked
    uint64_t dim = 0, rank_begin = 0;
    uint64_t* d_binom = nullptr;
    double* d_partials = nullptr;
    int grid = 0;
    uint64_t chunk = 1;
    float last_ms = 0.f;
    // representative basis of a symmetric ranked sweep (built at the second sweep of a handle, reused afterwards)
    uint64_t* d_rep = nullptr;
    uint32_t* d_mult = nullptr;
    uint64_t nrep = 0;
    int rep_state = 0;                 // 0 not built, 1 built, -1 not cacheable (too large / disabled)
    uint64_t sym_calls = 0, build_at = 3;  // build_at: the sweep at which storing the representatives has paid off
    float t_enum = 0.f;                    // kernel time of an enumerating sweep (faster of the first two)
    bool ms_pending = false;           // ev0/ev1 hold the last sweep's kernel time, not read yet
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<uint64_t> h_binom;     // host copy of the binomial table (ranked mode)
    unsigned* d_done = nullptr;        // arrival counter of the fused last-CTA reduction
    int enum_grid = 0, rep_grid = 0;   // launch shapes of the enumerating / stored-representative sweeps (cached)
    uint64_t enum_chunk = 1, rep_chunk = 1;
};


// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
extern "C" int gwz_context_create(int device, gwz_context** out) {
    GWZ_REQUIRE(out != nullptr, "gwz_context_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(GWZ_ERR_CUDA, std::string("no CUDA device available (this library has no CPU path): ") +
                                      cudaGetErrorString(e));
    GWZ_REQUIRE(device >= 0 && device < count, "gwz_context_create: device index out of range");
    gwz_context* c = new gwz_context();
    c->device = device;
    GWZ_CUDA(cudaSetDevice(device));
    GWZ_CUDA(cudaGetDeviceProperties(&c->prop, device));
    GWZ_CUDA(cudaDeviceGetAttribute(&c->clock_khz, cudaDevAttrClockRate, device));
    GWZ_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    // the error flag sits right behind the accumulator vector, on the device and in the pinned host mirror: a run
    // fetches both with one copy
    GWZ_CUDA(cudaMalloc(&c->d_acc, sizeof(double) * (NUM_VEC + 1)));
    GWZ_CUDA(cudaMalloc(&c->d_group, sizeof(double) * NUM_VEC * GROUP_ROWS));
    GWZ_CUDA(cudaMallocHost(&c->h_acc, sizeof(double) * (NUM_VEC + 1)));
    c->d_err = reinterpret_cast<int*>(c->d_acc + NUM_VEC);
    c->h_err = reinterpret_cast<int*>(c->h_acc + NUM_VEC);
    std::memset(c->h_acc, 0, sizeof(double) * (NUM_VEC + 1));
    GWZ_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(double) * (NUM_VEC + 1), c->stream));
    GWZ_CUDA(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
    GWZ_CUDA(cudaEventCreate(&c->ev0));
    GWZ_CUDA(cudaEventCreate(&c->ev1));
    GWZ_CUDA(cudaEventCreate(&c->ev2));
    GWZ_CUDA(cudaEventCreate(&c->ev3));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
    return GWZ_OK;
}

extern "C" int gwz_context_destroy(gwz_context* c) {
    if (c && c->magic != gwz::GWZ_TAG_CONTEXT) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_context_destroy: not a gwz_context handle");
    if (!c) return GWZ_OK;
    cudaSetDevice(c->device);
    if (c->comm) {
        const NcclApi* api = nccl_api(nullptr);
        if (api) api->CommDestroy(c->comm);
    }
    cudaFree(c->d_acc);
    cudaFree(c->d_group);
    cudaFreeHost(c->h_acc);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    if (c->stream) cudaStreamDestroy(c->stream);
    c->magic = 0;
    delete c;
    return GWZ_OK;
}

extern "C" int gwz_context_info(const gwz_context* c, int* device, int* sm_count, int* sm_clock_khz, int* cc_major,
                                int* cc_minor, size_t* global_mem_bytes) {
    GWZ_HANDLE(c, gwz::GWZ_TAG_CONTEXT, "gwz_context_info: NULL or foreign handle (expected a gwz_context)");
    GWZ_REQUIRE(c != nullptr, "gwz_context_info: ctx is NULL");
    if (device) *device = c->device;
    if (sm_count) *sm_count = c->prop.multiProcessorCount;
    if (sm_clock_khz) *sm_clock_khz = c->clock_khz;
    if (cc_major) *cc_major = c->prop.major;
    if (cc_minor) *cc_minor = c->prop.minor;
    if (global_mem_bytes) *global_mem_bytes = c->prop.totalGlobalMem;
    return GWZ_OK;
}

extern "C" int gwz_context_stream(const gwz_context* c, void** cuda_stream) {
    GWZ_HANDLE(c, gwz::GWZ_TAG_CONTEXT, "gwz_context_stream: NULL or foreign handle (expected a gwz_context)");
    GWZ_REQUIRE(c && cuda_stream, "gwz_context_stream: NULL argument");
    *cuda_stream = (void*)c->stream;
    return GWZ_OK;
}

extern "C" int gwz_context_synchronize(gwz_context* c) {
    GWZ_HANDLE(c, gwz::GWZ_TAG_CONTEXT, "gwz_context_synchronize: NULL or foreign handle (expected a gwz_context)");
    GWZ_REQUIRE(c != nullptr, "gwz_context_synchronize: ctx is NULL");
    if (int rc = use_device(c)) return rc;
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}

extern "C" int gwz_nccl_unique_id(uint8_t id[GWZ_NCCL_ID_BYTES]) {
    GWZ_REQUIRE(id != nullptr, "gwz_nccl_unique_id: id is NULL");
    const char* why = nullptr;
    const NcclApi* api = nccl_api(&why);
    if (!api) return fail(GWZ_ERR_NCCL, why ? why : "NCCL unavailable");
    ncclUniqueId uid;
    static_assert(sizeof(uid) == GWZ_NCCL_ID_BYTES, "ncclUniqueId size");
    GWZ_NCCL(api, api->GetUniqueId(&uid));
    std::memcpy(id, &uid, sizeof(uid));
    return GWZ_OK;
}

extern "C" int gwz_context_comm_init(gwz_context* c, int world_size, int rank, const uint8_t id[GWZ_NCCL_ID_BYTES]) {
    GWZ_HANDLE(c, gwz::GWZ_TAG_CONTEXT, "gwz_context_comm_init: NULL or foreign handle (expected a gwz_context)");
    GWZ_REQUIRE(c && id, "gwz_context_comm_init: NULL argument");
    GWZ_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "gwz_context_comm_init: bad rank/world_size");
    GWZ_REQUIRE(c->comm == nullptr, "gwz_context_comm_init: communicator already initialised");
    const char* why = nullptr;
    const NcclApi* api = nccl_api(&why);
    if (!api) return fail(GWZ_ERR_NCCL, why ? why : "NCCL unavailable");
    if (int rc = use_device(c)) return rc;
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    GWZ_NCCL(api, api->CommInitRank(&c->comm, world_size, uid, rank));
    c->world = world_size;
    c->rank = rank;
    return GWZ_OK;
}

extern "C" int gwz_context_comm_init_all(gwz_context* const* ctxs, int n) {
    GWZ_REQUIRE(ctxs && n >= 1, "gwz_context_comm_init_all: bad arguments");
    const char* why = nullptr;
    const NcclApi* api = nccl_api(&why);
    if (!api) return fail(GWZ_ERR_NCCL, why ? why : "NCCL unavailable");
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) {
        GWZ_HANDLE(ctxs[i], gwz::GWZ_TAG_CONTEXT, "gwz_context_comm_init_all: NULL or foreign handle (expected gwz_context)");
        GWZ_REQUIRE(ctxs[i]->comm == nullptr, "gwz_context_comm_init_all: already joined context");
        devs[i] = ctxs[i]->device;
    }
    GWZ_NCCL(api, api->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        ctxs[i]->comm = comms[i];
        ctxs[i]->world = n;
        ctxs[i]->rank = i;
    }
    return GWZ_OK;
}

extern "C" int gwz_context_comm_info(const gwz_context* c, int* world_size, int* rank) {
    GWZ_HANDLE(c, gwz::GWZ_TAG_CONTEXT, "gwz_context_comm_info: NULL or foreign handle (expected a gwz_context)");
    GWZ_REQUIRE(c != nullptr, "gwz_context_comm_info: ctx is NULL");
    if (world_size) *world_size = c->world;
    if (rank) *rank = c->rank;
    return GWZ_OK;
}

// ------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------
extern "C" int gwz_model_create_hubbard1d_gutzwiller(gwz_context* ctx, int N, int M, double u, double t,
                                                     gwz_model** out) {
    GWZ_HANDLE(ctx, gwz::GWZ_TAG_CONTEXT, "gwz_model_create_hubbard1d_gutzwiller: NULL or foreign handle (expected a gwz_context)");
    GWZ_REQUIRE(ctx && out, "gwz_model_create: NULL argument");
    GWZ_REQUIRE(N >= 1 && M >= 2, "gwz_model_create: need N >= 1 particles and M >= 2 modes");
    const int B = N + M - 1;
    GWZ_REQUIRE(B + 1 + KCAP_MAX <= 32 * MAXNW, "gwz_model_create: N+M-1 must be <= 250 bits");
    gwz_model* m = new gwz_model();
    m->ctx = ctx;
    m->N = N;
    m->M = M;
    m->B = B;
    m->W = (B + 63) / 64;
    m->nw = (B + 1 + KCAP_MAX + 31) / 32;
    m->u = u;
    m->t = t;
    DevModel& dm = m->dm;
    std::memset(&dm, 0, sizeof(dm));
    dm.N = N;
    dm.M = M;
    dm.B = B;
    dm.W = m->W;
    dm.ext_ws = (B + 1) >> 5;
    dm.ext_bs = (B + 1) & 31;
    dm.u = u;
    dm.t = t;
    for (int p = 0; p < M; ++p) dm.mmask[p >> 5] |= 1u << (p & 31);
    dm.mtopmask[(M - 1) >> 5] = 0xFFFFFFFFu;
    dm.mtop_b = (M - 1) & 31;
    dm.extmask_lo[dm.ext_ws] = 0xFFFFFFFFu;
    if (dm.ext_bs != 0 && dm.ext_ws + 1 < MAXNW) dm.extmask_hi[dm.ext_ws + 1] = 0xFFFFFFFFu;
    for (int p = 0; p <= B; ++p) dm.zmask[p >> 5] |= 1u << (p & 31);
    for (int p = 0; p < B; ++p) dm.amask[p >> 5] |= 1u << (p & 31);
    *out = m;
    return GWZ_OK;
}
extern "C" int gwz_model_destroy(gwz_model* m) {
    if (m && m->magic != gwz::GWZ_TAG_MODEL) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_model_destroy: not a gwz_model handle");
    m->magic = 0;
    delete m;
    return GWZ_OK;
}
extern "C" int gwz_model_info(const gwz_model* m, int* N, int* M, double* u, double* t, int* bits, int* words) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_model_info: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m != nullptr, "gwz_model_info: model is NULL");
    if (N) *N = m->N;
    if (M) *M = m->M;
    if (u) *u = m->u;
    if (t) *t = m->t;
    if (bits) *bits = m->B;
    if (words) *words = m->W;
    return GWZ_OK;
}
extern "C" int gwz_model_from_onr(const gwz_model* m, const int32_t* onr, uint64_t* a) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_model_from_onr: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && onr && a, "gwz_model_from_onr: NULL argument");
    int total = 0;
    for (int i = 0; i < m->M; ++i) {
        GWZ_REQUIRE(onr[i] >= 0, "gwz_model_from_onr: negative occupation");
        total += onr[i];
    }
    GWZ_REQUIRE(total == m->N, "gwz_model_from_onr: occupations do not sum to N");
    for (int w = 0; w < m->W; ++w) a[w] = 0;
    int pos = 0;
    for (int i = 0; i < m->M; ++i) {
        for (int k = 0; k < onr[i]; ++k, ++pos) a[pos >> 6] |= 1ull << (pos & 63);
        ++pos;
    }
    return GWZ_OK;
}
extern "C" int gwz_model_to_onr(const gwz_model* m, const uint64_t* a, int32_t* onr) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_model_to_onr: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && onr && a, "gwz_model_to_onr: NULL argument");
    int mode = 0;
    for (int i = 0; i < m->M; ++i) onr[i] = 0;
    for (int p = 0; p < m->B; ++p) {
        if ((a[p >> 6] >> (p & 63)) & 1ull) {
            GWZ_REQUIRE(mode < m->M, "gwz_model_to_onr: malformed address");
            onr[mode]++;
        } else {
            ++mode;
        }
    }
    return GWZ_OK;
}
extern "C" int gwz_model_near_uniform(const gwz_model* m, uint64_t* a) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_model_near_uniform: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && a, "gwz_model_near_uniform: NULL argument");
    std::vector<int32_t> onr(m->M);
    const int ff = m->N / m->M, extras = m->N % m->M;
    for (int i = 0; i < m->M; ++i) onr[i] = ff + (i < extras ? 1 : 0);
    return gwz_model_from_onr(m, onr.data(), a);
}
void gwz::binomial_table(std::vector<uint64_t>& tab) {
    tab.assign(65 * 65, 0);
    for (int n = 0; n <= 64; ++n) {
        tab[n * 65 + 0] = 1;
        for (int k = 1; k <= n; ++k) {
            const uint64_t a = tab[(n - 1) * 65 + k - 1], b = (k <= n - 1) ? tab[(n - 1) * 65 + k] : 0;
            tab[n * 65 + k] = (a > UINT64_MAX - b) ? UINT64_MAX : a + b;
        }
    }
}
extern "C" int gwz_model_dimension(const gwz_model* m, double* dim, uint64_t* dim_exact) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_model_dimension: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m != nullptr, "gwz_model_dimension: model is NULL");
    double r = 1.0;
    const int n = m->B, k = (m->N < m->M - 1) ? m->N : m->M - 1;
    for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
    if (dim) *dim = std::floor(r + 0.5);
    if (dim_exact) {
        *dim_exact = 0;
        if (m->B <= 64) {
            std::vector<uint64_t> tab;
            binomial_table(tab);
            const uint64_t v = tab[m->B * 65 + m->N];
            *dim_exact = (v == UINT64_MAX) ? 0 : v;
        }
    }
    return GWZ_OK;
}

// KC = bond classes per side handled by popcounts (4 or 5).  Measured on B200 (tools/quick_bench.py):
// for sampled walkers KC = 4 plus the compact generic-bond path is never slower than KC = 5 (whose 25
// pair-popcounts cost 1.45x), even at g*u ~ 0.66 where ~0.1 % of the sites hold >= 4 particles; a
// full-basis sweep visits every occupation pattern uniformly and is slightly faster with KC = 5.
// GWZ_KCAP=4|5 overrides (experiments).
static int choose_kc(const gwz_model* m, double g, bool uniform_measure) {
    (void)m;
    (void)g;
    if (const char* e = std::getenv("GWZ_KCAP")) {
        const int v = std::atoi(e);
        if (v == 4 || v == 5) return v;
    }
    return uniform_measure ? 5 : 4;
}

// parameter-dependent tables (O(KC^2) host work per call)
static void make_tables(const gwz_model* m, double g, int kc, DevTables* tb) {
    std::memset(tb, 0, sizeof(*tb));
    tb->g = g;
    tb->c = g * m->u;
    tb->kc = kc;
    for (int L = 0; L < kc; ++L)
        for (int R = 0; R < kc; ++R) {
            const int k = L * kc + R;
            const double rr = (L > 0) ? std::exp(-tb->c * (double)(R - L + 1)) : 0.0;
            const double rl = (R > 0) ? std::exp(-tb->c * (double)(L - R + 1)) : 0.0;
            const double er = (L > 0) ? std::sqrt((double)L * (double)(R + 1)) * rr : 0.0;
            const double el = (R > 0) ? std::sqrt((double)R * (double)(L + 1)) * rl : 0.0;
            tb->RR[k] = rr;
            tb->RL[k] = rl;
            tb->IRR[k] = rr > 0 ? 1.0 / rr : 0.0;
            tb->IRL[k] = rl > 0 ? 1.0 / rl : 0.0;
            tb->FS[k] = rr + rl;
            tb->FE[k] = er + el;
            tb->FD[k] = er * (double)(R - L + 1) + el * (double)(L - R + 1);
        }
}

int gwz::api_check_err_flag(gwz_context* c) {
    // caller has synchronised the stream after copying d_err -> h_err
    if (*c->h_err) {
        *c->h_err = 0;
        cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream);
        cudaStreamSynchronize(c->stream);
        return fail(GWZ_ERR_NONFINITE, "Infinite residence time");
    }
    return GWZ_OK;
}
static int check_err_flag(gwz_context* c) { return gwz::api_check_err_flag(c); }

// ------------------------------------------------------------------------------------------
// probe
// ------------------------------------------------------------------------------------------

extern "C" int gwz_probe_addresses(gwz_model* m, const uint64_t* addr_words, size_t n, const double* params,
                                   const double* u01, int kernel_kind, double* e_loc, double* residence_time,
                                   double* grad_ratio, int32_t* n_hops, double* rates, int32_t* chosen,
                                   uint64_t* next_addr) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_probe_addresses: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && params, "gwz_probe_addresses: NULL argument");
    GWZ_REQUIRE(m->v == 0.0, "gwz_probe_addresses: ExtendedHubbardReal1D models need the gwz_gen_* entry points");
    GWZ_REQUIRE(n == 0 || addr_words, "gwz_probe_addresses: addr_words is NULL");
    GWZ_REQUIRE(kernel_kind == GWZ_KERNEL_ENUMERATE || kernel_kind == GWZ_KERNEL_BITPARALLEL,
                "gwz_probe_addresses: unknown kernel_kind");
    GWZ_REQUIRE(!(rates && kernel_kind != GWZ_KERNEL_ENUMERATE),
                "gwz_probe_addresses: per-hop rates are only produced by GWZ_KERNEL_ENUMERATE");
    if (n == 0) return GWZ_OK;
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    cudaStream_t s = c->stream;
    const size_t W = m->W, twoM = 2 * (size_t)m->M;
    DevBuf<uint64_t> d_addr, d_next;
    DevBuf<double> d_u, d_e, d_t, d_g, d_r;
    DevBuf<int32_t> d_h, d_c;
    GWZ_CUDA(d_addr.alloc(n * W));
    GWZ_CUDA(cudaMemcpyAsync(d_addr.p, addr_words, sizeof(uint64_t) * n * W, cudaMemcpyHostToDevice, s));
    if (u01) {
        GWZ_CUDA(d_u.alloc(n));
        GWZ_CUDA(cudaMemcpyAsync(d_u.p, u01, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    }
    GWZ_CUDA(d_e.alloc(n));
    GWZ_CUDA(d_t.alloc(n));
    GWZ_CUDA(d_g.alloc(n));
    GWZ_CUDA(d_h.alloc(n));
    GWZ_CUDA(d_c.alloc(n));
    GWZ_CUDA(d_next.alloc(n * W));
    if (rates) GWZ_CUDA(d_r.alloc(n * twoM));
    ProbeLaunch L;
    L.dm = m->dm;
    make_tables(m, params[0], choose_kc(m, params[0], true), &L.tb);
    L.nw = m->nw;
    L.kind = kernel_kind;
    L.addr = d_addr.p;
    L.n = n;
    L.u01 = u01 ? d_u.p : nullptr;
    L.e_loc = d_e.p;
    L.tau = d_t.p;
    L.grad = d_g.p;
    L.n_hops = d_h.p;
    L.rates = rates ? d_r.p : nullptr;
    L.chosen = d_c.p;
    L.next_addr = d_next.p;
    L.err_flag = c->d_err;
    GWZ_LAUNCH(launch_probe(L, s));
    if (e_loc) GWZ_CUDA(cudaMemcpyAsync(e_loc, d_e.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    if (residence_time) GWZ_CUDA(cudaMemcpyAsync(residence_time, d_t.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    if (grad_ratio) GWZ_CUDA(cudaMemcpyAsync(grad_ratio, d_g.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    if (n_hops) GWZ_CUDA(cudaMemcpyAsync(n_hops, d_h.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    if (chosen) GWZ_CUDA(cudaMemcpyAsync(chosen, d_c.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    if (rates) GWZ_CUDA(cudaMemcpyAsync(rates, d_r.p, sizeof(double) * n * twoM, cudaMemcpyDeviceToHost, s));
    if (next_addr) GWZ_CUDA(cudaMemcpyAsync(next_addr, d_next.p, sizeof(uint64_t) * n * W, cudaMemcpyDeviceToHost, s));
    GWZ_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    GWZ_CUDA(cudaStreamSynchronize(s));
    return check_err_flag(c);
}

extern "C" int gwz_ansatz_val_and_grad(gwz_model* m, const uint64_t* addr_words, size_t n, const double* params,
                                       double* val, double* grad) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_ansatz_val_and_grad: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && params, "gwz_ansatz_val_and_grad: NULL argument");
    GWZ_REQUIRE(n == 0 || addr_words, "gwz_ansatz_val_and_grad: addr_words is NULL");
    if (n == 0) return GWZ_OK;
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    cudaStream_t s = c->stream;
    DevBuf<uint64_t> d_addr;
    DevBuf<double> d_v, d_g;
    GWZ_CUDA(d_addr.alloc(n * m->W));
    GWZ_CUDA(d_v.alloc(n));
    GWZ_CUDA(d_g.alloc(n));
    GWZ_CUDA(cudaMemcpyAsync(d_addr.p, addr_words, sizeof(uint64_t) * n * m->W, cudaMemcpyHostToDevice, s));
    GWZ_LAUNCH(launch_ansatz(m->dm, params[0], m->nw, d_addr.p, n, d_v.p, d_g.p, s));
    if (val) GWZ_CUDA(cudaMemcpyAsync(val, d_v.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    if (grad) GWZ_CUDA(cudaMemcpyAsync(grad, d_g.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    GWZ_CUDA(cudaStreamSynchronize(s));
    return GWZ_OK;
}

// ------------------------------------------------------------------------------------------
// walkers
// ------------------------------------------------------------------------------------------
extern "C" int gwz_walkers_create(gwz_model* m, uint64_t n_walkers, uint64_t global_offset, const uint64_t* start_addr,
                                  uint64_t seed, gwz_walkers** out) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_walkers_create: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && start_addr && out, "gwz_walkers_create: NULL argument");
    GWZ_REQUIRE(m->v == 0.0, "gwz_walkers_create: ExtendedHubbardReal1D models need the gwz_gen_* entry points");
    GWZ_REQUIRE(n_walkers >= 1, "gwz_walkers_create: need at least one walker");
    // validate the start address: exactly N ones within B bits
    {
        int ones = 0;
        for (int p = 0; p < 64 * m->W; ++p) {
            const int bit = (int)((start_addr[p >> 6] >> (p & 63)) & 1ull);
            GWZ_REQUIRE(!(bit && p >= m->B), "gwz_walkers_create: start address has bits beyond N+M-1");
            ones += bit;
        }
        GWZ_REQUIRE(ones == m->N, "gwz_walkers_create: start address does not hold N particles");
    }
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    gwz_walkers* w = new gwz_walkers();
    HandleGuard<gwz_walkers, gwz_walkers_destroy> guard(w);
    w->m = m;
    w->device = c->device;
    w->n = (size_t)n_walkers;
    w->offset = global_offset;
    w->seed = seed;
    w->steps_done = 0;
    w->acc_steps = 0;
    GWZ_CUDA(cudaMalloc(&w->d_addr, sizeof(uint64_t) * w->n * m->W));
    GWZ_CUDA(cudaMalloc(&w->d_acc, sizeof(double) * w->n * NUM_WALKER_ACC));
    GWZ_CUDA(cudaMalloc(&w->d_vec, sizeof(double) * NUM_VEC));
    GWZ_CUDA(cudaMemsetAsync(w->d_vec, 0, sizeof(double) * NUM_VEC, c->stream));
    GWZ_CUDA(cudaMalloc(&w->d_shift, sizeof(double) * 2));
    GWZ_CUDA(cudaMallocHost(&w->h_shift, sizeof(double) * 2));
    DevBuf<uint64_t> d_start;
    GWZ_CUDA(d_start.alloc(m->W));
    GWZ_CUDA(cudaMemcpyAsync(d_start.p, start_addr, sizeof(uint64_t) * m->W, cudaMemcpyHostToDevice, c->stream));
    GWZ_LAUNCH(launch_walker_init(w->d_addr, w->d_acc, w->n, m->W, d_start.p, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    *out = guard.release();
    return GWZ_OK;
}

extern "C" int gwz_walkers_destroy(gwz_walkers* w) {
    if (w && w->magic != gwz::GWZ_TAG_WALKERS) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_walkers_destroy: not a gwz_walkers handle");
    if (!w) return GWZ_OK;
    cudaSetDevice(w->device);
    cudaFree(w->d_addr);
    cudaFree(w->d_acc);
    cudaFree(w->d_partials);
    cudaFree(w->d_vec);
    cudaFree(w->d_shift);
    cudaFreeHost(w->h_shift);
    cudaFree(w->d_planes);
    cudaFree(w->d_ref_addr);
    cudaFree(w->d_ser_tau);
    cudaFree(w->d_ser_eloc);
    cudaFree(w->d_ser_v);
    cudaFree(w->d_blk);
    cudaFree(w->d_bpartials);
    w->magic = 0;
    delete w;
    return GWZ_OK;
}

extern "C" int gwz_walkers_count(const gwz_walkers* w, uint64_t* n_walkers, uint64_t* steps_done) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_walkers_count: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w != nullptr, "gwz_walkers_count: NULL handle");
    if (n_walkers) *n_walkers = w->n;
    if (steps_done) *steps_done = w->steps_done;
    return GWZ_OK;
}

extern "C" int gwz_walkers_get_addresses(gwz_walkers* w, uint64_t* addr_words) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_walkers_get_addresses: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w && addr_words, "gwz_walkers_get_addresses: NULL argument");
    gwz_context* c = w->m->ctx;
    if (int rc = use_device(c)) return rc;
    if (int rc = to_unary(w)) return rc;
    const int W = w->m->W;
    DevBuf<uint64_t> tmp;
    GWZ_CUDA(tmp.alloc(w->n * W));
    GWZ_LAUNCH(launch_transpose_addr(w->d_addr, tmp.p, w->n, W, false, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(addr_words, tmp.p, sizeof(uint64_t) * w->n * W, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}

extern "C" int gwz_walkers_set_addresses(gwz_walkers* w, const uint64_t* addr_words) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_walkers_set_addresses: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w && addr_words, "gwz_walkers_set_addresses: NULL argument");
    gwz_context* c = w->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int W = w->m->W;
    DevBuf<uint64_t> tmp;
    GWZ_CUDA(tmp.alloc(w->n * W));
    GWZ_CUDA(cudaMemcpyAsync(tmp.p, addr_words, sizeof(uint64_t) * w->n * W, cudaMemcpyHostToDevice, c->stream));
    GWZ_LAUNCH(launch_transpose_addr(tmp.p, w->d_addr, w->n, W, true, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    w->in_planes = false;  // the unary copy is now the valid one
    return GWZ_OK;
}

extern "C" int gwz_walkers_set_step_counter(gwz_walkers* w, uint64_t steps_done) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_walkers_set_step_counter: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w != nullptr, "gwz_walkers_set_step_counter: NULL handle");
    w->steps_done = steps_done;
    return GWZ_OK;
}

static int ensure_partials(gwz_walkers* w, int grid) {
    if (grid > w->partials_cap) {
        if (w->d_partials) cudaFree(w->d_partials);
        w->d_partials = nullptr;
        GWZ_CUDA(cudaMalloc(&w->d_partials, sizeof(double) * NUM_ACC * (size_t)grid));
        w->partials_cap = grid;
    }
    return GWZ_OK;
}

// make room for `rows` more steps of the (tau, E_loc) series behind the `keep_rows` rows that stay valid
static int ensure_series(gwz_walkers* w, uint64_t keep_rows, uint64_t rows) {
    gwz_context* c = w->m->ctx;
    const uint64_t need = keep_rows + rows;
    if (need <= w->ser_cap) return GWZ_OK;
    size_t free_b = 0, total_b = 0;
    GWZ_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const double bytes = 16.0 * (double)need * (double)w->n;
    const double have = (double)free_b + (keep_rows ? 0.0 : 16.0 * (double)w->ser_cap * (double)w->n);
    if (bytes > 0.92 * have)
        return gwz::api_fail(GWZ_ERR_NOMEM, "GWZ_RUN_SERIES: the (tau, E_loc) series of this run (16 B per walker-step) do "
                                            "not fit in the free device memory; use fewer steps or walkers per call");
    double *nt = nullptr, *ne = nullptr;
    if (keep_rows == 0) {  // nothing to carry over: release first
        cudaFree(w->d_ser_tau);
        cudaFree(w->d_ser_eloc);
        w->d_ser_tau = w->d_ser_eloc = nullptr;
        w->ser_cap = 0;
    }
    GWZ_CUDA(cudaMalloc(&nt, sizeof(double) * (size_t)need * w->n));
    if (cudaMalloc(&ne, sizeof(double) * (size_t)need * w->n) != cudaSuccess) {
        cudaFree(nt);
        return gwz::api_fail(GWZ_ERR_NOMEM, "GWZ_RUN_SERIES: out of device memory");
    }
    if (keep_rows) {
        GWZ_CUDA(cudaMemcpyAsync(nt, w->d_ser_tau, sizeof(double) * (size_t)keep_rows * w->n, cudaMemcpyDeviceToDevice, c->stream));
        GWZ_CUDA(cudaMemcpyAsync(ne, w->d_ser_eloc, sizeof(double) * (size_t)keep_rows * w->n, cudaMemcpyDeviceToDevice, c->stream));
        GWZ_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(w->d_ser_tau);
        cudaFree(w->d_ser_eloc);
    }
    w->d_ser_tau = nt;
    w->d_ser_eloc = ne;
    w->ser_cap = need;
    return GWZ_OK;
}

// enqueue burn-in + sampling + per-set reduction on the set's stream (no sync).  rec_tau / rec_eloc (+ rec_addr):
// caller's device buffers for a recording run; GWZ_RUN_SERIES in flags: the set's own series buffers instead.
static int enqueue_run(gwz_walkers* w, const DevTables& tb, uint64_t steps, uint64_t burn_in, int kind, unsigned flags,
                       double* rec_tau, double* rec_eloc, uint64_t* rec_addr, bool time_it, int32_t* rec_hop = nullptr,
                       const double* u01s = nullptr, double* reduce_dst = nullptr) {
    gwz_model* m = w->m;
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    const bool keep = (flags & GWZ_RUN_KEEP_ACC) != 0 && w->acc_steps > 0;
    const bool series = (flags & (GWZ_RUN_SERIES | GWZ_RUN_BLOCKING)) != 0 && rec_tau == nullptr;
    w->blk_valid = false;
    if (series) {
        GWZ_REQUIRE(!keep || w->ser_steps == w->acc_steps,
                    "GWZ_RUN_SERIES with GWZ_RUN_KEEP_ACC: the steps accumulated so far were not recorded");
        if (!keep) w->ser_steps = 0;
        if (int rc = ensure_series(w, w->ser_steps, steps)) return rc;
        rec_tau = w->d_ser_tau + (size_t)w->ser_steps * w->n;
        rec_eloc = w->d_ser_eloc + (size_t)w->ser_steps * w->n;
    } else {
        w->ser_steps = 0;  // whatever was kept no longer covers the estimators
    }
    const bool record = rec_tau != nullptr;
    // production sampling runs on the occupation bit-plane representation (the incremental kernel can keep the
    // (tau, E_loc) series); per-step addresses and the reference-order kernel need the unary strings step by step
    bool planes = (kind == GWZ_KERNEL_BITPARALLEL) && rec_addr == nullptr && tb.kc == 4 && planes_enabled();
    const bool fast = planes && fast_enabled() && walker_fast_supported(m->N, m->M);
    const bool incr = planes && !fast && incr_enabled() && walker_incr_supported(m->N, m->M);
    GWZ_REQUIRE(fast || (rec_hop == nullptr && u01s == nullptr),
                "hop traces / prescribed uniforms need the production kernel (GWZ_KERNEL_BITPARALLEL, 4 <= M <= 250)");
    if (record && !incr && !fast) planes = false;
    if (int rc = planes ? to_planes(w) : to_unary(w)) return rc;
    int grid = 0;
    // the persistent grid comes from an occupancy query: once per (walker set, kernel), not once per call
    const int grid_key = fast ? 1 : incr ? 2 : planes ? 3 : 4 + 4 * kind + 2 * (tb.kc == 5) + (record ? 1 : 0);
    if (w->grid_key == grid_key) {
        grid = w->grid_cached;
    } else {
        if (fast) GWZ_CUDA(walker_fast_grid(m->N, m->M, w->n, c->prop.multiProcessorCount, &grid));
        else if (incr) GWZ_CUDA(walker_incr_grid(m->N, m->M, w->n, c->prop.multiProcessorCount, &grid));
        else if (planes) GWZ_CUDA(walker_planes_grid(m->N, m->M, w->n, c->prop.multiProcessorCount, &grid));
        else GWZ_CUDA(walker_grid(m->nw, m->N, kind, tb.kc, record, w->n, c->prop.multiProcessorCount, &grid));
        w->grid_key = grid_key;
        w->grid_cached = grid;
    }
    if (int rc = ensure_partials(w, grid)) return rc;
    WalkerLaunch L;
    L.dm = m->dm;
    L.tb = tb;
    L.nw = m->nw;
    L.kind = kind;
    L.addr = w->d_addr;
    L.planes = w->d_planes;
    L.acc = w->d_acc;
    L.n = w->n;
    L.global_offset = w->offset;
    L.seed = w->seed;
    L.partials = w->d_partials;
    L.grid = grid;
    L.err_flag = c->d_err;
    L.shift_e = L.shift_g = 0.0;
    L.rec_tau = nullptr;
    L.rec_eloc = nullptr;
    L.rec_addr = nullptr;
    L.rec_hop = nullptr;
    L.u01s = nullptr;
    if (time_it) GWZ_CUDA(cudaEventRecord(c->ev0, c->stream));
    if (burn_in > 0) {
        L.step0 = w->steps_done;
        L.steps = burn_in;
        L.accumulate = 0;
        L.reset_acc = 0;
        if (fast) {
            GWZ_LAUNCH(launch_walkers_fast(L, c->stream));
        } else if (incr) {
            GWZ_LAUNCH(launch_walkers_incr(L, c->stream));
        } else if (planes) {
            GWZ_LAUNCH(launch_walkers_planes(L, c->stream));
        } else {
            int g2 = 0;
            GWZ_CUDA(walker_grid(m->nw, m->N, kind, tb.kc, false, w->n, c->prop.multiProcessorCount, &g2));
            L.grid = g2;
            GWZ_LAUNCH(launch_walkers(L, c->stream));
            L.grid = grid;
        }
        w->steps_done += burn_in;
    }
    if (!keep) {
        // New estimator: its reference point (any constant near the mean; it only limits cancellation in
        // the covariance) is the pooled (E, G) of the previous run, or -- the first time -- E_loc and the
        // grad ratio of walker 0's current address, evaluated by a one-thread kernel.
        if (w->have_next) {
            w->shift_e = w->next_e;
            w->shift_g = w->next_g;
        } else {
            const uint64_t* ref = w->d_addr;
            size_t stride = w->n;
            if (planes) {
                GWZ_LAUNCH(launch_planes_to_unary(w->d_planes, w->n, m->dm, 0, 1, w->d_ref_addr, c->stream));
                ref = w->d_ref_addr;
                stride = 1;
            }
            GWZ_LAUNCH(launch_walker_ref(m->dm, tb, m->nw, ref, stride, w->d_shift, c->stream));
            GWZ_CUDA(cudaMemcpyAsync(w->h_shift, w->d_shift, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
            GWZ_CUDA(cudaStreamSynchronize(c->stream));
            w->shift_e = w->h_shift[0];
            w->shift_g = w->h_shift[1];
            if (time_it) GWZ_CUDA(cudaEventRecord(c->ev0, c->stream));
        }
    }
    L.shift_e = w->shift_e;
    L.shift_g = w->shift_g;
    L.step0 = w->steps_done;
    L.steps = steps;
    L.accumulate = 1;
    L.reset_acc = keep ? 0 : 1;
    L.rec_tau = rec_tau;
    L.rec_eloc = rec_eloc;
    L.rec_addr = rec_addr;
    L.rec_hop = rec_hop;
    L.u01s = u01s;
    if (fast) GWZ_LAUNCH(launch_walkers_fast(L, c->stream));
    else if (incr) GWZ_LAUNCH(launch_walkers_incr(L, c->stream));
    else if (planes) GWZ_LAUNCH(launch_walkers_planes(L, c->stream));
    else GWZ_LAUNCH(launch_walkers(L, c->stream));
    if (time_it) GWZ_CUDA(cudaEventRecord(c->ev1, c->stream));
    w->steps_done += steps;
    w->acc_steps = keep ? w->acc_steps + steps : steps;
    if (series) w->ser_steps += steps;
    // reduce_dst: a single walker set reduces straight into the context's all-reduce / fetch buffer
    // ... and, without a communicator, into its pinned host mirror as well (no copy after the run)
    const bool mirror = reduce_dst != nullptr && c->comm == nullptr && !(flags & GWZ_RUN_BLOCKING);
    GWZ_LAUNCH(launch_reduce_partials(w->d_partials, grid, NUM_ACC, reduce_dst ? reduce_dst : w->d_vec, c->stream,
                                      mirror ? c->h_acc : nullptr, mirror ? c->d_err : nullptr, mirror ? c->h_err : nullptr));
    w->vec_in_ctx = reduce_dst != nullptr;
    w->vec_mirrored = mirror;
    return GWZ_OK;
}

// enqueue the per-walker resample + blocking analysis of the set's series and its reduction into d_vec[NUM_ACC..]
static int enqueue_blocking(gwz_walkers* w, uint64_t len, bool time_it) {
    gwz_model* m = w->m;
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    GWZ_REQUIRE(w->ser_steps >= 1, "blocking analysis: no series kept (run with GWZ_RUN_SERIES first)");
    if (len == 0) len = w->ser_steps;
    GWZ_REQUIRE(len < (1ull << BLK_MAX_LEVELS), "blocking analysis: resample length too large");
    int grid = 0;
    if (blocking_grid(len, w->n, c->prop.multiProcessorCount, &grid) != cudaSuccess)
        return fail(GWZ_ERR_INVALID, "blocking analysis: resample length too large for the per-level shared-memory store");
    if (grid > w->bpartials_cap) {
        cudaFree(w->d_bpartials);
        w->d_bpartials = nullptr;
        w->bpartials_cap = 0;
        GWZ_CUDA(cudaMalloc(&w->d_bpartials, sizeof(double) * NUM_BACC * (size_t)grid));
        w->bpartials_cap = grid;
    }
    if (!w->d_blk) GWZ_CUDA(cudaMalloc(&w->d_blk, sizeof(double) * 5 * w->n));
    // the resampled series pass through this buffer once: one row per resident thread of the kernel
    const size_t scratch_rows = (size_t)grid * BLOCKING_BLOCK;
    const size_t scratch_need = blocking_scratch_stride(len) * scratch_rows;
    if (scratch_need > w->scratch_cap) {
        cudaFree(w->d_ser_v);
        w->d_ser_v = nullptr;
        w->scratch_cap = 0;
        if (cudaMalloc(&w->d_ser_v, sizeof(double) * scratch_need) != cudaSuccess) {
            cudaGetLastError();
            return fail(GWZ_ERR_NOMEM, "blocking analysis: no device memory for the resampled-series scratch rows");
        }
        w->scratch_cap = scratch_need;
    }
    BlockingLaunch L;
    L.tau = w->d_ser_tau;
    L.eloc = w->d_ser_eloc;
    L.stride = w->n;
    L.n = w->n;
    L.steps = w->ser_steps;
    L.len = len;
    L.acc = (w->ser_steps == w->acc_steps) ? w->d_acc : nullptr;
    L.acc_stride = w->n;
    L.shift_e = w->shift_e;
    L.shift_g = w->shift_g;
    L.scratch = w->d_ser_v;
    L.scratch_per_walker = false;
    L.out = w->d_blk;
    L.out_stride = w->n;
    L.partials = w->d_bpartials;
    L.grid = grid;
    if (time_it) GWZ_CUDA(cudaEventRecord(c->ev2, c->stream));
    GWZ_LAUNCH(launch_series_blocking(L, c->stream));
    if (time_it) GWZ_CUDA(cudaEventRecord(c->ev3, c->stream));
    GWZ_LAUNCH(launch_reduce_partials(w->d_bpartials, grid, NUM_BACC, w->d_vec + NUM_ACC, c->stream));
    w->blk_valid = true;
    w->blk_len = len;
    return GWZ_OK;
}

static void fill_blocking(const double* b, uint64_t steps, uint64_t len, float ms, bool with_grad,
                          gwz_blocking_summary* o) {
    std::memset(o, 0, sizeof(*o));
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const double nw = b[BACC_WALKERS], nfail = b[BACC_FAIL];
    o->walkers = (uint64_t)std::llround(nw);
    o->failed = (uint64_t)std::llround(nfail);
    o->mean = b[BACC_MEAN] / nw;                       // state.jl:163 / :277
    o->err_ok = std::sqrt(b[BACC_ERR2]) / nw;          // sqrt(mean(errs)) / walkers with `mean` of a scalar = the scalar
    o->err = o->failed ? nan : o->err_ok;              // a failed walker contributes NaN to the reference's sum
    o->err_err = o->failed ? nan : std::sqrt(b[BACC_EE2]) / nw;
    o->grad[0] = with_grad ? b[BACC_GRAD] / nw : nan;
    o->k_min = o->k_max = -1;
    for (int k = 0; k < 32; ++k)
        if (b[BACC_KHIST + k] > 0.5) {
            if (o->k_min < 0) o->k_min = k;
            o->k_max = k;
        }
    o->series_steps = steps;
    o->resample_len = len;
    o->kernel_ms = ms;
}

static void fill_result(const double* a, float ms, unsigned flags, gwz_vqmc_result* out) {
    const double nw = a[ACC_WALKERS];
    const double val = a[ACC_VAL] / nw, grad = a[ACC_GRAD] / nw;
    out->val = val;
    out->grad[0] = grad;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    if (nw > 1.5) {
        const double vv = (a[ACC_VAL2] / nw - val * val) * nw / (nw - 1.0);
        const double gg = (a[ACC_GRAD2] / nw - grad * grad) * nw / (nw - 1.0);
        out->err = std::sqrt(std::fmax(vv, 0.0) / nw);
        out->grad_err[0] = std::sqrt(std::fmax(gg, 0.0) / nw);
    } else {
        out->err = nan;
        out->grad_err[0] = nan;
    }
    const double T = a[ACC_T];
    const double pe = a[ACC_TE] / T, pg = a[ACC_TG] / T;
    out->pooled_val = pe;
    out->pooled_grad[0] = 2.0 * (a[ACC_TGE] / T - pe * pg);
    out->pooled_var = a[ACC_TEE] / T - pe * pe;
    out->walker_var = a[ACC_VARW] / nw;
    out->total_time = T;
    out->walkers = (uint64_t)std::llround(nw);
    out->steps = (uint64_t)std::llround(a[ACC_STEPS]);
    out->hops = (uint64_t)std::llround(a[ACC_HOPS]);
    for (int i = 0; i < NUM_ACC; ++i) out->acc[i] = a[i];
    out->kernel_ms = ms;
    out->reserved = 0.f;
    std::memset(&out->blocking, 0, sizeof(out->blocking));
    if (flags & GWZ_RUN_POOLED) {  // report the pooled ratio estimator as (val, grad)
        out->val = out->pooled_val;
        out->grad[0] = out->pooled_grad[0];
    }
}

// combine the per-set vectors: device sum within a context, NCCL all-reduce across contexts/ranks
// nvec = NUM_ACC, or NUM_VEC when the blocking accumulators travel with the sampling accumulators
static int combine_and_fetch(gwz_walkers* const* ws, int n, unsigned flags, gwz_vqmc_result* out, int nvec = NUM_ACC) {
    // group the sets by context (order of first appearance)
    std::vector<gwz_context*> ctxs;
    for (int i = 0; i < n; ++i) {
        gwz_context* c = ws[i]->m->ctx;
        bool seen = false;
        for (auto* x : ctxs) seen |= (x == c);
        if (!seen) ctxs.push_back(c);
    }
    for (gwz_context* c : ctxs) {
        if (int rc = use_device(c)) return rc;
        int rows = 0;
        const gwz_walkers* only = nullptr;
        for (int i = 0; i < n; ++i)
            if (ws[i]->m->ctx == c) {
                only = ws[i];
                ++rows;
            }
        GWZ_REQUIRE(rows <= GROUP_ROWS, "gwz_vqmc_run_group: too many walker sets on one device");
        if (rows == 1) {  // the usual case: no second reduction stage
            if (only->vec_in_ctx) {  // the sampling sums are already in c->d_acc
                if (nvec > NUM_ACC)
                    GWZ_CUDA(cudaMemcpyAsync(c->d_acc + NUM_ACC, only->d_vec + NUM_ACC, sizeof(double) * (nvec - NUM_ACC),
                                             cudaMemcpyDeviceToDevice, c->stream));
            } else {
                GWZ_CUDA(cudaMemcpyAsync(c->d_acc, only->d_vec, sizeof(double) * nvec, cudaMemcpyDeviceToDevice, c->stream));
            }
            continue;
        }
        rows = 0;
        for (int i = 0; i < n; ++i)
            if (ws[i]->m->ctx == c) {
                GWZ_CUDA(cudaMemcpyAsync(c->d_group + (size_t)rows * nvec, ws[i]->d_vec, sizeof(double) * nvec,
                                         cudaMemcpyDeviceToDevice, c->stream));
                ++rows;
            }
        GWZ_LAUNCH(launch_reduce_partials(c->d_group, rows, nvec, c->d_acc, c->stream));
    }
    bool any_comm = false;
    for (gwz_context* c : ctxs) any_comm |= (c->comm != nullptr);
    GWZ_REQUIRE(ctxs.size() == 1 || any_comm,
                "gwz_vqmc_run_group: walker sets on several devices need gwz_context_comm_init_all first");
    if (any_comm) {
        const NcclApi* api = nccl_api(nullptr);
        if (!api) return fail(GWZ_ERR_NCCL, "NCCL unavailable");
        for (gwz_context* c : ctxs) GWZ_REQUIRE(c->comm != nullptr, "gwz_vqmc_run_group: context without communicator");
        GWZ_NCCL(api, api->GroupStart());
        for (gwz_context* c : ctxs) {
            // one ncclAllReduce(sum, float64, NUM_ACC) per run, on the sampling stream
            int r = api->AllReduce(c->d_acc, c->d_acc, nvec, kNcclFloat64, kNcclSum, c->comm, c->stream);
            if (r != kNcclSuccess) {
                api->GroupEnd();
                return fail(GWZ_ERR_NCCL, std::string("ncclAllReduce: ") + api->GetErrorString(r));
            }
        }
        GWZ_NCCL(api, api->GroupEnd());
    }
    for (gwz_context* c : ctxs) {
        if (int rc = use_device(c)) return rc;
        // sums and the error flag (which lives right behind them) in one copy -- unless the reduction kernel has
        // already written them to the pinned host mirror
        bool mirrored = (n == 1) && ws[0]->vec_mirrored && nvec == NUM_ACC;
        if (!mirrored)
            GWZ_CUDA(cudaMemcpyAsync(c->h_acc, c->d_acc, sizeof(double) * (NUM_VEC + 1), cudaMemcpyDeviceToHost, c->stream));
    }
    int status = GWZ_OK;
    for (gwz_context* c : ctxs) {
        if (int rc = use_device(c)) return rc;
        GWZ_CUDA(cudaStreamSynchronize(c->stream));
        if (int rc = check_err_flag(c)) status = rc;
    }
    if (status) return status;
    float ms = 0.f;
    {
        gwz_context* c0 = ws[0]->m->ctx;
        if (int rc = use_device(c0)) return rc;
        GWZ_CUDA(cudaEventElapsedTime(&ms, c0->ev0, c0->ev1));
    }
    {
        // pooled means of this run seed the reference point of the next estimator of every set
        const double* a = ctxs[0]->h_acc;
        for (int i = 0; i < n; ++i) {
            ws[i]->next_e = a[ACC_TE] / a[ACC_T];
            ws[i]->next_g = a[ACC_TG] / a[ACC_T];
            ws[i]->have_next = std::isfinite(ws[i]->next_e) && std::isfinite(ws[i]->next_g);
        }
    }
    if (out) fill_result(ctxs[0]->h_acc, ms, flags, out);
    if (out && nvec == NUM_VEC) {
        gwz_context* c0 = ws[0]->m->ctx;
        float bms = 0.f;
        GWZ_CUDA(cudaEventElapsedTime(&bms, c0->ev2, c0->ev3));
        bool with_grad = true;
        for (int i = 0; i < n; ++i) with_grad &= (ws[i]->ser_steps == ws[i]->acc_steps);
        fill_blocking(ctxs[0]->h_acc + NUM_ACC, ws[0]->ser_steps, ws[0]->blk_len, bms, with_grad, &out->blocking);
    }
    return GWZ_OK;
}

extern "C" int gwz_vqmc_run_group(gwz_walkers* const* ws, int n, const double* params, uint64_t steps_per_walker,
                                  uint64_t burn_in, int kernel_kind, unsigned flags, gwz_vqmc_result* out) {
    GWZ_REQUIRE(ws && n >= 1 && params, "gwz_vqmc_run_group: bad arguments");
    GWZ_REQUIRE(steps_per_walker >= 1, "gwz_vqmc_run: steps_per_walker must be >= 1");
    GWZ_REQUIRE(kernel_kind == GWZ_KERNEL_ENUMERATE || kernel_kind == GWZ_KERNEL_BITPARALLEL,
                "gwz_vqmc_run: unknown kernel_kind");
    for (int i = 0; i < n; ++i)
        GWZ_HANDLE(ws[i], gwz::GWZ_TAG_WALKERS, "gwz_vqmc_run_group: NULL or foreign handle (expected gwz_walkers)");
    GWZ_REQUIRE(std::isfinite(params[0]), "gwz_vqmc_run: non-finite parameter");
    DevTables tb;
    make_tables(ws[0]->m, params[0], choose_kc(ws[0]->m, params[0], false), &tb);
    bool timed_first = false;
    std::vector<gwz_context*> timed;
    for (int i = 0; i < n; ++i) {
        GWZ_REQUIRE(ws[i]->m->N == ws[0]->m->N && ws[i]->m->M == ws[0]->m->M && ws[i]->m->u == ws[0]->m->u &&
                        ws[i]->m->t == ws[0]->m->t,
                    "gwz_vqmc_run_group: walker sets belong to different models");
        const bool time_it = (i == 0);
        timed_first |= time_it;
        if (int rc = enqueue_run(ws[i], tb, steps_per_walker, burn_in, kernel_kind, flags, nullptr, nullptr, nullptr,
                                 time_it, nullptr, nullptr, n == 1 ? ws[i]->m->ctx->d_acc : nullptr))
            return rc;
        if (flags & GWZ_RUN_BLOCKING)
            if (int rc = enqueue_blocking(ws[i], 0, time_it)) return rc;
    }
    return combine_and_fetch(ws, n, flags, out, (flags & GWZ_RUN_BLOCKING) ? NUM_VEC : NUM_ACC);
}

extern "C" int gwz_vqmc_run(gwz_walkers* w, const double* params, uint64_t steps_per_walker, uint64_t burn_in,
                            int kernel_kind, unsigned flags, gwz_vqmc_result* out) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_vqmc_run: NULL or foreign handle (expected a gwz_walkers)");
    gwz_walkers* one[1] = {w};
    return gwz_vqmc_run_group(one, 1, params, steps_per_walker, burn_in, kernel_kind, flags, out);
}

extern "C" int gwz_walkers_estimates(gwz_walkers* w, double* val_w, double* grad_w) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_walkers_estimates: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w != nullptr, "gwz_walkers_estimates: NULL handle");
    GWZ_REQUIRE(w->acc_steps > 0, "gwz_walkers_estimates: no steps accumulated yet");
    gwz_context* c = w->m->ctx;
    if (int rc = use_device(c)) return rc;
    DevBuf<double> d_v, d_g;
    GWZ_CUDA(d_v.alloc(w->n));
    GWZ_CUDA(d_g.alloc(w->n));
    GWZ_LAUNCH(launch_walker_estimates(w->d_acc, w->n, w->shift_e, d_v.p, d_g.p, c->stream));
    if (val_w) GWZ_CUDA(cudaMemcpyAsync(val_w, d_v.p, sizeof(double) * w->n, cudaMemcpyDeviceToHost, c->stream));
    if (grad_w) GWZ_CUDA(cudaMemcpyAsync(grad_w, d_g.p, sizeof(double) * w->n, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}

extern "C" int gwz_vqmc_run_record(gwz_walkers* w, const double* params, uint64_t steps, int kernel_kind, unsigned flags,
                                   double* tau, double* eloc, uint64_t* addr, gwz_vqmc_result* out) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_vqmc_run_record: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w && params && tau && eloc, "gwz_vqmc_run_record: NULL argument");
    GWZ_REQUIRE(steps >= 1, "gwz_vqmc_run_record: steps_per_walker must be >= 1");
    GWZ_REQUIRE(kernel_kind == GWZ_KERNEL_ENUMERATE || kernel_kind == GWZ_KERNEL_BITPARALLEL,
                "gwz_vqmc_run_record: unknown kernel_kind");
    gwz_context* c = w->m->ctx;
    if (int rc = use_device(c)) return rc;
    const size_t cells = (size_t)steps * w->n;
    DevBuf<double> d_tau, d_eloc;
    DevBuf<uint64_t> d_addr;
    GWZ_CUDA(d_tau.alloc(cells));
    GWZ_CUDA(d_eloc.alloc(cells));
    if (addr) GWZ_CUDA(d_addr.alloc(cells * w->m->W));
    DevTables tb;
    make_tables(w->m, params[0], choose_kc(w->m, params[0], false), &tb);
    if (int rc = enqueue_run(w, tb, steps, 0, kernel_kind, flags, d_tau.p, d_eloc.p, addr ? d_addr.p : nullptr, true))
        return rc;
    GWZ_CUDA(cudaMemcpyAsync(tau, d_tau.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(eloc, d_eloc.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
    if (addr)
        GWZ_CUDA(cudaMemcpyAsync(addr, d_addr.p, sizeof(uint64_t) * cells * w->m->W, cudaMemcpyDeviceToHost, c->stream));
    gwz_walkers* one[1] = {w};
    return combine_and_fetch(one, 1, flags, out);
}

// parity surface of the production kernel: per step (tau, E_loc, selected hop), optionally with prescribed uniforms
extern "C" int gwz_vqmc_run_trace(gwz_walkers* w, const double* params, uint64_t steps, unsigned flags, const double* u01,
                                  double* tau, double* eloc, int32_t* hop, gwz_vqmc_result* out) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_vqmc_run_trace: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(params && tau && eloc && hop, "gwz_vqmc_run_trace: NULL argument");
    GWZ_REQUIRE(steps >= 1, "gwz_vqmc_run_trace: steps_per_walker must be >= 1");
    GWZ_REQUIRE(std::isfinite(params[0]), "gwz_vqmc_run_trace: non-finite parameter");
    gwz_context* c = w->m->ctx;
    if (int rc = use_device(c)) return rc;
    const size_t cells = (size_t)steps * w->n;
    DevBuf<double> d_tau, d_eloc, d_u;
    DevBuf<int32_t> d_hop;
    GWZ_CUDA(d_tau.alloc(cells));
    GWZ_CUDA(d_eloc.alloc(cells));
    GWZ_CUDA(d_hop.alloc(cells));
    if (u01) {
        GWZ_CUDA(d_u.alloc(cells));
        GWZ_CUDA(cudaMemcpyAsync(d_u.p, u01, sizeof(double) * cells, cudaMemcpyHostToDevice, c->stream));
    }
    DevTables tb;
    make_tables(w->m, params[0], choose_kc(w->m, params[0], false), &tb);
    if (int rc = enqueue_run(w, tb, steps, 0, GWZ_KERNEL_BITPARALLEL, flags & GWZ_RUN_KEEP_ACC, d_tau.p, d_eloc.p, nullptr,
                             true, d_hop.p, u01 ? d_u.p : nullptr))
        return rc;
    GWZ_CUDA(cudaMemcpyAsync(tau, d_tau.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(eloc, d_eloc.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(hop, d_hop.p, sizeof(int32_t) * cells, cudaMemcpyDeviceToHost, c->stream));
    gwz_walkers* one[1] = {w};
    return combine_and_fetch(one, 1, flags, out);
}

// ---- device-side error bars (state.jl:50-60, 150-164, 265-285) --------------------------------------------
// per-walker results [5][stride] -> the caller's arrays
static int fetch_blocking_rows(gwz_context* c, const double* d_blk, size_t n, double* mean_w, double* err_w,
                               double* err_err_w, int32_t* k_w, int64_t* blocks_w) {
    if (mean_w) GWZ_CUDA(cudaMemcpyAsync(mean_w, d_blk, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (err_w) GWZ_CUDA(cudaMemcpyAsync(err_w, d_blk + n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (err_err_w) GWZ_CUDA(cudaMemcpyAsync(err_err_w, d_blk + 2 * n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    std::vector<double> kb;
    if (k_w || blocks_w) {
        kb.resize(2 * n);
        GWZ_CUDA(cudaMemcpyAsync(kb.data(), d_blk + 3 * n, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, c->stream));
    }
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n && k_w; ++i) k_w[i] = (int32_t)kb[i];
    for (size_t i = 0; i < n && blocks_w; ++i) blocks_w[i] = (int64_t)kb[n + i];
    return GWZ_OK;
}

extern "C" int gwz_walkers_blocking(gwz_walkers* w, uint64_t resample_len, gwz_blocking_summary* out, double* mean_w,
                                    double* err_w, double* err_err_w, int32_t* k_w, int64_t* blocks_w) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_walkers_blocking: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w != nullptr, "gwz_walkers_blocking: NULL handle");
    gwz_context* c = w->m->ctx;
    if (int rc = use_device(c)) return rc;
    const uint64_t len = resample_len ? resample_len : w->ser_steps;
    if (!(w->blk_valid && w->blk_len == len)) {  // GWZ_RUN_BLOCKING already analysed the default length
        if (int rc = enqueue_blocking(w, len, true)) return rc;
    } else {
        // re-reduce nothing: d_vec[NUM_ACC..] still holds this set's sums
    }
    GWZ_CUDA(cudaMemcpyAsync(c->d_acc + NUM_ACC, w->d_vec + NUM_ACC, sizeof(double) * NUM_BACC, cudaMemcpyDeviceToDevice, c->stream));
    if (c->comm) {
        const NcclApi* api = nccl_api(nullptr);
        if (!api) return fail(GWZ_ERR_NCCL, "NCCL unavailable");
        GWZ_NCCL(api, api->AllReduce(c->d_acc + NUM_ACC, c->d_acc + NUM_ACC, NUM_BACC, kNcclFloat64, kNcclSum, c->comm, c->stream));
    }
    GWZ_CUDA(cudaMemcpyAsync(c->h_acc + NUM_ACC, c->d_acc + NUM_ACC, sizeof(double) * NUM_BACC, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    if (out) {
        float ms = 0.f;
        GWZ_CUDA(cudaEventElapsedTime(&ms, c->ev2, c->ev3));
        fill_blocking(c->h_acc + NUM_ACC, w->ser_steps, len, ms, w->ser_steps == w->acc_steps, out);
    }
    return fetch_blocking_rows(c, w->d_blk, w->n, mean_w, err_w, err_err_w, k_w, blocks_w);
}

extern "C" int gwz_walkers_series(gwz_walkers* w, uint64_t* steps, double* tau, double* eloc) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_walkers_series: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w && steps, "gwz_walkers_series: NULL argument");
    gwz_context* c = w->m->ctx;
    if (int rc = use_device(c)) return rc;
    *steps = w->ser_steps;
    const size_t cells = (size_t)w->ser_steps * w->n;
    if (tau && cells) GWZ_CUDA(cudaMemcpyAsync(tau, w->d_ser_tau, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
    if (eloc && cells) GWZ_CUDA(cudaMemcpyAsync(eloc, w->d_ser_eloc, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}

extern "C" int gwz_series_blocking(gwz_context* c, const double* tau, const double* eloc, uint64_t steps, uint64_t n,
                                   uint64_t resample_len, gwz_blocking_summary* out, double* mean_w, double* err_w,
                                   double* err_err_w, int32_t* k_w, int64_t* blocks_w, double* resampled) {
    GWZ_HANDLE(c, gwz::GWZ_TAG_CONTEXT, "gwz_series_blocking: NULL or foreign handle (expected a gwz_context)");
    GWZ_REQUIRE(c && tau && eloc, "gwz_series_blocking: NULL argument");
    GWZ_REQUIRE(steps >= 1 && n >= 1, "gwz_series_blocking: empty series");  // state.jl:195-197 / "Not enough data"
    if (int rc = use_device(c)) return rc;
    const uint64_t len = resample_len ? resample_len : steps;
    GWZ_REQUIRE(len < (1ull << BLK_MAX_LEVELS), "gwz_series_blocking: resample length too large");
    int grid = 0;
    if (blocking_grid(len, (size_t)n, c->prop.multiProcessorCount, &grid) != cudaSuccess)
        return fail(GWZ_ERR_INVALID, "gwz_series_blocking: resample length too large for the per-level shared-memory store");
    const size_t cells = (size_t)steps * (size_t)n;
    DevBuf<double> d_tau, d_eloc, d_out, d_part, d_vec, d_scratch;
    const size_t srow = blocking_scratch_stride(len);
    GWZ_CUDA(d_scratch.alloc(srow * (size_t)n));
    GWZ_CUDA(d_tau.alloc(cells));
    GWZ_CUDA(d_eloc.alloc(cells));
    GWZ_CUDA(d_out.alloc(5 * (size_t)n));
    GWZ_CUDA(d_part.alloc((size_t)grid * NUM_BACC));
    GWZ_CUDA(d_vec.alloc(NUM_BACC));
    GWZ_CUDA(cudaMemcpyAsync(d_tau.p, tau, sizeof(double) * cells, cudaMemcpyHostToDevice, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(d_eloc.p, eloc, sizeof(double) * cells, cudaMemcpyHostToDevice, c->stream));
    BlockingLaunch L;
    L.tau = d_tau.p;
    L.eloc = d_eloc.p;
    L.stride = (size_t)n;
    L.n = (size_t)n;
    L.steps = steps;
    L.len = len;
    L.acc = nullptr;
    L.acc_stride = 0;
    L.shift_e = L.shift_g = 0.0;
    L.scratch = d_scratch.p;
    L.scratch_per_walker = true;
    L.out = d_out.p;
    L.out_stride = (size_t)n;
    L.partials = d_part.p;
    L.grid = grid;
    GWZ_CUDA(cudaEventRecord(c->ev2, c->stream));
    GWZ_LAUNCH(launch_series_blocking(L, c->stream));
    GWZ_CUDA(cudaEventRecord(c->ev3, c->stream));
    GWZ_LAUNCH(launch_reduce_partials(d_part.p, grid, NUM_BACC, d_vec.p, c->stream));
    double hb[NUM_BACC];
    GWZ_CUDA(cudaMemcpyAsync(hb, d_vec.p, sizeof(double) * NUM_BACC, cudaMemcpyDeviceToHost, c->stream));
    std::vector<double> rows;
    if (resampled && len > 1) {
        rows.resize(srow * (size_t)n);
        GWZ_CUDA(cudaMemcpyAsync(rows.data(), d_scratch.p, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost, c->stream));
    }
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    if (out) {
        float ms = 0.f;
        GWZ_CUDA(cudaEventElapsedTime(&ms, c->ev2, c->ev3));
        fill_blocking(hb, steps, len, ms, false, out);
    }
    for (size_t w = 0; w < (size_t)n && !rows.empty(); ++w)  // walker-major scratch -> the caller's step-major layout
        for (size_t i = 0; i < (size_t)len; ++i) resampled[i * (size_t)n + w] = rows[w * srow + i];
    return fetch_blocking_rows(c, d_out.p, (size_t)n, mean_w, err_w, err_err_w, k_w, blocks_w);
}

// DVec(res) (state.jl:303-319) without shipping the per-step series to the host
extern "C" int gwz_vqmc_sample_vector(gwz_walkers* w, const double* params, uint64_t steps, int kernel_kind, unsigned flags,
                                      uint64_t cap, uint64_t* addr_out, double* val_out, uint64_t* count,
                                      gwz_vqmc_result* out) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_vqmc_sample_vector: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w && params && count, "gwz_vqmc_sample_vector: NULL argument");
    GWZ_REQUIRE(cap == 0 || (addr_out && val_out), "gwz_vqmc_sample_vector: NULL output buffer");
    GWZ_REQUIRE(steps >= 1, "gwz_vqmc_sample_vector: steps_per_walker must be >= 1");
    GWZ_REQUIRE(w->m->W == 1, "gwz_vqmc_sample_vector: needs single-word addresses (N+M-1 <= 64)");
    GWZ_REQUIRE(kernel_kind == GWZ_KERNEL_ENUMERATE || kernel_kind == GWZ_KERNEL_BITPARALLEL,
                "gwz_vqmc_sample_vector: unknown kernel_kind");
    gwz_context* c = w->m->ctx;
    if (int rc = use_device(c)) return rc;
    const size_t cells = (size_t)steps * w->n;
    double dimd = 0.0;
    uint64_t dime = 0;
    if (int rc = gwz_model_dimension(w->m, &dimd, &dime)) return rc;
    const double distinct = std::fmin((double)cells, dimd);
    size_t slots = 1024;
    while ((double)slots < 2.0 * distinct) slots <<= 1;
    DevBuf<double> d_tau, d_eloc, d_tv, d_ov;
    DevBuf<uint64_t> d_addr, d_tk, d_ok;
    DevBuf<unsigned long long> d_cnt;
    GWZ_CUDA(d_tau.alloc(cells));
    GWZ_CUDA(d_eloc.alloc(cells));
    GWZ_CUDA(d_addr.alloc(cells));
    GWZ_CUDA(d_tk.alloc(slots));
    GWZ_CUDA(d_tv.alloc(slots));
    GWZ_CUDA(d_ok.alloc(cap));
    GWZ_CUDA(d_ov.alloc(cap));
    GWZ_CUDA(d_cnt.alloc(1));
    DevTables tb;
    make_tables(w->m, params[0], choose_kc(w->m, params[0], false), &tb);
    if (int rc = enqueue_run(w, tb, steps, 0, kernel_kind, flags, d_tau.p, d_eloc.p, d_addr.p, true)) return rc;
    GWZ_CUDA(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long), c->stream));
    GWZ_LAUNCH(launch_hist_fill(d_tk.p, d_tv.p, slots, c->stream));
    GWZ_LAUNCH(launch_hist_insert(d_addr.p, d_tau.p, cells, d_tk.p, d_tv.p, (uint64_t)slots - 1, c->stream));
    GWZ_LAUNCH(launch_hist_compact(d_tk.p, d_tv.p, slots, d_ok.p, d_ov.p, d_cnt.p, cap, c->stream));
    unsigned long long n_unique = 0;
    GWZ_CUDA(cudaMemcpyAsync(&n_unique, d_cnt.p, sizeof(n_unique), cudaMemcpyDeviceToHost, c->stream));
    gwz_walkers* one[1] = {w};
    if (int rc = combine_and_fetch(one, 1, flags, out)) return rc;  // synchronises the stream
    *count = n_unique;
    const size_t m = (size_t)std::min<unsigned long long>(n_unique, cap);
    std::vector<uint64_t> k(m);
    std::vector<double> v(m);
    GWZ_CUDA(cudaMemcpy(k.data(), d_ok.p, sizeof(uint64_t) * m, cudaMemcpyDeviceToHost));
    GWZ_CUDA(cudaMemcpy(v.data(), d_ov.p, sizeof(double) * m, cudaMemcpyDeviceToHost));
    GWZ_REQUIRE(n_unique <= cap, "gwz_vqmc_sample_vector: more distinct addresses than the output capacity");
    // ascending addresses, normalize!(dv) (2-norm); O(distinct) host work on the histogram, not on the samples
    std::vector<size_t> order(m);
    for (size_t i = 0; i < m; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return k[a] < k[b]; });
    double nrm = 0.0;
    for (size_t i = 0; i < m; ++i) nrm += v[i] * v[i];
    nrm = std::sqrt(nrm);
    for (size_t i = 0; i < m; ++i) {
        addr_out[i] = k[order[i]];
        val_out[i] = v[order[i]] / nrm;
    }
    return GWZ_OK;
}

// ------------------------------------------------------------------------------------------
// sweep
// ------------------------------------------------------------------------------------------
extern "C" int gwz_sweep_create(gwz_model* m, const uint64_t* basis_words, uint64_t dim, uint64_t rank_begin,
                                uint64_t rank_end, gwz_sweep** out) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_sweep_create: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && out, "gwz_sweep_create: NULL argument");
    GWZ_REQUIRE(m->v == 0.0, "gwz_sweep_create: ExtendedHubbardReal1D models need the gwz_gen_* entry points");
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    gwz_sweep* s = new gwz_sweep();
    HandleGuard<gwz_sweep, gwz_sweep_destroy> guard(s);
    s->m = m;
    s->device = c->device;
    if (basis_words) {
        GWZ_REQUIRE(dim >= 1, "gwz_sweep_create: empty basis");
        s->dim = dim;
        const int W = m->W;
        GWZ_CUDA(cudaMalloc(&s->d_basis, sizeof(uint64_t) * dim * W));
        if (W == 1) {
            GWZ_CUDA(cudaMemcpyAsync(s->d_basis, basis_words, sizeof(uint64_t) * dim, cudaMemcpyHostToDevice, c->stream));
        } else {
            DevBuf<uint64_t> tmp;
            GWZ_CUDA(tmp.alloc(dim * W));
            GWZ_CUDA(cudaMemcpyAsync(tmp.p, basis_words, sizeof(uint64_t) * dim * W, cudaMemcpyHostToDevice, c->stream));
            GWZ_LAUNCH(launch_transpose_addr(tmp.p, s->d_basis, dim, W, true, c->stream));
            GWZ_CUDA(cudaStreamSynchronize(c->stream));
        }
    } else {
        GWZ_REQUIRE(m->B <= 64, "gwz_sweep_create: ranked enumeration needs N+M-1 <= 64 bits");
        std::vector<uint64_t> tab;
        binomial_table(tab);
        const uint64_t total = tab[m->B * 65 + m->N];
        GWZ_REQUIRE(total != UINT64_MAX, "gwz_sweep_create: dimension overflows 64 bits");
        if (rank_end == 0) rank_end = total;
        GWZ_REQUIRE(rank_begin < rank_end && rank_end <= total, "gwz_sweep_create: bad rank range");
        s->dim = rank_end - rank_begin;
        s->rank_begin = rank_begin;
        GWZ_CUDA(cudaMalloc(&s->d_binom, sizeof(uint64_t) * 65 * 65));
        GWZ_CUDA(cudaMemcpyAsync(s->d_binom, tab.data(), sizeof(uint64_t) * 65 * 65, cudaMemcpyHostToDevice, c->stream));
        GWZ_CUDA(cudaStreamSynchronize(c->stream));  // tab is copied from before it moves
        s->h_binom = std::move(tab);
    }
    GWZ_CUDA(sweep_grid(m->nw, choose_kc(m, 0.0, true), m->N, m->M, s->d_basis == nullptr, s->dim, c->prop.multiProcessorCount,
                        &s->grid, &s->chunk));
    GWZ_CUDA(cudaMalloc(&s->d_partials, sizeof(double) * 4 * (size_t)s->grid));
    GWZ_CUDA(cudaEventCreate(&s->ev0));
    GWZ_CUDA(cudaEventCreate(&s->ev1));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    *out = guard.release();
    return GWZ_OK;
}

extern "C" int gwz_sweep_destroy(gwz_sweep* s) {
    if (s && s->magic != gwz::GWZ_TAG_SWEEP) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_sweep_destroy: not a gwz_sweep handle");
    if (!s) return GWZ_OK;
    cudaSetDevice(s->device);
    cudaFree(s->d_basis);
    cudaFree(s->d_binom);
    cudaFree(s->d_partials);
    cudaFree(s->d_rep);
    cudaFree(s->d_mult);
    cudaFree(s->d_done);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    s->magic = 0;
    delete s;
    return GWZ_OK;
}

extern "C" int gwz_sweep_dim(const gwz_sweep* s, uint64_t* dim) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_SWEEP, "gwz_sweep_dim: NULL or foreign handle (expected a gwz_sweep)");
    GWZ_REQUIRE(s && dim, "gwz_sweep_dim: NULL argument");
    *dim = s->dim;
    return GWZ_OK;
}

// GWZ_SWEEP_SYM=0: evaluate every state of a ranked sweep instead of one representative per translation orbit
static bool sweep_sym_enabled() {
    const char* e = std::getenv("GWZ_SWEEP_SYM");
    return !(e && e[0] == '0');
}

// ---- symmetric ranked sweeps: which ranks can hold an orbit representative, and the stored representative basis ----
namespace gwz {
// An orbit representative carries the largest occupation in mode M, so n_M >= tmin = ceil(N/M): the top tmin bits of
// the unary string are set.  In rank (= numeric) order those states are the LAST C(B-tmin, N-tmin) ranks: the ranks
// below cannot hold a representative and are not visited at all.  At commensurate filling the only representative with
// n_M = tmin is the uniform state |tmin ... tmin>: it is handed over on its own (weight 1, in the shard that holds its
// rank) and the visited ranks start at n_M >= tmin + 1 (N = M = 16: 26 % of the states).
RepRange representative_range(const std::vector<uint64_t>& tab, int N, int M, int B, uint64_t rank_begin, uint64_t dim) {
    RepRange r{rank_begin, dim, 0ull, 0};
    int tmin = (N + M - 1) / M;
    const uint64_t total = tab[(size_t)B * 65 + N];
    if (total == UINT64_MAX) return r;
    if (tmin * M == N) {
        uint64_t x = 0, rk = 0;
        int j = 0;
        for (int i = 0; i < M; ++i)
            for (int q = 0; q < tmin; ++q) {
                const int pos = i * (tmin + 1) + q;
                x |= 1ull << pos;
                rk += tab[(size_t)pos * 65 + (++j)];
            }
        if (rk >= rank_begin && rk - rank_begin < dim) {
            r.extra_state = x;
            r.has_extra = 1;
        }
        ++tmin;
    }
    const uint64_t tail = tab[(size_t)(B - tmin) * 65 + (N - tmin)];
    if (tail == UINT64_MAX || tail > total) {
        r.has_extra = 0;
        return r;
    }
    const uint64_t hi = rank_begin + dim;
    r.lo = std::min(std::max(rank_begin, total - tail), hi);
    r.count = hi - r.lo;
    return r;
}

// The representatives of a rank range do not depend on the parameters: count them, collect (state, orbit size) pairs
// and sort them by state.  *stored = 0 when they are not worth / not possible to store (*stored = 1 with *nrep = 0:
// the range holds no representative).
// GWZ_SWEEP_CACHE=0: never store them (every sweep enumerates and tests its rank range).
int build_representative_basis(gwz_context* c, int N, int M, int B, const uint64_t* d_binom, const RepRange& r,
                               uint64_t** d_rep, uint32_t** d_mult, uint64_t* nrep, int* stored) {
    *d_rep = nullptr;
    *d_mult = nullptr;
    *nrep = 0;
    *stored = 0;
    if (const char* e = std::getenv("GWZ_SWEEP_CACHE"))
        if (e[0] == '0') return GWZ_OK;
    int grid = 1;
    uint64_t chunk = 8;
    GWZ_CUDA(sweep_grid(1, 4, N, M, true, r.count ? r.count : 1, c->prop.multiProcessorCount, &grid, &chunk));
    DevBuf<unsigned long long> d_cnt;
    GWZ_CUDA(d_cnt.alloc(1));
    unsigned long long n = 0;
    if (r.count) {
        GWZ_CUDA(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long), c->stream));
        GWZ_LAUNCH(launch_sweep_collect(N, M, B, r.count, r.lo, grid, chunk, d_binom, d_cnt.p, nullptr, nullptr, c->stream));
        GWZ_CUDA(cudaMemcpyAsync(&n, d_cnt.p, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
        GWZ_CUDA(cudaStreamSynchronize(c->stream));
    }
    const uint64_t total = n + (r.has_extra ? 1 : 0);
    size_t free_b = 0, total_b = 0;
    GWZ_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // (8 + 4) B per representative, twice during the sort; keep it to a quarter of the free memory
    if (total == 0) {  // a rank range without representatives: stored, and empty
        *stored = 1;
        return GWZ_OK;
    }
    if (total > 0x7fffffffull || total * 24 > free_b / 4) return GWZ_OK;
    uint64_t* rep = nullptr;
    uint32_t* mult = nullptr;
    GWZ_CUDA(cudaMalloc(&rep, sizeof(uint64_t) * total));
    if (cudaMalloc(&mult, sizeof(uint32_t) * total) != cudaSuccess) {
        cudaFree(rep);
        return api_fail(GWZ_ERR_NOMEM, "representative basis: out of device memory");
    }
    DevBuf<uint64_t> own_rep;   // released on an error path
    DevBuf<uint32_t> own_mult;
    own_rep.p = rep;
    own_mult.p = mult;
    if (n) {
        GWZ_CUDA(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long), c->stream));
        GWZ_LAUNCH(launch_sweep_collect(N, M, B, r.count, r.lo, grid, chunk, d_binom, d_cnt.p, rep, mult, c->stream));
    }
    if (r.has_extra) {
        const uint32_t one = 1u;
        GWZ_CUDA(cudaMemcpyAsync(rep + n, &r.extra_state, sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
        GWZ_CUDA(cudaMemcpyAsync(mult + n, &one, sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    }
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    api_count_launch();
    GWZ_CUDA(sort_representatives(rep, mult, total, c->stream));
    own_rep.p = nullptr;
    own_mult.p = nullptr;
    *d_rep = rep;
    *d_mult = mult;
    *nrep = total;
    *stored = 1;
    return GWZ_OK;
}
}  // namespace gwz

static int sweep_run(gwz_sweep* s, const double* params, double acc4[4]) {
    gwz_model* m = s->m;
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    GWZ_REQUIRE(std::isfinite(params[0]), "gwz_sweep: non-finite parameter");
    SweepLaunch L;
    std::memset(&L, 0, sizeof(L));
    L.dm = m->dm;
    make_tables(m, params[0], choose_kc(m, params[0], true), &L.tb);
    L.nw = m->nw;
    L.basis = s->d_basis;
    L.dim = s->dim;
    L.rank_begin = s->rank_begin;
    L.binom = s->d_binom;
    L.with_grad = 1;
    L.symmetric = sweep_sym_enabled() ? 1 : 0;
    L.partials = s->d_partials;
    L.grid = s->grid;
    L.chunk = s->chunk;
    int grid = s->grid;
    RepRange rr{s->rank_begin, s->dim, 0ull, 0};
    if (L.symmetric && s->d_basis == nullptr && m->N + m->M <= 64) {
        rr = representative_range(s->h_binom, m->N, m->M, m->B, s->rank_begin, s->dim);
        L.rank_begin = rr.lo;
        L.dim = rr.count;
        L.extra_state = rr.extra_state;
        L.has_extra = rr.has_extra;
        uint64_t chunk = 1;
        if (s->enum_grid == 0) {  // occupancy query: once per handle
            GWZ_CUDA(sweep_grid(m->nw, L.tb.kc, m->N, m->M, true, L.dim ? L.dim : 1, c->prop.multiProcessorCount,
                                &s->enum_grid, &s->enum_chunk));
        }
        grid = s->enum_grid;
        chunk = s->enum_chunk;
        if (grid <= s->grid) {
            L.grid = grid;
            L.chunk = chunk;
        } else {  // cannot happen (fewer ranks never need more CTAs); keep the partials buffer safe anyway
            grid = s->grid;
            L.chunk = (L.dim + (uint64_t)grid * SWEEP_BLOCK - 1) / ((uint64_t)grid * SWEEP_BLOCK) + 1;
        }
    }
    const bool sym_ranked = L.symmetric && s->d_basis == nullptr && m->N + m->M <= 64;
    if (sym_ranked && s->rep_state == 0 && ++s->sym_calls >= s->build_at && s->sym_calls >= 2) {
        s->rep_state = -1;
        int stored = 0;
        if (int rc = build_representative_basis(c, m->N, m->M, m->B, s->d_binom, rr, &s->d_rep, &s->d_mult, &s->nrep, &stored))
            return rc;
        if (stored && s->nrep) s->rep_state = 1;  // otherwise every sweep keeps enumerating (same sums either way)
    }
    if (sym_ranked && s->rep_state == 1) {
        // later sweeps of the same handle: evaluate the stored representatives, weighted with their orbit sizes
        L.basis = s->d_rep;
        L.dim = s->nrep;
        L.rank_begin = 0;
        L.weight = s->d_mult;
        L.symmetric = 0;
        L.has_extra = 0;
        uint64_t chunk = 1;
        if (s->rep_grid == 0) {
            GWZ_CUDA(sweep_grid(m->nw, L.tb.kc, m->N, m->M, false, L.dim, c->prop.multiProcessorCount, &s->rep_grid,
                                &s->rep_chunk));
        }
        grid = s->rep_grid;
        chunk = s->rep_chunk;
        if (grid > s->grid) {
            cudaFree(s->d_partials);
            s->d_partials = nullptr;
            GWZ_CUDA(cudaMalloc(&s->d_partials, sizeof(double) * 4 * (size_t)grid));
            s->grid = grid;
            L.partials = s->d_partials;
        }
        L.grid = grid;
        L.chunk = chunk;
    }
    // small grids of the state-by-state kernel reduce inside the sweep kernel (last CTA); without a communicator the
    // four sums go to the pinned host mirror as well, so that a repeated sweep is one launch and one synchronisation
    const bool sym_kernel = L.symmetric && L.basis == nullptr && m->N + m->M <= 64;
    const bool fused = !sym_kernel && grid <= 2048;
    if (fused) {
        if (!s->d_done) {
            GWZ_CUDA(cudaMalloc(&s->d_done, sizeof(unsigned)));
            GWZ_CUDA(cudaMemsetAsync(s->d_done, 0, sizeof(unsigned), c->stream));
        }
        L.done = s->d_done;
        L.out = c->d_acc;
        L.mirror = c->comm ? nullptr : c->h_acc;
    }
    GWZ_CUDA(cudaEventRecord(s->ev0, c->stream));
    GWZ_LAUNCH(launch_sweep(L, c->stream));
    GWZ_CUDA(cudaEventRecord(s->ev1, c->stream));
    if (!fused)
        GWZ_LAUNCH(launch_reduce_partials(s->d_partials, grid, 4, c->d_acc, c->stream, c->comm ? nullptr : c->h_acc));
    if (c->comm) {
        const NcclApi* api = nccl_api(nullptr);
        if (!api) return fail(GWZ_ERR_NCCL, "NCCL unavailable");
        GWZ_NCCL(api, api->AllReduce(c->d_acc, c->d_acc, 4, kNcclFloat64, kNcclSum, c->comm, c->stream));
        GWZ_CUDA(cudaMemcpyAsync(c->h_acc, c->d_acc, sizeof(double) * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    s->ms_pending = true;  // cudaEventElapsedTime costs 5 us: only when gwz_sweep_last_kernel_ms asks (a small sweep is 60 us)
    if (sym_ranked && s->rep_state == 0 && s->sym_calls <= 2) {
        // rent or buy: storing the representatives costs about 1 ms (allocations, sort) + three enumeration passes and
        // then saves about 70 % of an enumerating sweep per call; do it once the sweeps so far would have paid for it.
        // The sweep time is taken from the faster of the first two calls (the first one may include lazy loading).
        float t = 0.f;
        GWZ_CUDA(cudaEventElapsedTime(&t, s->ev0, s->ev1));
        s->last_ms = t;
        s->ms_pending = false;
        s->t_enum = s->sym_calls == 1 ? t : std::min(s->t_enum, t);
        const double T = std::max((double)s->t_enum, 1e-3);
        s->build_at = s->sym_calls == 1 ? 3 : std::max<uint64_t>(3, (uint64_t)std::ceil((1.0 + 3.0 * T) / (0.7 * T)));
        if (const char* e = std::getenv("GWZ_SWEEP_CACHE_AT")) s->build_at = (uint64_t)std::max(2, std::atoi(e));
    }
    for (int i = 0; i < 4; ++i) acc4[i] = c->h_acc[i];
    return GWZ_OK;
}

extern "C" int gwz_sweep_val_and_grad(gwz_sweep* s, const double* params, double* val, double* grad, double* acc4) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_SWEEP, "gwz_sweep_val_and_grad: NULL or foreign handle (expected a gwz_sweep)");
    GWZ_REQUIRE(s && params, "gwz_sweep_val_and_grad: NULL argument");
    double a[4];
    if (int rc = sweep_run(s, params, a)) return rc;
    const double num = a[0], den = a[1], ng = a[2], dg = a[3];
    if (val) *val = num / den;                                   // localenergy.jl:126
    if (grad) grad[0] = (ng * den - num * dg) / (den * den);      // localenergy.jl:127
    if (acc4)
        for (int i = 0; i < 4; ++i) acc4[i] = a[i];
    return GWZ_OK;
}

// host-only: which ranks a symmetric ranked sweep of [rank_begin, rank_begin + dim) visits (see representative_range)
extern "C" int gwz_sweep_representative_range(int N, int M, uint64_t rank_begin, uint64_t dim, uint64_t* first_rank,
                                              uint64_t* count, int32_t* has_uniform, uint64_t* uniform_state) {
    GWZ_REQUIRE(N >= 1 && M >= 2 && N + M <= 64, "gwz_sweep_representative_range: needs N >= 1, M >= 2, N + M <= 64");
    std::vector<uint64_t> tab;
    binomial_table(tab);
    const int B = N + M - 1;
    const uint64_t total = tab[(size_t)B * 65 + N];
    GWZ_REQUIRE(total != UINT64_MAX && rank_begin <= total && dim <= total - rank_begin,
                "gwz_sweep_representative_range: bad rank range");
    const RepRange r = representative_range(tab, N, M, B, rank_begin, dim);
    if (first_rank) *first_rank = r.lo;
    if (count) *count = r.count;
    if (has_uniform) *has_uniform = r.has_extra;
    if (uniform_state) *uniform_state = r.extra_state;
    return GWZ_OK;
}

extern "C" int gwz_sweep_val(gwz_sweep* s, const double* params, double* val) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_SWEEP, "gwz_sweep_val: NULL or foreign handle (expected a gwz_sweep)");
    GWZ_REQUIRE(val != nullptr, "gwz_sweep_val: val is NULL");
    return gwz_sweep_val_and_grad(s, params, val, nullptr, nullptr);
}

static int sweep_terms_impl(gwz_sweep* s, const double* params, uint64_t first, uint64_t count, double* out,
                            uint64_t* states) {
    GWZ_REQUIRE(first <= s->dim && count <= s->dim - first, "gwz_sweep_terms: range outside the sweep");
    if (count == 0) return GWZ_OK;
    gwz_model* m = s->m;
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    DevBuf<double> d_out;
    DevBuf<uint64_t> d_states;
    if (out) GWZ_CUDA(d_out.alloc(count * 4));
    if (states) GWZ_CUDA(d_states.alloc(count * m->W));
    SweepLaunch L;
    std::memset(&L, 0, sizeof(L));
    L.dm = m->dm;
    make_tables(m, params ? params[0] : 0.0, choose_kc(m, params ? params[0] : 0.0, true), &L.tb);
    L.nw = m->nw;
    L.basis = s->d_basis;
    L.dim = s->dim;
    L.rank_begin = s->rank_begin;
    L.binom = s->d_binom;
    L.first = first;
    L.count = count;
    L.terms_out = out ? d_out.p : nullptr;
    L.states_out = states ? d_states.p : nullptr;
    GWZ_LAUNCH(launch_sweep_terms(L, c->stream));
    if (out) GWZ_CUDA(cudaMemcpyAsync(out, d_out.p, sizeof(double) * count * 4, cudaMemcpyDeviceToHost, c->stream));
    if (states)
        GWZ_CUDA(cudaMemcpyAsync(states, d_states.p, sizeof(uint64_t) * count * m->W, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}

extern "C" int gwz_sweep_terms(gwz_sweep* s, const double* params, uint64_t first, uint64_t count, double* out) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_SWEEP, "gwz_sweep_terms: NULL or foreign handle (expected a gwz_sweep)");
    GWZ_REQUIRE(s && params && out, "gwz_sweep_terms: NULL argument");
    return sweep_terms_impl(s, params, first, count, out, nullptr);
}
extern "C" int gwz_sweep_states(gwz_sweep* s, uint64_t first, uint64_t count, uint64_t* out) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_SWEEP, "gwz_sweep_states: NULL or foreign handle (expected a gwz_sweep)");
    GWZ_REQUIRE(s && out, "gwz_sweep_states: NULL argument");
    return sweep_terms_impl(s, nullptr, first, count, nullptr, out);
}
extern "C" int gwz_sweep_last_kernel_ms(const gwz_sweep* s, float* ms) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_SWEEP, "gwz_sweep_last_kernel_ms: NULL or foreign handle (expected a gwz_sweep)");
    GWZ_REQUIRE(s && ms, "gwz_sweep_last_kernel_ms: NULL argument");
    if (s->ms_pending) {
        gwz_sweep* w = const_cast<gwz_sweep*>(s);
        if (int rc = use_device(s->m->ctx)) return rc;
        GWZ_CUDA(cudaEventElapsedTime(&w->last_ms, s->ev0, s->ev1));
        w->ms_pending = false;
    }
    *ms = s->last_ms;
    return GWZ_OK;
}

// ------------------------------------------------------------------------------------------
// optimiser driver (src/amsgrad.jl:117-226): O(np) host arithmetic around the GPU objective
// ------------------------------------------------------------------------------------------
extern "C" int gwz_gd_opts_default(gwz_gd_opts* o) {
    GWZ_REQUIRE(o != nullptr, "gwz_gd_opts_default: NULL argument");
    std::memset(o, 0, sizeof(*o));
    o->maxiter = 100;
    o->use_amsgrad = 1;
    o->early_stop = 1;
    o->alpha = 0.01;
    o->beta1 = 0.1;
    o->beta2 = 0.01;
    const double tol = std::sqrt(std::numeric_limits<double>::epsilon());
    o->grad_tol = o->param_tol = o->val_tol = tol;
    o->err_tol = 0.0;
    return GWZ_OK;
}

extern "C" int gwz_adaptive_gradient_descent(gwz_objective_fn f_vg, gwz_objective_fn f_veg, void* user, const double* p0,
                                             int np, const gwz_gd_opts* o, gwz_gd_trace* tr) {
    GWZ_REQUIRE(f_vg && f_veg && p0 && o && tr && np >= 1, "gwz_adaptive_gradient_descent: bad arguments");
    GWZ_REQUIRE(tr->param && tr->value && tr->error && tr->gradient && tr->first_moment && tr->second_moment &&
                    tr->param_delta,
                "gwz_adaptive_gradient_descent: trace buffers must be allocated by the caller");
    GWZ_REQUIRE(o->maxiter >= 0, "gwz_adaptive_gradient_descent: maxiter < 0");
    std::vector<double> p(p0, p0 + np), grad(np), m1(np), m2(np), dp(np);
    double val = 0.0, err = 0.0, old_val = std::numeric_limits<double>::infinity();
    if (int rc = f_vg(user, p.data(), np, &val, &err, grad.data())) return rc;  // amsgrad.jl:136
    for (int i = 0; i < np; ++i) {
        m1[i] = o->first_moment_init ? o->first_moment_init[i] : grad[i];
        m2[i] = o->second_moment_init ? o->second_moment_init[i] : 0.0;
    }
    int iter = 0, rows = 0;
    tr->converged = 0;
    tr->reason = 0;
    tr->np = np;
    while (iter <= o->maxiter) {  // amsgrad.jl:166 (maxiter+1 evaluations)
        ++iter;
        if (int rc = f_veg(user, p.data(), np, &val, &err, grad.data())) return rc;  // :168
        double ndp = 0.0, ngrad = 0.0;
        for (int i = 0; i < np; ++i) {
            if (o->use_amsgrad) {  // :260-261
                m1[i] = (1 - o->beta1) * m1[i] + o->beta1 * grad[i];
                const double cand = (1 - o->beta2) * m2[i] + o->beta2 * (grad[i] * grad[i]);
                m2[i] = cand > m2[i] ? cand : m2[i];
            } else {  // :241-242
                m1[i] = grad[i];
                m2[i] = 1.0;
            }
            const double nf = (o->fix_params && o->fix_params[i]) ? 0.0 : 1.0;
            grad[i] = grad[i] * nf;                                  // :171
            dp[i] = nf * -(o->alpha * m1[i] / std::sqrt(m2[i]));     // :173
            ndp += dp[i] * dp[i];
            ngrad += grad[i] * grad[i];
        }
        const double dval = old_val - val;
        old_val = val;
        std::memcpy(tr->param + (size_t)rows * np, p.data(), sizeof(double) * np);
        tr->value[rows] = val;
        tr->error[rows] = err;
        std::memcpy(tr->gradient + (size_t)rows * np, grad.data(), sizeof(double) * np);
        std::memcpy(tr->first_moment + (size_t)rows * np, m1.data(), sizeof(double) * np);
        std::memcpy(tr->second_moment + (size_t)rows * np, m2.data(), sizeof(double) * np);
        std::memcpy(tr->param_delta + (size_t)rows * np, dp.data(), sizeof(double) * np);
        ++rows;
        if (o->early_stop) {  // :185-210
            if (std::sqrt(ndp) < o->param_tol) { tr->reason = 1; tr->converged = 1; break; }
            if (std::sqrt(ngrad) < o->grad_tol) { tr->reason = 2; tr->converged = 1; break; }
            if (std::fabs(dval) < o->val_tol) { tr->reason = 3; tr->converged = 1; break; }
            if (std::fabs(dval) < err * o->err_tol) { tr->reason = 4; tr->converged = 1; break; }
        }
        for (int i = 0; i < np; ++i) p[i] = p[i] + dp[i];  // :212
    }
    tr->iterations = rows;
    return GWZ_OK;
}

static inline bool flags_has_blocking(unsigned f) { return (f & GWZ_RUN_BLOCKING) != 0; }
struct VqmcObjective {
    gwz_walkers* w;
    uint64_t steps, burn_in;
    int kind;
    unsigned flags;
    gwz_vqmc_result last;
};
// val_and_grad(qmc::KineticVQMC, p) (wrapper.jl:90-97): _reset! keeps walker positions, clears the estimators
static int vqmc_objective_vg(void* user, const double* p, int, double* val, double* err, double* grad) {
    VqmcObjective* o = (VqmcObjective*)user;
    const unsigned fl = o->flags & ~(GWZ_RUN_KEEP_ACC | GWZ_RUN_SERIES | GWZ_RUN_BLOCKING);
    if (int rc = gwz_vqmc_run(o->w, p, o->steps, o->burn_in, o->kind, fl, &o->last)) return rc;
    *val = o->last.val;
    *err = 0.0;
    grad[0] = o->last.grad[0];
    return GWZ_OK;
}
// val_err_and_grad(qmc::KineticVQMC, p) (wrapper.jl:98-105 -> state.jl:150-164).  With GWZ_RUN_BLOCKING the value, the
// error bar and the gradient are the reference's (per-walker resample + blocking on the device); without it the
// error is the cross-walker standard error of the per-walker estimates.
static int vqmc_objective_veg(void* user, const double* p, int, double* val, double* err, double* grad) {
    VqmcObjective* o = (VqmcObjective*)user;
    if (int rc = gwz_vqmc_run(o->w, p, o->steps, o->burn_in, o->kind, o->flags & ~GWZ_RUN_KEEP_ACC, &o->last)) return rc;
    const gwz_vqmc_result& r = o->last;
    if ((o->flags & GWZ_RUN_BLOCKING) && !(o->flags & GWZ_RUN_POOLED)) {
        *val = r.blocking.mean;
        *err = r.blocking.err;  // NaN when a walker's M test failed: `abs(dval) < err * err_tol` is then false (amsgrad.jl:204)
        grad[0] = r.blocking.grad[0];
    } else {
        *val = r.val;
        *err = (flags_has_blocking(o->flags) ? r.blocking.err : (std::isnan(r.err) ? 0.0 : r.err));
        grad[0] = r.grad[0];
    }
    return GWZ_OK;
}
extern "C" int gwz_amsgrad_vqmc(gwz_walkers* w, const double* p0, const gwz_gd_opts* opts, uint64_t steps_per_walker,
                                uint64_t burn_in, int kernel_kind, unsigned flags, gwz_gd_trace* trace,
                                gwz_vqmc_result* last) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_WALKERS, "gwz_amsgrad_vqmc: NULL or foreign handle (expected a gwz_walkers)");
    GWZ_REQUIRE(w && p0 && opts && trace, "gwz_amsgrad_vqmc: NULL argument");
    VqmcObjective o{w, steps_per_walker, burn_in, kernel_kind, flags, {}};
    const int rc = gwz_adaptive_gradient_descent(vqmc_objective_vg, vqmc_objective_veg, &o, p0, GWZ_NUM_PARAMS, opts, trace);
    if (last) *last = o.last;
    return rc;
}
static int sweep_objective(void* user, const double* p, int, double* val, double* err, double* grad) {
    *err = 0.0;  // abstract.jl:59-62
    return gwz_sweep_val_and_grad((gwz_sweep*)user, p, val, grad, nullptr);
}
extern "C" int gwz_amsgrad_sweep(gwz_sweep* s, const double* p0, const gwz_gd_opts* opts, gwz_gd_trace* trace) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_SWEEP, "gwz_amsgrad_sweep: NULL or foreign handle (expected a gwz_sweep)");
    GWZ_REQUIRE(s && p0 && opts && trace, "gwz_amsgrad_sweep: NULL argument");
    return gwz_adaptive_gradient_descent(sweep_objective, sweep_objective, s, p0, GWZ_NUM_PARAMS, opts, trace);
}
