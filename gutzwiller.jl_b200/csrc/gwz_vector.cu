// gwz_vector.cu -- VectorAnsatz (src/ansatz/vector.jl:1-12): a stored vector used as a 0-parameter ansatz,
// `va(addr, _) = va.vector[addr]` (0 for addresses the vector does not hold, like a Rimu DVec).
//
// The vector lives in HBM as an open-addressing hash table keyed by the BoseFS bitstring (W 64-bit words per key,
// linear probing).  kinetic_sample! (qmc.jl:14-44) then needs one table lookup per hop: one walker per thread, the
// address in registers, hops enumerated in the reference's offdiagonals order, move = the reference's
// pick_random_from_cumsum (qmc.jl:52-61), Philox stream as everywhere else.  Also: the Rayleigh quotient of the
// vector (LocalEnergyEvaluator(H, VectorAnsatz(v)) over basis = keys(v), localenergy.jl:131-155) and batched lookups.
// Kernels and their C ABI live in this one file.
#include <cmath>
#include <cstring>
#include <limits>

#include "gwz_handles.h"

using namespace gwz;

namespace gwz {


struct VecTable {
    const uint64_t* keys;  // [slots][W]
    const double* vals;    // [slots]; NaN marks an unused slot (stored amplitudes are finite)
    uint64_t mask;         // slots - 1
    int W;
};

__host__ __device__ inline uint64_t vec_hash(const uint64_t* k, int W) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < W; ++i) {
        uint64_t x = k[i] + h;
        x ^= x >> 33;
        x *= 0xff51afd7ed558ccdull;
        x ^= x >> 33;
        x *= 0xc4ceb9fe1a85ec53ull;
        x ^= x >> 33;
        h = x;
    }
    return h;
}

template <int W>
__device__ __forceinline__ double vec_lookup(const VecTable& t, const uint64_t (&k)[W]) {
    uint64_t h = vec_hash(k, W) & t.mask;
    for (;;) {
        const double v = t.vals[h];
        if (v != v) return 0.0;  // unused slot: the vector does not hold this address
        const uint64_t* s = t.keys + h * W;
        bool eq = true;
#pragma unroll
        for (int i = 0; i < W; ++i) eq = eq && (s[i] == k[i]);
        if (eq) return v;
        h = (h + 1) & t.mask;
    }
}

template <int W>
__device__ __forceinline__ void words_from_regs(const uint32_t (&a)[2 * W], uint64_t (&k)[W]) {
#pragma unroll
    for (int i = 0; i < W; ++i) k[i] = (uint64_t)a[2 * i] | ((uint64_t)a[2 * i + 1] << 32);
}

// one kinetic_sample! with table amplitudes; returns false when the current address is not in the vector
template <int W>
__device__ __forceinline__ bool vec_step(uint32_t (&a)[2 * W], const DevModel& dm, const VecTable& t, double u01, double& tau,
                                         double& eloc, int& nhops) {
    constexpr int NW = 2 * W;
    uint64_t k[W];
    words_from_regs<W>(a, k);
    const double v1 = vec_lookup<W>(t, k);
    if (v1 == 0.0) return false;
    double S = 0.0, num = 0.0;
    long long d = 0;
    int h = 0;
    enumerate_hops<NW>(a, dm, [&](int hop, int z, bool right, int nsrc, int ndst) {
        uint32_t b[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) b[i] = a[i];
        apply_hop<NW>(b, dm, z, right);
        uint64_t kb[W];
        words_from_regs<W>(b, kb);
        const double v2 = vec_lookup<W>(t, kb);
        S += fabs(v2);                                                          // qmc.jl:27
        num += -dm.t * sqrt((double)nsrc * (double)(ndst + 1)) * v2;             // melem * val2 (qmc.jl:28)
        if (right) d += (long long)nsrc * (nsrc - 1) / 2;
        h = hop;
        return false;
    });
    const double D = dm.u * (double)d;
    tau = fabs(v1) / S;               // qmc.jl:33
    eloc = (D * v1 + num) / v1;       // qmc.jl:21,37
    nhops = h;
    const double T = u01 * S;         // qmc.jl:52-61
    double cum = 0.0;
    int zsel = -1;
    bool rsel = true;
    enumerate_hops<NW>(a, dm, [&](int, int z, bool right, int, int) {
        uint32_t b[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) b[i] = a[i];
        apply_hop<NW>(b, dm, z, right);
        uint64_t kb[W];
        words_from_regs<W>(b, kb);
        const double r = fabs(vec_lookup<W>(t, kb));
        cum += r;
        if (r > 0.0) {  // a hop into an address outside the vector has rate 0 and is never chosen
            zsel = z;
            rsel = right;
        }
        return T < cum;
    });
    if (zsel >= 0) apply_hop<NW>(a, dm, zsel, rsel);
    return true;
}

template <int W>
__global__ void __launch_bounds__(128)
vec_walker_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ VecTable t, const __grid_constant__ PhiloxKeys pk,
                  uint64_t* __restrict__ addr /* SoA [W][n] */, double* __restrict__ acc /* [3][n] */, size_t n, uint64_t goff,
                  uint64_t step0, uint64_t steps, int accumulate, int reset_acc, double E0, double* __restrict__ partials,
                  double* __restrict__ rec_tau, double* __restrict__ rec_eloc, uint64_t* __restrict__ rec_addr,
                  int* __restrict__ err_flag) {
    constexpr int NW = 2 * W;
    __shared__ double sm[4][GEN_HEAD];
    double ps[GEN_HEAD];
#pragma unroll
    for (int i = 0; i < GEN_HEAD; ++i) ps[i] = 0.0;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (size_t w = tid; w < n; w += nthreads) {
        uint32_t a[NW];
        load_addr_soa<NW>(addr, n, w, W, a);
        double st = 0.0, ste = 0.0, stee = 0.0;
        if (accumulate && !reset_acc) {
            st = acc[w];
            ste = acc[n + w];
            stee = acc[2 * n + w];
        }
        const uint64_t wid = goff + (uint64_t)w;
        const uint32_t c2 = (uint32_t)wid, c3 = (uint32_t)(wid >> 32);
        uint32_t rnd[4] = {0u, 0u, 0u, 0u};
        unsigned long long hops = 0ull;
        for (uint64_t k = 0; k < steps; ++k) {
            const uint64_t step = step0 + k;
            if (k == 0 || (step & 1ull) == 0ull) {
                const uint64_t blk = step >> 1;
                philox4x32_10_keyed((uint32_t)blk, (uint32_t)(blk >> 32), c2, c3, pk, rnd);
            }
            const double u01 = (step & 1ull) ? u01_from_words(rnd[2], rnd[3]) : u01_from_words(rnd[0], rnd[1]);
            if (rec_addr) store_addr_aos<NW>(rec_addr, k * n + w, W, a);
            double tau = 0.0, eloc = 0.0;
            int nh = 0;
            if (!vec_step<W>(a, dm, t, u01, tau, eloc, nh)) {
                bad = true;
                break;
            }
            bad |= !(tau < 1.0e300);
            if (rec_tau) {
                rec_tau[k * n + w] = tau;
                rec_eloc[k * n + w] = eloc;
            }
            const double dE = eloc - E0, tE = tau * dE;
            st += tau;
            ste += tE;
            stee = fma(tE, dE, stee);
            hops += (unsigned)nh;
        }
        store_addr_soa<NW>(addr, n, w, W, a);
        if (accumulate) {
            acc[w] = st;
            acc[n + w] = ste;
            acc[2 * n + w] = stee;
            const double val_w = E0 + ste / st;
            ps[GS_WALKERS] += 1.0;
            ps[GS_VAL] += val_w;
            ps[GS_VAL2] = fma(val_w, val_w, ps[GS_VAL2]);
            ps[GS_T] += st;
            ps[GS_TE] += st * val_w;
            ps[GS_TEE] += st * E0 * E0 + 2.0 * E0 * ste + stee;
            ps[GS_HOPS] += (double)hops;
            ps[GS_STEPS] += (double)steps;
        }
    }
    if (bad) atomicOr(err_flag, 1);
    if (!accumulate || partials == nullptr) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < GEN_HEAD; ++i) {
        const double v = warp_sum(ps[i]);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < GEN_HEAD)
        partials[(size_t)blockIdx.x * GEN_HEAD + threadIdx.x] = sm[0][threadIdx.x] + sm[1][threadIdx.x] + sm[2][threadIdx.x] + sm[3][threadIdx.x];
}

// batched va(addr, _)
template <int W>
__global__ void vec_lookup_kernel(const __grid_constant__ VecTable t, const uint64_t* __restrict__ addr, size_t n,
                                  double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k[W];
#pragma unroll
    for (int q = 0; q < W; ++q) k[q] = addr[i * W + q];
    out[i] = vec_lookup<W>(t, k);
}

// Rayleigh quotient terms over the occupied slots: (sum_k v_k (D_k v_k + sum_l m_kl v_l), sum_k v_k^2)
template <int W>
__global__ void __launch_bounds__(128)
vec_rayleigh_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ VecTable t, size_t slots,
                    double* __restrict__ partials /* [grid][2] */) {
    constexpr int NW = 2 * W;
    __shared__ double sm[4][2];
    double num = 0.0, den = 0.0;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (size_t)gridDim.x * blockDim.x) {
        if (t.vals[s] != t.vals[s]) continue;  // unused slot
        uint32_t a[NW];
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const uint64_t kw = t.keys[s * W + i];
            a[2 * i] = (uint32_t)kw;
            a[2 * i + 1] = (uint32_t)(kw >> 32);
        }
        const double v1 = t.vals[s];
        double off = 0.0;
        long long d = 0;
        enumerate_hops<NW>(a, dm, [&](int, int z, bool right, int nsrc, int ndst) {
            uint32_t b[NW];
#pragma unroll
            for (int i = 0; i < NW; ++i) b[i] = a[i];
            apply_hop<NW>(b, dm, z, right);
            uint64_t kb[W];
            words_from_regs<W>(b, kb);
            off += -dm.t * sqrt((double)nsrc * (double)(ndst + 1)) * vec_lookup<W>(t, kb);
            if (right) d += (long long)nsrc * (nsrc - 1) / 2;
            return false;
        });
        num += v1 * (dm.u * (double)d * v1 + off);  // localenergy.jl:139-146
        den += v1 * v1;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    num = warp_sum(num);
    den = warp_sum(den);
    if (lane == 0) {
        sm[warp][0] = num;
        sm[warp][1] = den;
    }
    __syncthreads();
    if (threadIdx.x < 2) partials[(size_t)blockIdx.x * 2 + threadIdx.x] = sm[0][threadIdx.x] + sm[1][threadIdx.x] + sm[2][threadIdx.x] + sm[3][threadIdx.x];
}

}  // namespace gwz

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
struct gwz_vansatz {
    uint32_t magic = gwz::GWZ_TAG_VANSATZ;  // handle type tag, checked by every entry point
    int device = 0;  // destroy must not dereference parent handles (a host GC may release them in any order)
    gwz_model* m;
    uint64_t n = 0, slots = 0;
    uint64_t* d_keys = nullptr;
    double* d_vals = nullptr;
    VecTable t{};
};
struct gwz_vwalkers {
    uint32_t magic = gwz::GWZ_TAG_VWALKERS;  // handle type tag, checked by every entry point
    int device = 0;
    gwz_vansatz* a;
    size_t n;
    uint64_t offset, seed, steps_done, acc_steps;
    uint64_t* d_addr = nullptr;  // SoA [W][n]
    double* d_acc = nullptr;     // [3][n]
    double* d_partials = nullptr;
    int partials_cap = 0;
    double* d_vec = nullptr;
    double* h_vec = nullptr;
    double shift_e = 0.0;
    bool have_shift = false;
};

#define GWZ_VEC_DISPATCH(W_, CALL)                                 \
    switch (W_) {                                                  \
        case 1: { constexpr int W = 1; CALL; } break;              \
        case 2: { constexpr int W = 2; CALL; } break;              \
        case 3: { constexpr int W = 3; CALL; } break;              \
        default: { constexpr int W = 4; CALL; } break;             \
    }

extern "C" int gwz_vector_ansatz_create(gwz_model* m, const uint64_t* addr_words, const double* values, uint64_t n,
                                        gwz_vansatz** out) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_vector_ansatz_create: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && addr_words && values && out, "gwz_vector_ansatz_create: NULL argument");
    GWZ_REQUIRE(n >= 1, "gwz_vector_ansatz_create: empty vector");
    GWZ_REQUIRE(m->v == 0.0, "gwz_vector_ansatz_create: HubbardReal1D only");
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    const int W = m->W;
    uint64_t slots = 1024;
    while (slots < 2 * n) slots <<= 1;
    std::vector<uint64_t> keys((size_t)slots * W, 0ull);
    std::vector<double> vals((size_t)slots, std::numeric_limits<double>::quiet_NaN());
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t* k = addr_words + i * W;
        int ones = 0;
        for (int p = 0; p < 64 * W; ++p) {
            const int bit = (int)((k[p >> 6] >> (p & 63)) & 1ull);
            GWZ_REQUIRE(!(bit && p >= m->B), "gwz_vector_ansatz_create: address has bits beyond N+M-1");
            ones += bit;
        }
        GWZ_REQUIRE(ones == m->N, "gwz_vector_ansatz_create: address does not hold N particles");
        GWZ_REQUIRE(std::isfinite(values[i]), "gwz_vector_ansatz_create: non-finite amplitude");
        uint64_t h = vec_hash(k, W) & (slots - 1);
        for (;;) {
            uint64_t* s = keys.data() + h * W;
            if (vals[h] != vals[h]) {
                for (int q = 0; q < W; ++q) s[q] = k[q];
                vals[h] = values[i];
                break;
            }
            bool eq = true;
            for (int q = 0; q < W; ++q) eq = eq && (s[q] == k[q]);
            GWZ_REQUIRE(!eq, "gwz_vector_ansatz_create: duplicate address");
            h = (h + 1) & (slots - 1);
        }
    }
    gwz_vansatz* a = new gwz_vansatz();
    HandleGuard<gwz_vansatz, gwz_vector_ansatz_destroy> guard(a);
    a->m = m;
    a->device = m->ctx->device;
    a->n = n;
    a->slots = slots;
    GWZ_CUDA(cudaMalloc(&a->d_keys, sizeof(uint64_t) * keys.size()));
    GWZ_CUDA(cudaMalloc(&a->d_vals, sizeof(double) * vals.size()));
    GWZ_CUDA(cudaMemcpy(a->d_keys, keys.data(), sizeof(uint64_t) * keys.size(), cudaMemcpyHostToDevice));
    GWZ_CUDA(cudaMemcpy(a->d_vals, vals.data(), sizeof(double) * vals.size(), cudaMemcpyHostToDevice));
    a->t.keys = a->d_keys;
    a->t.vals = a->d_vals;
    a->t.mask = slots - 1;
    a->t.W = W;
    *out = guard.release();
    return GWZ_OK;
}
extern "C" int gwz_vector_ansatz_destroy(gwz_vansatz* a) {
    if (a && a->magic != gwz::GWZ_TAG_VANSATZ) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_vector_ansatz_destroy: not a gwz_vansatz handle");
    if (!a) return GWZ_OK;
    if (cudaSetDevice(a->device) == cudaSuccess) {
        cudaFree(a->d_keys);
        cudaFree(a->d_vals);
    }
    a->magic = 0;
    delete a;
    return GWZ_OK;
}
extern "C" int gwz_vector_ansatz_lookup(gwz_vansatz* a, const uint64_t* addr_words, size_t n, double* val) {
    GWZ_HANDLE(a, gwz::GWZ_TAG_VANSATZ, "gwz_vector_ansatz_lookup: NULL or foreign handle (expected a gwz_vansatz)");
    GWZ_REQUIRE(a && val && (n == 0 || addr_words), "gwz_vector_ansatz_lookup: NULL argument");
    if (n == 0) return GWZ_OK;
    gwz_context* c = a->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int Wm = a->m->W;
    DevBuf<uint64_t> d_a;
    DevBuf<double> d_v;
    GWZ_CUDA(d_a.alloc(n * Wm));
    GWZ_CUDA(d_v.alloc(n));
    GWZ_CUDA(cudaMemcpyAsync(d_a.p, addr_words, sizeof(uint64_t) * n * Wm, cudaMemcpyHostToDevice, c->stream));
    api_count_launch();
    GWZ_VEC_DISPATCH(Wm, (vec_lookup_kernel<W><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(a->t, d_a.p, n, d_v.p)));
    GWZ_CUDA(cudaGetLastError());
    GWZ_CUDA(cudaMemcpyAsync(val, d_v.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}
// LocalEnergyEvaluator(H, VectorAnsatz(v))() over basis = keys(v) (vector.jl:10, localenergy.jl:131-155)
extern "C" int gwz_vector_ansatz_rayleigh(gwz_vansatz* a, double* val) {
    GWZ_HANDLE(a, gwz::GWZ_TAG_VANSATZ, "gwz_vector_ansatz_rayleigh: NULL or foreign handle (expected a gwz_vansatz)");
    GWZ_REQUIRE(a && val, "gwz_vector_ansatz_rayleigh: NULL argument");
    gwz_context* c = a->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int grid = 4 * c->prop.multiProcessorCount, Wm = a->m->W;
    DevBuf<double> d_p, d_o;
    GWZ_CUDA(d_p.alloc((size_t)grid * 2));
    GWZ_CUDA(d_o.alloc(2));
    api_count_launch();
    GWZ_VEC_DISPATCH(Wm, (vec_rayleigh_kernel<W><<<grid, 128, 0, c->stream>>>(a->m->dm, a->t, (size_t)a->slots, d_p.p)));
    GWZ_CUDA(cudaGetLastError());
    GWZ_LAUNCH(launch_reduce_partials(d_p.p, grid, 2, d_o.p, c->stream));
    double h[2];
    GWZ_CUDA(cudaMemcpyAsync(h, d_o.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    *val = h[0] / h[1];
    return GWZ_OK;
}

extern "C" int gwz_vwalkers_create(gwz_vansatz* a, uint64_t n_walkers, uint64_t global_offset, const uint64_t* start_addr,
                                   uint64_t seed, gwz_vwalkers** out) {
    GWZ_HANDLE(a, gwz::GWZ_TAG_VANSATZ, "gwz_vwalkers_create: NULL or foreign handle (expected a gwz_vansatz)");
    GWZ_REQUIRE(a && start_addr && out, "gwz_vwalkers_create: NULL argument");
    GWZ_REQUIRE(n_walkers >= 1, "gwz_vwalkers_create: need at least one walker");
    gwz_model* m = a->m;
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    double v0 = 0.0;
    if (int rc = gwz_vector_ansatz_lookup(a, start_addr, 1, &v0)) return rc;
    GWZ_REQUIRE(v0 != 0.0, "gwz_vwalkers_create: the starting address is not in the vector");
    gwz_vwalkers* w = new gwz_vwalkers();
    HandleGuard<gwz_vwalkers, gwz_vwalkers_destroy> guard(w);
    w->a = a;
    w->device = a->device;
    w->n = (size_t)n_walkers;
    w->offset = global_offset;
    w->seed = seed;
    w->steps_done = w->acc_steps = 0;
    GWZ_CUDA(cudaMalloc(&w->d_addr, sizeof(uint64_t) * w->n * m->W));
    GWZ_CUDA(cudaMalloc(&w->d_acc, sizeof(double) * w->n * 3));
    GWZ_CUDA(cudaMalloc(&w->d_vec, sizeof(double) * GEN_HEAD));
    GWZ_CUDA(cudaMallocHost(&w->h_vec, sizeof(double) * GEN_HEAD));
    DevBuf<uint64_t> d_start;
    DevBuf<double> d_acc5;  // launch_walker_init clears NUM_WALKER_ACC rows; use a scratch of that size
    GWZ_CUDA(d_start.alloc(m->W));
    GWZ_CUDA(d_acc5.alloc(w->n * NUM_WALKER_ACC));
    GWZ_CUDA(cudaMemcpyAsync(d_start.p, start_addr, sizeof(uint64_t) * m->W, cudaMemcpyHostToDevice, c->stream));
    GWZ_LAUNCH(launch_walker_init(w->d_addr, d_acc5.p, w->n, m->W, d_start.p, c->stream));
    GWZ_CUDA(cudaMemsetAsync(w->d_acc, 0, sizeof(double) * w->n * 3, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    *out = guard.release();
    return GWZ_OK;
}
extern "C" int gwz_vwalkers_destroy(gwz_vwalkers* w) {
    if (w && w->magic != gwz::GWZ_TAG_VWALKERS) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_vwalkers_destroy: not a gwz_vwalkers handle");
    if (!w) return GWZ_OK;
    if (cudaSetDevice(w->device) == cudaSuccess) {
        cudaFree(w->d_addr);
        cudaFree(w->d_acc);
        cudaFree(w->d_partials);
        cudaFree(w->d_vec);
        if (w->h_vec) cudaFreeHost(w->h_vec);
    }
    w->magic = 0;
    delete w;
    return GWZ_OK;
}
extern "C" int gwz_vwalkers_get_addresses(gwz_vwalkers* w, uint64_t* addr_words) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_VWALKERS, "gwz_vwalkers_get_addresses: NULL or foreign handle (expected a gwz_vwalkers)");
    GWZ_REQUIRE(w && addr_words, "gwz_vwalkers_get_addresses: NULL argument");
    gwz_context* c = w->a->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int W = w->a->m->W;
    DevBuf<uint64_t> tmp;
    GWZ_CUDA(tmp.alloc(w->n * W));
    GWZ_LAUNCH(launch_transpose_addr(w->d_addr, tmp.p, w->n, W, false, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(addr_words, tmp.p, sizeof(uint64_t) * w->n * W, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}

static int vwalkers_launch(gwz_vwalkers* w, uint64_t step0, uint64_t steps, int accumulate, int reset_acc, int grid,
                           double* rec_tau, double* rec_eloc, uint64_t* rec_addr) {
    gwz_model* m = w->a->m;
    gwz_context* c = m->ctx;
    api_count_launch();
    GWZ_VEC_DISPATCH(m->W, (vec_walker_kernel<W><<<grid, 128, 0, c->stream>>>(
                               m->dm, w->a->t, philox_keys(w->seed), w->d_addr, w->d_acc, w->n, w->offset, step0, steps, accumulate,
                               reset_acc, accumulate ? w->shift_e : 0.0, w->d_partials, rec_tau, rec_eloc, rec_addr, c->d_err)));
    GWZ_CUDA(cudaGetLastError());
    return GWZ_OK;
}

static int vwalkers_run(gwz_vwalkers* w, uint64_t steps, uint64_t burn_in, unsigned flags, double* rec_tau, double* rec_eloc,
                        uint64_t* rec_addr, gwz_gen_result* out) {
    GWZ_REQUIRE(w != nullptr, "gwz_vwalkers_run: NULL handle");
    GWZ_REQUIRE(steps >= 1, "gwz_vwalkers_run: steps_per_walker must be >= 1");
    gwz_model* m = w->a->m;
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    const size_t need = (w->n + 127) / 128, resident = (size_t)c->prop.multiProcessorCount * 8;
    const int grid = (int)(need < resident ? need : resident);
    if (grid > w->partials_cap) {
        if (w->d_partials) cudaFree(w->d_partials);
        w->d_partials = nullptr;
        GWZ_CUDA(cudaMalloc(&w->d_partials, sizeof(double) * GEN_HEAD * (size_t)grid));
        w->partials_cap = grid;
    }
    GWZ_CUDA(cudaEventRecord(c->ev0, c->stream));
    if (burn_in > 0) {
        if (int rc = vwalkers_launch(w, w->steps_done, burn_in, 0, 0, grid, nullptr, nullptr, nullptr)) return rc;
        w->steps_done += burn_in;
    }
    const bool keep = (flags & GWZ_RUN_KEEP_ACC) != 0 && w->acc_steps > 0;
    if (!keep && !w->have_shift) {
        // reference point of the shifted sums: the Rayleigh quotient of the vector (exact when it is an eigenvector)
        if (int rc = gwz_vector_ansatz_rayleigh(w->a, &w->shift_e)) return rc;
        if (!std::isfinite(w->shift_e)) w->shift_e = 0.0;
        w->have_shift = true;
        GWZ_CUDA(cudaEventRecord(c->ev0, c->stream));
    }
    const size_t cells = (size_t)steps * w->n;
    DevBuf<double> d_tau, d_el;
    DevBuf<uint64_t> d_ra;
    if (rec_tau) {
        GWZ_CUDA(d_tau.alloc(cells));
        GWZ_CUDA(d_el.alloc(cells));
        if (rec_addr) GWZ_CUDA(d_ra.alloc(cells * m->W));
    }
    if (int rc = vwalkers_launch(w, w->steps_done, steps, 1, keep ? 0 : 1, grid, rec_tau ? d_tau.p : nullptr,
                                 rec_tau ? d_el.p : nullptr, rec_addr ? d_ra.p : nullptr))
        return rc;
    GWZ_CUDA(cudaEventRecord(c->ev1, c->stream));
    w->steps_done += steps;
    w->acc_steps = keep ? w->acc_steps + steps : steps;
    GWZ_LAUNCH(launch_reduce_partials(w->d_partials, grid, GEN_HEAD, w->d_vec, c->stream));
    if (c->comm) {
        const NcclApi* api = nccl_api(nullptr);
        if (!api) return api_fail(GWZ_ERR_NCCL, "NCCL unavailable");
        GWZ_NCCL(api, api->AllReduce(w->d_vec, w->d_vec, GEN_HEAD, kNcclFloat64, kNcclSum, c->comm, c->stream));
    }
    GWZ_CUDA(cudaMemcpyAsync(w->h_vec, w->d_vec, sizeof(double) * GEN_HEAD, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (rec_tau) {
        GWZ_CUDA(cudaMemcpyAsync(rec_tau, d_tau.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
        GWZ_CUDA(cudaMemcpyAsync(rec_eloc, d_el.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
        if (rec_addr) GWZ_CUDA(cudaMemcpyAsync(rec_addr, d_ra.p, sizeof(uint64_t) * cells * m->W, cudaMemcpyDeviceToHost, c->stream));
    }
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    if (int rc = api_check_err_flag(c)) return rc;
    float ms = 0.f;
    GWZ_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (out) {
        const double* v = w->h_vec;
        const double nw = v[GS_WALKERS], T = v[GS_T], val = v[GS_VAL] / nw, pe = v[GS_TE] / T;
        out->val = (flags & GWZ_RUN_POOLED) ? pe : val;
        out->err = nw > 1.5 ? std::sqrt(std::fmax((v[GS_VAL2] / nw - val * val) * nw / (nw - 1.0), 0.0) / nw)
                            : std::numeric_limits<double>::quiet_NaN();
        out->pooled_val = pe;
        out->pooled_var = std::fmax(v[GS_TEE] / T - pe * pe, 0.0);
        out->total_time = T;
        out->walkers = (uint64_t)(nw + 0.5);
        out->steps = (uint64_t)(v[GS_STEPS] + 0.5);
        out->hops = (uint64_t)(v[GS_HOPS] + 0.5);
        out->kernel_ms = ms;
        out->np = 0;
    }
    return GWZ_OK;
}
extern "C" int gwz_vwalkers_run(gwz_vwalkers* w, uint64_t steps_per_walker, uint64_t burn_in, unsigned flags,
                                gwz_gen_result* out) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_VWALKERS, "gwz_vwalkers_run: NULL or foreign handle (expected a gwz_vwalkers)");
    return vwalkers_run(w, steps_per_walker, burn_in, flags, nullptr, nullptr, nullptr, out);
}
extern "C" int gwz_vwalkers_run_record(gwz_vwalkers* w, uint64_t steps_per_walker, unsigned flags, double* tau, double* eloc,
                                       uint64_t* addr, gwz_gen_result* out) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_VWALKERS, "gwz_vwalkers_run_record: NULL or foreign handle (expected a gwz_vwalkers)");
    GWZ_REQUIRE(tau && eloc, "gwz_vwalkers_run_record: NULL series buffer");
    return vwalkers_run(w, steps_per_walker, 0, flags, tau, eloc, addr, out);
}
