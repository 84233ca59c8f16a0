// gwz_probe_sweep.cu -- per-address probe (kinetic_sample! for a batch of addresses) and the
// LocalEnergyEvaluator full-basis sweep (src/localenergy.jl:94-155).
#include <algorithm>
#include <cub/cub.cuh>
#include <cstdlib>
#include "gwz_internal.h"
#include "gwz_steps.cuh"

namespace gwz {

// ------------------------------------------------------------------------------------------
// probe: one thread per address
// ------------------------------------------------------------------------------------------
template <int NW, int KC>
__global__ void __launch_bounds__(STEP_BLOCK)
probe_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ DevTables tb, int kind,
             const uint64_t* __restrict__ addr, size_t n, const double* __restrict__ u01,
             double* __restrict__ e_loc, double* __restrict__ tau_out, double* __restrict__ grad_out,
             int32_t* __restrict__ n_hops, double* __restrict__ rates, int32_t* __restrict__ chosen_out,
             uint64_t* __restrict__ next_addr, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StepScratch<NW, KC> sc(smem_raw, dm.N);
    sc.fill_exptab(dm.N, tb.c);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a[NW];
    load_addr_aos<NW>(addr, i, dm.W, a);
    const double u = u01 ? u01[i] : 0.5;
    double tau, eloc, gr;
    int nh, chosen = 0;
    if (kind == 1) {
        // production path; the reference hop index of the selected move is recovered afterwards
        step_bitparallel<NW, KC>(a, dm, tb, sc, u, tau, eloc, gr, nh, &chosen);
    } else {
        double* r = rates ? rates + i * (size_t)(2 * dm.M) : nullptr;
        if (r)
            for (int k = 0; k < 2 * dm.M; ++k) r[k] = 0.0;
        step_enumerate<NW>(a, dm, tb, u, tau, eloc, gr, nh, chosen, r);
    }
    if (!(tau < 1.0e300)) atomicOr(err_flag, 1);
    if (e_loc) e_loc[i] = eloc;
    if (tau_out) tau_out[i] = tau;
    if (grad_out) grad_out[i] = gr;
    if (n_hops) n_hops[i] = nh;
    if (chosen_out) chosen_out[i] = chosen;
    if (next_addr) store_addr_aos<NW>(next_addr, i, dm.W, a);
}

// GutzwillerAnsatz value and gradient (gutzwiller.jl:25-36): val = exp(-g*diag), der = -diag*val
template <int NW>
__global__ void __launch_bounds__(128)
ansatz_kernel(const __grid_constant__ DevModel dm, double g, const uint64_t* __restrict__ addr, size_t n,
              double* __restrict__ val, double* __restrict__ grad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a[NW];
    load_addr_aos<NW>(addr, i, dm.W, a);
    long long d = 0;
    enumerate_hops<NW>(a, dm, [&](int, int, bool right, int nsrc, int) {
        if (right) d += (long long)nsrc * (nsrc - 1) / 2;
        return false;
    });
    const double diag = dm.u * (double)d;
    const double v = exp(-g * diag);
    if (val) val[i] = v;
    if (grad) grad[i] = -diag * v;
}

// ------------------------------------------------------------------------------------------
// sweep
// ------------------------------------------------------------------------------------------
// per-state terms of localenergy.jl:105-120 in closed form.  With r_l = psi_l/psi_k,
// Emat = sum_l melem_l r_l = -t*E and sum_l melem_l r_l Delta_l = -t*ED:
//   num      = psi_k^2 (D + Emat)
//   den      = psi_k^2
//   num_grad = psi_k^2 (-2 D (D + Emat) + u t ED)       [grad psi = -D psi, D_l = D + u Delta_l]
//   den_grad = -2 D psi_k^2
template <int NW, int KC>
__device__ __forceinline__ void state_terms(const uint32_t (&a)[NW], const DevModel& dm, const DevTables& tb,
                                            const StepScratch<NW, KC>& sc, double& num, double& den, double& num_grad,
                                            double& den_grad) {
    uint32_t e[NW];
    extend<NW, KC>(a, dm, e);
    BondMasks<NW, KC> bm;
    classify_bonds<NW, KC>(e, dm, bm);
    BondSums bs;
    double bigS;
    const UnaryBig<NW, KC> big{e, bm.big, dm, sc};
    bond_sums<NW, KC, true>(tb, bm, NoCumSink{}, big, bs, bigS);
    const double D = dm.u * (double)bs.d;
    const double psi2 = exp(-2.0 * tb.g * D);  // abs2(exp(-g*diag))  (localenergy.jl:108,111)
    const double eloc = D - dm.t * bs.E;
    num = psi2 * eloc;
    den = psi2;
    num_grad = psi2 * (-2.0 * D * eloc + dm.u * dm.t * bs.ED);
    den_grad = -2.0 * D * psi2;
}

template <int NW, int KC, bool RANKED>
__global__ void __launch_bounds__(SWEEP_BLOCK)
sweep_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ DevTables tb,
             const uint64_t* __restrict__ basis, uint64_t dim, uint64_t rank_begin, uint64_t chunk,
             const uint64_t* __restrict__ binom, double* __restrict__ partials,
             const uint32_t* __restrict__ weight /* stored mode, optional: orbit sizes of a representative basis */,
             unsigned* __restrict__ done /* optional: fused reduction by the last CTA (small grids) */,
             double* __restrict__ out, double* __restrict__ mirror) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StepScratch<NW, KC> sc(smem_raw, dm.N, typename StepScratch<NW, KC>::SweepLayout{});
    sc.fill_exptab(dm.N, tb.c);
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (RANKED) {
        // each thread owns runs of `chunk` consecutive ranks: unrank once, then Gosper steps
        for (uint64_t first = tid * chunk; first < dim; first += nthreads * chunk) {
            const uint64_t cnt = min(chunk, dim - first);
            uint64_t x = unrank_colex(rank_begin + first, dm.N, dm.B, binom);
            for (uint64_t k = 0; k < cnt; ++k) {
                uint32_t a[NW];
                a[0] = (uint32_t)x;
                if (NW > 1) a[1] = (uint32_t)(x >> 32);
#pragma unroll
                for (int i = 2; i < NW; ++i) a[i] = 0u;
                double num, den, ng, dg;
                state_terms<NW, KC>(a, dm, tb, sc, num, den, ng, dg);
                s0 += num;
                s1 += den;
                s2 += ng;
                s3 += dg;
                x = next_same_popcount(x);
            }
        }
    } else {
        for (uint64_t idx = tid; idx < dim; idx += nthreads) {
            uint32_t a[NW];
            load_addr_soa<NW>(basis, dim, idx, dm.W, a);
            double num, den, ng, dg;
            state_terms<NW, KC>(a, dm, tb, sc, num, den, ng, dg);
            const double w = weight ? (double)weight[idx] : 1.0;
            s0 = fma(w, num, s0);
            s1 = fma(w, den, s1);
            s2 = fma(w, ng, s2);
            s3 = fma(w, dg, s3);
        }
    }
    __shared__ double sm[SWEEP_BLOCK / 32][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    s3 = warp_sum(s3);
    if (lane == 0) {
        sm[warp][0] = s0;
        sm[warp][1] = s1;
        sm[warp][2] = s2;
        sm[warp][3] = s3;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < SWEEP_BLOCK / 32; ++q) v += sm[q][threadIdx.x];
        partials[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
    }
    if (done == nullptr) return;
    // Fused reduction for small grids (repeated sweeps over a few thousand stored representatives are latency bound:
    // a second launch costs as much as the sweep).  The CTA that finishes last sums the per-block partials in a fixed
    // order (component = warp, blocks strided over the lanes, shuffle tree): bitwise reproducible.
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(done, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!last) return;
    double v = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32u) v += __ldcg(partials + (size_t)b * 4 + warp);
    v = warp_sum(v);
    if (lane == 0) {
        out[warp] = v;
        if (mirror) mirror[warp] = v;
    }
    if (threadIdx.x == 0) *done = 0u;
}

// Symmetry-reduced ranked sweep.  HubbardReal1D is a periodic chain and the Gutzwiller amplitude only depends on
// the multiset of occupations, so the four per-state terms are identical for the M cyclic shifts of an occupation
// vector and for their mirror images: evaluate one representative per orbit of the dihedral group and weight it
// with the orbit size.  On the unary string
// closed by a virtual separator (L = N+M bits, bit L-1 = 0) a mode shift is a cyclic rotation that brings a 0 to the
// top; the representative is the largest rotation.  Every thread still steps through its ranks (Gosper) and tests
// each state (a few rotations on average, early exit); representatives are compacted through a per-warp queue so
// that the expensive evaluation always runs with full warps.  Needs N+M <= 64.
// T = uint32_t when L <= 32 (half the instructions), else uint64_t.  The largest rotation has the longest run of
// ones at the top: reject as soon as a longer run exists anywhere (t+1 shift-and steps), then compare only against
// the rotations that bring an equally long run to the top (ties).
template <typename T>
__device__ __forceinline__ bool orbit_representative(T e, int L, T maskL, int M, int& mult) {
    constexpr int BITS = 8 * (int)sizeof(T);
    // t = ones directly below the top separator = occupation of mode M (t <= L-2 < BITS)
    const T inv = ~(T)(e << (BITS - L + 1));  // bit L-2 of e moved to the top, zeros shifted in below
    const int t = (sizeof(T) == 8) ? __clzll((long long)inv) : __clz((int)inv);
    // ge[k] has a 1 at the bottom bit of every run of >= k ones; build ge_t and ge_{t+1}
    T ge = ~(T)0;
    for (int k = 0; k < t; ++k) ge &= (T)(e >> k);
    if (ge & (T)(e >> t) & maskL) return false;  // a run of >= t+1 ones: that rotation is larger
    // zeros (below the top) that sit directly above a run of exactly t ones: bottoms of such runs are ge & e..., the zero
    // above a run whose bottom is at position b is at b + t
    T ties = (T)((t == 0 ? ~e : (T)(ge << t)) & ~e) & (T)(maskL >> 1);
    int same = 0;
    while (ties) {
        const int z = (sizeof(T) == 8) ? (__ffsll((long long)ties) - 1) : (__ffs((int)ties) - 1);
        ties &= ties - 1;
        const T r = (T)(((T)(e >> (z + 1)) | (T)(e << (L - 1 - z))) & maskL);
        if (r > e) return false;
        same += (r == e);
    }
    mult = same ? M / (same + 1) : M;  // the stabiliser is trivial for almost every orbit: keep the division off the common path
    // Reflection (n_1..n_M -> n_M..n_1 = the bit-reversed string) is a symmetry of the periodic chain and of the
    // Gutzwiller amplitude too: keep e only if it is also >= every rotation of its mirror image, and count the mirror
    // orbit in the weight.  Only the ~1/M survivors of the rotation test get here.
    const T rev = (T)((sizeof(T) == 8 ? (T)__brevll((unsigned long long)e) : (T)__brev((unsigned)e)) >> (BITS - L));
    // rev = 0 1^t 0 ... read from the bottom; bring the zero above that run to the top: a rotation with t ones on top
    const int sh = L - 2 - t;
    const T m = sh ? (T)(((T)(rev << sh) | (T)(rev >> (L - sh))) & maskL) : rev;
    if (m > e) return false;
    if (m < e) {
        // other rotations of the mirror image with a run of t on top (ties): any of them larger than e?
        T gm = ~(T)0;
        for (int k = 0; k < t; ++k) gm &= (T)(m >> k);
        T tm = (T)((t == 0 ? ~m : (T)(gm << t)) & ~m) & (T)(maskL >> 1);
        bool equal = false;
        while (tm) {
            const int z = (sizeof(T) == 8) ? (__ffsll((long long)tm) - 1) : (__ffs((int)tm) - 1);
            tm &= tm - 1;
            const T r = (T)(((T)(m >> (z + 1)) | (T)(m << (L - 1 - z))) & maskL);
            if (r > e) return false;
            equal = equal || (r == e);
        }
        if (!equal) mult *= 2;  // the mirror orbit is a different rotation orbit
    }
    return true;
}

constexpr int SYM_BINOM_SLOTS = 32 * 65;  // rows p = 0..31 of the [65][65] table (16.6 KB): systems with N+M-1 <= 31

template <int NW, int KC>
__global__ void __launch_bounds__(SWEEP_BLOCK)
sweep_sym_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ DevTables tb, uint64_t dim,
                 uint64_t rank_begin, uint64_t chunk, const uint64_t* __restrict__ binom, double* __restrict__ partials,
                 uint64_t extra_state, int has_extra) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StepScratch<NW, KC> sc(smem_raw, dm.N, typename StepScratch<NW, KC>::SweepLayout{});
    sc.fill_exptab(dm.N, tb.c);
    __shared__ uint64_t qx[SWEEP_BLOCK / 32][64];
    __shared__ int qm[SWEEP_BLOCK / 32][64];
    // the unranking walks the binomial table with dependent loads: keep the needed corner C(p, i), p <= B, i <= N,
    // in shared memory when it is small (short runs: the unranking latency is a large part of a small sweep)
    __shared__ uint64_t sbin[SYM_BINOM_SLOTS];
    const bool small_table = (dm.B + 1) * 65 <= SYM_BINOM_SLOTS;
    if (small_table) {
        for (int i = threadIdx.x; i < (dm.B + 1) * 65; i += blockDim.x) sbin[i] = binom[i];
        __syncthreads();
    }
    const uint64_t* __restrict__ bt = small_table ? sbin : binom;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    const int L = dm.N + dm.M;
    const uint64_t maskL = (L >= 64) ? ~0ull : ((1ull << L) - 1ull);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int qn = 0;  // warp-uniform fill of the queue
    auto evaluate = [&](uint64_t x, double w) {
        uint32_t a[NW];
        a[0] = (uint32_t)x;
        if (NW > 1) a[1] = (uint32_t)(x >> 32);
#pragma unroll
        for (int i = 2; i < NW; ++i) a[i] = 0u;
        double num, den, ng, dg;
        state_terms<NW, KC>(a, dm, tb, sc, num, den, ng, dg);
        s0 = fma(w, num, s0);
        s1 = fma(w, den, s1);
        s2 = fma(w, ng, s2);
        s3 = fma(w, dg, s3);
    };
    if (has_extra && tid == 0) evaluate(extra_state, 1.0);  // the uniform state at commensurate filling (see sweep_run)
    for (uint64_t base = (tid - lane) * chunk; base < dim; base += nthreads * chunk) {  // warp-uniform trip count
        const uint64_t first = base + (uint64_t)lane * chunk;
        const uint64_t cnt = first < dim ? min(chunk, dim - first) : 0;
        uint64_t x = cnt ? unrank_colex(rank_begin + first, dm.N, dm.B, bt) : 0ull;
        for (uint64_t k = 0; k < chunk; ++k) {
            int mult = 0;
            const bool rep = (k < cnt) && (L <= 32 ? orbit_representative<uint32_t>((uint32_t)x, L, (uint32_t)maskL, dm.M, mult)
                                                   : orbit_representative<uint64_t>(x, L, maskL, dm.M, mult));
            const unsigned b = __ballot_sync(0xffffffffu, rep);
            if (rep) {
                const int pos = qn + __popc(b & ((1u << lane) - 1u));
                qx[warp][pos] = x;
                qm[warp][pos] = mult;
            }
            qn += __popc(b);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                evaluate(qx[warp][qn + lane], (double)qm[warp][qn + lane]);
                __syncwarp();
            }
            if (k + 1 < cnt) x = (L <= 32) ? (uint64_t)next_same_popcount32((uint32_t)x) : next_same_popcount(x);
        }
    }
    if (lane < qn) evaluate(qx[warp][lane], (double)qm[warp][lane]);
    __shared__ double sm[SWEEP_BLOCK / 32][4];
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    s3 = warp_sum(s3);
    if (lane == 0) {
        sm[warp][0] = s0;
        sm[warp][1] = s1;
        sm[warp][2] = s2;
        sm[warp][3] = s3;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < SWEEP_BLOCK / 32; ++q) v += sm[q][threadIdx.x];
        partials[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
    }
}

// The orbit representatives of a rank range do not depend on the parameters: LocalEnergyEvaluator sweeps the same
// basis again and again (the reference stores build_basis(H) once, localenergy.jl:86), so they are collected once
// -- (state, orbit size) pairs in HBM -- and later sweeps only evaluate them.  store == nullptr: count only.
__global__ void __launch_bounds__(SWEEP_BLOCK)
sweep_collect_kernel(int N, int M, int B, uint64_t dim, uint64_t rank_begin, uint64_t chunk,
                     const uint64_t* __restrict__ binom, unsigned long long* __restrict__ counter,
                     uint64_t* __restrict__ store, uint32_t* __restrict__ mult_out) {
    const int lane = threadIdx.x & 31;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    const int L = N + M;
    const uint64_t maskL = (L >= 64) ? ~0ull : ((1ull << L) - 1ull);
    unsigned long long local = 0ull;
    for (uint64_t base = (tid - lane) * chunk; base < dim; base += nthreads * chunk) {  // warp-uniform trip count
        const uint64_t first = base + (uint64_t)lane * chunk;
        const uint64_t cnt = first < dim ? min(chunk, dim - first) : 0;
        uint64_t x = cnt ? unrank_colex(rank_begin + first, N, B, binom) : 0ull;
        for (uint64_t k = 0; k < chunk; ++k) {
            int mult = 0;
            const bool rep = (k < cnt) && (L <= 32 ? orbit_representative<uint32_t>((uint32_t)x, L, (uint32_t)maskL, M, mult)
                                                   : orbit_representative<uint64_t>(x, L, maskL, M, mult));
            if (store) {
                const unsigned b = __ballot_sync(0xffffffffu, rep);
                unsigned long long pos = 0ull;
                if (lane == 0 && b) pos = atomicAdd(counter, (unsigned long long)__popc(b));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (rep) {
                    const unsigned long long at = pos + __popc(b & ((1u << lane) - 1u));
                    store[at] = x;
                    mult_out[at] = (uint32_t)mult;
                }
            } else {
                local += rep ? 1ull : 0ull;
            }
            if (k + 1 < cnt) x = (L <= 32) ? (uint64_t)next_same_popcount32((uint32_t)x) : next_same_popcount(x);
        }
    }
    if (!store) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
        if (lane == 0 && local) atomicAdd(counter, local);
    }
}

cudaError_t launch_sweep_collect(int N, int M, int B, uint64_t dim, uint64_t rank_begin, int grid, uint64_t chunk,
                                 const uint64_t* binom, unsigned long long* counter, uint64_t* store, uint32_t* mult,
                                 cudaStream_t s) {
    sweep_collect_kernel<<<grid, SWEEP_BLOCK, 0, s>>>(N, M, B, dim, rank_begin, chunk, binom, counter, store, mult);
    return cudaGetLastError();
}

// sort the collected pairs by state (the collection order depends on the scheduling): a fixed summation order
cudaError_t sort_representatives(uint64_t* keys, uint32_t* vals, uint64_t n, cudaStream_t s) {
    if (n < 2) return cudaSuccess;
    if (n > 0x7fffffffull) return cudaErrorInvalidValue;
    uint64_t* k2 = nullptr;
    uint32_t* v2 = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cudaMalloc(&k2, sizeof(uint64_t) * n);
    if (e == cudaSuccess) e = cudaMalloc(&v2, sizeof(uint32_t) * n);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, k2, vals, v2, (int)n, 0, 64, s);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, k2, vals, v2, (int)n, 0, 64, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(keys, k2, sizeof(uint64_t) * n, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(vals, v2, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(k2);
    cudaFree(v2);
    cudaFree(tmp);
    return e;
}

// per-state outputs for parity tests: states first..first+count-1
template <int NW, int KC, bool RANKED>
__global__ void __launch_bounds__(SWEEP_BLOCK)
sweep_terms_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ DevTables tb,
                   const uint64_t* __restrict__ basis, uint64_t dim, uint64_t rank_begin,
                   const uint64_t* __restrict__ binom, uint64_t first, uint64_t count,
                   double* __restrict__ out, uint64_t* __restrict__ states_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StepScratch<NW, KC> sc(smem_raw, dm.N, typename StepScratch<NW, KC>::SweepLayout{});
    sc.fill_exptab(dm.N, tb.c);
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t a[NW];
    if (RANKED) {
        const uint64_t x = unrank_colex(rank_begin + first + i, dm.N, dm.B, binom);
        a[0] = (uint32_t)x;
        if (NW > 1) a[1] = (uint32_t)(x >> 32);
#pragma unroll
        for (int q = 2; q < NW; ++q) a[q] = 0u;
    } else {
        load_addr_soa<NW>(basis, dim, first + i, dm.W, a);
    }
    if (out) {
        double num, den, ng, dg;
        state_terms<NW, KC>(a, dm, tb, sc, num, den, ng, dg);
        out[i * 4 + 0] = num;
        out[i * 4 + 1] = den;
        out[i * 4 + 2] = ng;
        out[i * 4 + 3] = dg;
    }
    if (states_out) store_addr_aos<NW>(states_out, i, dm.W, a);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// opt in to > 48 KB of dynamic shared memory where the scratch needs it
template <typename K>
static cudaError_t ensure_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// CALL sees `NW` and `KC` as compile-time constants
#define GWZ_DISPATCH_KC(kc, ...)                              \
    if ((kc) == 5) { constexpr int KC = 5; __VA_ARGS__; }     \
    else { constexpr int KC = 4; __VA_ARGS__; }
#define GWZ_DISPATCH_NW(nw, kc, ...)                                               \
    switch (nw) {                                                                  \
        case 1: { constexpr int NW = 1; GWZ_DISPATCH_KC(kc, __VA_ARGS__) } break;  \
        case 2: { constexpr int NW = 2; GWZ_DISPATCH_KC(kc, __VA_ARGS__) } break;  \
        case 3: { constexpr int NW = 3; GWZ_DISPATCH_KC(kc, __VA_ARGS__) } break;  \
        case 4: { constexpr int NW = 4; GWZ_DISPATCH_KC(kc, __VA_ARGS__) } break;  \
        case 5: { constexpr int NW = 5; GWZ_DISPATCH_KC(kc, __VA_ARGS__) } break;  \
        case 6: { constexpr int NW = 6; GWZ_DISPATCH_KC(kc, __VA_ARGS__) } break;  \
        case 7: { constexpr int NW = 7; GWZ_DISPATCH_KC(kc, __VA_ARGS__) } break;  \
        case 8: { constexpr int NW = 8; GWZ_DISPATCH_KC(kc, __VA_ARGS__) } break;  \
        default: return cudaErrorInvalidValue;                                     \
    }

cudaError_t launch_probe(const ProbeLaunch& L, cudaStream_t s) {
    if (L.n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((L.n + STEP_BLOCK - 1) / STEP_BLOCK);
    GWZ_DISPATCH_NW(L.nw, L.tb.kc, {
        const size_t smem = StepScratch<NW, KC>::bytes(L.dm.N);
        if (cudaError_t e = ensure_smem(probe_kernel<NW, KC>, smem)) return e;
        probe_kernel<NW, KC><<<grid, STEP_BLOCK, smem, s>>>(L.dm, L.tb, L.kind, L.addr, L.n, L.u01, L.e_loc, L.tau, L.grad,
                                                            L.n_hops, L.rates, L.chosen, L.next_addr, L.err_flag);
    });
    return cudaGetLastError();
}

cudaError_t launch_ansatz(const DevModel& dm, double g, int nw, const uint64_t* addr, size_t n, double* val,
                          double* grad, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 127) / 128);
    GWZ_DISPATCH_NW(nw, 4, { if (KC == 4) ansatz_kernel<NW><<<grid, 128, 0, s>>>(dm, g, addr, n, val, grad); });
    return cudaGetLastError();
}

#ifndef GWZ_SWEEP_WAVE_CTAS
#define GWZ_SWEEP_WAVE_CTAS 32
#endif
cudaError_t sweep_grid(int nw, int kc, int N, int M, bool ranked, uint64_t dim, int sm_count, int* grid, uint64_t* chunk) {
    (void)nw; (void)kc; (void)N; (void)M;
    uint64_t c = 1;
    uint64_t max_threads = (uint64_t)sm_count * 16 * SWEEP_BLOCK;
    if (ranked) {
        // Every thread Gosper-steps through runs of `chunk` consecutive ranks (the unranking is amortised over the
        // run).  The cost of a rank varies along the rank range (the orbit test exits early for most low ranks), so
        // runs are kept short and CTAs many: the block scheduler balances them (measured: one resident wave of long
        // runs is 1.8x slower on {16,16} and 3.4x on {18,18}).
        const uint64_t wave = (uint64_t)sm_count * GWZ_SWEEP_WAVE_CTAS * SWEEP_BLOCK;  // short runs, several waves: the block scheduler balances the rank-dependent cost
        c = (dim + wave - 1) / wave;
        // run length: >= 8 amortises the unranking; small sweeps (BASELINE config 2 after pruning: 3.5e5 ranks) are
        // latency bound and do better with more, shorter runs (measured: {12,12} kernel 32 -> 24 us)
        uint64_t cmin = dim < 1500000ull ? 4 : 8;
        if (const char* e = std::getenv("GWZ_SWEEP_MIN_CHUNK")) cmin = (uint64_t)std::max(1, std::atoi(e));
        c = c < cmin ? cmin : (c > 256 ? 256 : c);
        max_threads = (uint64_t)sm_count * 4096 * SWEEP_BLOCK;  // beyond that threads take several runs
    }
    uint64_t threads = (dim + c - 1) / c;
    if (threads > max_threads) threads = max_threads;
    uint64_t g = (threads + SWEEP_BLOCK - 1) / SWEEP_BLOCK;
    if (g < 1) g = 1;
    *grid = (int)g;
    *chunk = c;
    return cudaSuccess;
}

cudaError_t launch_sweep(const SweepLaunch& L, cudaStream_t s) {
    const bool ranked = (L.basis == nullptr);
    GWZ_DISPATCH_NW(L.nw, L.tb.kc, {
        const size_t smem = StepScratch<NW, KC>::bytes_sweep(L.dm.N);
        if (ranked && L.symmetric && L.dm.N + L.dm.M <= 64) {
            if (cudaError_t e = ensure_smem(sweep_sym_kernel<NW, KC>, smem)) return e;
            sweep_sym_kernel<NW, KC><<<L.grid, SWEEP_BLOCK, smem, s>>>(L.dm, L.tb, L.dim, L.rank_begin, L.chunk, L.binom,
                                                                      L.partials, L.extra_state, L.has_extra);
        } else if (ranked) {
            if (cudaError_t e = ensure_smem(sweep_kernel<NW, KC, true>, smem)) return e;
            sweep_kernel<NW, KC, true><<<L.grid, SWEEP_BLOCK, smem, s>>>(L.dm, L.tb, L.basis, L.dim, L.rank_begin, L.chunk,
                                                                         L.binom, L.partials, nullptr, L.done, L.out,
                                                                         L.mirror);
        } else {
            if (cudaError_t e = ensure_smem(sweep_kernel<NW, KC, false>, smem)) return e;
            sweep_kernel<NW, KC, false><<<L.grid, SWEEP_BLOCK, smem, s>>>(L.dm, L.tb, L.basis, L.dim, L.rank_begin, L.chunk,
                                                                          L.binom, L.partials, L.weight, L.done, L.out,
                                                                          L.mirror);
        }
    });
    return cudaGetLastError();
}

cudaError_t launch_sweep_terms(const SweepLaunch& L, cudaStream_t s) {
    if (L.count == 0) return cudaSuccess;
    const bool ranked = (L.basis == nullptr);
    const unsigned grid = (unsigned)((L.count + SWEEP_BLOCK - 1) / SWEEP_BLOCK);
    GWZ_DISPATCH_NW(L.nw, L.tb.kc, {
        const size_t smem = StepScratch<NW, KC>::bytes_sweep(L.dm.N);
        if (ranked) {
            if (cudaError_t e = ensure_smem(sweep_terms_kernel<NW, KC, true>, smem)) return e;
            sweep_terms_kernel<NW, KC, true><<<grid, SWEEP_BLOCK, smem, s>>>(L.dm, L.tb, L.basis, L.dim, L.rank_begin, L.binom,
                                                                             L.first, L.count, L.terms_out, L.states_out);
        } else {
            if (cudaError_t e = ensure_smem(sweep_terms_kernel<NW, KC, false>, smem)) return e;
            sweep_terms_kernel<NW, KC, false><<<grid, SWEEP_BLOCK, smem, s>>>(L.dm, L.tb, L.basis, L.dim, L.rank_begin,
                                                                              L.binom, L.first, L.count, L.terms_out,
                                                                              L.states_out);
        }
    });
    return cudaGetLastError();
}

}  // namespace gwz
