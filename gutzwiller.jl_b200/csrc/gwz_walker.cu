// gwz_walker.cu -- persistent kinetic-VQMC walker kernel (one walker per thread).
//
// Implements kinetic_vqmc!(st; steps) (src/kinetic-vqmc/qmc.jl:73-97) for every walker in
// parallel (the reference's Threads.@threads loop, qmc.jl:99-106), with kinetic_sample!
// (qmc.jl:14-44) evaluated in closed form for HubbardReal1D + GutzwillerAnsatz, and the
// per-walker estimator of state.jl:42-49 kept as running sums instead of per-step arrays.
#include "gwz_internal.h"
#include "gwz_steps.cuh"

namespace gwz {

// dynamic shared memory: [selection scratch (production kernel) | per-thread partial vectors]; the
// block-reduction buffer aliases the scratch region
// KIND: 0 = enumerate (reference hop order); 4 / 5 = bit-parallel with KC = KIND bond classes per side
template <int NW, int KIND>
__host__ __device__ constexpr size_t walker_scratch_bytes(int N) {
    return KIND != 0 ? StepScratch<NW, (KIND ? KIND : 4)>::bytes(N) : sizeof(double) * (WALKER_BLOCK / 32) * NUM_ACC;
}
template <int NW, int KIND>
__host__ __device__ constexpr size_t walker_smem_bytes(int N) {
    return walker_scratch_bytes<NW, KIND>(N) + sizeof(double) * NUM_ACC * WALKER_BLOCK;
}
#ifndef GWZ_WALKER_MIN_BLOCKS_WIDE
#define GWZ_WALKER_MIN_BLOCKS_WIDE 3
#endif
#ifndef GWZ_WALKER_MIN_BLOCKS
#define GWZ_WALKER_MIN_BLOCKS 4
#endif

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int NW, int KIND, bool RECORD>
__global__ void __launch_bounds__(WALKER_BLOCK, KIND == 0 ? 1 : ((NW <= 4 ? GWZ_WALKER_MIN_BLOCKS : GWZ_WALKER_MIN_BLOCKS_WIDE) - (KIND == 5 ? 1 : 0)))
walker_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ DevTables tb,
              uint64_t* __restrict__ addr, double* __restrict__ acc, size_t n, uint64_t global_offset,
              uint64_t seed, uint64_t step0, uint64_t steps, int accumulate, int reset_acc,
              double shift_e, double shift_g, double* __restrict__ rec_tau, double* __restrict__ rec_eloc,
              uint64_t* __restrict__ rec_addr, double* __restrict__ partials, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int KC = KIND ? KIND : 4;
    const StepScratch<NW, KC> sc(smem_raw, dm.N);
    if (KIND != 0) sc.fill_exptab(dm.N, tb.c);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    // common reference point of the running sums (any constant works; it only limits cancellation
    // in the covariance): E_loc and grad ratio of a representative address, see walker_ref_kernel
    const double E0 = accumulate ? shift_e : 0.0, G0 = accumulate ? shift_g : 0.0;

    // this thread's partial accumulator vector lives in shared memory (touched once per walker, not per
    // step) so that it does not occupy 24 registers across the step loop
    double* p = reinterpret_cast<double*>(smem_raw + walker_scratch_bytes<NW, KIND>(dm.N)) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NUM_ACC; ++i) p[i * WALKER_BLOCK] = 0.0;
    bool bad = false;

    for (size_t w = tid; w < n; w += nthreads) {
        uint32_t a[NW];
        load_addr_soa<NW>(addr, n, w, dm.W, a);
        double st, ste, stg, stge, stee;
        if (reset_acc || !accumulate) {
            st = ste = stg = stge = stee = 0.0;
        } else {
            st = acc[0 * n + w];
            ste = acc[1 * n + w];
            stg = acc[2 * n + w];
            stge = acc[3 * n + w];
            stee = acc[4 * n + w];
        }
        const uint64_t wid = global_offset + (uint64_t)w;
        const uint32_t c2 = (uint32_t)wid, c3 = (uint32_t)(wid >> 32);
        uint32_t rnd[4] = {0u, 0u, 0u, 0u};
        unsigned long long hops = 0ull;
        double tcheck = 0.0;

        for (uint64_t k = 0; k < steps; ++k) {
            const uint64_t step = step0 + k;
            if (k == 0 || (step & 1ull) == 0ull) {
                const uint64_t blk = step >> 1;
                philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), c2, c3, k0, k1, rnd);
            }
            const double u01 = (step & 1ull) ? u01_from_words(rnd[2], rnd[3]) : u01_from_words(rnd[0], rnd[1]);
            if (RECORD && rec_addr) store_addr_aos<NW>(rec_addr, (size_t)k * n + w, dm.W, a);
            double tau, eloc, gr;
            int nh;
            if (KIND != 0) {
                step_bitparallel<NW, KC>(a, dm, tb, sc, u01, tau, eloc, gr, nh);
            } else {
                int chosen;
                step_enumerate<NW>(a, dm, tb, u01, tau, eloc, gr, nh, chosen, nullptr);
            }
            hops += (unsigned)nh;
            if (RECORD) {
                rec_tau[(size_t)k * n + w] = tau;
                rec_eloc[(size_t)k * n + w] = eloc;
            }
            // shifted running sums: weights = residence times (state.jl:43-46)
            const double dE = eloc - E0, dG = gr - G0;
            const double tE = tau * dE;
            tcheck += tau;
            st += tau;
            ste += tE;
            stg = fma(tau, dG, stg);
            stge = fma(tE, dG, stge);
            stee = fma(tE, dE, stee);
        }
        bad |= !(tcheck < 1.0e300);  // isfinite(residence_time) for every step  (qmc.jl:34-36)

        store_addr_soa<NW>(addr, n, w, dm.W, a);
        if (accumulate) {
            acc[0 * n + w] = st;
            acc[1 * n + w] = ste;
            acc[2 * n + w] = stg;
            acc[3 * n + w] = stge;
            acc[4 * n + w] = stee;
            // per-walker estimates (state.jl:42-49)
            const double inv = 1.0 / st;
            const double mE = ste * inv, mG = stg * inv;
            const double val_w = E0 + mE;
            const double grad_w = 2.0 * (stge * inv - mE * mG);
            p[ACC_WALKERS * WALKER_BLOCK] += 1.0;
            p[ACC_VAL * WALKER_BLOCK] += val_w;
            p[ACC_VAL2 * WALKER_BLOCK] = fma(val_w, val_w, p[ACC_VAL2 * WALKER_BLOCK]);
            p[ACC_GRAD * WALKER_BLOCK] += grad_w;
            p[ACC_GRAD2 * WALKER_BLOCK] = fma(grad_w, grad_w, p[ACC_GRAD2 * WALKER_BLOCK]);
            // pooled sums, un-shifted
            p[ACC_T * WALKER_BLOCK] += st;
            p[ACC_TE * WALKER_BLOCK] += st * val_w;
            p[ACC_TG * WALKER_BLOCK] += st * (G0 + mG);
            p[ACC_TGE * WALKER_BLOCK] += st * G0 * E0 + G0 * ste + E0 * stg + stge;
            p[ACC_TEE * WALKER_BLOCK] += st * E0 * E0 + 2.0 * E0 * ste + stee;
            p[ACC_HOPS * WALKER_BLOCK] += (double)hops;
            p[ACC_STEPS * WALKER_BLOCK] += (double)steps;
            p[ACC_VARW * WALKER_BLOCK] += stee * inv - mE * mE;  // var(E_loc, weights tau), uncorrected
        }
    }

    if (bad) atomicOr(err_flag, 1);
    if (!accumulate || partials == nullptr) return;

    // block reduction in fixed order: lanes -> warps -> block (reuses the per-thread scratch region)
    __syncthreads();
    double(*sm)[NUM_ACC] = reinterpret_cast<double(*)[NUM_ACC]>(KIND != 0 ? (unsigned char*)sc.cum : smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NUM_ACC; ++i) {
        const double v = warp_sum(p[i * WALKER_BLOCK]);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NUM_ACC) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < WALKER_BLOCK / 32; ++q) v += sm[q][threadIdx.x];
        partials[(size_t)blockIdx.x * NUM_ACC + threadIdx.x] = v;
    }
}

// E_loc and grad ratio of walker 0's current address -> shift[0..1] (reference point of the sums)
template <int NW>
__global__ void walker_ref_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ DevTables tb,
                                  const uint64_t* __restrict__ addr, size_t n, double* __restrict__ shift) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t a[NW];
    load_addr_soa<NW>(addr, n, 0, dm.W, a);
    double tau, eloc, gr;
    int nh, chosen;
    step_enumerate<NW>(a, dm, tb, 0.5, tau, eloc, gr, nh, chosen, nullptr);
    shift[0] = eloc;
    shift[1] = gr;
}

// ------------------------------------------------------------------------------------------
// small helper kernels
// ------------------------------------------------------------------------------------------
// out[comp] (device) and, when mirror != nullptr, mirror[comp] (pinned host memory, written over the bus so that the
// host needs no copy after the stream is idle); block 0 also mirrors the error flag.
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nblocks, int ncomp,
                                       double* __restrict__ out, double* __restrict__ mirror,
                                       const int* __restrict__ err_dev, int* __restrict__ err_host) {
    // one warp per component, fixed summation order
    __shared__ double sm[32];
    const int comp = blockIdx.x;
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) v += partials[(size_t)b * ncomp + comp];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += sm[q];
        out[comp] = t;
        if (mirror) mirror[comp] = t;
        if (comp == 0 && err_host) *err_host = *err_dev;
    }
}

__global__ void walker_init_kernel(uint64_t* __restrict__ addr, double* __restrict__ acc, size_t n, int W,
                                   const uint64_t* __restrict__ start) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < W; ++w) addr[(size_t)w * n + i] = start[w];  // starting_address(H) (state.jl:26)
    for (int c = 0; c < NUM_WALKER_ACC; ++c) acc[(size_t)c * n + i] = 0.0;
}

__global__ void walker_estimates_kernel(const double* __restrict__ acc, size_t n, double shift_e,
                                        double* __restrict__ val_w, double* __restrict__ grad_w) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double st = acc[0 * n + i], ste = acc[1 * n + i], stg = acc[2 * n + i], stge = acc[3 * n + i];
    const double inv = 1.0 / st;
    const double mE = ste * inv, mG = stg * inv;
    val_w[i] = shift_e + mE;
    grad_w[i] = 2.0 * (stge * inv - mE * mG);
}

__global__ void transpose_addr_kernel(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, size_t n,
                                      int W, int to_soa) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < W; ++w) {
        if (to_soa) dst[(size_t)w * n + i] = src[i * (size_t)W + w];
        else dst[i * (size_t)W + w] = src[(size_t)w * n + i];
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int NW, int KIND, bool RECORD>
static cudaError_t walker_prepare(int N) {
    // opt in to > 48 KB of dynamic shared memory where the scratch needs it
    if (walker_smem_bytes<NW, KIND>(N) <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(walker_kernel<NW, KIND, RECORD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)walker_smem_bytes<NW, KIND>(N));
}
template <int NW, int KIND, bool RECORD>
static cudaError_t launch_one(const WalkerLaunch& L, cudaStream_t s) {
    if (cudaError_t e = walker_prepare<NW, KIND, RECORD>(L.dm.N); e != cudaSuccess) return e;
    walker_kernel<NW, KIND, RECORD><<<L.grid, WALKER_BLOCK, walker_smem_bytes<NW, KIND>(L.dm.N), s>>>(
        L.dm, L.tb, L.addr, L.acc, L.n, L.global_offset, L.seed, L.step0, L.steps, L.accumulate, L.reset_acc,
        L.shift_e, L.shift_g, L.rec_tau, L.rec_eloc, L.rec_addr, L.partials, L.err_flag);
    return cudaGetLastError();
}
template <int NW>
static cudaError_t launch_nw(const WalkerLaunch& L, cudaStream_t s) {
    const bool rec = (L.rec_tau != nullptr);
    if (L.kind == 1 && L.tb.kc == 5) return rec ? launch_one<NW, 5, true>(L, s) : launch_one<NW, 5, false>(L, s);
    if (L.kind == 1) return rec ? launch_one<NW, 4, true>(L, s) : launch_one<NW, 4, false>(L, s);
    return rec ? launch_one<NW, 0, true>(L, s) : launch_one<NW, 0, false>(L, s);
}
cudaError_t launch_walkers(const WalkerLaunch& L, cudaStream_t s) {
    switch (L.nw) {
        case 1: return launch_nw<1>(L, s);
        case 2: return launch_nw<2>(L, s);
        case 3: return launch_nw<3>(L, s);
        case 4: return launch_nw<4>(L, s);
        case 5: return launch_nw<5>(L, s);
        case 6: return launch_nw<6>(L, s);
        case 7: return launch_nw<7>(L, s);
        case 8: return launch_nw<8>(L, s);
    }
    return cudaErrorInvalidValue;
}

template <int NW, int KIND, bool RECORD>
static cudaError_t occ_one(int N, int* blocks) {
    if (cudaError_t e = walker_prepare<NW, KIND, RECORD>(N); e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, walker_kernel<NW, KIND, RECORD>, WALKER_BLOCK,
                                                         walker_smem_bytes<NW, KIND>(N));
}
template <int NW>
static cudaError_t occ_nw(int N, int kind, int kc, bool rec, int* blocks) {
    if (kind == 1 && kc == 5) return rec ? occ_one<NW, 5, true>(N, blocks) : occ_one<NW, 5, false>(N, blocks);
    if (kind == 1) return rec ? occ_one<NW, 4, true>(N, blocks) : occ_one<NW, 4, false>(N, blocks);
    return rec ? occ_one<NW, 0, true>(N, blocks) : occ_one<NW, 0, false>(N, blocks);
}
cudaError_t walker_grid(int nw, int N, int kind, int kc, bool record, size_t n, int sm_count, int* grid) {
    int blocks = 0;
    cudaError_t e = cudaErrorInvalidValue;
    switch (nw) {
        case 1: e = occ_nw<1>(N, kind, kc, record, &blocks); break;
        case 2: e = occ_nw<2>(N, kind, kc, record, &blocks); break;
        case 3: e = occ_nw<3>(N, kind, kc, record, &blocks); break;
        case 4: e = occ_nw<4>(N, kind, kc, record, &blocks); break;
        case 5: e = occ_nw<5>(N, kind, kc, record, &blocks); break;
        case 6: e = occ_nw<6>(N, kind, kc, record, &blocks); break;
        case 7: e = occ_nw<7>(N, kind, kc, record, &blocks); break;
        case 8: e = occ_nw<8>(N, kind, kc, record, &blocks); break;
    }
    if (e != cudaSuccess) return e;
    if (blocks < 1) blocks = 1;
    // persistent grid: one full wave of resident CTAs (multiple of the SM count), fewer if the
    // walkers do not fill it
    size_t resident = (size_t)blocks * (size_t)sm_count;
    size_t need = (n + WALKER_BLOCK - 1) / WALKER_BLOCK;
    *grid = (int)(need < resident ? need : resident);
    if (*grid < 1) *grid = 1;
    return cudaSuccess;
}

cudaError_t launch_walker_ref(const DevModel& dm, const DevTables& tb, int nw, const uint64_t* addr, size_t n,
                              double* shift, cudaStream_t s) {
    switch (nw) {
        case 1: walker_ref_kernel<1><<<1, 32, 0, s>>>(dm, tb, addr, n, shift); break;
        case 2: walker_ref_kernel<2><<<1, 32, 0, s>>>(dm, tb, addr, n, shift); break;
        case 3: walker_ref_kernel<3><<<1, 32, 0, s>>>(dm, tb, addr, n, shift); break;
        case 4: walker_ref_kernel<4><<<1, 32, 0, s>>>(dm, tb, addr, n, shift); break;
        case 5: walker_ref_kernel<5><<<1, 32, 0, s>>>(dm, tb, addr, n, shift); break;
        case 6: walker_ref_kernel<6><<<1, 32, 0, s>>>(dm, tb, addr, n, shift); break;
        case 7: walker_ref_kernel<7><<<1, 32, 0, s>>>(dm, tb, addr, n, shift); break;
        case 8: walker_ref_kernel<8><<<1, 32, 0, s>>>(dm, tb, addr, n, shift); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const double* partials, int nblocks, int ncomp, double* out, cudaStream_t s,
                                   double* mirror, const int* err_dev, int* err_host) {
    reduce_partials_kernel<<<ncomp, nblocks > 64 ? 256 : 64, 0, s>>>(partials, nblocks, ncomp, out, mirror, err_dev, err_host);
    return cudaGetLastError();
}
cudaError_t launch_walker_init(uint64_t* addr, double* acc, size_t n, int W, const uint64_t* start_dev,
                               cudaStream_t s) {
    walker_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(addr, acc, n, W, start_dev);
    return cudaGetLastError();
}
cudaError_t launch_walker_estimates(const double* acc, size_t n, double shift_e, double* val_w, double* grad_w,
                                    cudaStream_t s) {
    walker_estimates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(acc, n, shift_e, val_w, grad_w);
    return cudaGetLastError();
}
cudaError_t launch_transpose_addr(const uint64_t* src, uint64_t* dst, size_t n, int W, bool to_soa, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    transpose_addr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, dst, n, W, to_soa ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace gwz
