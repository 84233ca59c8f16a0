// gwz_walker_planes.cu -- production walker kernel on the occupation bit-plane representation
// (see gwz_planes.cuh).  Same contract as walker_kernel<NW, 4, false> in gwz_walker.cu:
// kinetic_vqmc!(st; steps) (src/kinetic-vqmc/qmc.jl:73-97) for every walker, running sums of
// state.jl:42-49, per-block partial accumulator vectors.
#include "gwz_internal.h"
#include "gwz_planes.cuh"

#ifndef GWZ_PLANES_MIN_BLOCKS
#define GWZ_PLANES_MIN_BLOCKS 4
#endif

namespace gwz {

template <int NWM, int NP>
__host__ __device__ constexpr size_t planes_stage_bytes() {
    return 0;  // the generic-bond path of the plane representation works from registers
}
template <int NWM, int NP>
__host__ __device__ constexpr size_t planes_smem_bytes(int N) {
    return StepScratch<NWM, 4>::bytes(N) + planes_stage_bytes<NWM, NP>() + sizeof(double) * NUM_ACC * STEP_BLOCK;
}

template <int NWM, int NP>
__global__ void __launch_bounds__(STEP_BLOCK, NWM == 1 ? 5 : (NWM == 2 ? GWZ_PLANES_MIN_BLOCKS : 3))
walker_planes_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ DevTables tb,
                     uint32_t* __restrict__ planes, double* __restrict__ acc, size_t n, uint64_t global_offset,
                     uint64_t seed, uint64_t step0, uint64_t steps, int accumulate, int reset_acc,
                     double shift_e, double shift_g, double* __restrict__ partials, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StepScratch<NWM, 4> sc(smem_raw, dm.N);
    sc.fill_exptab(dm.N, tb.c);
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw + StepScratch<NWM, 4>::bytes(dm.N)) + threadIdx.x;
    double* p = reinterpret_cast<double*>(smem_raw + StepScratch<NWM, 4>::bytes(dm.N) + planes_stage_bytes<NWM, NP>()) +
                threadIdx.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const double E0 = accumulate ? shift_e : 0.0, G0 = accumulate ? shift_g : 0.0;
#pragma unroll
    for (int i = 0; i < NUM_ACC; ++i) p[i * STEP_BLOCK] = 0.0;
    bool bad = false;

    for (size_t w = tid; w < n; w += nthreads) {
        uint32_t pl[NP][NWM];
#pragma unroll
        for (int k = 0; k < NP; ++k)
#pragma unroll
            for (int i = 0; i < NWM; ++i) pl[k][i] = planes[((size_t)k * NWM + i) * n + w];
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
            double tau, eloc, gr;
            int nh;
            step_planes<NWM, NP>(pl, dm, tb, sc, stage, u01, tau, eloc, gr, nh);
            hops += (unsigned)nh;
            const double dE = eloc - E0, dG = gr - G0;
            const double tE = tau * dE;
            tcheck += tau;
            st += tau;
            ste += tE;
            stg = fma(tau, dG, stg);
            stge = fma(tE, dG, stge);
            stee = fma(tE, dE, stee);
        }
        bad |= !(tcheck < 1.0e300);

#pragma unroll
        for (int k = 0; k < NP; ++k)
#pragma unroll
            for (int i = 0; i < NWM; ++i) planes[((size_t)k * NWM + i) * n + w] = pl[k][i];
        if (accumulate) {
            acc[0 * n + w] = st;
            acc[1 * n + w] = ste;
            acc[2 * n + w] = stg;
            acc[3 * n + w] = stge;
            acc[4 * n + w] = stee;
            const double inv = 1.0 / st;
            const double mE = ste * inv, mG = stg * inv;
            const double val_w = E0 + mE;
            const double grad_w = 2.0 * (stge * inv - mE * mG);
            p[ACC_WALKERS * STEP_BLOCK] += 1.0;
            p[ACC_VAL * STEP_BLOCK] += val_w;
            p[ACC_VAL2 * STEP_BLOCK] = fma(val_w, val_w, p[ACC_VAL2 * STEP_BLOCK]);
            p[ACC_GRAD * STEP_BLOCK] += grad_w;
            p[ACC_GRAD2 * STEP_BLOCK] = fma(grad_w, grad_w, p[ACC_GRAD2 * STEP_BLOCK]);
            p[ACC_T * STEP_BLOCK] += st;
            p[ACC_TE * STEP_BLOCK] += st * val_w;
            p[ACC_TG * STEP_BLOCK] += st * (G0 + mG);
            p[ACC_TGE * STEP_BLOCK] += st * G0 * E0 + G0 * ste + E0 * stg + stge;
            p[ACC_TEE * STEP_BLOCK] += st * E0 * E0 + 2.0 * E0 * ste + stee;
            p[ACC_HOPS * STEP_BLOCK] += (double)hops;
            p[ACC_STEPS * STEP_BLOCK] += (double)steps;
            p[ACC_VARW * STEP_BLOCK] += stee * inv - mE * mE;  // var(E_loc, weights tau), uncorrected
        }
    }

    if (bad) atomicOr(err_flag, 1);
    if (!accumulate || partials == nullptr) return;
    __syncthreads();
    double(*sm)[NUM_ACC] = reinterpret_cast<double(*)[NUM_ACC]>(sc.cum);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NUM_ACC; ++i) {
        const double v = warp_sum(p[i * STEP_BLOCK]);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NUM_ACC) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < STEP_BLOCK / 32; ++q) v += sm[q][threadIdx.x];
        partials[(size_t)blockIdx.x * NUM_ACC + threadIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------------
// layout conversion (rare: when walkers enter / leave the plane representation)
// ------------------------------------------------------------------------------------------
// unary SoA addr[W][n] -> planes[(k*nwm + i)*n + w]
__global__ void unary_to_planes_kernel(const uint64_t* __restrict__ addr, size_t n, int W, int B, int M, int nwm, int np,
                                       uint32_t* __restrict__ planes) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    uint32_t pl[8][8];
    for (int k = 0; k < 8; ++k)
        for (int i = 0; i < 8; ++i) pl[k][i] = 0u;
    int mode = 0, cnt = 0;
    for (int p = 0; p <= B; ++p) {
        const int bit = (p < B) ? (int)((addr[(size_t)(p >> 6) * n + w] >> (p & 63)) & 1ull) : 0;
        if (bit) {
            ++cnt;
        } else {
            if (mode < M)
                for (int k = 0; k < np; ++k)
                    if ((cnt >> k) & 1) pl[k][mode >> 5] |= 1u << (mode & 31);
            ++mode;
            cnt = 0;
        }
    }
    for (int k = 0; k < np; ++k)
        for (int i = 0; i < nwm; ++i) planes[((size_t)k * nwm + i) * n + w] = pl[k][i];
    (void)W;
}

// planes -> unary SoA addr[W][n_out] for walkers first..first+count-1
__global__ void planes_to_unary_kernel(const uint32_t* __restrict__ planes, size_t n, int W, int M, int nwm, int np,
                                       size_t first, size_t count, uint64_t* __restrict__ addr) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const size_t w = first + j;
    uint64_t a[4] = {0ull, 0ull, 0ull, 0ull};
    int pos = 0;
    for (int mode = 0; mode < M; ++mode) {
        int occ = 0;
        for (int k = 0; k < np; ++k) occ |= (int)((planes[((size_t)k * nwm + (mode >> 5)) * n + w] >> (mode & 31)) & 1u) << k;
        for (int q = 0; q < occ; ++q, ++pos) a[pos >> 6] |= 1ull << (pos & 63);
        ++pos;
    }
    for (int i = 0; i < W; ++i) addr[(size_t)i * count + j] = a[i];
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int NWM, int NP>
static cudaError_t planes_prepare(int N) {
    const size_t smem = planes_smem_bytes<NWM, NP>(N);
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(walker_planes_kernel<NWM, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <int NWM, int NP>
static cudaError_t planes_launch(const WalkerLaunch& L, cudaStream_t s) {
    if (cudaError_t e = planes_prepare<NWM, NP>(L.dm.N)) return e;
    walker_planes_kernel<NWM, NP><<<L.grid, STEP_BLOCK, planes_smem_bytes<NWM, NP>(L.dm.N), s>>>(
        L.dm, L.tb, L.planes, L.acc, L.n, L.global_offset, L.seed, L.step0, L.steps, L.accumulate, L.reset_acc, L.shift_e, L.shift_g,
        L.partials, L.err_flag);
    return cudaGetLastError();
}
template <int NWM, int NP>
static cudaError_t planes_occ(int N, int* blocks) {
    if (cudaError_t e = planes_prepare<NWM, NP>(N)) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, walker_planes_kernel<NWM, NP>, STEP_BLOCK,
                                                         planes_smem_bytes<NWM, NP>(N));
}

#define GWZ_PLANES_DISPATCH(nwm, np, FN, ...)                                  \
    switch ((nwm) * 16 + (np)) {                                               \
        case 1 * 16 + 4: return FN<1, 4>(__VA_ARGS__);                         \
        case 1 * 16 + 6: return FN<1, 6>(__VA_ARGS__);                         \
        case 1 * 16 + 8: return FN<1, 8>(__VA_ARGS__);                         \
        case 2 * 16 + 4: return FN<2, 4>(__VA_ARGS__);                         \
        case 2 * 16 + 6: return FN<2, 6>(__VA_ARGS__);                         \
        case 2 * 16 + 8: return FN<2, 8>(__VA_ARGS__);                         \
        case 3 * 16 + 4: return FN<3, 4>(__VA_ARGS__);                         \
        case 3 * 16 + 6: return FN<3, 6>(__VA_ARGS__);                         \
        case 3 * 16 + 8: return FN<3, 8>(__VA_ARGS__);                         \
        case 4 * 16 + 4: return FN<4, 4>(__VA_ARGS__);                         \
        case 4 * 16 + 6: return FN<4, 6>(__VA_ARGS__);                         \
        case 4 * 16 + 8: return FN<4, 8>(__VA_ARGS__);                         \
        case 5 * 16 + 4: return FN<5, 4>(__VA_ARGS__);                         \
        case 5 * 16 + 6: return FN<5, 6>(__VA_ARGS__);                         \
        case 5 * 16 + 8: return FN<5, 8>(__VA_ARGS__);                         \
        case 6 * 16 + 4: return FN<6, 4>(__VA_ARGS__);                         \
        case 6 * 16 + 6: return FN<6, 6>(__VA_ARGS__);                         \
        case 6 * 16 + 8: return FN<6, 8>(__VA_ARGS__);                         \
        case 7 * 16 + 4: return FN<7, 4>(__VA_ARGS__);                         \
        case 7 * 16 + 6: return FN<7, 6>(__VA_ARGS__);                         \
        case 7 * 16 + 8: return FN<7, 8>(__VA_ARGS__);                         \
        case 8 * 16 + 4: return FN<8, 4>(__VA_ARGS__);                         \
        case 8 * 16 + 6: return FN<8, 6>(__VA_ARGS__);                         \
        case 8 * 16 + 8: return FN<8, 8>(__VA_ARGS__);                         \
        default: return cudaErrorInvalidValue;                                 \
    }

void planes_shape(int N, int M, int* nwm, int* np) {
    *nwm = (M + 31) / 32;
    *np = N <= 15 ? 4 : (N <= 63 ? 6 : 8);
}

cudaError_t launch_walkers_planes(const WalkerLaunch& L, cudaStream_t s) {
    int nwm, np;
    planes_shape(L.dm.N, L.dm.M, &nwm, &np);
    GWZ_PLANES_DISPATCH(nwm, np, planes_launch, L, s);
}

static cudaError_t planes_occ_dispatch(int nwm, int np, int N, int* blocks) {
    GWZ_PLANES_DISPATCH(nwm, np, planes_occ, N, blocks);
}
cudaError_t walker_planes_grid(int N, int M, size_t n, int sm_count, int* grid) {
    int nwm, np, blocks = 0;
    planes_shape(N, M, &nwm, &np);
    if (cudaError_t e = planes_occ_dispatch(nwm, np, N, &blocks)) return e;
    if (blocks < 1) blocks = 1;
    const size_t resident = (size_t)blocks * (size_t)sm_count;
    const size_t need = (n + STEP_BLOCK - 1) / STEP_BLOCK;
    *grid = (int)(need < resident ? need : resident);
    if (*grid < 1) *grid = 1;
    return cudaSuccess;
}

cudaError_t launch_unary_to_planes(const uint64_t* addr, size_t n, const DevModel& dm, uint32_t* planes, cudaStream_t s) {
    int nwm, np;
    planes_shape(dm.N, dm.M, &nwm, &np);
    unary_to_planes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(addr, n, dm.W, dm.B, dm.M, nwm, np, planes);
    return cudaGetLastError();
}
cudaError_t launch_planes_to_unary(const uint32_t* planes, size_t n, const DevModel& dm, size_t first, size_t count,
                                   uint64_t* addr, cudaStream_t s) {
    int nwm, np;
    planes_shape(dm.N, dm.M, &nwm, &np);
    planes_to_unary_kernel<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(planes, n, dm.W, dm.M, nwm, np, first, count, addr);
    return cudaGetLastError();
}

}  // namespace gwz
