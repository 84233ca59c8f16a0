// gwz_walker_incr.cu -- production walker kernel: occupation bit-planes + incremental bond-class counts
// (see gwz_incr.cuh).  Same contract as walker_planes_kernel / walker_kernel<NW, 4, false>:
// kinetic_vqmc!(st; steps) (src/kinetic-vqmc/qmc.jl:73-97) for every walker, running sums of
// state.jl:42-49, per-block partial accumulator vectors.  The resident HBM format is the plane
// format of gwz_walker_planes.cu; occupation bytes and class counts are rebuilt per launch.
#include "gwz_internal.h"
#include "gwz_incr.cuh"

#ifndef GWZ_INCR_MIN_BLOCKS
#define GWZ_INCR_MIN_BLOCKS 6
#endif

namespace gwz {

template <int NWM>
__host__ __device__ constexpr size_t incr_smem_bytes(int N) {
    return IncrScratch<NWM>::bytes(N) + sizeof(double) * NUM_ACC * STEP_BLOCK;
}

// REC: also keep the (tau, E_loc) series of the launch in HBM, step-major [step][walker] (qmc.jl:89-92 stores them per
// walker; the device-side blocking analysis of gwz_blocking.cu reads them back)
template <int NWM, int NP, bool REC>
__global__ void __launch_bounds__(STEP_BLOCK, NWM <= 2 ? GWZ_INCR_MIN_BLOCKS : (NWM <= 4 ? 4 : 2))
walker_incr_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ DevTables tb,
                   const __grid_constant__ PhiloxKeys pk, uint32_t* __restrict__ planes, double* __restrict__ acc, size_t n, uint64_t global_offset,
                   uint64_t step0, uint64_t steps, int accumulate, int reset_acc, double shift_e, double shift_g,
                   double* __restrict__ partials, int* __restrict__ err_flag, double* __restrict__ rec_tau,
                   double* __restrict__ rec_eloc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const IncrScratch<NWM> sc(smem_raw, dm.N);
    sc.fill_tables(dm.N, tb);
    double* p = reinterpret_cast<double*>(smem_raw + IncrScratch<NWM>::bytes(dm.N)) + threadIdx.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const double E0 = shift_e, G0 = shift_g;  // the launcher passes 0 for burn-in launches
#pragma unroll
    for (int i = 0; i < NUM_ACC; ++i) p[i * STEP_BLOCK] = 0.0;
    bool bad = false;

    for (size_t w = tid; w < n; w += nthreads) {
        uint32_t pl[NP][NWM];
#pragma unroll
        for (int k = 0; k < NP; ++k)
#pragma unroll
            for (int i = 0; i < NWM; ++i) pl[k][i] = planes[((size_t)k * NWM + i) * n + w];
        incr_init<NWM, NP>(pl, dm, sc);
        BigCache bc;
        bc.S = bc.E = 0.0;
        bc.nocc = bc.d = 0;
        bc.valid = false;
        uint32_t bigP[NWM];
        big_plane<NWM, NP>(pl, bigP);
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
                philox4x32_10_keyed((uint32_t)blk, (uint32_t)(blk >> 32), c2, c3, pk, rnd);
            }
            const double u01 = (step & 1ull) ? u01_from_words(rnd[2], rnd[3]) : u01_from_words(rnd[0], rnd[1]);
            double tau, eloc, gr;
            int nh;
            step_incr<NWM, NP>(pl, dm, tb, sc, bc, bigP, u01, tau, eloc, gr, nh);
            hops += (unsigned)nh;
            if (REC) {
                __stcs(rec_tau + (size_t)k * n + w, tau);
                __stcs(rec_eloc + (size_t)k * n + w, eloc);
            }
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
    double(*sm)[NUM_ACC] = reinterpret_cast<double(*)[NUM_ACC]>(sc.onr);  // occupation bytes are dead by now
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
// launchers
// ------------------------------------------------------------------------------------------
template <int NWM, int NP, bool REC>
static cudaError_t incr_prepare(int N) {
    const size_t smem = incr_smem_bytes<NWM>(N);
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(walker_incr_kernel<NWM, NP, REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <int NWM, int NP>
static cudaError_t incr_launch(const WalkerLaunch& L, cudaStream_t s) {
    const bool rec = L.rec_tau != nullptr && L.accumulate;
    if (cudaError_t e = rec ? incr_prepare<NWM, NP, true>(L.dm.N) : incr_prepare<NWM, NP, false>(L.dm.N)) return e;
    if (rec)
        walker_incr_kernel<NWM, NP, true><<<L.grid, STEP_BLOCK, incr_smem_bytes<NWM>(L.dm.N), s>>>(
            L.dm, L.tb, philox_keys(L.seed), L.planes, L.acc, L.n, L.global_offset, L.step0, L.steps, L.accumulate,
            L.reset_acc, L.shift_e, L.shift_g, L.partials, L.err_flag, L.rec_tau, L.rec_eloc);
    else
        walker_incr_kernel<NWM, NP, false><<<L.grid, STEP_BLOCK, incr_smem_bytes<NWM>(L.dm.N), s>>>(
            L.dm, L.tb, philox_keys(L.seed), L.planes, L.acc, L.n, L.global_offset, L.step0, L.steps, L.accumulate,
            L.reset_acc, L.accumulate ? L.shift_e : 0.0, L.accumulate ? L.shift_g : 0.0, L.partials, L.err_flag, nullptr,
            nullptr);
    return cudaGetLastError();
}
template <int NWM, int NP>
static cudaError_t incr_occ(int N, int* blocks) {
    if (cudaError_t e = incr_prepare<NWM, NP, false>(N)) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, walker_incr_kernel<NWM, NP, false>, STEP_BLOCK,
                                                         incr_smem_bytes<NWM>(N));
}

#define GWZ_INCR_DISPATCH(nwm, np, FN, ...)                \
    switch ((nwm) * 16 + (np)) {                           \
        case 1 * 16 + 4: return FN<1, 4>(__VA_ARGS__);     \
        case 1 * 16 + 6: return FN<1, 6>(__VA_ARGS__);     \
        case 1 * 16 + 8: return FN<1, 8>(__VA_ARGS__);     \
        case 2 * 16 + 4: return FN<2, 4>(__VA_ARGS__);     \
        case 2 * 16 + 6: return FN<2, 6>(__VA_ARGS__);     \
        case 2 * 16 + 8: return FN<2, 8>(__VA_ARGS__);     \
        case 3 * 16 + 4: return FN<3, 4>(__VA_ARGS__);     \
        case 3 * 16 + 6: return FN<3, 6>(__VA_ARGS__);     \
        case 3 * 16 + 8: return FN<3, 8>(__VA_ARGS__);     \
        case 4 * 16 + 4: return FN<4, 4>(__VA_ARGS__);     \
        case 4 * 16 + 6: return FN<4, 6>(__VA_ARGS__);     \
        case 4 * 16 + 8: return FN<4, 8>(__VA_ARGS__);     \
        case 5 * 16 + 4: return FN<5, 4>(__VA_ARGS__);     \
        case 5 * 16 + 6: return FN<5, 6>(__VA_ARGS__);     \
        case 5 * 16 + 8: return FN<5, 8>(__VA_ARGS__);     \
        case 6 * 16 + 4: return FN<6, 4>(__VA_ARGS__);     \
        case 6 * 16 + 6: return FN<6, 6>(__VA_ARGS__);     \
        case 6 * 16 + 8: return FN<6, 8>(__VA_ARGS__);     \
        case 7 * 16 + 4: return FN<7, 4>(__VA_ARGS__);     \
        case 7 * 16 + 6: return FN<7, 6>(__VA_ARGS__);     \
        case 7 * 16 + 8: return FN<7, 8>(__VA_ARGS__);     \
        case 8 * 16 + 4: return FN<8, 4>(__VA_ARGS__);     \
        case 8 * 16 + 6: return FN<8, 6>(__VA_ARGS__);     \
        case 8 * 16 + 8: return FN<8, 8>(__VA_ARGS__);     \
        default: return cudaErrorInvalidValue;             \
    }

// the incremental bookkeeping needs the four modes z-1 .. z+2 around a bond to be distinct
bool walker_incr_supported(int N, int M) {
    (void)N;
    return M >= 4;
}

cudaError_t launch_walkers_incr(const WalkerLaunch& L, cudaStream_t s) {
    int nwm, np;
    planes_shape(L.dm.N, L.dm.M, &nwm, &np);
    GWZ_INCR_DISPATCH(nwm, np, incr_launch, L, s);
}

static cudaError_t incr_occ_dispatch(int nwm, int np, int N, int* blocks) {
    GWZ_INCR_DISPATCH(nwm, np, incr_occ, N, blocks);
}
cudaError_t walker_incr_grid(int N, int M, size_t n, int sm_count, int* grid) {
    int nwm, np, blocks = 0;
    planes_shape(N, M, &nwm, &np);
    if (cudaError_t e = incr_occ_dispatch(nwm, np, N, &blocks)) return e;
    if (blocks < 1) blocks = 1;
    const size_t resident = (size_t)blocks * (size_t)sm_count;
    const size_t need = (n + STEP_BLOCK - 1) / STEP_BLOCK;
    *grid = (int)(need < resident ? need : resident);
    if (*grid < 1) *grid = 1;
    return cudaSuccess;
}

}  // namespace gwz
