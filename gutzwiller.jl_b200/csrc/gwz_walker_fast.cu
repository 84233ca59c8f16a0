// gwz_walker_fast.cu -- production walker kernel (round 2): kinetic_vqmc!(st; steps) (src/kinetic-vqmc/qmc.jl:73-97)
// for every walker with the step of gwz_fast.cuh (integer hop counts per rate class, hop masks, table-driven updates), the
// running sums of state.jl:42-49 and per-block partial accumulator vectors.  Same contract and same resident HBM format
// (occupation bit-planes) as walker_incr_kernel; its own u -> hop map (see gwz_fast.cuh) and its own use of the Philox
// stream (one 32-bit word per step instead of 53 bits).
//
// REC: the recording instantiation keeps the (tau, E_loc) series of the launch in HBM, step-major [step][walker]
// (qmc.jl:89-92; read back by the blocking kernel, gwz_blocking.cu), optionally the selected hop of every step (the
// parity surface: a trajectory is its start address plus its hops) and optionally takes the uniforms from the caller
// instead of the Philox stream (one kinetic_sample! per address with a prescribed u: the stratified-u test).
#include "gwz_internal.h"
#include "gwz_fast.cuh"

namespace gwz {

template <int NWM>
__host__ __device__ constexpr size_t fast_smem_bytes(int N) {
    return FastScratch<NWM>::bytes(N) + sizeof(double) * NUM_ACC * (fast_block(NWM) / 32);
}

template <int NWM, int NP, bool REC>
__global__ void __launch_bounds__(fast_block(NWM), fast_min_blocks(NWM))
walker_fast_kernel(const __grid_constant__ DevModel dm, const __grid_constant__ DevTables tb,
                   const __grid_constant__ FastRates fr, const __grid_constant__ PhiloxKeys pk,
                   uint32_t* __restrict__ planes, double* __restrict__ acc, size_t n, uint64_t global_offset, uint64_t step0,
                   uint64_t steps, int accumulate, int reset_acc, double shift_e, double shift_g,
                   double* __restrict__ partials, int* __restrict__ err_flag, double* __restrict__ rec_tau,
                   double* __restrict__ rec_eloc, int32_t* __restrict__ rec_hop, const double* __restrict__ u01s) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FastScratch<NWM> sc(smem_raw, dm.N);
    sc.fill_tables(dm.N, tb);
    const FastAddr ad = fast_addr<FastScratch<NWM>::VW>(smem_raw);
    // per-warp partial accumulator vectors (shared memory): every finished batch of 32 walkers is reduced by shuffles
    double(*psum)[NUM_ACC] = reinterpret_cast<double(*)[NUM_ACC]>(smem_raw + FastScratch<NWM>::bytes(dm.N));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < NUM_ACC) psum[warp][lane] = 0.0;
    __syncwarp();
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const double E0 = shift_e, G0 = shift_g;  // the launcher passes 0 for burn-in launches
    bool bad = false;

    for (size_t w0 = (size_t)blockIdx.x * blockDim.x; w0 < n; w0 += nthreads) {  // uniform per CTA
        const size_t w = w0 + threadIdx.x;
        const bool live = w < n;
        double contrib[NUM_ACC];
#pragma unroll
        for (int i = 0; i < NUM_ACC; ++i) contrib[i] = 0.0;
        if (live) {
            FastState s;
            {
                uint32_t pl[NP][NWM];
#pragma unroll
                for (int k = 0; k < NP; ++k)
#pragma unroll
                    for (int i = 0; i < NWM; ++i) pl[k][i] = planes[((size_t)k * NWM + i) * n + w];
                fast_init<NWM, NP>(pl, dm, sc, ad.tb, s);
            }
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
            unsigned long long occ = 0ull;  // sum over the steps of the number of occupied modes (= hops / 2)
            const double st0 = st;
            unsigned occ4 = 0u;

            // one kinetic_sample! and its contribution to the running sums; rw = the 32-bit random word of the step
            auto one_step = [&](uint64_t step, unsigned step32, uint32_t rw) {
                double u01 = (double)rw * 0x1.0p-32;
                const size_t k = (size_t)(step - step0);  // recording instantiation only
                if (REC && u01s) u01 = fmin(u01s[k * n + w], 1.0 - 0x1.0p-32);
                double tau, eloc, gr;
                int no, hop;
                const bool refresh = (step32 & (FAST_REFRESH - 1)) == FAST_REFRESH - 1;
                step_fast<NWM>(s, dm, fr, sc, ad, u01, refresh, tau, eloc, gr, no, hop);
                occ4 += (unsigned)no;
                if (REC) {
                    if (rec_tau) __stcs(rec_tau + k * n + w, tau);
                    if (rec_eloc) __stcs(rec_eloc + k * n + w, eloc);
                    if (rec_hop) rec_hop[k * n + w] = hop;
                }
                const double dE = eloc - E0, dG = gr - G0;
                const double tE = tau * dE;
                st += tau;
                ste += tE;
                stg = fma(tau, dG, stg);
                stge = fma(tE, dG, stge);
                stee = fma(tE, dE, stee);
            };
            // one Philox4x32-10 block (key = seed, counter = (step / 4, global walker id)) feeds four steps: the uniform
            // of a step is one 32-bit word, u = word / 2^32; the words are consumed by rotating the four registers.  A
            // launch may start and end inside a block.  The four steps are NOT unrolled: four copies of the step body
            // overflow the instruction cache (measured: 5 % fewer instructions per step, 3 - 30 % slower).
            const uint64_t end = step0 + steps;
            uint64_t step = step0;
            while (step < end) {
                uint32_t rnd[4];
                const uint64_t blk = step >> 2;
                philox4x32_10_keyed((uint32_t)blk, (uint32_t)(blk >> 32), c2, c3, pk, rnd);
                int q = (int)((unsigned)step & 3u);
                const uint64_t left = end - step;
                const int qe = left < (uint64_t)(4 - q) ? q + (int)left : 4;
#pragma unroll 1
                for (int r = 0; r < q; ++r) {  // a launch that starts inside a block: skip the words already used
                    rnd[0] = rnd[1];
                    rnd[1] = rnd[2];
                    rnd[2] = rnd[3];
                }
                occ4 = 0u;
                const int nq = qe - q;
                const unsigned s32 = (unsigned)step;
#pragma unroll 1
                for (int r = 0; r < nq; ++r) {
                    one_step(step + (unsigned)r, s32 + (unsigned)r, rnd[0]);
                    rnd[0] = rnd[1];
                    rnd[1] = rnd[2];
                    rnd[2] = rnd[3];
                }
                step += (unsigned)nq;
                occ += occ4;
            }
            bad |= !(st - st0 < 1.0e300);  // isfinite(residence_time) for every step (qmc.jl:34-36)

            {
                uint32_t pl[NP][NWM];
                fast_store_planes<NWM, NP>(sc, s, pl);
#pragma unroll
                for (int k = 0; k < NP; ++k)
#pragma unroll
                    for (int i = 0; i < NWM; ++i) planes[((size_t)k * NWM + i) * n + w] = pl[k][i];
            }
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
                contrib[ACC_WALKERS] = 1.0;
                contrib[ACC_VAL] = val_w;
                contrib[ACC_VAL2] = val_w * val_w;
                contrib[ACC_GRAD] = grad_w;
                contrib[ACC_GRAD2] = grad_w * grad_w;
                contrib[ACC_T] = st;
                contrib[ACC_TE] = st * val_w;
                contrib[ACC_TG] = st * (G0 + mG);
                contrib[ACC_TGE] = st * G0 * E0 + G0 * ste + E0 * stg + stge;
                contrib[ACC_TEE] = st * E0 * E0 + 2.0 * E0 * ste + stee;
                contrib[ACC_HOPS] = 2.0 * (double)occ;
                contrib[ACC_STEPS] = (double)steps;
                contrib[ACC_VARW] = stee * inv - mE * mE;  // var(E_loc, weights tau), uncorrected
            }
        }
        if (accumulate) {  // uniform: every lane of the warp is here (dead lanes contribute zeros)
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NUM_ACC; ++i) {
                const double v = warp_sum(contrib[i]);
                if (lane == 0) psum[warp][i] += v;
            }
        }
    }

    if (bad) atomicOr(err_flag, 1);
    if (!accumulate || partials == nullptr) return;
    __syncthreads();
    if (threadIdx.x < NUM_ACC) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < fast_block(NWM) / 32; ++q) v += psum[q][threadIdx.x];
        partials[(size_t)blockIdx.x * NUM_ACC + threadIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <int NWM, int NP, bool REC>
static cudaError_t fast_prepare(int N) {
    const size_t smem = fast_smem_bytes<NWM>(N);
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(walker_fast_kernel<NWM, NP, REC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <int NWM, int NP>
static cudaError_t fast_launch(const WalkerLaunch& L, cudaStream_t s) {
    const bool rec = (L.rec_tau != nullptr || L.rec_hop != nullptr || L.u01s != nullptr) && L.accumulate;
    if (cudaError_t e = rec ? fast_prepare<NWM, NP, true>(L.dm.N) : fast_prepare<NWM, NP, false>(L.dm.N)) return e;
    const FastRates fr = fast_rates(L.tb.c);
    if (rec)
        walker_fast_kernel<NWM, NP, true><<<L.grid, fast_block(NWM), fast_smem_bytes<NWM>(L.dm.N), s>>>(
            L.dm, L.tb, fr, philox_keys(L.seed), L.planes, L.acc, L.n, L.global_offset, L.step0, L.steps, L.accumulate,
            L.reset_acc, L.shift_e, L.shift_g, L.partials, L.err_flag, L.rec_tau, L.rec_eloc, L.rec_hop, L.u01s);
    else
        walker_fast_kernel<NWM, NP, false><<<L.grid, fast_block(NWM), fast_smem_bytes<NWM>(L.dm.N), s>>>(
            L.dm, L.tb, fr, philox_keys(L.seed), L.planes, L.acc, L.n, L.global_offset, L.step0, L.steps, L.accumulate,
            L.reset_acc, L.accumulate ? L.shift_e : 0.0, L.accumulate ? L.shift_g : 0.0, L.partials, L.err_flag, nullptr,
            nullptr, nullptr, nullptr);
    return cudaGetLastError();
}
template <int NWM, int NP>
static cudaError_t fast_occ(int N, int* blocks) {
    if (cudaError_t e = fast_prepare<NWM, NP, false>(N)) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, walker_fast_kernel<NWM, NP, false>, fast_block(NWM),
                                                         fast_smem_bytes<NWM>(N));
}

#define GWZ_FAST_DISPATCH(nwm, np, FN, ...)                \
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

// the bookkeeping needs the four modes z-1 .. z+2 around a bond to be distinct, and every count in its field
bool walker_fast_supported(int N, int M) {
    (void)N;
    return M >= 4 && M <= 250;
}

cudaError_t launch_walkers_fast(const WalkerLaunch& L, cudaStream_t s) {
    int nwm, np;
    planes_shape(L.dm.N, L.dm.M, &nwm, &np);
    GWZ_FAST_DISPATCH(nwm, np, fast_launch, L, s);
}

static cudaError_t fast_occ_dispatch(int nwm, int np, int N, int* blocks) {
    GWZ_FAST_DISPATCH(nwm, np, fast_occ, N, blocks);
}
cudaError_t walker_fast_grid(int N, int M, size_t n, int sm_count, int* grid) {
    int nwm, np, blocks = 0;
    planes_shape(N, M, &nwm, &np);
    if (cudaError_t e = fast_occ_dispatch(nwm, np, N, &blocks)) return e;
    if (blocks < 1) blocks = 1;
    const size_t resident = (size_t)blocks * (size_t)sm_count;
    const size_t fb = (size_t)fast_block(nwm);
    const size_t need = (n + fb - 1) / fb;
    *grid = (int)(need < resident ? need : resident);
    if (*grid < 1) *grid = 1;
    return cudaSuccess;
}

}  // namespace gwz
