// gwz_blocking.cu -- val_err_and_grad(::KineticVQMCWalkerState) (src/kinetic-vqmc/state.jl:50-60) for every
// walker at once: blocking_analysis(resample(residence_times, local_energies)) on the (tau, E_loc) series the
// walker kernels keep in HBM (step-major [step][walker]: a warp reads 32 neighbouring walkers of one step as one
// 256-byte segment), followed by the cross-walker combination of state.jl:150-164 / 265-285.
//
// One walker per thread, two lockstep phases per walker (gwz_blocking.cuh).  Phase A reads step k of every walker's
// series at the same time (coalesced 256-byte segments per warp) and writes the equal-time resampled values to a
// scratch column; how many slots a step completes differs between threads, so only the (fire-and-forget) stores
// diverge.  Phase B reads the scratch row back in lockstep and folds the values into the per-level moments, which
// live in shared memory in a [level][field][thread] layout (level 0 in registers); all threads are at the same value
// index, so the carry chain of the pair averaging is uniform across the warp.
// HBM-bound by design: algorithmic bytes = 16 per recorded walker-step (+ 80 per walker: running sums in, five
// results out); the scratch round trip adds 16 more per step, which stay in the 126 MB L2: a thread reuses its scratch
// row for every walker it handles (rows of all resident threads: 58 MB at 64 slots).
#include "gwz_internal.h"
#include "gwz_blocking.cuh"

namespace gwz {

__constant__ double c_chi2_q99[30] = GWZ_CHI2_Q99_TABLE;

struct SmemLevels {
    double* p;  // already offset by threadIdx.x
    __device__ __forceinline__ double& operator()(int level, int field) const {
        return p[(level * BLK_FIELDS + field) * BLOCKING_BLOCK];
    }
};

// Interval k of this thread's series.  Calls come with k = 0, 1, 2, ... and all threads of a warp are at the same k
// (lockstep).  The rows travel through a ring in shared memory filled by asynchronous copies (cp.async, 8 bytes per
// thread and array: a warp's copies of one row are two coalesced 256-byte requests), RING_ROWS - 1 rows ahead of the
// one being consumed: the depth of the memory pipeline costs shared memory, not registers.  (With the rows in
// registers, one batch of four ahead, the kernel spent 76 % of its long-scoreboard stalls waiting for that batch:
// ncu source page, profiles/r02l.)
#ifndef GWZ_BLK_RING
#define GWZ_BLK_RING 16
#endif
#ifndef GWZ_BLK_MINB
#define GWZ_BLK_MINB 1
#endif
constexpr int RING_ROWS = GWZ_BLK_RING;            // a power of two
// The ring: RING_ROWS slots of 16 bytes per thread, [slot][lane] inside the warp's region.  Phase A keeps (tau, E_loc)
// of one row in a slot, phase B one half of a 4-value group: in both layouts a thread touches only its own 16-byte
// column, so warps (and lanes) need no synchronisation when they change phase.
constexpr uint32_t RING_SLOT_BYTES = 16u * 32;
constexpr uint32_t RING_WARP_BYTES = RING_ROWS * RING_SLOT_BYTES;
constexpr size_t RING_BYTES = (size_t)RING_WARP_BYTES * (BLOCKING_BLOCK / 32);

__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {  // through L2 only (.cg)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}

struct SeriesRing {
    const double* __restrict__ pt;  // the next row to request
    const double* __restrict__ pe;
    size_t stride;
    uint64_t n, issued;
    uint32_t ring;  // shared address of this thread's column: slot r at ring + r RING_SLOT_BYTES = (tau, E_loc) of a row
    __device__ __forceinline__ void request() {  // one commit group per row (empty past the end: uniform accounting)
        if (issued < n) {
            const uint32_t dst = ring + ((uint32_t)issued & (RING_ROWS - 1)) * RING_SLOT_BYTES;
            cp_async8(dst, pt);
            cp_async8(dst + 8u, pe);
            pt += stride;
            pe += stride;
        }
        cp_async_commit();
        ++issued;
    }
    __device__ __forceinline__ void prime() {
#pragma unroll 1
        for (int r = 0; r < RING_ROWS - 1; ++r) request();
    }
    __device__ __forceinline__ void operator()(uint64_t k, double& t, double& e) {
        request();                       // row k + RING_ROWS - 1, into the slot of row k - 1
        cp_async_wait<RING_ROWS - 1>();  // everything but the RING_ROWS - 1 youngest rows has landed: row k is there
        const uint32_t src = ring + ((uint32_t)k & (RING_ROWS - 1)) * RING_SLOT_BYTES;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t), "=d"(e) : "r"(src));
    }
};
// slot i of the resampled series; calls come with i = 0, 1, 2, ...  The scratch is walker-major (each thread fills
// its own contiguous row): threads reach a given slot at different times, and consecutive 8-byte stores of ONE thread
// complete a 32-byte sector within four slots, whereas a step-major layout leaves partially written sectors in L2
// until the slowest neighbour arrives (evicted half-filled for long series).
struct ScratchStore {
    double* __restrict__ v;
    __device__ __forceinline__ void operator()(uint64_t, double x) { *v++ = x; }
};

__global__ void __launch_bounds__(BLOCKING_BLOCK, GWZ_BLK_MINB)
series_blocking_kernel(const double* __restrict__ tau, const double* __restrict__ eloc, size_t stride, size_t n,
                       uint64_t steps, uint64_t len, int levels, const double* __restrict__ acc, size_t acc_stride,
                       double shift_e, double shift_g, double* __restrict__ scratch, size_t scratch_stride,
                       int scratch_per_walker, double* __restrict__ out, size_t out_stride, double* __restrict__ partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* lv = reinterpret_cast<double*>(smem_raw) + threadIdx.x;
    const size_t lv_bytes = sizeof(double) * (size_t)(levels > 0 ? levels : 1) * BLK_FIELDS * BLOCKING_BLOCK;
    int* khist = reinterpret_cast<int*>(smem_raw + lv_bytes);
    // the copy ring of this warp (16-byte aligned)
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(smem_raw + lv_bytes + 128) + (threadIdx.x >> 5) * RING_WARP_BYTES;
    const uint32_t wl = threadIdx.x & 31u;  // lane inside the warp
    if (threadIdx.x < 32) khist[threadIdx.x] = 0;
    __syncthreads();
    double a_walkers = 0.0, a_mean = 0.0, a_err2 = 0.0, a_ee2 = 0.0, a_fail = 0.0, a_grad = 0.0;

    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < n; w += nthreads) {
        const double* tw = tau + w;
        const double* ew = eloc + w;
        // total time and the centre of the moments (the time-weighted mean): from the walker's running sums when
        // they cover exactly this series, else by one pass over the series (same order as the reference's sum)
        double total, centre, grad_fix_g = 0.0, grad_w = 0.0, val_w = 0.0;
        if (acc) {
            const double st = acc[w], ste = acc[acc_stride + w], stg = acc[2 * acc_stride + w],
                         stge = acc[3 * acc_stride + w];
            const double inv = 1.0 / st;
            total = st;
            val_w = shift_e + ste * inv;
            centre = val_w;
            grad_w = 2.0 * (stge * inv - (ste * inv) * (stg * inv));  // state.jl:42-49
            grad_fix_g = shift_g + stg * inv;                          // mean of the gradient ratio
        } else {
            total = 0.0;
            double te = 0.0;
            for (uint64_t k = 0; k < steps; ++k) {
                const double t = tw[k * stride];
                total = bl_add(total, t);
                te = fma(t, ew[k * stride], te);
            }
            centre = te / total;
        }
        BlockingOut o;
        if (len == 1) {  // blocking_analysis of a single value
            o.mean = centre;
            o.err = o.err_err = o.p_cov = NAN;
            o.k = 0;
            o.blocks = 1;
        } else {
            // phase A: resample, driven by the input intervals (all threads at the same step of their series)
            // this thread's scratch row: reused for every walker the thread handles (a few hundred KB per SM in
            // flight: the stores and the reads of phase B hit L2), unless the caller wants the rows back
            double* vw = scratch + (scratch_per_walker ? w : (size_t)blockIdx.x * blockDim.x + threadIdx.x) * scratch_stride;
            SeriesRing src;
            src.pt = tw;
            src.pe = ew;
            src.stride = stride;
            src.n = steps;
            src.issued = 0;
            src.ring = ring0 + 16u * wl;
            src.prime();
            resample_by_input(steps, len, total, src, ScratchStore{vw});
            cp_async_wait<0>();  // rows requested past the point where the last slot filled
            __threadfence_block();  // the scratch row is read back by asynchronous copies: order them after its stores
            // phase B: the resampled values, in lockstep, into the per-level moments
            for (int i = 0; i < levels * BLK_FIELDS; ++i) lv[i * BLOCKING_BLOCK] = 0.0;
            BlockingStream<SmemLevels> bs(SmemLevels{lv}, levels, centre);
            {
                const double* pv = vw;
                uint64_t i = 0;
                if (levels >= 2) {
                    // groups of four values (one 32-byte sector of this thread's row) come back from L2 through the
                    // same ring, GROUPS - 1 groups ahead; levels 0 and 1 of the pair averaging stay in registers
                    constexpr int GROUPS = RING_ROWS / 2;                  // 32 bytes per thread and group
                    constexpr uint32_t HALF_BYTES = RING_SLOT_BYTES;       // one 16-byte half of a group = one slot
                    const uint32_t gring = ring0 + 16u * wl;
                    const uint64_t ngroups = len / 4;
                    uint64_t req = 0;
                    auto request = [&]() {
                        if (req < ngroups) {
                            const uint32_t dst = gring + ((uint32_t)req & (GROUPS - 1)) * (2u * HALF_BYTES);
                            cp_async16(dst, pv + 4 * req);
                            cp_async16(dst + HALF_BYTES, pv + 4 * req + 2);
                        }
                        cp_async_commit();
                        ++req;
                    };
#pragma unroll 1
                    for (int r = 0; r < GROUPS - 1; ++r) request();
                    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0, b4 = 0.0;
#pragma unroll 1
                    for (uint64_t gi = 0; gi < ngroups; ++gi) {
                        request();
                        cp_async_wait<GROUPS - 1>();
                        const uint32_t srcg = gring + ((uint32_t)gi & (GROUPS - 1)) * (2u * HALF_BYTES);
                        double x0, x1, x2, x3;
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(srcg));
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x2), "=d"(x3) : "r"(srcg + HALF_BYTES));
                        bs.push4(x0, x1, x2, x3, b0, b1, b2, b3, b4);
                    }
                    cp_async_wait<0>();
                    bs.flush_level1(b0, b1, b2, b3, b4);
                    i = 4 * ngroups;
                    pv += i;
                }
                for (; i < len; ++i, ++pv) bs.push(*pv);
            }
            bs.finalize(len, c_chi2_q99, o);
        }
        // state.jl:57-58: the gradient with val = the blocking mean instead of the weighted mean
        const double grad_b = grad_w + 2.0 * (val_w - o.mean) * grad_fix_g;
        if (out) {
            out[w] = o.mean;
            out[out_stride + w] = o.err;
            out[2 * out_stride + w] = o.err_err;
            out[3 * out_stride + w] = (double)o.k;
            out[4 * out_stride + w] = (double)o.blocks;
        }
        a_walkers += 1.0;
        a_mean += o.mean;
        a_grad += grad_b;
        if (o.k > 0) {
            a_err2 = fma(o.err, o.err, a_err2);
            a_ee2 = fma(o.err_err, o.err_err, a_ee2);
            atomicAdd(&khist[o.k < 32 ? o.k : 31], 1);
        } else {
            a_fail += 1.0;
        }
    }
    if (partials == nullptr) return;
    // block partials in a fixed order (deterministic): warp shuffle -> shared -> thread 0..NUM_BACC-1
    __syncthreads();
    double* red = reinterpret_cast<double*>(smem_raw);  // the level store is dead by now
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double vals[6] = {a_walkers, a_mean, a_err2, a_ee2, a_fail, a_grad};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double v = warp_sum(vals[i]);
        if (lane == 0) red[warp * 6 + i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NUM_BACC) {
        double v = 0.0;
        if (threadIdx.x < 6) {
            for (int q = 0; q < BLOCKING_BLOCK / 32; ++q) v += red[q * 6 + threadIdx.x];
        } else {
            v = (double)khist[threadIdx.x - 6];
        }
        partials[(size_t)blockIdx.x * NUM_BACC + threadIdx.x] = v;
    }
}

// doubles per walker row of the scratch: the resample length rounded up to whole 32-byte sectors
size_t blocking_scratch_stride(uint64_t len) { return (size_t)((len + 3) & ~3ull); }

static size_t blocking_smem(int levels) {
    const size_t lv = sizeof(double) * (size_t)(levels > 0 ? levels : 1) * BLK_FIELDS * BLOCKING_BLOCK;
    const size_t red = sizeof(double) * 6 * (BLOCKING_BLOCK / 32);
    return (lv > red ? lv : red) + 128 /* khist */ + RING_BYTES;
}

cudaError_t blocking_grid(uint64_t len, size_t n, int sm_count, int* grid) {
    const int levels = blocking_levels(len);
    const size_t smem = blocking_smem(levels);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        if (cudaError_t e = cudaFuncSetAttribute(series_blocking_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
            return e;
    }
    int blocks = 0;
    if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, series_blocking_kernel, BLOCKING_BLOCK, smem))
        return e;
    if (blocks < 1) blocks = 1;
    const size_t resident = (size_t)blocks * (size_t)sm_count;
    const size_t need = (n + BLOCKING_BLOCK - 1) / BLOCKING_BLOCK;
    *grid = (int)(need < resident ? need : resident);
    if (*grid < 1) *grid = 1;
    return cudaSuccess;
}

cudaError_t launch_series_blocking(const BlockingLaunch& L, cudaStream_t s) {
    const int levels = blocking_levels(L.len);
    series_blocking_kernel<<<L.grid, BLOCKING_BLOCK, blocking_smem(levels), s>>>(
        L.tau, L.eloc, L.stride, L.n, L.steps, L.len, levels, L.acc, L.acc_stride, L.shift_e, L.shift_g, L.scratch,
        blocking_scratch_stride(L.len), L.scratch_per_walker ? 1 : 0, L.out, L.out_stride, L.partials);
    return cudaGetLastError();
}

}  // namespace gwz
