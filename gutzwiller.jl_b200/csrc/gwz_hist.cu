// gwz_hist.cu -- DVec(result) on the device (src/kinetic-vqmc/state.jl:303-319): the residence time of every
// recorded step is deposited on its address.  Open-addressing hash table in HBM (linear probing, 64-bit
// atomicCAS on the key, atomicAdd on the value); single-word addresses (N+M-1 <= 64).
#include "gwz_internal.h"

namespace gwz {

constexpr uint64_t HIST_EMPTY = ~0ull;  // never a valid address: a BoseFS string has at most 63 ones in 64 bits

__device__ __forceinline__ uint64_t hist_mix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__global__ void hist_fill_kernel(uint64_t* __restrict__ tkeys, double* __restrict__ tvals, size_t slots) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (size_t)gridDim.x * blockDim.x) {
        tkeys[i] = HIST_EMPTY;
        tvals[i] = 0.0;
    }
}

__global__ void hist_insert_kernel(const uint64_t* __restrict__ keys, const double* __restrict__ tau, size_t cells,
                                   uint64_t* __restrict__ tkeys, double* __restrict__ tvals, uint64_t mask) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        uint64_t h = hist_mix(k) & mask;
        for (;;) {
            const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(tkeys + h),
                                                      (unsigned long long)HIST_EMPTY, (unsigned long long)k);
            if (prev == HIST_EMPTY || prev == k) {
                atomicAdd(tvals + h, tau[i]);
                break;
            }
            h = (h + 1) & mask;  // the table has at least twice as many slots as distinct keys
        }
    }
}

__global__ void hist_compact_kernel(const uint64_t* __restrict__ tkeys, const double* __restrict__ tvals, size_t slots,
                                    uint64_t* __restrict__ out_keys, double* __restrict__ out_vals,
                                    unsigned long long* __restrict__ count, uint64_t cap) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += (size_t)gridDim.x * blockDim.x) {
        if (tkeys[i] != HIST_EMPTY) {
            const unsigned long long pos = atomicAdd(count, 1ull);
            if (pos < cap) {
                out_keys[pos] = tkeys[i];
                out_vals[pos] = tvals[i];
            }
        }
    }
}

cudaError_t launch_hist_fill(uint64_t* tkeys, double* tvals, size_t slots, cudaStream_t s) {
    hist_fill_kernel<<<1184, 256, 0, s>>>(tkeys, tvals, slots);
    return cudaGetLastError();
}
cudaError_t launch_hist_insert(const uint64_t* keys, const double* tau, size_t cells, uint64_t* tkeys, double* tvals,
                               uint64_t mask, cudaStream_t s) {
    hist_insert_kernel<<<1184, 256, 0, s>>>(keys, tau, cells, tkeys, tvals, mask);
    return cudaGetLastError();
}
cudaError_t launch_hist_compact(const uint64_t* tkeys, const double* tvals, size_t slots, uint64_t* out_keys,
                                double* out_vals, unsigned long long* count, uint64_t cap, cudaStream_t s) {
    hist_compact_kernel<<<1184, 256, 0, s>>>(tkeys, tvals, slots, out_keys, out_vals, count, cap);
    return cudaGetLastError();
}

}  // namespace gwz
