// gwz_device.cuh -- device-side building blocks shared by the probe, walker and sweep kernels.
//
// State representation: a BoseFS{N,M} address is a unary ("stars and bars") bitstring of
// B = N+M-1 bits (Rimu BitString).  On the device it lives in NW 32-bit
// registers per thread (statically indexed, fully unrolled), extended to
//     e = a | bit B = 0 (virtual separator closing mode M) | K low bits of a above it
// so that the periodic bond (M,1) looks like any other bond.
//
// A *bond* is a 0 bit of e at position z in [0,B]: L(z) = run of ones directly below it =
// occupation of mode i, R(z) = run directly above = occupation of mode i+1.  The two hops of
// offdiagonals(H,a) (Rimu hopnextneighbour; App. A.4) that cross the bond are
//     right  i -> i+1   if L>0:  Delta = R-L+1,  melem = -t*sqrt(L*(R+1))
//     left   i+1 -> i   if R>0:  Delta = L-R+1,  melem = -t*sqrt(R*(L+1))
// and for GutzwillerAnsatz psi(new)/psi(a) = exp(-g*u*Delta) (gutzwiller.jl:34-36 with
// diagonal_element = u*sum n(n-1)/2).  kinetic_sample! (qmc.jl:14-44) then reduces to sums over
// bonds of functions of (L,R).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gwz {

// Bond classes with L,R in 0..KC-1 are counted bit-parallel; KC (4 or 5) is a template parameter
// chosen per launch from the expected occupation statistics (class id = L*KC + R).
constexpr int KCAP_MAX = 5;
constexpr int MAXCLS = KCAP_MAX * KCAP_MAX;
constexpr int MAXNW = 8;            // 32-bit words of the extended string (B + 1 + KCAP_MAX <= 256)

struct DevModel {
    int N, M, B;   // particles, modes, bits
    int W;         // 64-bit words per stored address
    int ext_ws;    // (B+1) >> 5 : word of the first extension bit
    int ext_bs;    // (B+1) & 31
    double u, t;
    uint32_t zmask[MAXNW];  // bits 0..B   (bond positions)
    uint32_t amask[MAXNW];  // bits 0..B-1 (address bits)
    uint32_t extmask_lo[MAXNW];  // all ones for word ext_ws, else 0
    uint32_t extmask_hi[MAXNW];  // all ones for word ext_ws+1 (if ext_bs != 0), else 0
    // occupation bit-plane representation (gwz_planes.cuh): M bits per plane
    uint32_t mmask[MAXNW];     // bits 0..M-1
    uint32_t mtopmask[MAXNW];  // all ones for the word holding bit M-1, else 0
    int mtop_b;                // (M-1) & 31
    int pad_;
};

constexpr int STEP_BLOCK = 128;  // threads per CTA of every kernel that selects moves (scratch layout)

// Everything that depends on the variational parameter g; passed by value as a kernel
// argument so that statically indexed entries become constant-bank operands.
struct DevTables {
    double g;            // params[1]
    double c;            // g*u
    int kc;              // the KC these tables are laid out for (index = L*kc + R)
    int pad_;
    double FS[MAXCLS];     // sum of the (<=2) hop rates of a bond of class (L,R)
    double FE[MAXCLS];     // sum of sqrt(n_src*(n_dst+1))*rate
    double FD[MAXCLS];     // sum of sqrt(..)*rate*Delta        (gradient of the sweep)
    double RR[MAXCLS];     // rate of the right hop (0 if L==0)
    double RL[MAXCLS];     // rate of the left hop  (0 if R==0)
    double IRR[MAXCLS];    // 1/RR (0 if RR==0)
    double IRL[MAXCLS];
};

// Per-CTA shared-memory scratch of the bond evaluation.  Values that are *dynamically indexed* (class
// prefix sums and class masks during move selection; the bitstring itself in the rare generic-bond
// path) are staged here so that the register arrays stay statically indexed.  Per-thread entries use
// the layout [entry][thread] (conflict-free); exptab is shared by the CTA.
template <int NW, int KC>
struct StepScratch {
    static constexpr int NC = KC * KC;
    double* exptab;  // [2N+2]: exp(-c*(i-N)), i.e. the hop rate for Delta = i-N in [-N+1, N+1]
    double* sqrtab;  // [N+2]:  sqrt(i)
    double* cum;     // [NC][STEP_BLOCK]
    uint32_t* el;    // [KC][NW][STEP_BLOCK]
    uint32_t* er;    // [KC][NW][STEP_BLOCK]
    uint32_t* ew;    // [NW][STEP_BLOCK] extended bitstring   (staged only when a generic bond exists)
    uint32_t* bw;    // [NW][STEP_BLOCK] generic-bond mask
    __host__ __device__ static constexpr size_t exptab_bytes(int N) { return sizeof(double) * (size_t)(3 * N + 4); }
    __host__ __device__ static constexpr size_t bytes(int N) {
        return exptab_bytes(N) + sizeof(double) * NC * STEP_BLOCK +
               sizeof(uint32_t) * (2 * KC * NW + 2 * NW) * STEP_BLOCK;
    }
    __device__ __forceinline__ StepScratch(unsigned char* base, int N) {
        exptab = reinterpret_cast<double*>(base);
        sqrtab = exptab + (2 * N + 2);
        cum = sqrtab + (N + 2);
        el = reinterpret_cast<uint32_t*>(cum + NC * STEP_BLOCK);
        er = el + KC * NW * STEP_BLOCK;
        ew = er + KC * NW * STEP_BLOCK;
        bw = ew + NW * STEP_BLOCK;
    }
    // sweep kernels never select a move: no prefix sums, no staged class masks (3 KB instead of ~40 KB per CTA)
    struct SweepLayout {};
    __host__ __device__ static constexpr size_t bytes_sweep(int N) {
        return exptab_bytes(N) + sizeof(uint32_t) * (2 * NW) * STEP_BLOCK;
    }
    __device__ __forceinline__ StepScratch(unsigned char* base, int N, SweepLayout) {
        exptab = reinterpret_cast<double*>(base);
        sqrtab = exptab + (2 * N + 2);
        cum = nullptr;
        el = er = nullptr;
        ew = reinterpret_cast<uint32_t*>(sqrtab + (N + 2));
        bw = ew + NW * STEP_BLOCK;
    }
    // CTA-wide: tabulate the hop rates once per launch (keeps exp() out of the step loop)
    __device__ __forceinline__ void fill_exptab(int N, double c) const {
        for (int i = threadIdx.x; i < 2 * N + 2; i += blockDim.x) exptab[i] = exp(-c * (double)(i - N));
        for (int i = threadIdx.x; i < N + 2; i += blockDim.x) sqrtab[i] = sqrt((double)i);
        __syncthreads();
    }
};

// ------------------------------------------------------------------------------------------
// multi-word helpers (NW compile-time, arrays stay in registers)
// ------------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void shl1(const uint32_t (&x)[NW], uint32_t (&o)[NW]) {
#pragma unroll
    for (int i = NW - 1; i > 0; --i) o[i] = __funnelshift_l(x[i - 1], x[i], 1);
    o[0] = x[0] << 1;
}
template <int NW>
__device__ __forceinline__ void shr1(const uint32_t (&x)[NW], uint32_t (&o)[NW]) {
#pragma unroll
    for (int i = 0; i < NW - 1; ++i) o[i] = __funnelshift_r(x[i], x[i + 1], 1);
    o[NW - 1] = x[NW - 1] >> 1;
}
// flip bit p (dynamic) of a register array
template <int NW>
__device__ __forceinline__ void flip_bit(uint32_t (&x)[NW], int p) {
    const int w = p >> 5;
    const uint32_t m = 1u << (p & 31);
#pragma unroll
    for (int i = 0; i < NW; ++i) x[i] ^= (i == w) ? m : 0u;
}
template <int NW>
__device__ __forceinline__ int get_bit(const uint32_t (&x)[NW], int p) {
    const int w = p >> 5;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) v = (i == w) ? x[i] : v;
    return (v >> (p & 31)) & 1;
}
// j-th (0-based) set bit of a 32-bit word; requires j < popc(x)
__device__ __forceinline__ int nth_set_bit32(uint32_t x, int j) {
    int pos = 0;
    int c;
    c = __popc(x & 0xFFFFu); if (j >= c) { j -= c; pos += 16; x >>= 16; }
    c = __popc(x & 0xFFu);   if (j >= c) { j -= c; pos += 8;  x >>= 8; }
    c = __popc(x & 0xFu);    if (j >= c) { j -= c; pos += 4;  x >>= 4; }
    c = __popc(x & 0x3u);    if (j >= c) { j -= c; pos += 2;  x >>= 2; }
    c = x & 1u;              if (j >= c) { pos += 1; }
    return pos;
}
__device__ __forceinline__ double int_to_double_fast(int c) {
#ifdef GWZ_I2F_CONVERT
    return (double)c;  // I2F.F64.S32 on the XU pipe
#else
    // exact for 0 <= c < 2^31: one DADD instead of an XU-pipe I2F.F64
    return __hiloint2double(0x43300000, c) - 4503599627370496.0;
#endif
}

// load a stored address (W 64-bit words, word-major SoA: words[w*stride + idx]) into NW regs
template <int NW>
__device__ __forceinline__ void load_addr_soa(const uint64_t* __restrict__ words, size_t stride, size_t idx,
                                              int W, uint32_t (&a)[NW]) {
#pragma unroll
    for (int i = 0; i < (NW + 1) / 2; ++i) {
        uint64_t v = (i < W) ? words[(size_t)i * stride + idx] : 0ull;
        a[2 * i] = (uint32_t)v;
        if (2 * i + 1 < NW) a[2 * i + 1] = (uint32_t)(v >> 32);
    }
}
template <int NW>
__device__ __forceinline__ void store_addr_soa(uint64_t* __restrict__ words, size_t stride, size_t idx, int W,
                                               const uint32_t (&a)[NW]) {
#pragma unroll
    for (int i = 0; i < (NW + 1) / 2; ++i) {
        uint64_t v = a[2 * i];
        if (2 * i + 1 < NW) v |= (uint64_t)a[2 * i + 1] << 32;
        if (i < W) words[(size_t)i * stride + idx] = v;
    }
}
// array-of-addresses layout (n x W), used at the API boundary
template <int NW>
__device__ __forceinline__ void load_addr_aos(const uint64_t* __restrict__ words, size_t idx, int W,
                                              uint32_t (&a)[NW]) {
#pragma unroll
    for (int i = 0; i < (NW + 1) / 2; ++i) {
        uint64_t v = (i < W) ? words[idx * (size_t)W + i] : 0ull;
        a[2 * i] = (uint32_t)v;
        if (2 * i + 1 < NW) a[2 * i + 1] = (uint32_t)(v >> 32);
    }
}
template <int NW>
__device__ __forceinline__ void store_addr_aos(uint64_t* __restrict__ words, size_t idx, int W,
                                               const uint32_t (&a)[NW]) {
#pragma unroll
    for (int i = 0; i < (NW + 1) / 2; ++i) {
        uint64_t v = a[2 * i];
        if (2 * i + 1 < NW) v |= (uint64_t)a[2 * i + 1] << 32;
        if (i < W) words[idx * (size_t)W + i] = v;
    }
}

// e = a with the K low bits of a replicated above the virtual separator at bit B
template <int NW, int KC>
__device__ __forceinline__ void extend(const uint32_t (&a)[NW], const DevModel& dm, uint32_t (&e)[NW]) {
    const uint32_t low = a[0] & ((1u << KC) - 1u);
    const uint32_t lo_part = low << dm.ext_bs;
    const uint32_t hi_part = __funnelshift_l(low, 0u, dm.ext_bs);  // low >> (32 - ext_bs), 0 when ext_bs == 0
#pragma unroll
    for (int i = 0; i < NW; ++i) e[i] = a[i] | (lo_part & dm.extmask_lo[i]) | (hi_part & dm.extmask_hi[i]);
}

// ------------------------------------------------------------------------------------------
// hop application (App. A.1): bond at zero position z, direction right (i -> i+1) or left
// ------------------------------------------------------------------------------------------
template <int NW>
__device__ __forceinline__ void apply_hop(uint32_t (&a)[NW], const DevModel& dm, int z, bool right) {
    if (z != dm.B) {
        // right: bits (z-1, z) = (1,0) -> (0,1);  left: bits (z, z+1) = (0,1) -> (1,0)
        const int p = right ? z - 1 : z;
        flip_bit<NW>(a, p);
        flip_bit<NW>(a, p + 1);
    } else if (right) {
        // mode M -> mode 1: ((a << 1) | 1) & mask(B)
        uint32_t o[NW];
        shl1<NW>(a, o);
        o[0] |= 1u;
#pragma unroll
        for (int i = 0; i < NW; ++i) a[i] = o[i] & dm.amask[i];
    } else {
        // mode 1 -> mode M: (a >> 1) | (1 << (B-1))
        uint32_t o[NW];
        shr1<NW>(a, o);
#pragma unroll
        for (int i = 0; i < NW; ++i) a[i] = o[i];
        flip_bit<NW>(a, dm.B - 1);
    }
}

// ------------------------------------------------------------------------------------------
// Generic bonds (a run of >= KCAP particles on either side): rare, so they are handled by ONE small
// out-of-line routine that reads the staged bitstring from shared memory (dynamic indexing is free
// there) and the tabulated rates.  Keeping this path compact matters: the step loop must stay
// resident in the 32 KB L1.5 instruction cache even when some warps take it.
// ------------------------------------------------------------------------------------------
#ifndef GWZ_BIGPASS_INLINE
#define GWZ_BIGPASS_INLINE __forceinline__
#endif
struct BigOut {
    double S, E, ED;  // mode 0: sums over the generic bonds
    long long d;
    int nocc;
    int z, right;     // mode 1: selected hop
};

// ew/bw/exptab: shared memory (ew, bw already offset by threadIdx.x; stride STEP_BLOCK).
// mode 0: accumulate; mode 1: find the hop whose interval of the running sum contains t.
static __device__ GWZ_BIGPASS_INLINE BigOut big_pass(const uint32_t* ew, const uint32_t* bw, int nw, int B, int N,
                                                      const double* exptab, const double* sqrtab, int mode, double t) {
    BigOut o;
    o.S = o.E = o.ED = 0.0;
    o.d = 0;
    o.nocc = 0;
    o.z = -1;
    o.right = 1;
    double acc = 0.0;
    for (int i = 0; i < nw; ++i) {
        uint32_t m = bw[i * STEP_BLOCK];
        while (m) {
            const int z = 32 * i + (__ffs(m) - 1);
            m &= m - 1;
            // ones directly below z
            int L = 0;
            for (int p = z - 1; p >= 0;) {
                const int b = p & 31;
                const uint32_t x = ~(ew[(p >> 5) * STEP_BLOCK] << (31 - b));
                int c = x ? __clz(x) : 32;
                c = min(c, b + 1);
                L += c;
                if (c < b + 1) break;
                p -= c;
            }
            // ones directly above z; for the periodic bond (z == B) that is mode 1 = trailing ones.
            // Bit B of the staged string is 0, so the scan stops at B at the latest.
            int R = 0;
            for (int p = (z == B) ? 0 : z + 1; p < B;) {
                const int b = p & 31;
                const uint32_t x = ~(ew[(p >> 5) * STEP_BLOCK] >> b);
                int c = x ? (__ffs(x) - 1) : 32;
                c = min(min(c, 32 - b), B - p);
                R += c;
                if (c < 32 - b || p + c >= B) break;
                p += c;
            }
            const double rr = (L > 0) ? exptab[(R - L + 1) + N] : 0.0;
            const double rl = (R > 0) ? exptab[(L - R + 1) + N] : 0.0;
            if (mode == 0) {
                // sqrt(n_src (n_dst+1)) as a product of tabulated roots (<= 1 ulp from the fused root)
                const double er = sqrtab[L] * sqrtab[R + 1] * rr;
                const double el = sqrtab[R] * sqrtab[L + 1] * rl;
                o.S += rr;
                o.S += rl;
                o.E += er + el;
                o.ED += er * (double)(R - L + 1) + el * (double)(L - R + 1);
                o.nocc += (L > 0);
                o.d += (long long)L * (L - 1) / 2;
            } else {
                if (rr > 0.0) {
                    acc += rr;
                    o.z = z;
                    o.right = 1;
                    if (t < acc) return o;
                }
                if (rl > 0.0) {
                    acc += rl;
                    o.z = z;
                    o.right = 0;
                    if (t < acc) return o;
                }
            }
        }
    }
    return o;
}

// ------------------------------------------------------------------------------------------
// Bond classification: EL[l] / ER[r] = bonds whose lower / upper run has length exactly l / r
// (l, r < KCAP); big = bonds with a run >= KCAP on either side.
// ------------------------------------------------------------------------------------------
template <int NW, int KC>
struct BondMasks {
    uint32_t EL[KC][NW];
    uint32_t ER[KC][NW];
    uint32_t big[NW];
};

template <int NW, int KC>
__device__ __forceinline__ void classify_bonds(const uint32_t (&e)[NW], const DevModel& dm, BondMasks<NW, KC>& bm) {
    uint32_t Z[NW], cur[NW], nxt[NW], sh[NW], bigL[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) Z[i] = ~e[i] & dm.zmask[i];
    // lower runs: L_1 = e << 1, L_{k+1} = L_k & (L_k << 1)
    shl1<NW>(e, cur);
#pragma unroll
    for (int i = 0; i < NW; ++i) bm.EL[0][i] = Z[i] & ~cur[i];
#pragma unroll
    for (int k = 1; k < KC; ++k) {
        shl1<NW>(cur, sh);
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            nxt[i] = cur[i] & sh[i];
            bm.EL[k][i] = Z[i] & cur[i] & ~nxt[i];
            cur[i] = nxt[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) bigL[i] = Z[i] & cur[i];
    // upper runs
    shr1<NW>(e, cur);
#pragma unroll
    for (int i = 0; i < NW; ++i) bm.ER[0][i] = Z[i] & ~cur[i];
#pragma unroll
    for (int k = 1; k < KC; ++k) {
        shr1<NW>(cur, sh);
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            nxt[i] = cur[i] & sh[i];
            bm.ER[k][i] = Z[i] & cur[i] & ~nxt[i];
            cur[i] = nxt[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) bm.big[i] = bigL[i] | (Z[i] & cur[i]);
}

// Per-address quantities of kinetic_sample! / LocalEnergyEvaluator
struct BondSums {
    double S;        // sum of hop rates             -> residence time 1/S      (qmc.jl:27,33)
    double E;        // sum sqrt(n_src (n_dst+1)) * rate -> E_loc = D - t*E     (qmc.jl:28,37)
    double ED;       // sum sqrt(..) * rate * Delta  (sweep gradient only)
    long long d;     // sum_i n_i (n_i - 1) / 2      -> D = u*d                 (App. A.3)
    int nocc;        // occupied modes, n_hops = 2*nocc                         (qmc.jl:15)
};

// How the rare generic bonds are evaluated is a policy of the state representation.  Unary bitstring:
// stage the extended string + generic-bond mask of this thread in shared memory, run big_pass.
template <int NW, int KC>
struct UnaryBig {
    const uint32_t (&e)[NW];
    const uint32_t (&big)[NW];
    const DevModel& dm;
    const StepScratch<NW, KC>& sc;
    __device__ __forceinline__ bool any() const {
        uint32_t a = 0u;
#pragma unroll
        for (int i = 0; i < NW; ++i) a |= big[i];
        return a != 0u;
    }
    // mode 0: stage + sums; mode 1: selection (words are still staged from mode 0 of the same step)
    __device__ __forceinline__ BigOut pass(int mode, double t) const {
        const int tid = threadIdx.x;
        if (mode == 0) {
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                sc.ew[i * STEP_BLOCK + tid] = e[i];
                sc.bw[i * STEP_BLOCK + tid] = big[i];
            }
        }
        return big_pass(sc.ew + tid, sc.bw + tid, NW, dm.B, dm.N, sc.exptab, sc.sqrtab, mode, t);
    }
};

// Sum over all bonds.  The running sum of class weights cnt[k]*FS[k] (k ascending) after class k is
// handed to `sink(k, S)` (kept for the class-ordered move selection; a no-op for the sweep).  Class 0
// (L = R = 0: an empty bond) carries no hop and is skipped.  WITH_GRAD adds the ED sum needed by
// the sweep gradient.  Returns the total weight of the regular classes.
struct NoCumSink {
    __device__ __forceinline__ void operator()(int, double) const {}
};
struct SmemCumSink {
    double* cum;  // already offset by threadIdx.x
    __device__ __forceinline__ void operator()(int k, double v) const { cum[k * STEP_BLOCK] = v; }
};

template <int NW, int KC, bool WITH_GRAD, typename SINK, typename BIG>
__device__ __forceinline__ double bond_sums(const DevTables& tb, const BondMasks<NW, KC>& bm, SINK&& sink,
                                            const BIG& big, BondSums& out, double& big_S) {
    double S = 0.0, E = 0.0, ED = 0.0;
    int nocc = 0;
    int dd = 0;
    sink(0, 0.0);
#pragma unroll
    for (int l = 0; l < KC; ++l) {
        int rowcnt = 0;
#pragma unroll
        for (int r = 0; r < KC; ++r) {
            const int k = l * KC + r;
            if (k == 0) continue;
            int c = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) c += __popc(bm.EL[l][i] & bm.ER[r][i]);
            const double cd = int_to_double_fast(c);
            S = fma(cd, tb.FS[k], S);
            E = fma(cd, tb.FE[k], E);
            if (WITH_GRAD) ED = fma(cd, tb.FD[k], ED);
            sink(k, S);
            rowcnt += c;
        }
        if (l > 0) nocc += rowcnt;
        dd += rowcnt * (l * (l - 1) / 2);
    }
    const double reg_total = S;
    long long d = dd;
    // rare: bonds next to a mode with >= KC particles
    double bS = 0.0;
    if (big.any()) {
        const BigOut o = big.pass(0, 0.0);
        bS = o.S;
        E += o.E;
        if (WITH_GRAD) ED += o.ED;
        nocc += o.nocc;
        d += o.d;
    }
    big_S = bS;
    out.S = reg_total + bS;
    out.E = E;
    out.ED = ED;
    out.d = d;
    out.nocc = nocc;
    return reg_total;
}

// stage the class masks for the dynamic (l, r) lookup of the move selection
template <int NW, int KC>
__device__ __forceinline__ void stage_masks(const BondMasks<NW, KC>& bm, const StepScratch<NW, KC>& sc) {
    const int t = threadIdx.x;
#pragma unroll
    for (int q = 0; q < KC; ++q)
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            sc.el[(q * NW + i) * STEP_BLOCK + t] = bm.EL[q][i];
            sc.er[(q * NW + i) * STEP_BLOCK + t] = bm.ER[q][i];
        }
}

// Class-ordered move selection: partition [0,S) into consecutive intervals, one per hop, ordered
// by (class id, direction right-then-left, bond position), big bonds last.  T in [0,S).
// Needs the prefix sums (SmemCumSink) and masks (stage_masks) of this thread in `sc`.
// Returns the bond position z and the direction.
template <int NW, int KC, typename BIG>
__device__ __forceinline__ void select_move(const DevTables& tb, const StepScratch<NW, KC>& sc, const BIG& big,
                                            double reg_total, double big_S, double T, int& z_out, bool& right_out) {
    constexpr int NC = KC * KC;
    constexpr int BISECT = (NC - 1 <= 16) ? 4 : 5;  // ceil(log2(NC-1)) for KC = 4, 5
    static_assert(NC - 1 <= 32, "bisection depth");
    if (T < reg_total || !(big_S > 0.0)) {
        if (!(T < reg_total)) T = __longlong_as_double(__double_as_longlong(reg_total) - 1);
        const double* cum = sc.cum + threadIdx.x;
        // first class k in [1,NC-1] with T < cum[k] (cum[NC-1] = reg_total > T) by bisection
        int lo = 1, hi = NC - 1;
#pragma unroll
        for (int it = 0; it < BISECT; ++it) {
            const int mid = (lo + hi) >> 1;
            const bool below = T < cum[mid * STEP_BLOCK];
            hi = below ? mid : hi;
            lo = below ? lo : min(mid + 1, hi);
        }
        const int idx = lo;
        const double base = cum[(idx - 1) * STEP_BLOCK];
        const int l = idx / KC, r = idx - l * KC;
        // masks of class (l,r), their popcounts
        uint32_t m[NW];
        int pc[NW];
        int c = 0;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            m[i] = sc.el[(l * NW + i) * STEP_BLOCK + threadIdx.x] & sc.er[(r * NW + i) * STEP_BLOCK + threadIdx.x];
            pc[i] = __popc(m[i]);
            c += pc[i];
        }
        const double cd = int_to_double_fast(c);
        const double t = T - base;
        const double wR = cd * tb.RR[idx];
        const bool right = (t < wR) || (r == 0);
        int j = right ? (int)(t * tb.IRR[idx]) : (int)((t - wR) * tb.IRL[idx]);
        j = max(0, min(j, c - 1));
        // j-th bond of the class
        int pos = 0;
        uint32_t word = m[NW - 1];
        bool found = false;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            const bool here = !found && (j < pc[i]);
            word = here ? m[i] : word;
            pos = here ? 32 * i : pos;
            found = found || here;
            j -= found ? 0 : pc[i];
        }
        z_out = pos + nth_set_bit32(word, j);
        right_out = right;
    } else {
        // walk the generic bonds in the same order as bond_sums (their words are still staged)
        const BigOut o = big.pass(1, T - reg_total);
        z_out = o.z;
        right_out = o.right != 0;
    }
}

// 1-based reference hop index (Rimu offdiagonals order) of the hop (z, right) -- probe only
template <int NW>
__device__ __forceinline__ int reference_hop_index(const uint32_t (&e)[NW], const DevModel& dm, int z, bool right) {
    // tops of runs: e(p)=1 and e(p+1)=0, p in [0,B-1]
    uint32_t sh[NW];
    shr1<NW>(e, sh);
    int rank = 0;  // occupied modes whose top is below the limit
    const int limit = right ? z - 1 : z;  // tops at positions <= limit belong to modes <= i
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        uint32_t tops = e[i] & ~sh[i] & dm.amask[i];
        const int lo = 32 * i;
        uint32_t keep;
        if (limit < lo) keep = 0u;
        else if (limit >= lo + 31) keep = 0xFFFFFFFFu;
        else keep = (2u << (limit - lo)) - 1u;
        rank += __popc(tops & keep);
    }
    if (right) return 2 * rank - 1;          // source = mode i (its top is at z-1, counted)
    if (z == dm.B) return 2;                 // source = mode 1 hopping left (wraps to M)
    return 2 * (rank + 1);                   // source = mode i+1 = next occupied mode after the tops <= z
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11): key = seed, counter = (step >> 1, walker id)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// same generator with the ten round keys precomputed on the host (kernel-argument constant bank):
// removes the key schedule from the step loop
struct PhiloxKeys {
    uint32_t k[20];  // k[2r], k[2r+1] = round-r keys
};
__host__ __device__ inline PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys pk;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        pk.k[2 * r] = k0;
        pk.k[2 * r + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return pk;
}
__device__ __forceinline__ void philox4x32_10_keyed(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    const PhiloxKeys& pk, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ pk.k[2 * r], n2 = hi0 ^ c3 ^ pk.k[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01_from_words(uint32_t lo, uint32_t hi) {
    const uint64_t r = ((uint64_t)hi << 32) | lo;
    return (double)(r >> 11) * 0x1.0p-53;
}

// ------------------------------------------------------------------------------------------
// Enumerating evaluation in the reference's offdiagonals order (Rimu hopnextneighbour order:
// occupied modes ascending, right hop then left hop; App. A.4).  Calls
//     f(hop_index_1based, z, right, n_src, n_dst) -> bool (true = stop)
// for every hop.  Walks the address bit by bit (uniform trip count); generic in the occupations.
// ------------------------------------------------------------------------------------------
template <int NW, typename F>
__device__ __forceinline__ void enumerate_hops(const uint32_t (&a)[NW], const DevModel& dm, F&& f) {
    // occupation of mode M (ones directly below bit B) and of mode 1 (trailing ones)
    int nM = 0, n1 = 0;
    {
        bool run = true;
        for (int p = dm.B - 1; p >= 0 && run; --p) {
            run = get_bit<NW>(a, p) != 0;
            nM += run ? 1 : 0;
        }
        run = true;
        for (int p = 0; p < dm.B && run; ++p) {
            run = get_bit<NW>(a, p) != 0;
            n1 += run ? 1 : 0;
        }
    }
    // stream the modes with a one-mode delay so that (n_prev, n_cur, n_next) are all known
    int n_run = 0;    // ones read so far in the run being scanned
    int n_prev = nM;  // occupation of the mode before the pending one
    int n_pend = -1;  // occupation of the pending mode, -1 = none yet
    int z_pend = 0;   // separator that closes the pending mode
    int hop = 0;
    bool stop = false;
    auto emit = [&](int nprev, int ncur, int nnext, int zc) {
        // the pending mode's run sits at bits zc-ncur .. zc-1; the bond below it is the separator
        // at zc-ncur-1, or the periodic bond B for mode 1
        if (ncur > 0 && !stop) {
            ++hop;
            stop = f(hop, zc, true, ncur, nnext);
            if (!stop) {
                ++hop;
                const int zl = zc - ncur - 1;
                stop = f(hop, zl < 0 ? dm.B : zl, false, ncur, nprev);
            }
        }
    };
    for (int p = 0; p <= dm.B && !stop; ++p) {
        const int bit = (p < dm.B) ? get_bit<NW>(a, p) : 0;
        if (bit) {
            ++n_run;
        } else {
            if (n_pend >= 0) {
                emit(n_prev, n_pend, n_run, z_pend);
                n_prev = n_pend;
            }
            n_pend = n_run;
            z_pend = p;
            n_run = 0;
        }
    }
    // the last pending mode is mode M; its right neighbour is mode 1
    if (!stop && n_pend >= 0) emit(n_prev, n_pend, n1, z_pend);
}

// colex unranking: the rank-th N-subset of bit positions, as a 64-bit word
__device__ __forceinline__ uint64_t unrank_colex(uint64_t rank, int N, int B, const uint64_t* __restrict__ binom) {
    uint64_t x = 0;
    int p = B - 1;
    for (int i = N; i >= 1; --i) {
        while (binom[p * 65 + i] > rank) --p;  // largest p with C(p,i) <= rank
        x |= 1ull << p;
        rank -= binom[p * 65 + i];
        --p;
    }
    return x;
}
// next integer with the same popcount (Gosper)
__device__ __forceinline__ uint64_t next_same_popcount(uint64_t x) {
    const uint64_t t = x | (x - 1);
    return (t + 1) | (((~t & (0 - ~t)) - 1) >> (__ffsll((long long)x)));
}

__device__ __forceinline__ uint32_t next_same_popcount32(uint32_t x) {
    const uint32_t t = x | (x - 1u);
    return (t + 1u) | (((~t & (0u - ~t)) - 1u) >> (__ffs((int)x)));
}
// warp / block reductions of doubles (deterministic order)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace gwz
