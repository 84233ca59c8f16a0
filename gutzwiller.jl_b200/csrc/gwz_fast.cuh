// gwz_fast.cuh -- production kinetic_sample! (src/kinetic-vqmc/qmc.jl:14-44) with everything a hop does not change
// carried from step to step as exact integers and bit masks.
//
// A hop across bond z changes the occupations of modes z and z+1, hence the classes (L,R) of the three bonds z-1, z,
// z+1 and nothing else.  Regular bonds (both occupations < 4) fall into 16 classes (l, r); their two hops have the
// rates exp(-c (r-l+1)) (right, if l > 0) and exp(-c (l-r+1)) (left, if r > 0): only SIX distinct rates r_g,
// g = Delta + 2 = 0..5.  The walker carries
//   * np   -- the number of regular hops per rate class g, six 10-bit fields of one 64-bit integer (register);
//   * E    -- sum over regular bonds of sqrt(n_src (n_dst + 1)) x rate (the off-diagonal part of E_loc), a double
//             that is updated by table deltas and re-summed from the occupations every FAST_REFRESH steps and
//             whenever it has fallen 16-fold below its peak since the last re-summation, which bounds its rounding
//             drift below 5e-13 relative for any dynamic range of the rates; dsum = sum n (n-1) / 2 (int) and the
//             numbers of occupied modes / of modes with >= 4 particles (two 16-bit fields of one int);
//   * the occupations as bytes (shared memory, [word of 4 modes][thread]);
//   * HOP MASKS over the bonds (shared memory, [row][thread][words]): HR_g = {z : the right hop across z is regular
//     and of rate class g}, HL_g likewise for the left hops, g = 0..5, plus the mask of the bonds that touch a mode
//     with >= 4 particles ("generic" bonds).
// Per step: S = sum_g n_g r_g is evaluated from the integers in a fixed order (6 FMA: a pure function of the
// address).  The move is drawn in two levels.  (1) the rate class g by the running sums of S.  (2) all hops of a rate
// class are equally likely, so the j-th of them in the order (direction, bond position) is taken: the j-th set bit of
// the row pair (HR_g, HL_g) -- two vector loads, no logic.  The hop then updates the three touched bonds from tables
// indexed by the capped occupations before and after: 16-byte entries (dE, dnp) plus one word with the four hop-mask
// rows the bond leaves and enters, i.e. three additions to (E, np) and twelve single-bit shared-memory XORs; a
// fourth table gives the change of the occupied / >= 4 counts.  Bonds touching a mode with >= 4 particles are rare;
// their sums are cached and re-evaluated from the occupation bytes only when a hop touches such a mode, as in
// gwz_incr.cuh.
//
// What bounds the kernel is the shared-memory data pipe (wavefronts per step), then instruction issue: tables that are
// indexed by data (different entries per lane) cost a wavefront per distinct bank group, so they are kept dense and
// small (16-byte entries, row ids as bytes), and everything that depends only on the bond position is computed.
//
// The partition of [0, S) into hops is ordered by (rate class, direction, bond position), generic bonds last: the same
// distribution as the reference's pick_random_from_cumsum (qmc.jl:52-61) with a different u -> hop map (tests:
// stratified u per address, recorded trajectories re-evaluated per step, energies against the exact sweep).
// Needs 4 <= M <= 250 (the four modes around a bond are distinct; counts fit their fields).
#pragma once
#include "gwz_incr.cuh"

namespace gwz {

// threads per CTA of the production kernel (its own scratch layout) and resident CTAs per SM: 768 threads per SM with
// up to 85 registers for M <= 64 (1024 threads with 64 registers measured slower: the kernel is bound by the
// shared-memory pipe and by issue, not by latency)
#ifndef GWZ_FAST_FB
#define GWZ_FAST_FB 256
#endif
#ifndef GWZ_FAST_MINB
#define GWZ_FAST_MINB 3
#endif
__host__ __device__ constexpr int fast_block(int nwm) { return nwm <= 4 ? GWZ_FAST_FB : 256; }
__host__ __device__ constexpr int fast_min_blocks(int nwm) { return nwm <= 2 ? GWZ_FAST_MINB : (nwm <= 4 ? 512 / fast_block(nwm) : 1); }
constexpr int FAST_REFRESH = 64;   // steps between re-summations of E (a power of two)
// hop-mask rows per thread: [0..5] HR_g | [6..11] HL_g | [12] dummy (absorbs the XORs of absent hops) | [13] generic bonds
constexpr int HM_R = 0, HM_L = 6, HM_DUMMY = 12, HM_BIG = 13, HM_ROWS = 14;

struct FastEntry {  // what one bond's class change adds to (E, np)
    double dE;
    unsigned long long dn;
};
static_assert(sizeof(FastEntry) == 16, "FastEntry must be one 16-byte shared-memory access");

// bit position of the 10-bit field of rate class g inside np (three fields per 32-bit half)
__host__ __device__ constexpr int fast_field_pos(int g) { return g < 3 ? 10 * g : 32 + 10 * (g - 3); }

// rates of the six regular rate classes, passed as a kernel argument (constant bank)
struct FastRates {
    double r[6];   // exp(-c (g - 2))
    double rs[6];  // r[g] / 2^(field position inside its 32-bit half): multiplies the masked, unshifted field
};

// 32-bit shared-window addressing for the step loop: one base register, every region at a compile-time offset
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds64(uint32_t a, uint32_t& v0, uint32_t& v1) {
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(a));
}
__device__ __forceinline__ void lds128(uint32_t a, uint32_t& v0, uint32_t& v1, uint32_t& v2, uint32_t& v3) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a));
}
__device__ __forceinline__ uint32_t ldsu8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void stsu8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ double lds64f(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds_entry_sums(uint32_t a, double& dE, unsigned long long& dn) {
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(*reinterpret_cast<unsigned long long*>(&dE)), "=l"(dn) : "r"(a));
}
__device__ __forceinline__ void red_xor(uint32_t a, uint32_t v) { asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// 1 / x for a normal positive x: hardware seed + two Newton steps (relative error < 2^-51)
__device__ __forceinline__ double rcp_newton(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

template <int NWM>
struct FastScratch {
    // hop-mask rows: NWV words per thread (NWM rounded up to a power of two), stored as NCH chunks of VW words so that
    // a row is read by NCH conflict-free vector loads: [row][chunk][thread][VW]
    static constexpr int FB = fast_block(NWM);
    static constexpr int NWV = NWM <= 1 ? 1 : (NWM <= 2 ? 2 : (NWM <= 4 ? 4 : 8));
    static constexpr int VW = NWV < 4 ? NWV : 4;
    static constexpr int NCH = NWV / VW;
    static constexpr uint32_t CH_BYTES = 4u * VW * FB;
    static constexpr uint32_t ROW_BYTES = CH_BYTES * NCH;
    static constexpr int ONR_WORDS = 8 * NWM;                 // occupation bytes, 4 modes per word: [word][thread]
    static constexpr uint32_t ONR_STRIDE = 4u * FB;
    static constexpr int NTL = 5 * 15, NTM = 15 * 15, NTR = 15 * 5;
    // byte offsets from the start of the dynamic shared memory (step_fast addresses everything as base + constant)
    static constexpr uint32_t OFF_TL = 0,                        // [a4][ix]   bond z-1: (a, x) -> (a, x')
                              OFF_TM = OFF_TL + 16u * NTL,       // [ix][iy]   bond z:   (x, y) -> (x', y')
                              OFF_TR = OFF_TM + 16u * NTM,       // [iy][b4]   bond z+1: (y, b) -> (y', b)
                              // the hop-mask rows of the same transitions, one byte each: right hop before | left hop
                              // before | right hop after | left hop after (same indices as TL, TM, TR)
                              OFF_RL = OFF_TR + 16u * NTR, OFF_RM = OFF_RL + 4u * 76, OFF_RR = OFF_RM + 4u * 228,
                              OFF_TC = OFF_RR + 4u * 76,         // [ix][iy]   int32: d(occupied) + 65536 d(modes >= 4)
                              OFF_IRATE = OFF_TC + 4u * 228,     // [8]        1 / r_g
                              OFF_FE5 = OFF_IRATE + 8u * 8,      // [5][5]     E weight of a bond of capped class (l, r)
                              OFF_NTH = OFF_FE5 + 8u * 26,       // [256]      positions of the set bits of a byte
                              OFF_ROWS = OFF_NTH + 4u * 256,     // hop-mask rows
                              OFF_ONR = OFF_ROWS + HM_ROWS * ROW_BYTES, OFF_EXP = OFF_ONR + ONR_WORDS * ONR_STRIDE;
    __host__ __device__ static constexpr size_t bytes(int N) { return OFF_EXP + sizeof(double) * (size_t)(3 * N + 4); }

    unsigned char* base;
    double* exptab;  // [2N+2]  generic-bond path
    double* sqrtab;  // [N+2]
    __device__ __forceinline__ FastScratch(unsigned char* b, int N) : base(b) {
        exptab = reinterpret_cast<double*>(b + OFF_EXP);
        sqrtab = exptab + (2 * N + 2);
    }
    // word i of hop-mask row `row` of this thread / occupation word q of this thread (generic pointers: set-up code)
    __device__ __forceinline__ uint32_t* row_word(int row, int i) const {
        return reinterpret_cast<uint32_t*>(base + OFF_ROWS + (uint32_t)row * ROW_BYTES + (uint32_t)(i / VW) * CH_BYTES +
                                           4u * VW * threadIdx.x + 4u * (uint32_t)(i % VW));
    }
    __device__ __forceinline__ uint32_t* onr_word(int q) const {
        return reinterpret_cast<uint32_t*>(base + OFF_ONR + (uint32_t)q * ONR_STRIDE + 4u * threadIdx.x);
    }
    __host__ __device__ static constexpr uint32_t onr_off(int m) { return (uint32_t)(m >> 2) * ONR_STRIDE + (uint32_t)(m & 3); }
    __host__ __device__ static constexpr uint32_t bond_off(int z) {
        return (uint32_t)((z >> 5) / VW) * CH_BYTES + 4u * (uint32_t)((z >> 5) % VW);
    }
    static_assert(ROW_BYTES * HM_ROWS < (1u << 24) && (ROW_BYTES & 0xFFu) == 0, "row offsets are ids x ROW_BYTES");

    struct Contrib {  // one bond of capped class (l, r): its part of (E, np) and the rows holding its two hops
        double e;
        unsigned long long n;
        int rR, rL;
    };
    __device__ static __forceinline__ Contrib contrib(const DevTables& tb, int l, int r) {
        Contrib c;
        c.e = 0.0;
        c.n = 0ull;
        c.rR = c.rL = HM_DUMMY;
        if (l < 4 && r < 4 && l >= 0 && r >= 0) {
            c.e = tb.FE[4 * l + r];
            if (l >= 1) {
                c.n += 1ull << fast_field_pos(r - l + 3);
                c.rR = HM_R + (r - l + 3);
            }
            if (r >= 1) {
                c.n += 1ull << fast_field_pos(l - r + 3);
                c.rL = HM_L + (l - r + 3);
            }
        }
        return c;
    }
    __device__ static __forceinline__ FastEntry delta(const Contrib& a, const Contrib& b, uint32_t& rows) {  // b -> a
        FastEntry e;
        e.dE = a.e - b.e;
        e.dn = a.n - b.n;
        rows = (uint32_t)b.rR | ((uint32_t)b.rL << 8) | ((uint32_t)a.rR << 16) | ((uint32_t)a.rL << 24);
        return e;
    }
    __device__ __forceinline__ void fill_tables(int N, const DevTables& tb) const {
        FastEntry* TL = reinterpret_cast<FastEntry*>(base + OFF_TL);
        FastEntry* TM = reinterpret_cast<FastEntry*>(base + OFF_TM);
        FastEntry* TR = reinterpret_cast<FastEntry*>(base + OFF_TR);
        int32_t* TC = reinterpret_cast<int32_t*>(base + OFF_TC);
        uint32_t* RL = reinterpret_cast<uint32_t*>(base + OFF_RL);
        uint32_t* RM = reinterpret_cast<uint32_t*>(base + OFF_RM);
        uint32_t* RR = reinterpret_cast<uint32_t*>(base + OFF_RR);
        double* irate = reinterpret_cast<double*>(base + OFF_IRATE);
        double* FE5 = reinterpret_cast<double*>(base + OFF_FE5);
        uint32_t* nth = reinterpret_cast<uint32_t*>(base + OFF_NTH);
        for (int i = threadIdx.x; i < 2 * N + 2; i += blockDim.x) exptab[i] = exp(-tb.c * (double)(i - N));
        for (int i = threadIdx.x; i < N + 2; i += blockDim.x) sqrtab[i] = sqrt((double)i);
        // ix = 3 x4 + (x' - x4 + 1) encodes the capped occupation before and after (x' in {x4-1, x4, x4+1})
        for (int i = threadIdx.x; i < NTL; i += blockDim.x) {
            const int a = i / 15, ix = i % 15, x = ix / 3, x2 = x + ix % 3 - 1;
            TL[i] = delta(contrib(tb, a, x2), contrib(tb, a, x), RL[i]);
        }
        for (int i = threadIdx.x; i < NTM; i += blockDim.x) {
            const int ix = i / 15, iy = i % 15, x = ix / 3, x2 = x + ix % 3 - 1, y = iy / 3, y2 = y + iy % 3 - 1;
            TM[i] = delta(contrib(tb, x2, y2), contrib(tb, x, y), RM[i]);
            const int docc = (int)(x == 0 && x2 > 0) - (int)(x > 0 && x2 == 0) + (int)(y == 0 && y2 > 0) - (int)(y > 0 && y2 == 0);
            const int dbig = (int)(x2 == 4) - (int)(x == 4) + (int)(y2 == 4) - (int)(y == 4);
            TC[i] = docc + 65536 * dbig;
        }
        for (int i = threadIdx.x; i < NTR; i += blockDim.x) {
            const int iy = i / 5, b = i % 5, y = iy / 3, y2 = y + iy % 3 - 1;
            TR[i] = delta(contrib(tb, y2, b), contrib(tb, y, b), RR[i]);
        }
        if (threadIdx.x < 8) irate[threadIdx.x] = exp(tb.c * (double)((int)threadIdx.x - 2));
        if (threadIdx.x < 25) {
            const int l = threadIdx.x / 5, r = threadIdx.x % 5;
            FE5[threadIdx.x] = (l < 4 && r < 4) ? tb.FE[4 * l + r] : 0.0;
        }
        for (int b = threadIdx.x; b < 256; b += blockDim.x) {
            uint32_t e = 0u;
            int j = 0;
            for (int q = 0; q < 8; ++q)
                if ((b >> q) & 1) e |= (uint32_t)q << (4 * j++);
            nth[b] = e;
        }
        __syncthreads();
    }
};

__host__ inline FastRates fast_rates(double c) {
    FastRates fr;
    for (int g = 0; g < 6; ++g) {
        fr.r[g] = exp(-c * (double)(g - 2));
        fr.rs[g] = ldexp(fr.r[g], -(fast_field_pos(g) & 31));
    }
    return fr;
}

// shared-window addresses kept in registers for the whole kernel: made opaque to the compiler, which would otherwise
// recompute them from the special registers several times per step
struct FastAddr {
    uint32_t sb;    // start of the dynamic shared memory
    uint32_t tb;    // sb + 4 threadIdx.x:      this thread's column of the [word][thread] regions
    uint32_t rows;  // sb + 4 VW threadIdx.x:   this thread's column of the hop-mask rows (+ OFF_ROWS)
};
template <int VW>
__device__ __forceinline__ FastAddr fast_addr(const void* smem) {
    FastAddr a;
    a.sb = smem_addr(smem);
    a.tb = a.sb + 4u * threadIdx.x;
    a.rows = a.sb + 4u * VW * threadIdx.x;
    asm volatile("" : "+r"(a.sb), "+r"(a.tb), "+r"(a.rows));
    return a;
}

struct FastState {
    unsigned long long np;  // regular hops per rate class, 10-bit fields
    double E;
    int dsum;
    int cnt;  // occupied modes + 65536 x modes with >= 4 particles
    int epk;  // high word of the largest E since the last re-summation (E >= 0: the high words order like the values)
    double bigS, bigE;  // cached sums over the generic bonds
    bool big_valid;
};

// sums / selection over the generic bonds (a neighbour with >= 4 particles), from the occupation bytes;
// same order and arithmetic as PlanesBig::pass
template <int NWM>
__device__ __forceinline__ BigOut fast_big_pass(const uint32_t (&bigm)[NWM], const DevModel& dm, const FastScratch<NWM>& sc,
                                                int mode, double t) {
    using SC = FastScratch<NWM>;
    const unsigned char* ob = sc.base + SC::OFF_ONR + 4u * threadIdx.x;
    BigOut o;
    o.S = o.E = o.ED = 0.0;
    o.d = 0;
    o.nocc = 0;
    o.z = -1;
    o.right = 1;
    double acc = 0.0;
    bool done = false;
    const int N = dm.N, M = dm.M;
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
        uint32_t m = bigm[i];
        while (m && !done) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int z = 32 * i + b;
            const int zn = (z + 1 == M) ? 0 : z + 1;
            const int L = ob[SC::onr_off(z)], R = ob[SC::onr_off(zn)];
            const double rr = (L > 0) ? sc.exptab[(R - L + 1) + N] : 0.0;
            const double rl = (R > 0) ? sc.exptab[(L - R + 1) + N] : 0.0;
            if (mode == 0) {
                const double er = sc.sqrtab[L] * sc.sqrtab[R + 1] * rr;
                const double el = sc.sqrtab[R] * sc.sqrtab[L + 1] * rl;
                o.S += rr;
                o.S += rl;
                o.E += er + el;
            } else {
                if (rr > 0.0) {
                    acc += rr;
                    o.z = z;
                    o.right = 1;
                    if (t < acc) done = true;
                }
                if (!done && rl > 0.0) {
                    acc += rl;
                    o.z = z;
                    o.right = 0;
                    if (t < acc) done = true;
                }
            }
        }
    }
    return o;
}

// E re-summed from the occupation bytes (the weights of the bonds M-1, 0, 1, ..., M-2 in this order)
template <int NWM>
__device__ __forceinline__ double fast_resum_E(uint32_t tb, int M) {
    using SC = FastScratch<NWM>;
    const uint32_t sb = tb - 4u * threadIdx.x;
    const uint32_t ob = tb + SC::OFF_ONR, fe = sb + SC::OFF_FE5;
    // E = sum over m of FE[n_(m-1)][n_m]: the bond (M-1, 0) first, then the bonds 0 .. M-2; full words without
    // per-byte guards, then the last M & 3 modes
    const int nfull = M >> 2, last = M - 1;
    const uint32_t wl = lds32(ob + (uint32_t)(last >> 2) * SC::ONR_STRIDE);
    uint32_t prev = fe + 40u * min((wl >> (8 * (last & 3))) & 0xFFu, 4u);  // row n of the weight table
    double E = 0.0;
#pragma unroll 1
    for (int q = 0; q < nfull; ++q) {
        const uint32_t w = lds32(ob + (uint32_t)q * SC::ONR_STRIDE);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t n = min(__byte_perm(w, 0u, 0x4440u + r), 4u);
            E += lds64f(prev + 8u * n);
            prev = fe + 40u * n;
        }
    }
    if (M & 3) {
        const uint32_t w = lds32(ob + (uint32_t)nfull * SC::ONR_STRIDE);
#pragma unroll
        for (int r = 0; r < 3; ++r)
            if (r < (M & 3)) {
                const uint32_t n = min(__byte_perm(w, 0u, 0x4440u + r), 4u);
                E += lds64f(prev + 8u * n);
                prev = fe + 40u * n;
            }
    }
    return E;
}

// (re)build the carried state of a walker from its planes
template <int NWM, int NP>
__device__ __forceinline__ void fast_init(const uint32_t (&p)[NP][NWM], const DevModel& dm, const FastScratch<NWM>& sc,
                                          uint32_t tb, FastState& s) {
    using SC = FastScratch<NWM>;
    uint32_t BIG[NWM];
    big_plane<NWM, NP>(p, BIG);
    // occupation bytes: planes 0 and 1 everywhere, the higher planes only where n >= 4
#pragma unroll
    for (int i = 0; i < NWM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t n0 = (p[0][i] >> (4 * j)) & 0xFu, n1 = (p[1][i] >> (4 * j)) & 0xFu;
            uint32_t word = ((n0 * 0x00204081u) & 0x01010101u) | (((n1 * 0x00204081u) & 0x01010101u) << 1);
            uint32_t hi = (BIG[i] >> (4 * j)) & 0xFu;
            while (hi) {
                const int t = __ffs(hi) - 1;
                hi &= hi - 1;
                const int b = 4 * j + t;
                uint32_t n = 0u;
#pragma unroll
                for (int k = 2; k < NP; ++k) n |= ((p[k][i] >> b) & 1u) << k;
                word |= n << (8 * t);
            }
            *sc.onr_word(8 * i + j) = word;
        }
    // one-hot masks of the occupations and of the right neighbours' occupations
    uint32_t A[5][NWM], Q[5][NWM];
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
        const uint32_t ok = ~BIG[i] & dm.mmask[i];
        A[0][i] = ~p[0][i] & ~p[1][i] & ok;
        A[1][i] = p[0][i] & ~p[1][i] & ok;
        A[2][i] = ~p[0][i] & p[1][i] & ok;
        A[3][i] = p[0][i] & p[1][i] & ok;
        A[4][i] = BIG[i];
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) ring_next<NWM>(A[q], dm, Q[q]);
    // hop masks: right hop of class g from l = 1..3 particles onto r = l + g - 3; left hop from r = 1..3 onto l = r + g - 3
    unsigned long long np = 0ull;
#pragma unroll
    for (int g = 0; g < 6; ++g) {
        int c = 0;
#pragma unroll
        for (int i = 0; i < SC::NWV; ++i) {
            uint32_t hr = 0u, hl = 0u;
            if (i < NWM) {
#pragma unroll
                for (int l = 1; l < 4; ++l) {
                    const int o = l + g - 3;
                    if (o >= 0 && o < 4) {
                        hr |= A[l][i < NWM ? i : 0] & Q[o][i < NWM ? i : 0];
                        hl |= A[o][i < NWM ? i : 0] & Q[l][i < NWM ? i : 0];
                    }
                }
            }
            *sc.row_word(HM_R + g, i) = hr;
            *sc.row_word(HM_L + g, i) = hl;
            c += __popc(hr) + __popc(hl);
        }
        np += (unsigned long long)c << fast_field_pos(g);
    }
#pragma unroll
    for (int i = 0; i < SC::NWV; ++i) {
        *sc.row_word(HM_DUMMY, i) = 0u;
        *sc.row_word(HM_BIG, i) = i < NWM ? ((A[4][i < NWM ? i : 0] | Q[4][i < NWM ? i : 0]) & dm.mmask[i < NWM ? i : 0]) : 0u;
    }
    // sum n (n - 1) / 2, occupied modes, modes with >= 4 particles
    int n1 = 0, n2 = 0, n3 = 0, dbig = 0, nbig = 0;
    const unsigned char* ob = sc.base + SC::OFF_ONR + 4u * threadIdx.x;
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
        n1 += __popc(A[1][i]);
        n2 += __popc(A[2][i]);
        n3 += __popc(A[3][i]);
        uint32_t m = BIG[i];
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int n = ob[SC::onr_off(32 * i + b)];
            dbig += n * (n - 1) / 2;
            ++nbig;
        }
    }
    s.np = np;
    s.dsum = n2 + 3 * n3 + dbig;
    s.cnt = (n1 + n2 + n3 + nbig) + 65536 * nbig;
    // E by class counts, in the order of bond_sums (class id 4 l + r ascending)
    {
        const double* FE5 = reinterpret_cast<const double*>(sc.base + SC::OFF_FE5);
        double E = 0.0;
#pragma unroll
        for (int l = 0; l < 4; ++l)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (l == 0 && r == 0) continue;
                int c = 0;
#pragma unroll
                for (int i = 0; i < NWM; ++i) c += __popc(A[l][i] & Q[r][i]);
                E = fma(int_to_double_fast(c), FE5[5 * l + r], E);
            }
        s.E = E;
    }
    s.epk = __double2hiint(s.E);
    s.bigS = s.bigE = 0.0;
    s.big_valid = nbig == 0;
}

// the planes of the walker from its occupation bytes
template <int NWM, int NP>
__device__ __forceinline__ void fast_store_planes(const FastScratch<NWM>& sc, const FastState& s, uint32_t (&p)[NP][NWM]) {
    const bool any_big = s.cnt > 0xFFFF;
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
#pragma unroll
        for (int k = 0; k < NP; ++k) p[k][i] = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t w = *sc.onr_word(8 * i + j);
            // bit 0 (bit 1) of the four bytes gathered into a nibble
            p[0][i] |= (((w & 0x01010101u) * 0x01020408u) >> 24) << (4 * j);
            p[1][i] |= ((((w >> 1) & 0x01010101u) * 0x01020408u) >> 24) << (4 * j);
            if (any_big && (w & 0xFCFCFCFCu)) {
#pragma unroll
                for (int k = 2; k < NP; ++k) p[k][i] |= ((((w >> k) & 0x01010101u) * 0x01020408u) >> 24) << (4 * j);
            }
        }
    }
}

// j-th (0-based) set bit of x, j < popc(x) (nth_set_bit32_lut with the byte table addressed in the shared window)
__device__ __forceinline__ int fast_nth_set_bit(uint32_t x, int j, uint32_t nth) {
    uint32_t b = x - ((x >> 1) & 0x55555555u);
    b = (b & 0x33333333u) + ((b >> 2) & 0x33333333u);
    b = (b + (b >> 4)) & 0x0F0F0F0Fu;
    const uint32_t pre = b * 0x01010100u;  // byte k = number of set bits below byte k
    const uint32_t f = ((((uint32_t)j * 0x01010101u) | 0x80808080u) - pre) & 0x80808080u;  // byte k: j >= pre_k
    const int sh = 8 * (__popc(f) - 1);
    const uint32_t e = lds32(nth + 4u * ((x >> sh) & 0xFFu));
    const int jj = j - (int)((pre >> sh) & 0xFFu);
    return sh + (int)((e >> (4 * jj)) & 7u);
}

// hop-mask row at shared address a (this thread's column) into NWV registers
template <int NWM>
__device__ __forceinline__ void fast_load_row(uint32_t a, uint32_t* w) {
    using SC = FastScratch<NWM>;
#pragma unroll
    for (int c = 0; c < SC::NCH; ++c) {
        if (SC::VW == 1) w[c] = lds32(a + c * SC::CH_BYTES);
        if (SC::VW == 2) lds64(a + c * SC::CH_BYTES, w[2 * c], w[2 * c + 1]);
        if (SC::VW == 4) lds128(a + c * SC::CH_BYTES, w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
    }
}

// shared address of the occupation byte of mode m (ob = this thread's column): word m >> 2, byte m & 3
template <int NWM>
__device__ __forceinline__ uint32_t fast_onr_addr(uint32_t ob, int m) {
    using SC = FastScratch<NWM>;
    // (m >> 2) STRIDE + (m & 3) = m (STRIDE / 4) - (m & 3) (STRIDE / 4 - 1)
    return ((uint32_t)m & 3u) * (1u - SC::ONR_STRIDE / 4u) + ((uint32_t)m * (SC::ONR_STRIDE / 4u) + ob);
}
// offset of the hop-mask word of bond z inside a row of this thread
template <int NWM>
__device__ __forceinline__ uint32_t fast_bond_off(int z) {
    using SC = FastScratch<NWM>;
    if (SC::NWV == 1) return 0u;
    if (SC::NCH == 1) return ((uint32_t)z >> 3) & (4u * (SC::NWV - 1));  // 4 (z >> 5)
    const uint32_t w = (uint32_t)z >> 5;
    return (w / SC::VW) * SC::CH_BYTES + 4u * (w % SC::VW);
}

// one kinetic_sample! (qmc.jl:14-44).  hop_code: 2 z + (right ? 1 : 0) of the selected hop, -1 if none.
// ad = shared-window addresses of the dynamic shared memory (fast_addr); 0 <= u01 < 1 - 2^-33;
// refresh: re-sum E after this step (every FAST_REFRESH-th step of the walk).
template <int NWM>
__device__ __forceinline__ void step_fast(FastState& s, const DevModel& dm, const FastRates& fr,
                                          const FastScratch<NWM>& sc, const FastAddr& ad, double u01, bool refresh, double& tau,
                                          double& eloc, double& gratio, int& nocc, int& hop_code) {
    using SC = FastScratch<NWM>;
    constexpr int NWV = SC::NWV;
    const uint32_t tb = ad.tb, sb = ad.sb;
    const uint32_t rows = ad.rows + SC::OFF_ROWS;  // this thread's column of the hop-mask rows
    const uint32_t ob = tb + SC::OFF_ONR;          // ... of the occupation bytes

    // ---- S and its running sums over the rate classes, from the integer hop counts ----
    const uint32_t nlo = (uint32_t)s.np, nhi = (uint32_t)(s.np >> 32);
    double c0, c1, c2, c3, c4, c5;
    {
        // the masked, unshifted fields (rs carries the 2^-position): one conversion each on the XU pipe
        const double f0 = (double)(nlo & 0x3FFu), f1 = (double)(nlo & (0x3FFu << 10)), f2 = (double)(nlo & (0x3FFu << 20));
        const double f3 = (double)(nhi & 0x3FFu), f4 = (double)(nhi & (0x3FFu << 10)), f5 = (double)(nhi & (0x3FFu << 20));
        c0 = f0 * fr.rs[0];
        c1 = fma(f1, fr.rs[1], c0);
        c2 = fma(f2, fr.rs[2], c1);
        c3 = fma(f3, fr.rs[3], c2);
        c4 = fma(f4, fr.rs[4], c3);
        c5 = fma(f5, fr.rs[5], c4);
    }
    const double reg_total = c5;

    // ---- generic bonds: cached sums (zero while no mode has >= 4 particles), re-evaluated after a hop touched one ----
    if (!s.big_valid) {
        uint32_t bw[NWV], bigm[NWM];
        fast_load_row<NWM>(rows + HM_BIG * SC::ROW_BYTES, bw);
#pragma unroll
        for (int i = 0; i < NWM; ++i) bigm[i] = bw[i];
        const BigOut o = fast_big_pass<NWM>(bigm, dm, sc, 0, 0.0);
        s.bigS = o.S;
        s.bigE = o.E;
        s.big_valid = true;
    }
    const double Stot = reg_total + s.bigS;
    const double D = dm.u * (double)s.dsum;
    tau = rcp_newton(Stot);
    eloc = D - dm.t * (s.E + s.bigE);
    gratio = -D;
    nocc = s.cnt & 0xFFFF;
    const double T = u01 * Stot;  // < Stot: u01 is at least 2^-33 below 1

    int z;
    bool right;
    if (T < reg_total) {
        // level 1: rate class g = number of running sums <= T
        int g = 0;
        double base = 0.0;
        if (!(T < c0)) { g = 1; base = c0; }
        if (!(T < c1)) { g = 2; base = c1; }
        if (!(T < c2)) { g = 3; base = c2; }
        if (!(T < c3)) { g = 4; base = c3; }
        if (!(T < c4)) { g = 5; base = c4; }
        const int ng = (int)((s.np >> (g < 3 ? 10 * g : 2 + 10 * g)) & 0x3FFull);  // fast_field_pos(g)
        int j = (int)((T - base) * lds64f(sb + SC::OFF_IRATE + 8u * (uint32_t)g));
        j = max(0, min(j, ng - 1));
        // level 2: the j-th hop of the rate class in the order (right hops by position, left hops by position)
        uint32_t W[2 * NWV];
        const uint32_t ra = rows + (uint32_t)g * SC::ROW_BYTES;
        fast_load_row<NWM>(ra + HM_R * SC::ROW_BYTES, W);
        fast_load_row<NWM>(ra + HM_L * SC::ROW_BYTES, W + NWV);
        uint32_t word = W[0];
        int jb = j, zb = 0, pre = 0, nR = 0;
#pragma unroll
        for (int i = 0; i + 1 < 2 * NWV; ++i) {
            pre += __popc(W[i]);
            if (i + 1 == NWV) nR = pre;
            if (j >= pre) {  // monotone in i: the last word whose prefix count is <= j holds the j-th hop
                word = W[i + 1];
                jb = j - pre;
                zb = 32 * ((i + 1) % NWV);
            }
        }
        right = j < nR;
        z = zb + fast_nth_set_bit(word, jb, sb + SC::OFF_NTH);
    } else {
        uint32_t bw[NWV], bigm[NWM];
        fast_load_row<NWM>(rows + HM_BIG * SC::ROW_BYTES, bw);
#pragma unroll
        for (int i = 0; i < NWM; ++i) bigm[i] = bw[i];
        const BigOut o = fast_big_pass<NWM>(bigm, dm, sc, 1, T - reg_total);
        z = o.z;
        right = o.right != 0;
        if (z < 0) {  // cannot happen: T >= reg_total implies a generic bond with a hop
            hop_code = -1;
            return;
        }
    }
    hop_code = 2 * z + (right ? 1 : 0);

    // ---- the hop: bonds z-1 = (a, x), z = (x, y), z+1 = (y, b) ----
    const int M = dm.M;
    const int zn = (z + 1 == M) ? 0 : z + 1;
    const int zm = (z == 0) ? M - 1 : z - 1;
    const int zp = (zn + 1 == M) ? 0 : zn + 1;
    const uint32_t az = fast_onr_addr<NWM>(ob, z), an = fast_onr_addr<NWM>(ob, zn);
    const int a = (int)ldsu8(fast_onr_addr<NWM>(ob, zm)), x = (int)ldsu8(az), y = (int)ldsu8(an),
              b = (int)ldsu8(fast_onr_addr<NWM>(ob, zp));
    const int dxy = right ? -1 : 1;
    const int x2 = x + dxy, y2 = y - dxy;
    stsu8(az, (uint32_t)x2);
    stsu8(an, (uint32_t)y2);
    const int a4 = min(a, 4), b4 = min(b, 4), x4 = min(x, 4), y4 = min(y, 4), x24 = min(x2, 4), y24 = min(y2, 4);
    const int ix = 2 * x4 + x24 + 1, iy = 2 * y4 + y24 + 1;  // 3 x4 + (x24 - x4 + 1)
    const uint32_t il = (uint32_t)(a4 * 15 + ix), im = (uint32_t)(ix * 15 + iy), ir = (uint32_t)(iy * 5 + b4);
    double eL, eM, eR;
    unsigned long long nL, nM, nR2;
    lds_entry_sums(sb + SC::OFF_TL + 16u * il, eL, nL);
    lds_entry_sums(sb + SC::OFF_TM + 16u * im, eM, nM);
    lds_entry_sums(sb + SC::OFF_TR + 16u * ir, eR, nR2);
    const uint32_t rl = lds32(sb + SC::OFF_RL + 4u * il), rm = lds32(sb + SC::OFF_RM + 4u * im),
                   rr = lds32(sb + SC::OFF_RR + 4u * ir);
    s.E += (eL + eM) + eR;
    s.np += (nL + nM) + nR2;
    s.cnt += (int)lds32(sb + SC::OFF_TC + 4u * im);
    s.dsum += (right ? y - x : x - y) + 1;  // n_dst - n_src + 1
    {  // hop masks: each of the three bonds leaves two rows and enters two (absent hops go to the dummy row)
        // 1 << (s & 31) as the high word of (1:0) << s (one funnel shift, the hardware wraps the shift amount)
        const uint32_t bz = __funnelshift_l(0u, 1u, (uint32_t)z), bm = __funnelshift_l(0u, 1u, (uint32_t)zm),
                       bn = __funnelshift_l(0u, 1u, (uint32_t)zn);
        const uint32_t xm = rows + fast_bond_off<NWM>(zm), xz = rows + fast_bond_off<NWM>(z), xn = rows + fast_bond_off<NWM>(zn);
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // byte q of the row words, zero-extended, times ROW_BYTES
            red_xor(xm + __byte_perm(rl, 0u, 0x4440u + q) * SC::ROW_BYTES, bm);
            red_xor(xz + __byte_perm(rm, 0u, 0x4440u + q) * SC::ROW_BYTES, bz);
            red_xor(xn + __byte_perm(rr, 0u, 0x4440u + q) * SC::ROW_BYTES, bn);
        }
        // a mode with >= 4 particles in the window (before or after): the generic-bond sums and mask change
        if (max(max(a4, x4), max(max(y4, b4), max(x24, y24))) == 4) {
            s.big_valid = false;
            const uint32_t big = HM_BIG * SC::ROW_BYTES;
            if ((a4 == 4 || x4 == 4) != (a4 == 4 || x24 == 4)) red_xor(xm + big, bm);
            if ((x4 == 4 || y4 == 4) != (x24 == 4 || y24 == 4)) red_xor(xz + big, bz);
            if ((y4 == 4 || b4 == 4) != (y24 == 4 || b4 == 4)) red_xor(xn + big, bn);
        }
    }
    // rounding drift of the carried E: re-sum from the occupations periodically (refresh: the caller's step counter,
    // uniform over the CTA), and as soon as E has fallen 16-fold below its largest value since the last re-summation
    // (the rounding errors it carries scale with that peak)
    const int ehi = __double2hiint(s.E);
    s.epk = max(s.epk, ehi);
    if (refresh || ehi + (4 << 20) < s.epk) {
        s.E = fast_resum_E<NWM>(tb, dm.M);
        s.epk = __double2hiint(s.E);
    }
}

}  // namespace gwz
