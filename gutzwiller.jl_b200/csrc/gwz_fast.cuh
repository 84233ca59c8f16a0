// gwz_fast.cuh -- production kinetic_sample! (src/kinetic-vqmc/qmc.jl:14-44) with everything a hop does not change
// carried from step to step as exact integers.
//
// A hop across bond z changes the occupations of modes z and z+1, hence the classes (L,R) of the three bonds z-1, z,
// z+1 and nothing else.  Regular bonds (both occupations < 4) fall into 16 classes (l, r); their two hops have the
// rates exp(-c (r-l+1)) (right, if l > 0) and exp(-c (l-r+1)) (left, if r > 0): only SIX distinct rates r_g,
// g = Delta + 2 = 0..5.  The walker carries
//   * np   -- the number of regular hops per rate class g, six 10-bit fields of one 64-bit integer (register);
//   * E    -- sum over regular bonds of sqrt(n_src (n_dst + 1)) x rate (the off-diagonal part of E_loc), a double
//             that is updated by table deltas and re-summed from the integer state every FAST_REFRESH steps and
//             whenever it has fallen 16-fold below its peak since the last re-summation, which bounds its rounding
//             drift below 2e-13 relative for any dynamic range of the rates; dsum = sum n (n-1) / 2, the number of occupied
//             modes and the number of modes with >= 4 particles (ints);
//   * the occupations as bytes and as ONE-HOT bit masks over the modes, A_s = {z : n_z = s} and its ring rotation
//     Q_s = {z : n_(z+1) = s} for s = 0..3, plus A_4 / Q_4 for "n >= 4" (shared memory, [row][word][thread]).
// Per step: S = sum_g n_g r_g is evaluated from the integers in a fixed order (6 FMA: a pure function of the
// address).  The move is drawn in two levels.  (1) the rate class g by the running sums of S.  (2) all hops of a rate
// class are equally likely, so the j-th of them in the order (direction, bond position) is taken: the bonds with a
// right hop of rate class g are  U_s A_s & Q_(s+g-3)  (source occupation s = 1..3), those with a left hop
// U_s A_(s+g-3) & Q_s; the rows are stored with zero rows around them, so the out-of-range members of the unions read
// as empty and the row index needs no branch -- twelve shared-memory loads and six logic operations per 32 modes.
// The hop then updates np / E by three table look-ups (one per touched bond: 16-byte entries (dE, dnp) indexed by the
// capped occupations before and after), two occupation bytes and eight one-hot bits (shared-memory reductions).
// Bonds touching a mode with >= 4 particles ("generic" bonds) are rare; their sums are cached and re-evaluated from
// the occupation bytes only when a hop touches such a mode, as in gwz_incr.cuh.
//
// The partition of [0, S) into hops is ordered by (rate class, direction, bond position), generic bonds last: the same
// distribution as the reference's pick_random_from_cumsum (qmc.jl:52-61) with a different u -> hop map (tests:
// stratified u per address, recorded trajectories re-evaluated per step, energies against the exact sweep).
// Needs 4 <= M <= 250 (the four modes around a bond are distinct; counts fit their fields).
#pragma once
#include "gwz_incr.cuh"

namespace gwz {

constexpr int FAST_REFRESH = 32;
// one-hot rows per thread (each NWM words): [0,1] zero | [2..5] A_0..A_3 | [6,7] zero | [8..11] Q_0..Q_3 | [12,13] zero |
// [14] A_4 (n >= 4) | [15] Q_4.  A row index q + 2 (resp. q + 8 for Q) is valid for every q in -2..5.
constexpr int OH_ROWS = 16, OH_A = 2, OH_Q = 8, OH_A4 = 14, OH_Q4 = 15;

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

template <int NWM>
struct FastScratch {
    static constexpr int ONR_WORDS = 8 * NWM;  // 4 modes per word
    static constexpr int NTL = 5 * 15, NTM = 15 * 15, NTR = 15 * 5;
    // shared by the CTA
    FastEntry* TL;  // [a4][ix]      bond z-1: (a, x) -> (a, x')
    FastEntry* TM;  // [ix][iy]      bond z:   (x, y) -> (x', y')
    FastEntry* TR;  // [iy][b4]      bond z+1: (y, b) -> (y', b)
    double* irate;  // [8]  1 / r_g
    double* FE;     // [16] class weights for the re-summation of E, index 4 l + r
    uint32_t* nth;  // [256]
    // per thread, [entry][thread]
    uint32_t* oh;   // [OH_ROWS][NWM][STEP_BLOCK]
    uint32_t* onr;  // [ONR_WORDS][STEP_BLOCK]
    // generic-bond path
    double* exptab;  // [2N+2]
    double* sqrtab;  // [N+2]
    // byte offsets of the fixed-size regions from the start of the dynamic shared memory (step_fast addresses them
    // as 32-bit shared-window addresses: base + compile-time constant)
    static constexpr uint32_t OFF_TL = 0, OFF_TM = OFF_TL + 16 * NTL, OFF_TR = OFF_TM + 16 * NTM,
                              OFF_IRATE = OFF_TR + 16 * NTR, OFF_FE = OFF_IRATE + 8 * 8, OFF_NTH = OFF_FE + 8 * 16,
                              OFF_OH = OFF_NTH + 4 * 256, OFF_ONR = OFF_OH + 4 * OH_ROWS * NWM * STEP_BLOCK,
                              OFF_EXP = OFF_ONR + 4 * ONR_WORDS * STEP_BLOCK;
    static constexpr uint32_t ROW_BYTES = 4 * NWM * STEP_BLOCK;  // one one-hot row of all threads
    __host__ __device__ static constexpr size_t bytes(int N) { return OFF_EXP + sizeof(double) * (size_t)(3 * N + 4); }
    __device__ __forceinline__ FastScratch(unsigned char* base, int N) {
        TL = reinterpret_cast<FastEntry*>(base + OFF_TL);
        TM = reinterpret_cast<FastEntry*>(base + OFF_TM);
        TR = reinterpret_cast<FastEntry*>(base + OFF_TR);
        irate = reinterpret_cast<double*>(base + OFF_IRATE);
        FE = reinterpret_cast<double*>(base + OFF_FE);
        nth = reinterpret_cast<uint32_t*>(base + OFF_NTH);
        oh = reinterpret_cast<uint32_t*>(base + OFF_OH);
        onr = reinterpret_cast<uint32_t*>(base + OFF_ONR);
        exptab = reinterpret_cast<double*>(base + OFF_EXP);
        sqrtab = exptab + (2 * N + 2);
    }
    // contribution of one bond of capped class (l, r) to (E, np); zero for non-regular classes
    __device__ static __forceinline__ FastEntry contrib(const DevTables& tb, int l, int r) {
        FastEntry e;
        e.dE = 0.0;
        e.dn = 0ull;
        if (l < 4 && r < 4 && l >= 0 && r >= 0) {
            e.dE = tb.FE[4 * l + r];
            if (l >= 1) e.dn += 1ull << fast_field_pos(r - l + 3);
            if (r >= 1) e.dn += 1ull << fast_field_pos(l - r + 3);
        }
        return e;
    }
    __device__ static __forceinline__ FastEntry delta(const FastEntry& a, const FastEntry& b) {  // a - b
        FastEntry e;
        e.dE = a.dE - b.dE;
        e.dn = a.dn - b.dn;
        return e;
    }
    __device__ __forceinline__ void fill_tables(int N, const DevTables& tb) const {
        for (int i = threadIdx.x; i < 2 * N + 2; i += blockDim.x) exptab[i] = exp(-tb.c * (double)(i - N));
        for (int i = threadIdx.x; i < N + 2; i += blockDim.x) sqrtab[i] = sqrt((double)i);
        // ix = 3 x4 + (x' - x4 + 1) encodes the capped occupation before and after (x' in {x4-1, x4, x4+1})
        for (int i = threadIdx.x; i < NTL; i += blockDim.x) {
            const int a = i / 15, ix = i % 15, x = ix / 3, x2 = x + ix % 3 - 1;
            TL[i] = delta(contrib(tb, a, x2), contrib(tb, a, x));
        }
        for (int i = threadIdx.x; i < NTM; i += blockDim.x) {
            const int ix = i / 15, iy = i % 15, x = ix / 3, x2 = x + ix % 3 - 1, y = iy / 3, y2 = y + iy % 3 - 1;
            TM[i] = delta(contrib(tb, x2, y2), contrib(tb, x, y));
        }
        for (int i = threadIdx.x; i < NTR; i += blockDim.x) {
            const int iy = i / 5, b = i % 5, y = iy / 3, y2 = y + iy % 3 - 1;
            TR[i] = delta(contrib(tb, y2, b), contrib(tb, y, b));
        }
        if (threadIdx.x < 8) irate[threadIdx.x] = exp(tb.c * (double)((int)threadIdx.x - 2));
        if (threadIdx.x < 16) FE[threadIdx.x] = tb.FE[threadIdx.x];
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

struct FastState {
    unsigned long long np;  // regular hops per rate class, 10-bit fields
    double E;
    int dsum, nocc, nbig, age;
    int epk;  // high word of the largest E since the last re-summation (E >= 0: the high words order like the values)
    double bigS, bigE;  // cached sums over the generic bonds
    bool big_valid;
};

// byte address of mode m inside the thread's occupation bytes (ob = byte pointer to word 0 of this thread)
__device__ __forceinline__ int fast_off(int m) { return (m >> 2) * (STEP_BLOCK * 4) + (m & 3); }
// word offset (in uint32 units) of one-hot row `row`, word i, for this thread's column
template <int NWM>
__device__ __forceinline__ int oh_at(int row, int i) { return (row * NWM + i) * STEP_BLOCK; }
// one-hot rows of the capped occupation s (0..4): A block 2, 3, 4, 5, 14; Q block 8, 9, 10, 11, 15.  sel = s | 0x7770:
// byte s of the table into byte 0, a zero byte (byte 7 of the pair) into the others -- one PRMT each
__device__ __forceinline__ uint32_t oh_row_a(uint32_t sel) { return __byte_perm(0x05040302u, (uint32_t)OH_A4, sel); }
__device__ __forceinline__ uint32_t oh_row_q(uint32_t sel) { return __byte_perm(0x0B0A0908u, (uint32_t)OH_Q4, sel); }

// 32-bit shared-window addressing for the step loop: one base register, every region at a compile-time offset
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
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
__device__ __forceinline__ FastEntry lds_entry(uint32_t a) {
    FastEntry e;
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(*reinterpret_cast<unsigned long long*>(&e.dE)), "=l"(e.dn) : "r"(a));
    return e;
}
__device__ __forceinline__ void red_xor(uint32_t a, uint32_t v) { asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// sums / selection over the generic bonds (a neighbour with >= 4 particles), from the occupation bytes;
// same order and arithmetic as PlanesBig::pass
template <int NWM>
__device__ __forceinline__ BigOut fast_big_pass(const uint32_t (&bigm)[NWM], const DevModel& dm, const unsigned char* ob,
                                                const double* exptab, const double* sqrtab, int mode, double t) {
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
            const int L = ob[fast_off(z)], R = ob[fast_off(zn)];
            const double rr = (L > 0) ? exptab[(R - L + 1) + N] : 0.0;
            const double rl = (R > 0) ? exptab[(L - R + 1) + N] : 0.0;
            if (mode == 0) {
                const double er = sqrtab[L] * sqrtab[R + 1] * rr;
                const double el = sqrtab[R] * sqrtab[L + 1] * rl;
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

// E re-summed from the one-hot rows (class counts by popcount), in the order of bond_sums (class id 4 l + r ascending)
template <int NWM>
__device__ __forceinline__ double fast_resum_E(const FastScratch<NWM>& sc) {
    const uint32_t* oh = sc.oh + threadIdx.x;
    uint32_t A[4][NWM], Q[4][NWM];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < NWM; ++i) {
            A[s][i] = oh[oh_at<NWM>(OH_A + s, i)];
            Q[s][i] = oh[oh_at<NWM>(OH_Q + s, i)];
        }
    double E = 0.0;
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (l == 0 && r == 0) continue;
            int c = 0;
#pragma unroll
            for (int i = 0; i < NWM; ++i) c += __popc(A[l][i] & Q[r][i]);
            E = fma(int_to_double_fast(c), sc.FE[4 * l + r], E);
        }
    return E;
}

// (re)build the carried state of a walker from its planes
template <int NWM, int NP>
__device__ __forceinline__ void fast_init(const uint32_t (&p)[NP][NWM], const DevModel& dm, const FastScratch<NWM>& sc,
                                          FastState& s) {
    const int tid = threadIdx.x;
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
            sc.onr[(8 * i + j) * STEP_BLOCK + tid] = word;
        }
    // one-hot rows and their ring rotations
    uint32_t* oh = sc.oh + tid;
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
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
        oh[oh_at<NWM>(0, i)] = oh[oh_at<NWM>(1, i)] = oh[oh_at<NWM>(6, i)] = oh[oh_at<NWM>(7, i)] = 0u;
        oh[oh_at<NWM>(12, i)] = oh[oh_at<NWM>(13, i)] = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            oh[oh_at<NWM>(OH_A + q, i)] = A[q][i];
            oh[oh_at<NWM>(OH_Q + q, i)] = Q[q][i];
        }
        oh[oh_at<NWM>(OH_A4, i)] = A[4][i];
        oh[oh_at<NWM>(OH_Q4, i)] = Q[4][i];
    }
    // hop counts per rate class, sum n (n - 1) / 2, occupied modes, modes with >= 4 particles
    unsigned long long np = 0ull;
    int n1 = 0, n2 = 0, n3 = 0, dbig = 0, nbig = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int c = 0;
#pragma unroll
            for (int i = 0; i < NWM; ++i) c += __popc(A[l][i] & Q[r][i]);
            if (l >= 1) np += (unsigned long long)c << fast_field_pos(r - l + 3);
            if (r >= 1) np += (unsigned long long)c << fast_field_pos(l - r + 3);
        }
    const unsigned char* ob = reinterpret_cast<const unsigned char*>(sc.onr + tid);
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
        n1 += __popc(A[1][i]);
        n2 += __popc(A[2][i]);
        n3 += __popc(A[3][i]);
        uint32_t m = BIG[i];
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int n = ob[fast_off(32 * i + b)];
            dbig += n * (n - 1) / 2;
            ++nbig;
        }
    }
    s.np = np;
    s.dsum = n2 + 3 * n3 + dbig;
    s.nocc = n1 + n2 + n3 + nbig;
    s.nbig = nbig;
    s.E = fast_resum_E<NWM>(sc);
    s.epk = __double2hiint(s.E);
    s.age = 0;
    s.bigS = s.bigE = 0.0;
    s.big_valid = false;
}

// the planes of the walker from its one-hot rows (n < 4) and occupation bytes (n >= 4)
template <int NWM, int NP>
__device__ __forceinline__ void fast_store_planes(const FastScratch<NWM>& sc, uint32_t (&p)[NP][NWM]) {
    const uint32_t* oh = sc.oh + threadIdx.x;
    const unsigned char* ob = reinterpret_cast<const unsigned char*>(sc.onr + threadIdx.x);
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
        const uint32_t a1 = oh[oh_at<NWM>(OH_A + 1, i)], a2 = oh[oh_at<NWM>(OH_A + 2, i)], a3 = oh[oh_at<NWM>(OH_A + 3, i)];
        p[0][i] = a1 | a3;
        p[1][i] = a2 | a3;
#pragma unroll
        for (int k = 2; k < NP; ++k) p[k][i] = 0u;
        uint32_t m = oh[oh_at<NWM>(OH_A4, i)];
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t n = ob[fast_off(32 * i + b)];
#pragma unroll
            for (int k = 0; k < NP; ++k) p[k][i] |= ((n >> k) & 1u) << b;
        }
    }
}

// one kinetic_sample! (qmc.jl:14-44).  hop_code: 2 z + (right ? 1 : 0) of the selected hop, -1 if none.
// sb = shared-window address of the start of the dynamic shared memory (smem_addr(smem_raw)).
template <int NWM>
__device__ __forceinline__ void step_fast(FastState& s, const DevModel& dm, const FastRates& fr,
                                          const FastScratch<NWM>& sc, uint32_t sb, double u01, double& tau, double& eloc,
                                          double& gratio, int& nhops, int& hop_code) {
    using SC = FastScratch<NWM>;
    const uint32_t st = sb + 4u * threadIdx.x;  // this thread's column in the [entry][thread] regions
    const uint32_t oh = st + SC::OFF_OH;        // one-hot row r, word i: oh + (r NWM + i) 512
    const uint32_t ob = st + SC::OFF_ONR;       // occupation byte of mode m: ob + (m >> 2) 512 + (m & 3)

    // ---- S and its running sums over the rate classes, from the integer hop counts ----
    const uint32_t nlo = (uint32_t)s.np, nhi = (uint32_t)(s.np >> 32);
    double c0, c1, c2, c3, c4, c5;
    {
        const double magic = 4503599627370496.0;  // 2^52: (double)x for 0 <= x < 2^32 is hiloint2double(0x43300000, x) - 2^52
        const double f0 = __hiloint2double(0x43300000, (int)(nlo & 0x3FFu)) - magic;
        const double f1 = __hiloint2double(0x43300000, (int)(nlo & (0x3FFu << 10))) - magic;
        const double f2 = __hiloint2double(0x43300000, (int)(nlo & (0x3FFu << 20))) - magic;
        const double f3 = __hiloint2double(0x43300000, (int)(nhi & 0x3FFu)) - magic;
        const double f4 = __hiloint2double(0x43300000, (int)(nhi & (0x3FFu << 10))) - magic;
        const double f5 = __hiloint2double(0x43300000, (int)(nhi & (0x3FFu << 20))) - magic;
        c0 = f0 * fr.rs[0];
        c1 = fma(f1, fr.rs[1], c0);
        c2 = fma(f2, fr.rs[2], c1);
        c3 = fma(f3, fr.rs[3], c2);
        c4 = fma(f4, fr.rs[4], c3);
        c5 = fma(f5, fr.rs[5], c4);
    }
    const double reg_total = c5;

    // ---- generic bonds: cached sums ----
    uint32_t bigm[NWM];
#pragma unroll
    for (int i = 0; i < NWM; ++i) bigm[i] = 0u;
    double bigS = 0.0, bigE = 0.0;
    const unsigned char* const obp = reinterpret_cast<const unsigned char*>(sc.onr + threadIdx.x);  // generic-bond path only
    if (s.nbig) {
#pragma unroll
        for (int i = 0; i < NWM; ++i)
            bigm[i] = (lds32(oh + (OH_A4 * NWM + i) * 512u) | lds32(oh + (OH_Q4 * NWM + i) * 512u)) & dm.mmask[i];
        if (!s.big_valid) {
            const BigOut o = fast_big_pass<NWM>(bigm, dm, obp, sc.exptab, sc.sqrtab, 0, 0.0);
            s.bigS = o.S;
            s.bigE = o.E;
            s.big_valid = true;
        }
        bigS = s.bigS;
        bigE = s.bigE;
    }
    const double Stot = reg_total + bigS;
    const double D = dm.u * (double)s.dsum;
    tau = 1.0 / Stot;
    eloc = D - dm.t * (s.E + bigE);
    gratio = -D;
    nhops = 2 * s.nocc;
    double T = u01 * Stot;
    if (!(T < Stot)) T = __longlong_as_double(__double_as_longlong(Stot) - 1);

    int z;
    bool right;
    if (T < reg_total || !(bigS > 0.0)) {
        if (!(T < reg_total)) T = __longlong_as_double(__double_as_longlong(reg_total) - 1);
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
        //   right hop from a mode with s particles (s = 1..3) to a neighbour with s + g - 3: A_s & Q_(s+g-3)
        //   left hop from a neighbour with s particles onto a mode with s + g - 3:         A_(s+g-3) & Q_s
        const uint32_t dyn = oh + (uint32_t)(g - 2) * SC::ROW_BYTES;  // + OH_Q rows: Q_(1+g-3); + OH_A rows: A_(1+g-3)
        uint32_t mR[NWM], mL[NWM];
        int nR = 0;
#pragma unroll
        for (int i = 0; i < NWM; ++i) {
            const uint32_t a1 = lds32(oh + ((OH_A + 1) * NWM + i) * 512u), a2 = lds32(oh + ((OH_A + 2) * NWM + i) * 512u),
                           a3 = lds32(oh + ((OH_A + 3) * NWM + i) * 512u);
            const uint32_t q1 = lds32(oh + ((OH_Q + 1) * NWM + i) * 512u), q2 = lds32(oh + ((OH_Q + 2) * NWM + i) * 512u),
                           q3 = lds32(oh + ((OH_Q + 3) * NWM + i) * 512u);
            const uint32_t dq0 = lds32(dyn + ((OH_Q + 0) * NWM + i) * 512u), dq1 = lds32(dyn + ((OH_Q + 1) * NWM + i) * 512u),
                           dq2 = lds32(dyn + ((OH_Q + 2) * NWM + i) * 512u);
            const uint32_t da0 = lds32(dyn + ((OH_A + 0) * NWM + i) * 512u), da1 = lds32(dyn + ((OH_A + 1) * NWM + i) * 512u),
                           da2 = lds32(dyn + ((OH_A + 2) * NWM + i) * 512u);
            mR[i] = (a1 & dq0) | (a2 & dq1) | (a3 & dq2);
            mL[i] = (q1 & da0) | (q2 & da1) | (q3 & da2);
            nR += __popc(mR[i]);
        }
        right = j < nR;
        int jb = right ? j : j - nR;
        int pos = 0;
        uint32_t word = right ? mR[NWM - 1] : mL[NWM - 1];
        bool fnd = false;
#pragma unroll
        for (int i = 0; i < NWM; ++i) {
            const uint32_t m = right ? mR[i] : mL[i];
            const int pc = __popc(m);
            const bool here = !fnd && (jb < pc);
            word = here ? m : word;
            pos = here ? 32 * i : pos;
            fnd = fnd || here;
            jb -= fnd ? 0 : pc;
        }
        if (!fnd) jb = max(0, __popc(word) - 1);  // cannot happen while np matches the masks
        z = word ? pos + nth_set_bit32_lut(word, jb, sc.nth) : -1;
    } else {
        const BigOut o = fast_big_pass<NWM>(bigm, dm, obp, sc.exptab, sc.sqrtab, 1, T - reg_total);
        z = o.z;
        right = o.right != 0;
    }
    hop_code = z < 0 ? -1 : 2 * z + (right ? 1 : 0);
    if (z < 0) return;

    // ---- the hop: bonds z-1 = (a, x), z = (x, y), z+1 = (y, b) ----
    const int M = dm.M;
    const int zn = (z + 1 == M) ? 0 : z + 1;
    const int zm = (z == 0) ? M - 1 : z - 1;
    const int zp = (zn + 1 == M) ? 0 : zn + 1;
    const uint32_t az = ob + (uint32_t)fast_off(z), an = ob + (uint32_t)fast_off(zn);
    const int a = (int)ldsu8(ob + (uint32_t)fast_off(zm)), x = (int)ldsu8(az), y = (int)ldsu8(an),
              b = (int)ldsu8(ob + (uint32_t)fast_off(zp));
    const int x2 = right ? x - 1 : x + 1, y2 = right ? y + 1 : y - 1;
    stsu8(az, (uint32_t)x2);
    stsu8(an, (uint32_t)y2);
    const int a4 = min(a, 4), b4 = min(b, 4), x4 = min(x, 4), y4 = min(y, 4), x24 = min(x2, 4), y24 = min(y2, 4);
    const int ix = 2 * x4 + x24 + 1, iy = 2 * y4 + y24 + 1;  // 3 x4 + (x24 - x4 + 1)
    const FastEntry eL = lds_entry(sb + SC::OFF_TL + 16u * (uint32_t)(a4 * 15 + ix)),
                    eM = lds_entry(sb + SC::OFF_TM + 16u * (uint32_t)(ix * 15 + iy)),
                    eR = lds_entry(sb + SC::OFF_TR + 16u * (uint32_t)(iy * 5 + b4));
    s.E += (eL.dE + eM.dE) + eR.dE;
    s.np += (eL.dn + eM.dn) + eR.dn;
    const int nsrc = right ? x : y, ndst = right ? y : x;
    s.dsum += ndst - nsrc + 1;
    s.nocc += (int)(ndst == 0) - (int)(nsrc == 1);
    s.nbig += (int)(x24 == 4) - (int)(x4 == 4) + (int)(y24 == 4) - (int)(y4 == 4);
    // a mode with >= 4 particles in the window (before or after): the generic-bond sums change
    if (max(max(a4, x4), max(max(y4, b4), max(x24, y24))) == 4) s.big_valid = false;
    {  // one-hot rows: mode z leaves row x4 and enters x24 (A at bit z, Q at bit z-1), mode z+1 likewise (A: z+1, Q: z)
        const uint32_t bz = 1u << (z & 31), bn = 1u << (zn & 31), bm = 1u << (zm & 31);
        const uint32_t wz = oh + (uint32_t)(z >> 5) * 512u, wn = oh + (uint32_t)(zn >> 5) * 512u,
                       wm = oh + (uint32_t)(zm >> 5) * 512u;
        const uint32_t sx = (uint32_t)x4 | 0x7770u, sx2 = (uint32_t)x24 | 0x7770u, sy = (uint32_t)y4 | 0x7770u,
                       sy2 = (uint32_t)y24 | 0x7770u;
        red_xor(wz + oh_row_a(sx) * SC::ROW_BYTES, bz);
        red_xor(wz + oh_row_a(sx2) * SC::ROW_BYTES, bz);
        red_xor(wm + oh_row_q(sx) * SC::ROW_BYTES, bm);
        red_xor(wm + oh_row_q(sx2) * SC::ROW_BYTES, bm);
        red_xor(wn + oh_row_a(sy) * SC::ROW_BYTES, bn);
        red_xor(wn + oh_row_a(sy2) * SC::ROW_BYTES, bn);
        red_xor(wz + oh_row_q(sy) * SC::ROW_BYTES, bz);
        red_xor(wz + oh_row_q(sy2) * SC::ROW_BYTES, bz);
    }
    // rounding drift of the carried E: re-sum from the integers periodically, and as soon as E has fallen 16-fold
    // below its largest value since the last re-summation (the rounding errors it carries scale with that peak)
    const int ehi = __double2hiint(s.E);
    s.epk = max(s.epk, ehi);
    if (++s.age >= FAST_REFRESH || ehi + (4 << 20) < s.epk) {
        s.E = fast_resum_E<NWM>(sc);
        s.epk = __double2hiint(s.E);
        s.age = 0;
    }
}

}  // namespace gwz
