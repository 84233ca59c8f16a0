// gwz_incr.cuh -- incremental bond-class bookkeeping on top of the occupation bit-planes.
//
// A hop across bond z changes the occupations of modes z and z+1 only, hence the classes (L,R) of
// the three bonds z-1, z, z+1 and nothing else.  The walker therefore carries, next to its planes,
//   * the 15 (+1) class counts c(l,r), l,r in 0..3, packed as bytes (byte r of row word l), and
//   * its occupation numbers as bytes (4 modes per 32-bit word)
// in shared memory ([entry][thread] layout: every lane owns a bank, any dynamic index is conflict
// free).  A step costs O(1) bookkeeping instead of the O(classes x words) pair-popcounts of
// step_planes: six class-count updates, two occupation bytes, the plane ripple.  Everything that
// enters the estimators stays a pure function of the address and is evaluated in exactly the same
// floating-point order as bond_sums / select_move (S and E are re-summed from the integer counts
// each step; only integers are carried from step to step), so this path is bitwise identical to
// step_planes and to the unary kernel (tests/test_gpu_walkers.py::test_planes_kernel_is_bitwise_identical_to_unary runs
// the three of them with GWZ_FAST=0).  Round 2's production kernel is gwz_fast.cuh; this one stays as its
// independently validated counterpart (GWZ_FAST=0).  Reference: kinetic_sample!, src/kinetic-vqmc/qmc.jl:14-44.
#pragma once
#include "gwz_planes.cuh"

namespace gwz {

constexpr int INCR_ROWS = 5;  // rows 0..3 + a sink row for bonds whose lower mode holds >= 4 particles

template <int NWM>
struct IncrScratch {
    static constexpr int ONR_WORDS = 8 * NWM;  // 4 modes per word
    double* exptab;  // [2N+2] hop rate for Delta = i-N
    double* sqrtab;  // [N+2]
    double* FS;      // [16] per-class tables for the dynamically indexed part of the move selection
    double* RR;      // [16]
    double* IRR;     // [16]
    double* IRL;     // [16]
    uint32_t* nth;   // [256] nibble j of nth[b] = position of the j-th set bit of the byte b
    uint32_t* cnt;   // [INCR_ROWS][STEP_BLOCK]
    uint32_t* onr;   // [ONR_WORDS][STEP_BLOCK]
    __host__ __device__ static constexpr size_t bytes(int N) {
        return sizeof(double) * (size_t)(3 * N + 4) + sizeof(double) * 64 + sizeof(uint32_t) * 256 +
               sizeof(uint32_t) * (size_t)(INCR_ROWS + ONR_WORDS) * STEP_BLOCK;
    }
    // fixed-size regions first: their shared-memory offsets are compile-time constants (immediates of LDS / STS); the
    // N-dependent rate tables (only used by the rare generic-bond path) come last
    __device__ __forceinline__ IncrScratch(unsigned char* base, int N) {
        FS = reinterpret_cast<double*>(base);
        RR = FS + 16;
        IRR = RR + 16;
        IRL = IRR + 16;
        nth = reinterpret_cast<uint32_t*>(IRL + 16);
        cnt = nth + 256;
        onr = cnt + INCR_ROWS * STEP_BLOCK;
        exptab = reinterpret_cast<double*>(onr + ONR_WORDS * STEP_BLOCK);
        sqrtab = exptab + (2 * N + 2);
    }
    __device__ __forceinline__ void fill_tables(int N, const DevTables& tb) const {
        for (int i = threadIdx.x; i < 2 * N + 2; i += blockDim.x) exptab[i] = exp(-tb.c * (double)(i - N));
        for (int i = threadIdx.x; i < N + 2; i += blockDim.x) sqrtab[i] = sqrt((double)i);
        if (threadIdx.x < 16) {
            FS[threadIdx.x] = tb.FS[threadIdx.x];
            RR[threadIdx.x] = tb.RR[threadIdx.x];
            IRR[threadIdx.x] = tb.IRR[threadIdx.x];
            IRL[threadIdx.x] = tb.IRL[threadIdx.x];
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

// 1 << (8*r) for r in 0..3, 0 for r == 4 (PTX shl clamps the shift amount)
__device__ __forceinline__ uint32_t count_unit(int r4) {
    uint32_t d;
    const uint32_t one = 1u;
    asm("shl.b32 %0, %1, %2;" : "=r"(d) : "r"(one), "r"(8 * r4));
    return d;
}
__device__ __forceinline__ int byte_of(uint32_t w, int r) { return (int)((w >> (8 * r)) & 0xFFu); }
// (double) of byte r of w: one I2F.F64.U8 with a byte selector on the XU pipe (r static), instead of
// shift + mask + magic-number add on the ALU / FP64 pipes that bound this kernel
__device__ __forceinline__ double byte_to_double(uint32_t w, int r) {
    double d;
    asm("cvt.rn.f64.u8 %0, %1;" : "=d"(d) : "r"(w >> (8 * r)));
    return d;
}

// j-th (0-based) set bit of x, j < popc(x): byte-parallel popcounts, their prefix by one multiply, a
// borrow-free per-byte compare, then a 256-entry table for the position inside the byte
__device__ __forceinline__ int nth_set_bit32_lut(uint32_t x, int j, const uint32_t* nth) {
    uint32_t b = x - ((x >> 1) & 0x55555555u);
    b = (b & 0x33333333u) + ((b >> 2) & 0x33333333u);
    b = (b + (b >> 4)) & 0x0F0F0F0Fu;
    const uint32_t pre = b * 0x01010100u;  // byte k = number of set bits below byte k
    const uint32_t f = ((((uint32_t)j * 0x01010101u) | 0x80808080u) - pre) & 0x80808080u;  // byte k: j >= pre_k
    const int sh = 8 * (__popc(f) - 1);
    const uint32_t e = nth[(x >> sh) & 0xFFu];
    const int jj = j - (int)((pre >> sh) & 0xFFu);
    return sh + (int)((e >> (4 * jj)) & 7u);
}

// (re)build the shared-memory side of a walker from its planes: occupation bytes and packed class counts
template <int NWM, int NP>
__device__ __forceinline__ void incr_init(const uint32_t (&p)[NP][NWM], const DevModel& dm, const IncrScratch<NWM>& sc) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int i = 0; i < NWM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t word = 0u;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const uint32_t nib = (p[k][i] >> (4 * j)) & 0xFu;
                word += ((nib * 0x00204081u) & 0x01010101u) << k;  // bit t of nib -> byte t
            }
            sc.onr[(8 * i + j) * STEP_BLOCK + tid] = word;
        }
    BondMasks<NWM, 4> bm;
    classify_planes<NWM, NP>(p, dm, bm);
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        uint32_t row = 0u;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int c = 0;
#pragma unroll
            for (int i = 0; i < NWM; ++i) c += __popc(bm.EL[l][i] & bm.ER[r][i]);
            row |= (uint32_t)c << (8 * r);
        }
        sc.cnt[l * STEP_BLOCK + tid] = row;
    }
    sc.cnt[4 * STEP_BLOCK + tid] = 0u;
}

// Sums over the generic bonds (a neighbour with >= 4 particles), kept from step to step: they only
// change when a hop touches one of the modes around such a bond, so the (divergent, out-of-line)
// evaluation runs on a few percent of the steps instead of whenever any lane of the warp owns one.
// The cached numbers are the unmodified results of PlanesBig::pass(0), i.e. pure functions of the address.
struct BigCache {
    double S, E;
    int nocc, d;
    bool valid;
};

// modes with >= 4 particles (OR of the planes >= 2): carried with the walker, refreshed only when a hop ripples into
// those planes
template <int NWM, int NP>
__device__ __forceinline__ void big_plane(const uint32_t (&p)[NP][NWM], uint32_t (&bigP)[NWM]) {
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
        uint32_t b = p[2][i];
#pragma unroll
        for (int k = 3; k < NP; ++k) b |= p[k][i];
        bigP[i] = b;
    }
}

// one kinetic_sample! (qmc.jl:14-44) with incremental class counts
template <int NWM, int NP>
__device__ __forceinline__ void step_incr(uint32_t (&p)[NP][NWM], const DevModel& dm, const DevTables& tb,
                                          const IncrScratch<NWM>& sc, BigCache& bc, uint32_t (&bigP)[NWM], double u01,
                                          double& tau, double& eloc, double& gratio, int& nhops) {
    const int tid = threadIdx.x;
    uint32_t* const cntp = sc.cnt + tid;
    // ---- S, E, D, occupied modes from the class counts (same order as bond_sums) ----
    uint32_t roww[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) roww[l] = cntp[l * STEP_BLOCK];
    double S = 0.0, E = 0.0;
    double rowcum[4];
    int nocc = 0, dd = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int k = 4 * l + r;
            if (k == 0) continue;
            const double cd = byte_to_double(roww[l], r);
            S = fma(cd, tb.FS[k], S);
            E = fma(cd, tb.FE[k], E);
        }
        rowcum[l] = S;
        const int rowcnt = (int)((roww[l] * 0x01010101u) >> 24);
        if (l > 0) nocc += rowcnt;
        dd += rowcnt * (l * (l - 1) / 2);
    }
    const double reg_total = S;
    // ---- generic bonds (a neighbour with >= 4 particles): rare, evaluated from the planes ----
    uint32_t bigm[NWM];
    uint32_t anyb = 0u;
#pragma unroll
    for (int i = 0; i < NWM; ++i) anyb |= bigP[i];
#pragma unroll
    for (int i = 0; i < NWM; ++i) bigm[i] = 0u;
    long long d = dd;
    double bigS = 0.0;
    const PlanesBig<NWM, NP> big{p, bigm, dm, nullptr, sc.exptab, sc.sqrtab};
    if (anyb) {
        uint32_t bigQ[NWM];
        ring_next<NWM>(bigP, dm, bigQ);
#pragma unroll
        for (int i = 0; i < NWM; ++i) bigm[i] = (bigP[i] | bigQ[i]) & dm.mmask[i];
        if (!bc.valid) {
            const BigOut o = big.pass(0, 0.0);
            bc.S = o.S;
            bc.E = o.E;
            bc.nocc = o.nocc;
            bc.d = (int)o.d;
            bc.valid = true;
        }
        bigS = bc.S;
        E += bc.E;
        nocc += bc.nocc;
        d += bc.d;
    }
    const double Stot = reg_total + bigS;
    const double D = dm.u * (double)d;
    tau = 1.0 / Stot;
    eloc = D - dm.t * E;
    gratio = -D;
    nhops = 2 * nocc;
    double T = u01 * Stot;
    if (!(T < Stot)) T = __longlong_as_double(__double_as_longlong(Stot) - 1);

    // ---- class-ordered move selection (same partition of [0,S) as select_move) ----
    int z;
    bool right;
    if (T < reg_total || !(bigS > 0.0)) {
        if (!(T < reg_total)) T = __longlong_as_double(__double_as_longlong(reg_total) - 1);
        // row: first l with T < rowcum[l]
        const bool g0 = !(T < rowcum[0]), g1 = !(T < rowcum[1]), g2 = !(T < rowcum[2]);
        const int l = (int)g0 + (int)g1 + (int)g2;
        const double q = g2 ? rowcum[2] : (g1 ? rowcum[1] : (g0 ? rowcum[0] : 0.0));
        const uint32_t rw = g2 ? roww[3] : (g1 ? roww[2] : (g0 ? roww[1] : roww[0]));
        // class inside the row: first r with T < running sum (the sums are non-decreasing)
        const double* fs = sc.FS + 4 * l;
        const double q0 = fma(byte_to_double(rw, 0), fs[0], q);
        const double q1 = fma(byte_to_double(rw, 1), fs[1], q0);
        const double q2 = fma(byte_to_double(rw, 2), fs[2], q1);
        const bool h0 = !(T < q0), h1 = !(T < q1), h2 = !(T < q2);
        const int r = (int)h0 + (int)h1 + (int)h2;
        const double base = h2 ? q2 : (h1 ? q1 : (h0 ? q0 : q));
        const int c = (int)((rw >> (8 * r)) & 0xFFu);
        const int idx = 4 * l + r;
        const double cd = int_to_double_fast(c);
        const double t = T - base;
        const double wR = cd * sc.RR[idx];
        right = (t < wR) || (r == 0);
        int j = right ? (int)(t * sc.IRR[idx]) : (int)((t - wR) * sc.IRL[idx]);
        j = max(0, min(j, c - 1));
        // mask of class (l,r): modes with n == l whose right neighbour has n == r
        const uint32_t l0 = (l & 1) ? 0u : ~0u, l1 = (l & 2) ? 0u : ~0u;
        const uint32_t r0 = (r & 1) ? 0u : ~0u, r1 = (r & 2) ? 0u : ~0u;
        uint32_t A[NWM], Bm[NWM], Q[NWM];
#pragma unroll
        for (int i = 0; i < NWM; ++i) {
            const uint32_t ok = ~bigP[i] & dm.mmask[i];
            A[i] = (p[0][i] ^ l0) & (p[1][i] ^ l1) & ok;
            Bm[i] = (p[0][i] ^ r0) & (p[1][i] ^ r1) & ok;
        }
        ring_next<NWM>(Bm, dm, Q);
        int pos = 0;
        uint32_t word = A[NWM - 1] & Q[NWM - 1];
        bool fnd = false;
#pragma unroll
        for (int i = 0; i < NWM; ++i) {
            const uint32_t m = A[i] & Q[i];
            const int pc = __popc(m);
            const bool here = !fnd && (j < pc);
            word = here ? m : word;
            pos = here ? 32 * i : pos;
            fnd = fnd || here;
            j -= fnd ? 0 : pc;
        }
        z = pos + nth_set_bit32_lut(word, j, sc.nth);
    } else {
        const BigOut o = big.pass(1, T - reg_total);
        z = o.z;
        right = o.right != 0;
    }
    if (z < 0) return;

    // ---- the hop: occupation bytes, class counts of the bonds z-1, z, z+1, planes ----
    const int M = dm.M;
    const int zn = (z + 1 == M) ? 0 : z + 1;
    const int zm = (z == 0) ? M - 1 : z - 1;
    const int zp = (zn + 1 == M) ? 0 : zn + 1;
    unsigned char* const ob = reinterpret_cast<unsigned char*>(sc.onr + tid);
    auto off = [](int m) { return (m >> 2) * (STEP_BLOCK * 4) + (m & 3); };
    const int a = ob[off(zm)], x = ob[off(z)], y = ob[off(zn)], b = ob[off(zp)];
    const int x2 = right ? x - 1 : x + 1, y2 = right ? y + 1 : y - 1;
    ob[off(z)] = (unsigned char)x2;
    ob[off(zn)] = (unsigned char)y2;
    const int a4 = min(a, 4), x4 = min(x, 4), y4 = min(y, 4), b4 = min(b, 4), x24 = min(x2, 4), y24 = min(y2, 4);
    // a mode with >= 4 particles in the window (before or after): the generic-bond sums change
    if (max(max(a4, x4), max(max(y4, b4), max(x24, y24))) == 4) bc.valid = false;
    const uint32_t ub = count_unit(b4);
    cntp[a4 * STEP_BLOCK] += count_unit(x24) - count_unit(x4);  // bond z-1: (a,x)  -> (a,x')
    cntp[x4 * STEP_BLOCK] -= count_unit(y4);                    // bond z:   (x,y)  -> (x',y')
    cntp[x24 * STEP_BLOCK] += count_unit(y24);
    cntp[y4 * STEP_BLOCK] -= ub;                                // bond z+1: (y,b)  -> (y',b)
    cntp[y24 * STEP_BLOCK] += ub;
    if (apply_hop_planes<NWM, NP>(p, dm, z, right)) big_plane<NWM, NP>(p, bigP);
}

}  // namespace gwz
