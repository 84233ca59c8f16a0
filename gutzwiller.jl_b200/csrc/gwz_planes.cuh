// gwz_planes.cuh -- occupation bit-plane representation of a walker (production sampling path).
//
// Instead of the unary BoseFS bitstring (N+M-1 bits) the resident walker keeps NP bit-planes of M
// bits: bit i of plane k = bit k of the occupation n_i of mode i.  A bond is then simply a mode
// index i (between mode i and mode i+1 mod M), the bond classes (L,R) = (min(n_i,4), min(n_{i+1},4))
// come from two planes and a one-bit ring rotation (which also makes the periodic bond an ordinary
// one), and every multi-word operation runs over ceil(M/32) words instead of ceil((N+M+4)/32):
// half the words at unit filling.  The class order, the order of bonds inside a class and the
// floating-point summation order are identical to the unary path, so both paths produce bitwise
// identical trajectories and estimators (tests/test_gpu_walkers.py::test_planes_kernel_is_bitwise_
// identical_to_unary).  The unary string stays the exchange format at the C ABI.
#pragma once
#include "gwz_device.cuh"

namespace gwz {

// Q(i) = P(i+1 mod M): rotate right by one inside the M-bit ring
template <int NWM>
__device__ __forceinline__ void ring_next(const uint32_t (&p)[NWM], const DevModel& dm, uint32_t (&q)[NWM]) {
    shr1<NWM>(p, q);  // bits above M-1 are zero, so bit M-1 of q is free
    const uint32_t wrap = (p[0] & 1u) << dm.mtop_b;
#pragma unroll
    for (int i = 0; i < NWM; ++i) q[i] |= wrap & dm.mtopmask[i];
}

// bond classes from the planes: EL[l] = modes with n_i == l, ER[r] = modes whose right neighbour has
// n == r (l, r in 0..3); big = bonds touching an occupation >= 4
template <int NWM, int NP>
__device__ __forceinline__ void classify_planes(const uint32_t (&p)[NP][NWM], const DevModel& dm,
                                                BondMasks<NWM, 4>& bm) {
    static_assert(NP >= 3, "need at least three planes");
    uint32_t bigP[NWM], q0[NWM], q1[NWM], bigQ[NWM];
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
        uint32_t b = p[2][i];
#pragma unroll
        for (int k = 3; k < NP; ++k) b |= p[k][i];
        bigP[i] = b;
    }
    ring_next<NWM>(p[0], dm, q0);
    ring_next<NWM>(p[1], dm, q1);
    ring_next<NWM>(bigP, dm, bigQ);
#pragma unroll
    for (int i = 0; i < NWM; ++i) {
        const uint32_t okP = ~bigP[i] & dm.mmask[i], okQ = ~bigQ[i] & dm.mmask[i];
        bm.EL[0][i] = ~p[1][i] & ~p[0][i] & okP;
        bm.EL[1][i] = ~p[1][i] & p[0][i] & okP;
        bm.EL[2][i] = p[1][i] & ~p[0][i] & okP;
        bm.EL[3][i] = p[1][i] & p[0][i] & okP;
        bm.ER[0][i] = ~q1[i] & ~q0[i] & okQ;
        bm.ER[1][i] = ~q1[i] & q0[i] & okQ;
        bm.ER[2][i] = q1[i] & ~q0[i] & okQ;
        bm.ER[3][i] = q1[i] & q0[i] & okQ;
        bm.big[i] = (bigP[i] | bigQ[i]) & dm.mmask[i];
    }
}

template <int NWM, int NP>
struct PlanesBig {
    const uint32_t (&p)[NP][NWM];
    const uint32_t (&big)[NWM];
    const DevModel& dm;
    uint32_t* stage;  // shared memory, already offset by threadIdx.x: [(NP+1)*NWM][STEP_BLOCK]
    const double* exptab;
    const double* sqrtab;
    __device__ __forceinline__ bool any() const {
        uint32_t a = 0u;
#pragma unroll
        for (int i = 0; i < NWM; ++i) a |= big[i];
        return a != 0u;
    }
    // Register-only evaluation: the word index is static (unrolled), the bit index comes from ffs; the
    // right neighbour's occupation is read from the ring-rotated planes at the same bit.
    __device__ __forceinline__ BigOut pass(int mode, double t) const {
        uint32_t q[NP][NWM];
#pragma unroll
        for (int k = 0; k < NP; ++k) ring_next<NWM>(p[k], dm, q[k]);
        BigOut o;
        o.S = o.E = o.ED = 0.0;
        o.d = 0;
        o.nocc = 0;
        o.z = -1;
        o.right = 1;
        double acc = 0.0;
        bool done = false;
        int dsum = 0;
        const int N = dm.N;
#pragma unroll
        for (int i = 0; i < NWM; ++i) {
            uint32_t m = big[i];
            while (m && !done) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                int L = 0, R = 0;
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    L |= (int)((p[k][i] >> b) & 1u) << k;
                    R |= (int)((q[k][i] >> b) & 1u) << k;
                }
                const double rr = (L > 0) ? exptab[(R - L + 1) + N] : 0.0;
                const double rl = (R > 0) ? exptab[(L - R + 1) + N] : 0.0;
                if (mode == 0) {
                    const double er = sqrtab[L] * sqrtab[R + 1] * rr;
                    const double el = sqrtab[R] * sqrtab[L + 1] * rl;
                    o.S += rr;
                    o.S += rl;
                    o.E += er + el;
                    o.ED += er * (double)(R - L + 1) + el * (double)(L - R + 1);
                    o.nocc += (L > 0);
                    dsum += L * (L - 1) / 2;  // N <= 250: fits easily
                } else {
                    if (rr > 0.0) {
                        acc += rr;
                        o.z = 32 * i + b;
                        o.right = 1;
                        if (t < acc) done = true;
                    }
                    if (!done && rl > 0.0) {
                        acc += rl;
                        o.z = 32 * i + b;
                        o.right = 0;
                        if (t < acc) done = true;
                    }
                }
            }
        }
        o.d = dsum;
        return o;
    }
};

// the single bit `pos` as a register array
template <int NWM>
__device__ __forceinline__ void bit_words(int pos, uint32_t (&c)[NWM]) {
    if (NWM == 1) {
        c[0] = 1u << pos;
    } else if (NWM == 2) {
        const unsigned long long m = 1ull << pos;
        c[0] = (uint32_t)m;
        c[NWM - 1] = (uint32_t)(m >> 32);
    } else {
        const int w = pos >> 5;
        const uint32_t b = 1u << (pos & 31);
#pragma unroll
        for (int i = 0; i < NWM; ++i) c[i] = (i == w) ? b : 0u;
    }
}

// n_pos += 1 / -= 1 by a bit-sliced ripple; the carry rarely travels beyond plane 1
// (both return whether the ripple reached plane 2 or higher, i.e. whether the ">= 4 particles" plane changed)
template <int NWM, int NP>
__device__ __forceinline__ bool planes_inc(uint32_t (&p)[NP][NWM], int pos) {
    uint32_t c[NWM];
    bit_words<NWM>(pos, c);
    bool high = false;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        if (k >= 2) {
            uint32_t any = 0u;
#pragma unroll
            for (int i = 0; i < NWM; ++i) any |= c[i];
            if (!any) break;
            high = true;
        }
#pragma unroll
        for (int i = 0; i < NWM; ++i) {
            const uint32_t carry = p[k][i] & c[i];
            p[k][i] ^= c[i];
            c[i] = carry;
        }
    }
    return high;
}
template <int NWM, int NP>
__device__ __forceinline__ bool planes_dec(uint32_t (&p)[NP][NWM], int pos) {
    uint32_t c[NWM];
    bit_words<NWM>(pos, c);
    bool high = false;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        if (k >= 2) {
            uint32_t any = 0u;
#pragma unroll
            for (int i = 0; i < NWM; ++i) any |= c[i];
            if (!any) break;
            high = true;
        }
#pragma unroll
        for (int i = 0; i < NWM; ++i) {
            const uint32_t borrow = ~p[k][i] & c[i];
            p[k][i] ^= c[i];
            c[i] = borrow;
        }
    }
    return high;
}

// hop across bond z (between mode z and mode z+1 mod M): right = z -> z+1, left = z+1 -> z
template <int NWM, int NP>
__device__ __forceinline__ bool apply_hop_planes(uint32_t (&p)[NP][NWM], const DevModel& dm, int z, bool right) {
    const int zn = (z + 1 == dm.M) ? 0 : z + 1;
    const bool a = planes_dec<NWM, NP>(p, right ? z : zn);
    const bool b = planes_inc<NWM, NP>(p, right ? zn : z);
    return a || b;
}

// one kinetic_sample! (qmc.jl:14-44) on the plane representation
template <int NWM, int NP>
__device__ __forceinline__ void step_planes(uint32_t (&p)[NP][NWM], const DevModel& dm, const DevTables& tb,
                                            const StepScratch<NWM, 4>& sc, uint32_t* stage, double u01, double& tau,
                                            double& eloc, double& gratio, int& nhops) {
    BondMasks<NWM, 4> bm;
    classify_planes<NWM, NP>(p, dm, bm);
    BondSums bs;
    double bigS;
    const PlanesBig<NWM, NP> big{p, bm.big, dm, stage, sc.exptab, sc.sqrtab};
    const double reg_total = bond_sums<NWM, 4, false>(tb, bm, SmemCumSink{sc.cum + threadIdx.x}, big, bs, bigS);
    stage_masks<NWM, 4>(bm, sc);
    const double D = dm.u * (double)bs.d;
    tau = 1.0 / bs.S;
    eloc = D - dm.t * bs.E;
    gratio = -D;
    nhops = 2 * bs.nocc;
    double T = u01 * bs.S;
    if (!(T < bs.S)) T = __longlong_as_double(__double_as_longlong(bs.S) - 1);
    int z;
    bool right;
    select_move<NWM, 4>(tb, sc, big, reg_total, bigS, T, z, right);
    if (z >= 0) apply_hop_planes<NWM, NP>(p, dm, z, right);
}

}  // namespace gwz
