// gwz_generic.cuh -- warp-cooperative engine for the reference's other single-component ansaetze
// (SURVEY.md 8f rank 2): ExtendedGutzwiller, Multinomial, DensityProfile, RelativeJastrow, Jastrow --
// and GutzwillerAnsatz itself, on HubbardReal1D / ExtendedHubbardReal1D.
//
// All of them are "log-linear": psi(n) = exp(-phi(n)),
//     phi(n) = 1/2 n^T Q n + b^T n + s * (sum_k log n_k! - log K),     d log psi / d p_k = -F_k(n),
// with a symmetric M x M matrix Q, a vector b and a scalar s that are linear in the parameters
// (built on the host per call, make_gen_tables in gwz_api_generic.cu) and integer- or log-valued
// features F_k.  The amplitude ratio of a hop i -> j needs no second amplitude evaluation
// (the reference evaluates ansatz(addr2) in full for every hop, qmc.jl:25-31):
//     psi(n')/psi(n) = exp(-(V_j - V_i + 1/2 Q_jj + 1/2 Q_ii - Q_ij + s (log(n_j + 1) - log n_i))),   V = Q n + b.
//
// One walker (or basis state, or probed address) per WARP: the occupation numbers, the potential
// V and the hop rates live in a per-warp slice of shared memory; lane l owns the modes
// [l*MPL, (l+1)*MPL) for the rates -- so a warp inclusive scan visits the hops in the reference's
// offdiagonals order (occupied modes ascending, right hop then left hop) and the
// move is picked by the reference's own rule (first i with u*S < cumsum_i, qmc.jl:52-61) -- and the
// parameters p = l + 32 q for features, gradient ratios and their accumulators.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gwz_device.cuh"

namespace gwz {

enum GenKind {
    GEN_GUTZWILLER = 0,
    GEN_EXT_GUTZWILLER = 1,
    GEN_MULTINOMIAL = 2,
    GEN_DENSITY_PROFILE = 3,
    GEN_RELATIVE_JASTROW = 4,
    GEN_JASTROW = 5
};

constexpr int GEN_BLOCK = 32;                // one warp per CTA: the shared-memory slice sits at fixed offsets
constexpr int GEN_WARPS = GEN_BLOCK / 32;
constexpr int GEN_HEAD = 8;                  // scalar slots of the accumulator vector
constexpr int GEN_REFRESH = 256;             // steps between from-scratch re-evaluations of V and F

// accumulator vector: GEN_HEAD scalars, then four P-vectors
enum GenSlot { GS_WALKERS = 0, GS_VAL = 1, GS_VAL2 = 2, GS_T = 3, GS_TE = 4, GS_TEE = 5, GS_HOPS = 6, GS_STEPS = 7 };
__host__ __device__ inline int gen_vec_len(int P) { return GEN_HEAD + 4 * P; }

struct GenModel {
    int N, M, P, kind;
    int B, W;   // address bits / 64-bit words
    int MP;     // 32 * MPL >= M
    int MPL;    // modes per lane
    double u, v, t;
    double s, logK;        // multinomial exponent (params[0]) and log of its normalisation; s = 0 otherwise
    const double* Qt;      // [M*M]
    const double* bvec;    // [M]
    const double* cR;      // [M]  1/2 Q_rr + 1/2 Q_kk - Q_kr, r = k+1
    const double* cL;      // [M]  same for l = k-1
    const double* logtab;  // [N+2] log(i), logtab[0] = 0
    const double* lgam;    // [N+2] log(i!)
    const double* sqrtab;  // [N+2]
    const uint16_t* pair_i;  // [P] Jastrow pair (i <= j) of parameter p (jastrow.jl:34-40 order)
    const uint16_t* pair_j;
    // translation-invariant couplings (Q circulant, s = 0: every ansatz but Multinomial and Jastrow): a hop changes
    // EVERY exponent, but by an amount that depends only on the distance to the hop -- facX[a] = exp(+D2[a]),
    // facY[a] = exp(-D2[a]), D2 = second difference of the first row of Q (see gen_rates_ti)
    int circ;
    int pad_;
    const double* facX;  // [M + GEN_FAC_PAD]: facX[i] = exp(+D2[(i - 1) mod M]) (one entry of wrap-around in front, the
    const double* facY;  //                      rest behind: the modes of a lane read consecutive entries, no index wrap)
};
constexpr int GEN_FAC_PAD = 16;

// CombinationAnsatz (src/ansatz/combination.jl): psi = alpha psi_1 + (1 - alpha) psi_2 with two log-linear components;
// parameters (p_1..., p_2..., alpha).  Kernels: gwz_generic_combo.cu.
struct GenCombo {
    GenModel g1, g2;
    double alpha;
    int P;  // g1.P + g2.P + 1
};

// per-warp shared-memory slice
struct GenWarp {
    int* occ;      // [MP]   occupation numbers (zero padded)
    double* V;     // [MP]   potential Q n + b
    double* rate;  // [2*MP] psi(n')/psi(n) of the right (2k) and left (2k+1) hop out of mode k; 0 = no hop
    double* mr;    // [2*MP] matrix element * rate (sweep)
    // combination kernels only: second potential, per-hop exponents of the two components, per-hop alpha-gradient
    double* V2;    // [MP]
    double* e1;    // [2*MP]
    double* e2;    // [2*MP]
    double* ga;    // [2*MP]
    __host__ __device__ static size_t bytes(int MP) { return (size_t)MP * (sizeof(int) + 5 * sizeof(double)); }
    __host__ __device__ static size_t bytes_combo(int MP) { return (size_t)MP * (sizeof(int) + 12 * sizeof(double)); }
    __device__ GenWarp(unsigned char* base, int MP, int warp) {
        unsigned char* p = base + (size_t)warp * bytes(MP);
        V = reinterpret_cast<double*>(p);
        rate = V + MP;
        mr = rate + 2 * MP;
        occ = reinterpret_cast<int*>(mr + 2 * MP);
        V2 = e1 = e2 = ga = nullptr;
    }
    // combination layout (one warp per CTA)
    struct ComboTag {};
    __device__ GenWarp(unsigned char* base, int MP, ComboTag) {
        V = reinterpret_cast<double*>(base);
        rate = V + MP;
        mr = rate + 2 * MP;
        V2 = mr + 2 * MP;
        e1 = V2 + MP;
        e2 = e1 + 2 * MP;
        ga = e2 + 2 * MP;
        occ = reinterpret_cast<int*>(ga + 2 * MP);
    }
    // the same slice seen by component 2 (its potential in place of V)
    __device__ GenWarp second() const {
        GenWarp w = *this;
        w.V = V2;
        return w;
    }
};

// Reductions / scans over a group of G consecutive lanes (G = 32: the warp; the walker kernel also runs 2 or 4
// walkers per warp in groups of 16 or 8 lanes, which amortises the per-step scalar work).  `lane` arguments of the
// helpers below are lane indices within the group.
template <int G = 32>
__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    return v;
}
template <int G = 32>
__device__ __forceinline__ long long warp_sum_all(long long v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    return v;
}
template <int G = 32>
__device__ __forceinline__ int warp_sum_all(int v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    return v;
}
template <int G = 32>
__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o, G);
        if (lane >= o) v += t;
    }
    return v;
}

// unary BoseFS words (W words in `a`, warp-uniform pointer) -> occupation numbers in shared memory
template <int G = 32>
__device__ __forceinline__ void gen_decode(const GenModel& gm, const uint64_t* a, const GenWarp& ws, int lane) {
    for (int k = lane; k < gm.MP; k += G) ws.occ[k] = 0;
    __syncwarp();
    for (int p = lane; p < gm.B; p += G) {
        const int w = p >> 6, b = p & 63;
        if ((a[w] >> b) & 1ull) {
            int zeros = __popcll(~a[w] & ((1ull << b) - 1ull));
            for (int i = 0; i < w; ++i) zeros += __popcll(~a[i]);
            atomicAdd(&ws.occ[zeros], 1);
        }
    }
    __syncwarp();
}
// occupation numbers -> unary words (lane 0 writes `out`, W words)
__device__ __forceinline__ void gen_encode(const GenModel& gm, const GenWarp& ws, int lane, uint64_t* out) {
    if (lane == 0) {
        uint64_t a[4] = {0ull, 0ull, 0ull, 0ull};
        int pos = 0;
        for (int k = 0; k < gm.M; ++k) {
            for (int q = 0; q < ws.occ[k]; ++q, ++pos) a[pos >> 6] |= 1ull << (pos & 63);
            ++pos;
        }
        for (int i = 0; i < gm.W; ++i) out[i] = a[i];
    }
}

// V = Q n + b from scratch (rows of Q are read coalesced: Q is symmetric)
template <int G = 32>
__device__ __forceinline__ void gen_potential(const GenModel& gm, const GenWarp& ws, int lane) {
    const int M = gm.M;
    for (int k = lane; k < M; k += G) {
        double acc = gm.bvec[k];
        for (int a = 0; a < M; ++a) {
            const int na = ws.occ[a];
            if (na) acc = fma(__ldg(gm.Qt + (size_t)a * M + k), (double)na, acc);
        }
        ws.V[k] = acc;
    }
    __syncwarp();
}

// sum n(n-1)/2, sum n_k n_{k+1} (periodic; Rimu extended_hubbard_interaction), sum log n_k!
template <int G = 32>
__device__ __forceinline__ void gen_totals(const GenModel& gm, const GenWarp& ws, int lane, long long& reg,
                                           long long& ext, double& lgsum) {
    long long r = 0, e = 0;
    double lg = 0.0;
    const int M = gm.M;
    for (int k = lane; k < M; k += G) {
        const int n = ws.occ[k];
        const int nr = ws.occ[k + 1 == M ? 0 : k + 1];
        r += (long long)n * (n - 1);
        e += (long long)n * nr;
        lg += gm.lgam[n];
    }
    reg = warp_sum_all<G>(r) / 2;
    ext = warp_sum_all<G>(e);
    lgsum = warp_sum_all<G>(lg);
}

// change of sum n_k n_{k+1} when one particle moves i -> j (nearest neighbours)
__device__ __forceinline__ int gen_dext(const GenModel& gm, const GenWarp& ws, int i, int j) {
    const int M = gm.M;
    if (M == 2) return 2 * (ws.occ[i] - ws.occ[j] - 1);
    const int jp = j + 1 == M ? 0 : j + 1, jm = j == 0 ? M - 1 : j - 1;
    const int ip = i + 1 == M ? 0 : i + 1, im = i == 0 ? M - 1 : i - 1;
    return ws.occ[jp] + ws.occ[jm] - ws.occ[ip] - ws.occ[im] - 1;
}

// feature F_p(n) = -d log psi / d p_p, from scratch
__device__ __forceinline__ double gen_feature(const GenModel& gm, const GenWarp& ws, int p, long long reg, long long ext,
                                              double lgsum) {
    switch (gm.kind) {
        case GEN_GUTZWILLER: return gm.u * (double)reg + gm.v * (double)ext;
        case GEN_EXT_GUTZWILLER: return p == 0 ? (double)reg : (double)ext;
        case GEN_MULTINOMIAL: return lgsum - gm.logK;
        case GEN_DENSITY_PROFILE: return (double)ws.occ[p];
        case GEN_RELATIVE_JASTROW: {
            long long s = 0;
            const int M = gm.M;
            for (int k = 0; k < M; ++k) {
                int kd = k + p;
                if (kd >= M) kd -= M;
                s += (long long)ws.occ[k] * ws.occ[kd];
            }
            return (double)s;
        }
        default: return (double)(ws.occ[gm.pair_i[p]] * ws.occ[gm.pair_j[p]]);
    }
}

// F_p(n') - F_p(n) for the hop i -> j, from the occupations BEFORE the hop
__device__ __forceinline__ double gen_dfeature(const GenModel& gm, const GenWarp& ws, int p, int i, int j) {
    const int ni = ws.occ[i], nj = ws.occ[j];
    switch (gm.kind) {
        case GEN_GUTZWILLER: return gm.u * (double)(nj - ni + 1) + gm.v * (double)gen_dext(gm, ws, i, j);
        case GEN_EXT_GUTZWILLER: return p == 0 ? (double)(nj - ni + 1) : (double)gen_dext(gm, ws, i, j);
        case GEN_MULTINOMIAL: return gm.logtab[nj + 1] - gm.logtab[ni];
        case GEN_DENSITY_PROFILE: return (double)((p == j) - (p == i));
        case GEN_RELATIVE_JASTROW: {
            // C_d(n') - C_d(n) = n_{j+d} + n_{j-d} - n_{i+d} - n_{i-d} + 2[d=0] - [d = i-j] - [d = j-i]  (mod M)
            const int M = gm.M, d = p;
            int jp = j + d, jm = j - d, ip = i + d, im = i - d;
            jp -= (jp >= M) ? M : 0;
            ip -= (ip >= M) ? M : 0;
            jm += (jm < 0) ? M : 0;
            im += (im < 0) ? M : 0;
            int dij = i - j, dji = j - i;
            dij += (dij < 0) ? M : 0;
            dji += (dji < 0) ? M : 0;
            const int r = ws.occ[jp] + ws.occ[jm] - ws.occ[ip] - ws.occ[im] + 2 * (d == 0) - (d == dij) - (d == dji);
            return (double)r;
        }
        default: {
            const int a = gm.pair_i[p], b = gm.pair_j[p];
            const int na = ws.occ[a], nb = ws.occ[b];
            const int na2 = na + (a == j) - (a == i), nb2 = nb + (b == j) - (b == i);
            return (double)(na2 * nb2 - na * nb);
        }
    }
}

// phi(n) = 1/2 n^T Q n + b^T n + s (sum log n! - log K)  (psi = exp(-phi))
template <int G = 32>
__device__ __forceinline__ double gen_phi(const GenModel& gm, const GenWarp& ws, int lane, double lgsum) {
    double acc = 0.0;
    for (int k = lane; k < gm.M; k += G) {
        const int n = ws.occ[k];
        if (n) acc = fma((double)n, 0.5 * (ws.V[k] + gm.bvec[k]), acc);
    }
    return warp_sum_all<G>(acc) + gm.s * (lgsum - gm.logK);
}

// hop rates of the modes this lane owns, in reference order; lane sums of the rates, of
// sqrt(n_src (n_dst+1)) * rate, and the number of occupied modes
template <int MPL, bool STORE_MR>
__device__ __forceinline__ void gen_rates(const GenModel& gm, const GenWarp& ws, int lane, double& S_lane,
                                          double& E_lane, int& nocc_lane) {
    const int M = gm.M;
    double S = 0.0, E = 0.0;
    int nocc = 0;
#pragma unroll
    for (int q = 0; q < MPL; ++q) {
        const int k = lane * MPL + q;
        if (k < M) {
            double rr = 0.0, rl = 0.0, mrr = 0.0, mrl = 0.0;
            const int n = ws.occ[k];
            if (n > 0) {
                const int kr = k + 1 == M ? 0 : k + 1, kl = k == 0 ? M - 1 : k - 1;
                const int nr = ws.occ[kr], nl = ws.occ[kl];
                const double vk = ws.V[k];
                double er = ws.V[kr] - vk + gm.cR[k];
                double el = ws.V[kl] - vk + gm.cL[k];
                if (gm.s != 0.0) {
                    const double ln = gm.logtab[n];
                    er += gm.s * (gm.logtab[nr + 1] - ln);
                    el += gm.s * (gm.logtab[nl + 1] - ln);
                }
                rr = exp(-er);
                rl = exp(-el);
                const double sn = gm.sqrtab[n];
                mrr = sn * gm.sqrtab[nr + 1] * rr;
                mrl = sn * gm.sqrtab[nl + 1] * rl;
                S += rr;
                S += rl;
                E += mrr;
                E += mrl;
                ++nocc;
            }
            ws.rate[2 * k] = rr;
            ws.rate[2 * k + 1] = rl;
            if (STORE_MR) {
                ws.mr[2 * k] = -gm.t * mrr;
                ws.mr[2 * k + 1] = -gm.t * mrl;
            }
        }
    }
    S_lane = S;
    E_lane = E;
    nocc_lane = nocc;
}

// ---- translation-invariant couplings: hop rates carried from step to step ----------------------------------------
// exp(-e) of both hops out of EVERY mode this lane owns (occupied or not), from the potential V
template <int MPL>
__device__ __forceinline__ void gen_raw_rates(const GenModel& gm, const GenWarp& ws, int lane, double (&rR)[MPL],
                                              double (&rL)[MPL]) {
    const int M = gm.M;
#pragma unroll
    for (int q = 0; q < MPL; ++q) {
        const int k = lane * MPL + q;
        rR[q] = rL[q] = 0.0;
        if (k < M) {
            const int kr = k + 1 == M ? 0 : k + 1, kl = k == 0 ? M - 1 : k - 1;
            const double vk = ws.V[k];
            rR[q] = exp(-(ws.V[kr] - vk + gm.cR[k]));  // the arithmetic of gen_rates
            rL[q] = exp(-(ws.V[kl] - vk + gm.cL[k]));
        }
    }
}
// matrix-element weights sqrt(n_k (n_kr + 1)), sqrt(n_k (n_kl + 1)) of the hops out of mode k = lane MPL + q (the products
// gen_rates forms first) and the occupied bit of the mode
template <int MPL>
__device__ __forceinline__ void gen_weight_ti(const GenModel& gm, const GenWarp& ws, int lane, int q, double& wR, double& wL,
                                              unsigned& occmask) {
    const int M = gm.M, k = lane * MPL + q;
    const int n = ws.occ[k];
    const int kr = k + 1 == M ? 0 : k + 1, kl = k == 0 ? M - 1 : k - 1;
    const double sn = gm.sqrtab[n];
    wR = sn * gm.sqrtab[ws.occ[kr] + 1];
    wL = sn * gm.sqrtab[ws.occ[kl] + 1];
    occmask = (occmask & ~(1u << q)) | ((n > 0 ? 1u : 0u) << q);
}
// the sums of gen_rates from the carried rates and weights (same order of additions and products); eR / eL: the rates of
// the existing hops
template <int MPL>
__device__ __forceinline__ void gen_rates_ti(const double (&rR)[MPL], const double (&rL)[MPL], const double (&wR)[MPL],
                                             const double (&wL)[MPL], unsigned occmask, double (&eR)[MPL], double (&eL)[MPL],
                                             double& S_lane, double& E_lane) {
    double S = 0.0, E = 0.0;
#pragma unroll
    for (int q = 0; q < MPL; ++q) {
        const bool ex = (occmask >> q) & 1u;
        eR[q] = ex ? rR[q] : 0.0;
        eL[q] = ex ? rL[q] : 0.0;
        S += eR[q];
        S += eL[q];
        E += wR[q] * eR[q];
        E += wL[q] * eL[q];
    }
    S_lane = S;
    E_lane = E;
}
// the hop src -> src +- 1 multiplies the rates out of mode k by factors that depend on a = k - src (mod M) only:
// right hop done: e_R(k) -= D2[a], e_L(k) += D2[a-1];  left hop done: e_R(k) += D2[a+1], e_L(k) -= D2[a]
template <int MPL>
__device__ __forceinline__ void gen_rates_ti_hop(const GenModel& gm, int lane, int src, int right, double (&rR)[MPL],
                                                 double (&rL)[MPL]) {
    const int M = gm.M;
    int a0 = lane * MPL - src;  // distance of this lane's first mode, in [0, M)
    a0 += a0 < 0 ? M : 0;
    // table entry i + 1 holds the factor of distance i: X[a], Y[a-1] after a right hop; Y[a+1], X[a] after a left hop
    const double* pR = (right ? gm.facX + 1 : gm.facY + 2) + a0;
    const double* pL = (right ? gm.facY : gm.facX + 1) + a0;
#pragma unroll
    for (int q = 0; q < MPL; ++q) {
        if (lane * MPL + q < M) {
            rR[q] *= __ldg(pR + q);
            rL[q] *= __ldg(pL + q);
        }
    }
}
// gen_select over the carried rates (registers instead of the shared-memory rate array): the running sum of
// gen_select (same additions in the same order; absent hops add 0), the hop = the number of running sums <= T
template <int MPL, int G = 32>
__device__ __forceinline__ void gen_select_ti(int lane, const double (&eR)[MPL], const double (&eL)[MPL], unsigned occmask,
                                              double S_lane, double cum_incl, double T, int& src, int& right) {
    unsigned hit = __ballot_sync(0xffffffffu, (T < cum_incl) && (S_lane > 0.0));
    unsigned live = __ballot_sync(0xffffffffu, S_lane > 0.0);
    if (G < 32) {  // this group's share of the ballots
        const int shift = (int)(threadIdx.x & 31u & ~(unsigned)(G - 1));
        hit = (hit >> shift) & ((1u << (G & 31)) - 1u);
        live = (live >> shift) & ((1u << (G & 31)) - 1u);
    }
    const int sel_lane = hit ? (__ffs(hit) - 1) : (31 - __clz(live));  // clamp as gen_select
    double acc = cum_incl - S_lane;
    int h = 0;
#pragma unroll
    for (int q = 0; q < MPL; ++q) {
        acc += eR[q];
        h += !(T < acc);
        acc += eL[q];
        h += !(T < acc);
    }
    // rounding left T at or above every running sum: the last hop of this lane (left hop of its last occupied mode)
    if (h >= 2 * MPL) h = 2 * (31 - __clz(occmask | 1u)) + 1;
    src = __shfl_sync(0xffffffffu, lane * MPL + (h >> 1), sel_lane, G);
    right = __shfl_sync(0xffffffffu, (h & 1) ^ 1, sel_lane, G);
}

// pick_random_from_cumsum (qmc.jl:52-61) over the hops in reference order: first hop whose inclusive
// running sum exceeds T = u01 * S.  Returns the source mode and direction (warp-uniform).
template <int MPL, int G = 32>
__device__ __forceinline__ void gen_select(const GenModel& gm, const GenWarp& ws, int lane, double S_lane, double cum_incl,
                                           double T, int& src, int& right) {
    unsigned hit = __ballot_sync(0xffffffffu, (T < cum_incl) && (S_lane > 0.0));
    unsigned live = __ballot_sync(0xffffffffu, S_lane > 0.0);
    if (G < 32) {  // this group's share of the ballots
        const int shift = (int)(threadIdx.x & 31u & ~(unsigned)(G - 1));
        hit = (hit >> shift) & ((1u << (G & 31)) - 1u);
        live = (live >> shift) & ((1u << (G & 31)) - 1u);
    }
    // rounding can leave T >= every running sum: clamp to the last hop
    const int sel_lane = hit ? (__ffs(hit) - 1) : (31 - __clz(live));
    int mode = -1, dir = 1;
    if (lane == sel_lane) {
        double acc = cum_incl - S_lane;
#pragma unroll
        for (int q = 0; q < MPL; ++q) {
            const int k = lane * MPL + q;
            if (k >= gm.M || mode >= 0) break;
            const double rr = ws.rate[2 * k], rl = ws.rate[2 * k + 1];
            if (rr > 0.0 || rl > 0.0) {
                acc += rr;
                if (T < acc) {
                    mode = k;
                    dir = 1;
                    break;
                }
                acc += rl;
                if (T < acc) {
                    mode = k;
                    dir = 0;
                    break;
                }
            }
        }
        if (mode < 0) {  // clamp: last hop of this lane
            for (int q = MPL - 1; q >= 0 && mode < 0; --q) {
                const int k = lane * MPL + q;
                if (k < gm.M && (ws.rate[2 * k] > 0.0 || ws.rate[2 * k + 1] > 0.0)) {
                    mode = k;
                    dir = 0;
                }
            }
        }
    }
    src = __shfl_sync(0xffffffffu, mode, sel_lane, G);
    right = __shfl_sync(0xffffffffu, dir, sel_lane, G);
}

}  // namespace gwz
