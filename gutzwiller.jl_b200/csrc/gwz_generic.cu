// gwz_generic.cu -- kernels of the ansatz-generic engine (gwz_generic.cuh): per-address probe /
// ansatz evaluation, the kinetic-VQMC walker kernel and the LocalEnergyEvaluator sweep, one warp per
// walker / state.  Reference: src/kinetic-vqmc/qmc.jl:14-106, src/kinetic-vqmc/state.jl:42-49,
// src/localenergy.jl:94-155, src/ansatz/*.jl.
#include "gwz_generic.cuh"
#include "gwz_internal.h"

namespace gwz {

// runtime modes-per-lane -> compile-time (probe and sweep; the walker kernel is templated on it)
#define GEN_MPL_SWITCH(mpl, CALL)       \
    switch (mpl) {                      \
        case 1: { constexpr int MPL_ = 1; CALL; } break; \
        case 2: { constexpr int MPL_ = 2; CALL; } break; \
        case 4: { constexpr int MPL_ = 4; CALL; } break; \
        default: { constexpr int MPL_ = 8; CALL; } break; \
    }

// ------------------------------------------------------------------------------------------
// layout conversion: unary addresses (AoS n x W) <-> occupation bytes (n x Mpad)
// ------------------------------------------------------------------------------------------
__global__ void gen_unary_to_occ_kernel(const uint64_t* __restrict__ addr, size_t n, int W, int B, int M, int Mpad,
                                        uint8_t* __restrict__ occ, int broadcast) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const uint64_t* a = addr + (broadcast ? 0 : w * (size_t)W);
    int mode = 0, cnt = 0;
    for (int p = 0; p <= B; ++p) {
        const int bit = (p < B) ? (int)((a[p >> 6] >> (p & 63)) & 1ull) : 0;
        if (bit) {
            ++cnt;
        } else {
            if (mode < M) occ[w * Mpad + mode] = (uint8_t)cnt;
            ++mode;
            cnt = 0;
        }
    }
    for (int k = M; k < Mpad; ++k) occ[w * Mpad + k] = 0;
}
__global__ void gen_occ_to_unary_kernel(const uint8_t* __restrict__ occ, size_t n, int W, int M, int Mpad,
                                        uint64_t* __restrict__ addr) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    uint64_t a[4] = {0ull, 0ull, 0ull, 0ull};
    int pos = 0;
    for (int k = 0; k < M; ++k) {
        const int c = occ[w * Mpad + k];
        for (int q = 0; q < c; ++q, ++pos) a[pos >> 6] |= 1ull << (pos & 63);
        ++pos;
    }
    for (int i = 0; i < W; ++i) addr[w * (size_t)W + i] = a[i];
}

// ------------------------------------------------------------------------------------------
// probe: one kinetic_sample! (qmc.jl:14-44) and / or val_and_grad(ansatz, addr, p) per address
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEN_BLOCK)
gen_probe_kernel(const __grid_constant__ GenModel gm, const uint64_t* __restrict__ addr, size_t n,
                 const double* __restrict__ u01, double* __restrict__ val_out, double* __restrict__ grad_out,
                 double* __restrict__ e_loc, double* __restrict__ tau_out, double* __restrict__ grad_ratio,
                 int* __restrict__ n_hops, double* __restrict__ rates, int* __restrict__ chosen,
                 uint64_t* __restrict__ next_addr, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const GenWarp ws(smem_raw, gm.MP, warp);
    const size_t nwarps = (size_t)gridDim.x * GEN_WARPS;
    const int M = gm.M, P = gm.P;
    for (size_t idx = (size_t)blockIdx.x * GEN_WARPS + warp; idx < n; idx += nwarps) {
        gen_decode(gm, addr + idx * (size_t)gm.W, ws, lane);
        gen_potential(gm, ws, lane);
        long long reg, ext;
        double lgsum;
        gen_totals(gm, ws, lane, reg, ext, lgsum);
        if (val_out || grad_out) {
            const double val = exp(-gen_phi(gm, ws, lane, lgsum));
            if (val_out && lane == 0) val_out[idx] = val;
            if (grad_out)
                for (int p = lane; p < P; p += 32) {
                    const double F = gen_feature(gm, ws, p, reg, ext, lgsum);
                    // MultinomialAnsatz: val * log(y) (multinomial.jl:58); the others: -(feature * val)
                    grad_out[idx * (size_t)P + p] = (gm.kind == GEN_MULTINOMIAL) ? val * -F : -(F * val);
                }
        }
        if (!(e_loc || tau_out || grad_ratio || n_hops || rates || chosen || next_addr)) continue;
        double S_lane, E_lane;
        int nocc_lane;
        GEN_MPL_SWITCH(gm.MPL, (gen_rates<MPL_, false>(gm, ws, lane, S_lane, E_lane, nocc_lane)));
        __syncwarp();
        const double cum = warp_scan_incl(S_lane, lane);
        const double S = __shfl_sync(0xffffffffu, cum, 31);
        const double Etot = warp_sum_all(E_lane);
        const int nocc = warp_sum_all(nocc_lane);
        const double tau = 1.0 / S;
        const double D = gm.u * (double)reg + gm.v * (double)ext;
        if (lane == 0) {
            if (e_loc) e_loc[idx] = D - gm.t * Etot;
            if (tau_out) tau_out[idx] = tau;
            if (n_hops) n_hops[idx] = 2 * nocc;
            if (!(tau < 1.0e300) || !(tau == tau)) atomicOr(err_flag, 1);
        }
        if (grad_ratio)
            for (int p = lane; p < P; p += 32) grad_ratio[idx * (size_t)P + p] = -gen_feature(gm, ws, p, reg, ext, lgsum);
        if (rates) {
            double* row = rates + idx * (size_t)(2 * M);
            for (int h = lane; h < 2 * M; h += 32) row[h] = 0.0;
            __syncwarp();
            // exclusive prefix of occupied modes over the lanes (lanes own contiguous mode ranges)
            int incl = nocc_lane;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int h = 2 * (incl - nocc_lane);
            for (int q = 0; q < gm.MPL; ++q) {
                const int k = lane * gm.MPL + q;
                if (k < M && ws.occ[k] > 0) {
                    row[h] = ws.rate[2 * k];
                    row[h + 1] = ws.rate[2 * k + 1];
                    h += 2;
                }
            }
        }
        if (chosen || next_addr) {
            const double u = u01 ? u01[idx] : 0.5;
            double T = u * S;
            if (!(T < S)) T = __longlong_as_double(__double_as_longlong(S) - 1);
            int src, right;
            GEN_MPL_SWITCH(gm.MPL, (gen_select<MPL_>(gm, ws, lane, S_lane, cum, T, src, right)));
            if (chosen) {
                int below = 0;
                for (int k = lane; k < src; k += 32) below += ws.occ[k] > 0;
                below = warp_sum_all(below);
                if (lane == 0) chosen[idx] = 2 * below + (right ? 1 : 2);
            }
            if (next_addr) {
                const int dst = right ? (src + 1 == M ? 0 : src + 1) : (src == 0 ? M - 1 : src - 1);
                __syncwarp();
                if (lane == 0) {
                    ws.occ[src] -= 1;
                    ws.occ[dst] += 1;
                }
                __syncwarp();
                gen_encode(gm, ws, lane, next_addr + idx * (size_t)gm.W);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// walkers: kinetic_vqmc!(st; steps) (qmc.jl:73-97), running sums of state.jl:42-49 per walker
// ------------------------------------------------------------------------------------------
#ifndef GWZ_GEN_MIN_BLOCKS
#define GWZ_GEN_MIN_BLOCKS 24
#endif
// G lanes per walker (32 / G walkers per warp): lane gl of a group owns the modes [gl*MPL, (gl+1)*MPL) for the rates and
// the parameters gl + G q; the groups of a warp advance in lockstep (same number of steps), each with its own walker.
template <int G, int PPL, int MPL, bool TI>
__global__ void __launch_bounds__(GEN_BLOCK, PPL <= 2 ? GWZ_GEN_MIN_BLOCKS : (PPL == 4 ? 12 : 8))
gen_walker_kernel(const __grid_constant__ GenModel gm, const __grid_constant__ PhiloxKeys pk, uint8_t* __restrict__ occ_g,
                  int Mpad, double* __restrict__ acc_g, size_t n, uint64_t goff, uint64_t step0, uint64_t steps,
                  int accumulate, int reset_acc, const double* __restrict__ shift, double* __restrict__ partials,
                  double* __restrict__ rec_tau, double* __restrict__ rec_eloc, uint64_t* __restrict__ rec_addr,
                  int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int GROUPS = 32 / G;  // walkers per warp (one warp per CTA)
    int lane = threadIdx.x & (G - 1);
    const int group = (threadIdx.x & 31) / G;
    GenWarp ws(smem_raw, G * MPL, group);  // compile-time slice layout: offsets fold into the LDS/STS immediates
    // keep the lane index and the slice pointers in registers: left alone, the compiler recomputes them from the special
    // registers several times per step (50 of 436 instructions per walker-step, ncu source page r02m)
    asm volatile("" : "+r"(lane), "+l"(ws.occ), "+l"(ws.V), "+l"(ws.rate));
    const size_t gwarp = (size_t)blockIdx.x * GROUPS + group, nwarps = (size_t)gridDim.x * GROUPS;
    const int M = gm.M, P = gm.P;
    const int AW = 3 + 2 * P;  // per-walker sums in HBM: st, stE, stEE, stG[P], stGE[P]
    const double E0 = (accumulate && shift) ? shift[0] : 0.0;
    double G0[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
        const int p = lane + G * q;
        G0[q] = (accumulate && shift && p < P) ? shift[1 + p] : 0.0;
    }
    double ps[GEN_HEAD];
    double pv[4][PPL];
#pragma unroll
    for (int i = 0; i < GEN_HEAD; ++i) ps[i] = 0.0;
#pragma unroll
    for (int q = 0; q < PPL; ++q) pv[0][q] = pv[1][q] = pv[2][q] = pv[3][q] = 0.0;
    bool bad = false;

    for (size_t wb = (size_t)blockIdx.x * GROUPS; wb < n; wb += nwarps) {  // warp-uniform trip count
        const bool live = wb + group < n;
        const size_t w = live ? wb + group : n - 1;  // a spare group repeats the last walker and writes nothing
        for (int k = lane; k < G * MPL; k += G) ws.occ[k] = (k < M) ? (int)occ_g[w * (size_t)Mpad + k] : 0;
        __syncwarp();
        gen_potential<G>(gm, ws, lane);
        long long reg, ext;
        double lgsum;
        gen_totals<G>(gm, ws, lane, reg, ext, lgsum);
        double F[PPL], aG[PPL], aGE[PPL];
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            const int p = lane + G * q;
            F[q] = (p < P) ? gen_feature(gm, ws, p, reg, ext, lgsum) : 0.0;
            aG[q] = aGE[q] = 0.0;
        }
        double st = 0.0, ste = 0.0, stee = 0.0;
        if (accumulate && !reset_acc) {
            const double* a = acc_g + w * (size_t)AW;
            st = a[0];
            ste = a[1];
            stee = a[2];
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int p = lane + G * q;
                if (p < P) {
                    aG[q] = a[3 + p];
                    aGE[q] = a[3 + P + p];
                }
            }
        }
        const uint64_t wid = goff + (uint64_t)w;
        const uint32_t c2 = (uint32_t)wid, c3 = (uint32_t)(wid >> 32);
        uint32_t rnd[4] = {0u, 0u, 0u, 0u};
        unsigned long long hops = 0ull;
        int nocc = 0;
        for (int k = lane; k < M; k += G) nocc += ws.occ[k] > 0;
        nocc = warp_sum_all<G>(nocc);
        // translation-invariant couplings: exp(-e) of every hop out of this lane's modes, carried from step to step
        double rR[MPL], rL[MPL], eR[MPL], eL[MPL], wR[MPL], wL[MPL];
        unsigned occmask = 0u;
        // RelativeJastrow: the occupations twice in a row (in the slice's unused rate array), so that n_(j+-d) needs no
        // index wrap, and the part of gen_dfeature that does not depend on the hop (|i - j| = 1 on the ring)
        int* o2 = reinterpret_cast<int*>(ws.rate);
        const bool rj = TI && gm.kind == GEN_RELATIVE_JASTROW && M > 2;
        int corr[PPL];
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            const int p = lane + G * q;
            corr[q] = 2 * (p == 0) - (p == 1) - (p == M - 1);
        }
        if (TI) {
            for (int k = lane; k < 2 * M; k += G) o2[k] = ws.occ[k < M ? k : k - M];
            gen_raw_rates<MPL>(gm, ws, lane, rR, rL);
#pragma unroll
            for (int q = 0; q < MPL; ++q) {
                wR[q] = wL[q] = 0.0;
                if (lane * MPL + q < M) gen_weight_ti<MPL>(gm, ws, lane, q, wR[q], wL[q], occmask);
            }
        }

        for (uint64_t k = 0; k < steps; ++k) {
            const uint64_t step = step0 + k;
            if (k != 0 && (k % GEN_REFRESH) == 0) {  // bound the rounding drift of the incremental V and F
                gen_potential<G>(gm, ws, lane);
                gen_totals<G>(gm, ws, lane, reg, ext, lgsum);
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    const int p = lane + G * q;
                    if (p < P) F[q] = gen_feature(gm, ws, p, reg, ext, lgsum);
                }
                if (TI) gen_raw_rates<MPL>(gm, ws, lane, rR, rL);  // ... and of the multiplicatively carried rates
            }
            // one Philox block serves two steps; the G lanes of the walker evaluate G consecutive blocks at once
            // (lane l: block (blk & ~(G-1)) + l) and hand their words to the group when their step comes
            const uint64_t blk = step >> 1;
            if (k == 0 || ((step & 1ull) == 0ull && (blk & (uint64_t)(G - 1)) == 0ull)) {
                const uint64_t mine = (blk & ~(uint64_t)(G - 1)) + (uint64_t)lane;
                philox4x32_10_keyed((uint32_t)mine, (uint32_t)(mine >> 32), c2, c3, pk, rnd);
            }
            const int rsrc = (int)(blk & (uint64_t)(G - 1));
            const uint32_t wlo = __shfl_sync(0xffffffffu, (step & 1ull) ? rnd[2] : rnd[0], rsrc, G);
            const uint32_t whi = __shfl_sync(0xffffffffu, (step & 1ull) ? rnd[3] : rnd[1], rsrc, G);
            const double u01 = u01_from_words(wlo, whi);
            double S_lane, E_lane;
            int nocc_lane;
            if (TI)
                gen_rates_ti<MPL>(rR, rL, wR, wL, occmask, eR, eL, S_lane, E_lane);
            else
                gen_rates<MPL, false>(gm, ws, lane, S_lane, E_lane, nocc_lane);
            const double cum = warp_scan_incl<G>(S_lane, lane);
            const double S = __shfl_sync(0xffffffffu, cum, G - 1, G);
            const double Etot = warp_sum_all<G>(E_lane);
            const double tau = 1.0 / S;
            const double eloc = gm.u * (double)reg + gm.v * (double)ext - gm.t * Etot;
            bad |= !(tau < 1.0e300);
            if (rec_tau && lane == 0 && live) {
                rec_tau[k * n + w] = tau;
                rec_eloc[k * n + w] = eloc;
            }
            if (rec_addr && live) gen_encode(gm, ws, lane, rec_addr + (k * n + w) * (size_t)gm.W);
            const double dE = eloc - E0;
            const double tE = tau * dE;
            st += tau;
            ste += tE;
            stee = fma(tE, dE, stee);
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const double dG = -F[q] - G0[q];
                aG[q] = fma(tau, dG, aG[q]);
                aGE[q] = fma(tE, dG, aGE[q]);
            }
            hops += (unsigned)(2 * nocc);
            double T = u01 * S;
            if (!(T < S)) T = __longlong_as_double(__double_as_longlong(S) - 1);
            int src, right;
            if (TI)
                gen_select_ti<MPL, G>(lane, eR, eL, occmask, S_lane, cum, T, src, right);
            else
                gen_select<MPL, G>(gm, ws, lane, S_lane, cum, T, src, right);
            const int dst = right ? (src + 1 == M ? 0 : src + 1) : (src == 0 ? M - 1 : src - 1);
            nocc += (ws.occ[dst] == 0) - (ws.occ[src] == 1);  // occupied modes, carried exactly
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int p = lane + G * q;
                if (p < P) {
                    if (rj)  // C_d(n') - C_d(n) = n_(j+d) + n_(j-d) - n_(i+d) - n_(i-d) + 2[d=0] - [d=1] - [d=M-1]
                        F[q] += (double)(o2[dst + p] + o2[dst - p + M] - o2[src + p] - o2[src - p + M] + corr[q]);
                    else
                        F[q] += gen_dfeature(gm, ws, p, src, dst);
                }
            }
            reg += ws.occ[dst] - ws.occ[src] + 1;
            ext += gen_dext(gm, ws, src, dst);
            __syncwarp();  // every lane has read the pre-hop occupations
            if (TI) {  // the potential is only needed at the next re-evaluation, which rebuilds it
                gen_rates_ti_hop<MPL>(gm, lane, src, right, rR, rL);
            } else {
                const double* qd = gm.Qt + (size_t)dst * M;
                const double* qs = gm.Qt + (size_t)src * M;
#pragma unroll
                for (int q = 0; q < MPL; ++q) {
                    const int m = lane + G * q;
                    if (m < M) ws.V[m] += __ldg(qd + m) - __ldg(qs + m);
                }
            }
            if (lane == 0) {
                const int ns = ws.occ[src] - 1, nd = ws.occ[dst] + 1;
                ws.occ[src] = ns;
                ws.occ[dst] = nd;
                if (TI) {
                    o2[src] = o2[src + M] = ns;
                    o2[dst] = o2[dst + M] = nd;
                }
            }
            __syncwarp();
            if (TI) {
                // The weights of the four modes around the hop change: first .. first + 3 (mod M), first = lo - 1, lo = the
                // lower end of the bond (M-1 for the bond M-1 -> 0).  Lanes 0..3 of the walker's group each evaluate one
                // of them -- one pass of gen_weight for every walker of the warp at once, instead of one (predicated)
                // pass per mode on the one or two lanes that own them -- and the owners fetch the results by shuffle.
                const int lo = right ? src : dst;
                int first = lo - 1;
                first += first < 0 ? M : 0;
                double nwR = 0.0, nwL = 0.0;
                int nex = 0;
                if (lane < 4) {
                    int k = first + lane;
                    k -= k >= M ? M : 0;
                    const int n = ws.occ[k];
                    const int kr = k + 1 == M ? 0 : k + 1, kl = k == 0 ? M - 1 : k - 1;
                    const double sn = gm.sqrtab[n];
                    nwR = sn * gm.sqrtab[ws.occ[kr] + 1];
                    nwL = sn * gm.sqrtab[ws.occ[kl] + 1];
                    nex = n > 0;
                }
                int d = lane * MPL - first;  // distance of this lane's first mode from the window's first, mod M
                d += d < 0 ? M : 0;
#pragma unroll
                for (int q = 0; q < MPL; ++q) {
                    const int from = d & 3;  // (any lane when the mode is outside the window: the value is not used)
                    const double vR = __shfl_sync(0xffffffffu, nwR, from, G), vL = __shfl_sync(0xffffffffu, nwL, from, G);
                    const int vx = __shfl_sync(0xffffffffu, nex, from, G);
                    if (d < 4 && lane * MPL + q < M) {
                        wR[q] = vR;
                        wL[q] = vL;
                        occmask = (occmask & ~(1u << q)) | ((unsigned)vx << q);
                    }
                    d = d + 1 == M ? 0 : d + 1;
                }
            }
        }

        if (live)
            for (int k = lane; k < M; k += G) occ_g[w * (size_t)Mpad + k] = (uint8_t)ws.occ[k];
        if (accumulate && live) {
            double* a = acc_g + w * (size_t)AW;
            if (lane == 0) {
                a[0] = st;
                a[1] = ste;
                a[2] = stee;
            }
            const double inv = 1.0 / st;
            const double mE = ste * inv;
            const double val_w = E0 + mE;
            ps[GS_WALKERS] += 1.0;
            ps[GS_VAL] += val_w;
            ps[GS_VAL2] = fma(val_w, val_w, ps[GS_VAL2]);
            ps[GS_T] += st;
            ps[GS_TE] += st * val_w;
            ps[GS_TEE] += st * E0 * E0 + 2.0 * E0 * ste + stee;
            ps[GS_HOPS] += (double)hops;
            ps[GS_STEPS] += (double)steps;
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int p = lane + G * q;
                if (p < P) {
                    a[3 + p] = aG[q];
                    a[3 + P + p] = aGE[q];
                    const double mG = aG[q] * inv;
                    const double grad_w = 2.0 * (aGE[q] * inv - mE * mG);
                    pv[0][q] += grad_w;
                    pv[1][q] = fma(grad_w, grad_w, pv[1][q]);
                    pv[2][q] += st * G0[q] + aG[q];
                    pv[3][q] += st * G0[q] * E0 + G0[q] * ste + E0 * aG[q] + aGE[q];
                }
            }
        }
        __syncwarp();
    }
    if (bad && lane == 0) atomicOr(err_flag, 1);
    if (!accumulate || partials == nullptr) return;
    double* out = partials + gwarp * (size_t)gen_vec_len(P);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < GEN_HEAD; ++i) out[i] = ps[i];
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
        const int p = lane + G * q;
        if (p < P) {
            out[GEN_HEAD + p] = pv[0][q];
            out[GEN_HEAD + P + p] = pv[1][q];
            out[GEN_HEAD + 2 * P + p] = pv[2][q];
            out[GEN_HEAD + 3 * P + p] = pv[3][q];
        }
    }
}

// val_w[n], grad_w[n x P] from the per-walker sums (state.jl:42-49)
__global__ void gen_estimates_kernel(const double* __restrict__ acc, size_t n, int P, const double* __restrict__ shift,
                                     double* __restrict__ val_w, double* __restrict__ grad_w) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const double* a = acc + w * (size_t)(3 + 2 * P);
    const double inv = 1.0 / a[0];
    const double mE = a[1] * inv;
    if (val_w) val_w[w] = shift[0] + mE;
    if (grad_w)
        for (int p = 0; p < P; ++p) grad_w[w * (size_t)P + p] = 2.0 * (a[3 + P + p] * inv - mE * (a[3 + p] * inv));
}

// ------------------------------------------------------------------------------------------
// LocalEnergyEvaluator (localenergy.jl:94-155): one warp per basis state
// ------------------------------------------------------------------------------------------
// G lanes per state (32 / G states per warp); gm.MPL / gm.MP are the per-group values set by the launcher
template <int G, int PPL>
__global__ void __launch_bounds__(GEN_BLOCK)
gen_sweep_kernel(const __grid_constant__ GenModel gm, const uint64_t* __restrict__ basis, uint64_t dim, uint64_t rank_begin,
                 const uint64_t* __restrict__ binom, uint64_t chunk, int with_grad, double* __restrict__ partials,
                 double* __restrict__ terms_out, uint64_t* __restrict__ states_out, uint64_t first, uint64_t count,
                 const uint32_t* __restrict__ weight /* stored mode, optional: orbit sizes of a representative basis */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int GROUPS = 32 / G;
    const int lane = threadIdx.x & (G - 1), group = (threadIdx.x & 31) / G;
    const GenWarp ws(smem_raw, gm.MP, group);
    uint64_t* aw = reinterpret_cast<uint64_t*>(smem_raw + (size_t)GROUPS * GenWarp::bytes(gm.MP)) + 4 * group;
    const uint64_t gwarp = (uint64_t)blockIdx.x * GROUPS + group;
    const int M = gm.M, P = gm.P;
    const bool per_state = terms_out != nullptr || states_out != nullptr;
    const uint64_t end = per_state ? first + count : dim;
    uint64_t lo = per_state ? first + gwarp * chunk : gwarp * chunk;
    if (lo > end) lo = end;
    const uint64_t hi = (lo + chunk < end) ? lo + chunk : end;
    double num = 0.0, den = 0.0, ng[PPL], dg[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) ng[q] = dg[q] = 0.0;
    uint64_t x = 0;
    for (uint64_t it = 0; it < chunk; ++it) {  // warp-uniform trip count; a group past its range idles on a valid state
        const uint64_t sidx = lo + it;
        const bool live = sidx < hi;
        if (live) {
            if (basis) {
                if (lane < gm.W) aw[lane] = basis[(uint64_t)lane * dim + sidx];
            } else {
                x = (it == 0) ? unrank_colex(rank_begin + sidx, gm.N, gm.B, binom) : next_same_popcount(x);
                if (lane == 0) aw[0] = x;
            }
        } else if (it == 0 && lane == 0) {  // any valid state: all particles in mode 1
            for (int i = 0; i < gm.W; ++i) aw[i] = 0ull;
            for (int b = 0; b < gm.N; ++b) aw[b >> 6] |= 1ull << (b & 63);
        }
        __syncwarp();
        if (states_out && live && lane < gm.W) states_out[(sidx - first) * (uint64_t)gm.W + lane] = aw[lane];
        gen_decode<G>(gm, aw, ws, lane);
        gen_potential<G>(gm, ws, lane);
        long long reg, ext;
        double lgsum;
        gen_totals<G>(gm, ws, lane, reg, ext, lgsum);
        double S_lane, E_lane;
        int nocc_lane;
        GEN_MPL_SWITCH(gm.MPL, (gen_rates<MPL_, true>(gm, ws, lane, S_lane, E_lane, nocc_lane)));
        __syncwarp();
        const double Etot = warp_sum_all<G>(E_lane);
        const double D = gm.u * (double)reg + gm.v * (double)ext;
        const double eloc = D - gm.t * Etot;
        const double psi2 = exp(-2.0 * gen_phi<G>(gm, ws, lane, lgsum));
        const double s_num = psi2 * eloc;
        double wgt = 1.0;
        if (per_state) {
            if (terms_out && live && lane == 0) {
                terms_out[(sidx - first) * (uint64_t)(2 + 2 * P)] = s_num;
                terms_out[(sidx - first) * (uint64_t)(2 + 2 * P) + 1] = psi2;
            }
        } else if (live) {
            wgt = weight ? (double)weight[sidx] : 1.0;
            num = fma(wgt, s_num, num);
            den = fma(wgt, psi2, den);
        }
        if (with_grad) {
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int p = lane + G * q;
                if (p < P) {
                    const double Fp = gen_feature(gm, ws, p, reg, ext, lgsum);
                    double cross = 0.0;  // sum over hops of melem * ratio * (F_p(n') - F_p(n))
                    for (int k = 0; k < M; ++k) {
                        if (ws.occ[k] > 0) {
                            const int kr = k + 1 == M ? 0 : k + 1, kl = k == 0 ? M - 1 : k - 1;
                            cross = fma(ws.mr[2 * k], gen_dfeature(gm, ws, p, k, kr), cross);
                            cross = fma(ws.mr[2 * k + 1], gen_dfeature(gm, ws, p, k, kl), cross);
                        }
                    }
                    // localenergy.jl:109-118 with k_grad = -F psi_k, l_grad = -F(l) psi_l
                    const double s_ng = -psi2 * (2.0 * Fp * eloc + cross);
                    const double s_dg = -2.0 * Fp * psi2;
                    if (per_state) {
                        if (terms_out && live) {
                            terms_out[(sidx - first) * (uint64_t)(2 + 2 * P) + 2 + p] = s_ng;
                            terms_out[(sidx - first) * (uint64_t)(2 + 2 * P) + 2 + P + p] = s_dg;
                        }
                    } else if (live) {
                        ng[q] = fma(wgt, s_ng, ng[q]);
                        dg[q] = fma(wgt, s_dg, dg[q]);
                    }
                }
            }
        }
        __syncwarp();
    }
    if (per_state || partials == nullptr) return;
    double* out = partials + gwarp * (uint64_t)(2 + 2 * P);
    if (lane == 0) {
        out[0] = num;
        out[1] = den;
    }
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
        const int p = lane + G * q;
        if (p < P) {
            out[2 + p] = ng[q];
            out[2 + P + p] = dg[q];
        }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static size_t gen_smem(const GenModel& gm) { return (size_t)GEN_WARPS * GenWarp::bytes(gm.MP) + GEN_WARPS * 4 * sizeof(uint64_t); }

template <typename K>
static cudaError_t gen_prepare(K kernel, size_t smem) {
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

int gen_ppl(int P) {
    const int need = (P + 31) / 32;
    return need <= 1 ? 1 : (need <= 2 ? 2 : (need <= 4 ? 4 : (need <= 8 ? 8 : 0)));
}

// one resident wave, from the occupancy of the kernel that will run
template <typename K>
static cudaError_t gen_grid_for(K kernel, const GenModel& gm, size_t n, int sm_count, int* grid) {
    const size_t smem = gen_smem(gm);
    if (cudaError_t e = gen_prepare(kernel, smem)) return e;
    int per_sm = 0;
    if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, GEN_BLOCK, smem)) return e;
    if (per_sm < 1) per_sm = 1;
    const size_t resident = (size_t)per_sm * sm_count;
    const size_t need = (n + GEN_WARPS - 1) / GEN_WARPS;
    *grid = (int)(need < resident ? need : resident);
    if (*grid < 1) *grid = 1;
    return cudaSuccess;
}
// Lanes per walker: the smallest group with <= 4 modes and <= 2 parameters per lane, else <= 4 modes and <= 8
// parameters, else the whole warp.  Measured on B200 (tools/gen_bench.py): with 8 modes per lane the serial per-lane work
// outweighs the amortised scalar work ({50,50}: G=8 0.92e9, G=16 1.21e9, G=32 1.14e9 walker-steps/s), and many parameter
// slots per lane cost more than the group saves (DensityProfile {50,50}, P=50: G=16 0.99e9, G=32 1.14e9).
int gen_group(int M, int P) {
    for (int pass = 0; pass < 2; ++pass)
        for (int G = 8; G <= 32; G *= 2)
            if ((M + G - 1) / G <= 4 && (P + G - 1) / G <= (pass == 0 ? 2 : 8)) return G;
    return (M <= 256 && P <= 256) ? 32 : 0;
}
static int gen_slots(int need) { return need <= 1 ? 1 : (need <= 2 ? 2 : (need <= 4 ? 4 : 8)); }

// only the (G, PPL, MPL) combinations gen_group can produce are instantiated
template <int G, int PPL, int MPL>
constexpr bool gen_walker_combo_used() {
    return MPL <= 4 || G == 32;  // groups narrower than the warp are only chosen with <= 4 modes per lane
}

// GWZ_GEN_TI=0: re-evaluate every rate with exp() per step also for translation-invariant couplings (A/B, tests)
static bool gen_ti_disabled() {
    const char* e = std::getenv("GWZ_GEN_TI");
    return e && e[0] == '0';
}

struct GenWalkerOp {
    const GenWalkerLaunch* L;  // null: occupancy query
    const GenModel* gm;
    size_t n;
    int sm_count;
    int* grid;
    cudaStream_t s;
};
template <int G, int PPL, int MPL, bool TI>
static cudaError_t gen_walker_do(const GenWalkerOp& op) {
    if constexpr (!gen_walker_combo_used<G, PPL, MPL>()) {
        return cudaErrorInvalidValue;
    } else {
        const size_t smem = (size_t)(32 / G) * GenWarp::bytes(G * MPL);  // one slice per walker of the warp
        if (cudaError_t e = gen_prepare(gen_walker_kernel<G, PPL, MPL, TI>, smem)) return e;
        if (op.L == nullptr) {
            int per_sm = 0;
            if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gen_walker_kernel<G, PPL, MPL, TI>, GEN_BLOCK, smem))
                return e;
            if (per_sm < 1) per_sm = 1;
            const size_t resident = (size_t)per_sm * op.sm_count;
            const size_t need = (op.n + (32 / G) - 1) / (32 / G);
            *op.grid = (int)(need < resident ? need : resident);
            if (*op.grid < 1) *op.grid = 1;
            return cudaSuccess;
        }
        const GenWalkerLaunch& L = *op.L;
        gen_walker_kernel<G, PPL, MPL, TI><<<L.grid, GEN_BLOCK, smem, op.s>>>(L.gm, philox_keys(L.seed), L.occ, L.Mpad, L.acc, L.n,
                                                                         L.global_offset, L.step0, L.steps, L.accumulate, L.reset_acc,
                                                                         L.shift, L.partials, L.rec_tau, L.rec_eloc, L.rec_addr,
                                                                         L.err_flag);
        return cudaGetLastError();
    }
}
template <int G, int PPL>
static cudaError_t gen_walker_mpl(int mpl, const GenWalkerOp& op) {
    const bool ti = op.gm->circ != 0 && op.gm->M >= 4 && !gen_ti_disabled();
    switch (mpl) {
        case 1: return ti ? gen_walker_do<G, PPL, 1, true>(op) : gen_walker_do<G, PPL, 1, false>(op);
        case 2: return ti ? gen_walker_do<G, PPL, 2, true>(op) : gen_walker_do<G, PPL, 2, false>(op);
        case 4: return ti ? gen_walker_do<G, PPL, 4, true>(op) : gen_walker_do<G, PPL, 4, false>(op);
        case 8: return ti ? gen_walker_do<G, PPL, 8, true>(op) : gen_walker_do<G, PPL, 8, false>(op);
        default: return cudaErrorInvalidValue;
    }
}
template <int G>
static cudaError_t gen_walker_ppl(int ppl, int mpl, const GenWalkerOp& op) {
    switch (ppl) {
        case 1: return gen_walker_mpl<G, 1>(mpl, op);
        case 2: return gen_walker_mpl<G, 2>(mpl, op);
        case 4: return gen_walker_mpl<G, 4>(mpl, op);
        case 8: return gen_walker_mpl<G, 8>(mpl, op);
        default: return cudaErrorInvalidValue;
    }
}
static cudaError_t gen_walker_dispatch(const GenModel& gm, const GenWalkerOp& op) {
    const int G = gen_group(gm.M, gm.P);
    if (G == 0) return cudaErrorInvalidValue;
    const int ppl = gen_slots((gm.P + G - 1) / G), mpl = gen_slots((gm.M + G - 1) / G);
    switch (G) {
        case 8: return gen_walker_ppl<8>(ppl, mpl, op);
        case 16: return gen_walker_ppl<16>(ppl, mpl, op);
        default: return gen_walker_ppl<32>(ppl, mpl, op);
    }
}
cudaError_t gen_walker_grid(const GenModel& gm, size_t n, int sm_count, int* grid, int* walkers_per_cta) {
    GenWalkerOp op{nullptr, &gm, n, sm_count, grid, nullptr};
    *walkers_per_cta = 32 / gen_group(gm.M, gm.P);
    return gen_walker_dispatch(gm, op);
}
// per-group geometry of a sweep launch
static GenModel gen_sweep_model(const GenModel& gm, int G) {
    GenModel g = gm;
    g.MPL = gen_slots((gm.M + G - 1) / G);
    g.MP = G * g.MPL;
    return g;
}
struct GenSweepOp {
    const GenSweepLaunch* L;  // null: occupancy query
    const GenModel* gm;
    size_t n;
    int sm_count;
    int* grid;
    cudaStream_t s;
};
template <int G, int PPL>
static cudaError_t gen_sweep_do(const GenSweepOp& op) {
    const GenModel g = gen_sweep_model(*op.gm, G);
    const size_t smem = (size_t)(32 / G) * (GenWarp::bytes(g.MP) + 4 * sizeof(uint64_t));
    if (cudaError_t e = gen_prepare(gen_sweep_kernel<G, PPL>, smem)) return e;
    if (op.L == nullptr) {
        int per_sm = 0;
        if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gen_sweep_kernel<G, PPL>, GEN_BLOCK, smem)) return e;
        if (per_sm < 1) per_sm = 1;
        const size_t resident = (size_t)per_sm * op.sm_count;
        const size_t need = (op.n + (32 / G) - 1) / (32 / G);
        *op.grid = (int)(need < resident ? need : resident);
        if (*op.grid < 1) *op.grid = 1;
        return cudaSuccess;
    }
    const GenSweepLaunch& L = *op.L;
    gen_sweep_kernel<G, PPL><<<L.grid, GEN_BLOCK, smem, op.s>>>(g, L.basis, L.dim, L.rank_begin, L.binom, L.chunk, L.with_grad,
                                                               L.partials, L.terms_out, L.states_out, L.first, L.count, L.weight);
    return cudaGetLastError();
}
template <int G>
static cudaError_t gen_sweep_ppl(int ppl, const GenSweepOp& op) {
    switch (ppl) {
        case 1: return gen_sweep_do<G, 1>(op);
        case 2: return gen_sweep_do<G, 2>(op);
        case 4: return gen_sweep_do<G, 4>(op);
        case 8: return gen_sweep_do<G, 8>(op);
        default: return cudaErrorInvalidValue;
    }
}
static cudaError_t gen_sweep_dispatch(const GenModel& gm, const GenSweepOp& op) {
    const int G = gen_group(gm.M, gm.P);
    if (G == 0) return cudaErrorInvalidValue;
    const int ppl = gen_slots((gm.P + G - 1) / G);
    switch (G) {
        case 8: return gen_sweep_ppl<8>(ppl, op);
        case 16: return gen_sweep_ppl<16>(ppl, op);
        default: return gen_sweep_ppl<32>(ppl, op);
    }
}
cudaError_t gen_sweep_grid(const GenModel& gm, size_t n, int sm_count, int* grid, int* states_per_cta) {
    GenSweepOp op{nullptr, &gm, n, sm_count, grid, nullptr};
    *states_per_cta = 32 / gen_group(gm.M, gm.P);
    return gen_sweep_dispatch(gm, op);
}

cudaError_t launch_gen_unary_to_occ(const uint64_t* addr, size_t n, const GenModel& gm, int Mpad, uint8_t* occ,
                                    bool broadcast, cudaStream_t s) {
    gen_unary_to_occ_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(addr, n, gm.W, gm.B, gm.M, Mpad, occ, broadcast ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t launch_gen_occ_to_unary(const uint8_t* occ, size_t n, const GenModel& gm, int Mpad, uint64_t* addr,
                                    cudaStream_t s) {
    gen_occ_to_unary_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(occ, n, gm.W, gm.M, Mpad, addr);
    return cudaGetLastError();
}

cudaError_t launch_gen_probe(const GenProbeLaunch& L, cudaStream_t s) {
    const size_t smem = gen_smem(L.gm);
    if (cudaError_t e = gen_prepare(gen_probe_kernel, smem)) return e;
    size_t grid = (L.n + GEN_WARPS - 1) / GEN_WARPS;
    if (grid > 65535u * 16u) grid = 65535u * 16u;
    if (grid < 1) grid = 1;
    gen_probe_kernel<<<(unsigned)grid, GEN_BLOCK, smem, s>>>(L.gm, L.addr, L.n, L.u01, L.val, L.grad, L.e_loc, L.tau,
                                                            L.grad_ratio, L.n_hops, L.rates, L.chosen, L.next_addr, L.err_flag);
    return cudaGetLastError();
}

cudaError_t launch_gen_walkers(const GenWalkerLaunch& L, cudaStream_t s) {
    GenWalkerOp op{&L, &L.gm, L.n, 0, nullptr, s};
    return gen_walker_dispatch(L.gm, op);
}
cudaError_t launch_gen_estimates(const double* acc, size_t n, int P, const double* shift, double* val_w, double* grad_w,
                                 cudaStream_t s) {
    gen_estimates_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(acc, n, P, shift, val_w, grad_w);
    return cudaGetLastError();
}

cudaError_t launch_gen_sweep(const GenSweepLaunch& L, cudaStream_t s) {
    GenSweepOp op{&L, &L.gm, 0, 0, nullptr, s};
    return gen_sweep_dispatch(L.gm, op);
}

}  // namespace gwz
