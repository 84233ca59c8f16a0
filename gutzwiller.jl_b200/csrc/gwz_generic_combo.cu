// gwz_generic_combo.cu -- CombinationAnsatz (src/ansatz/combination.jl:1-51) on the ansatz-generic engine:
//     psi = alpha psi_1 + (1 - alpha) psi_2,   grad = (alpha grad psi_1, (1 - alpha) grad psi_2, psi_1 - psi_2),
// both components log-linear (gwz_generic.cuh).  With lr = phi_1 - phi_2 (so psi_2/psi_1 = e^lr) and the scaled weights
//     a1 = alpha e^-m,  a2 = (1 - alpha) e^(lr - m),  m = max(0, lr),  den = a1 + a2        (psi = e^(m - phi_1) den)
// a hop with component exponents e1, e2 has psi'/psi = (a1 e^-e1 + a2 e^-e2) / den, its lr' = lr + e1 - e2, and
//     grad psi / psi = (-F_1 a1/den, -F_2 a2/den, (e^-m - e^(lr-m)) / den).
// Same warp-per-walker organisation, hop order and selection rule as gwz_generic.cu; not tuned beyond that.
#include "gwz_generic.cuh"
#include "gwz_internal.h"

namespace gwz {

struct ComboW {
    double a1, a2, den, ga;  // ga = d psi / d alpha / psi
};
__device__ __forceinline__ ComboW combo_weights(double alpha, double lr) {
    const double m = fmax(0.0, lr);
    const double em = exp(-m), el = exp(lr - m);
    ComboW w;
    w.a1 = alpha * em;
    w.a2 = (1.0 - alpha) * el;
    w.den = w.a1 + w.a2;
    w.ga = (em - el) / w.den;
    return w;
}

// exponents of the right / left hop out of an occupied mode k for one component
__device__ __forceinline__ void comp_exponents(const GenModel& g, const GenWarp& w, int k, int n, int kr, int kl, int nr,
                                               int nl, double& er, double& el) {
    const double vk = w.V[k];
    er = w.V[kr] - vk + g.cR[k];
    el = w.V[kl] - vk + g.cL[k];
    if (g.s != 0.0) {
        const double ln = g.logtab[n];
        er += g.s * (g.logtab[nr + 1] - ln);
        el += g.s * (g.logtab[nl + 1] - ln);
    }
}

// rates |psi'/psi|, signed matrix-element-weighted ratios and the per-hop exponents of the modes this lane owns
__device__ __forceinline__ void combo_rates(const GenCombo& cb, const GenWarp& ws, const GenWarp& ws2, const ComboW& cw,
                                            int lane, double& S_lane, double& E_lane, int& nocc_lane) {
    const GenModel& g1 = cb.g1;
    const int M = g1.M;
    double S = 0.0, E = 0.0;
    int nocc = 0;
    for (int q = 0; q < g1.MPL; ++q) {
        const int k = lane * g1.MPL + q;
        if (k < M) {
            double rr = 0.0, rl = 0.0, mrr = 0.0, mrl = 0.0, e1r = 0.0, e1l = 0.0, e2r = 0.0, e2l = 0.0;
            const int n = ws.occ[k];
            if (n > 0) {
                const int kr = k + 1 == M ? 0 : k + 1, kl = k == 0 ? M - 1 : k - 1;
                const int nr = ws.occ[kr], nl = ws.occ[kl];
                comp_exponents(g1, ws, k, n, kr, kl, nr, nl, e1r, e1l);
                comp_exponents(cb.g2, ws2, k, n, kr, kl, nr, nl, e2r, e2l);
                const double qr = (cw.a1 * exp(-e1r) + cw.a2 * exp(-e2r)) / cw.den;  // psi'/psi, signed
                const double ql = (cw.a1 * exp(-e1l) + cw.a2 * exp(-e2l)) / cw.den;
                rr = fabs(qr);  // qmc.jl:27 abs(val2)
                rl = fabs(ql);
                const double sn = g1.sqrtab[n];
                mrr = sn * g1.sqrtab[nr + 1] * qr;
                mrl = sn * g1.sqrtab[nl + 1] * ql;
                S += rr;
                S += rl;
                E += mrr;
                E += mrl;
                ++nocc;
            }
            ws.rate[2 * k] = rr;
            ws.rate[2 * k + 1] = rl;
            ws.mr[2 * k] = -g1.t * mrr;
            ws.mr[2 * k + 1] = -g1.t * mrl;
            ws.e1[2 * k] = e1r;
            ws.e1[2 * k + 1] = e1l;
            ws.e2[2 * k] = e2r;
            ws.e2[2 * k + 1] = e2l;
        }
    }
    S_lane = S;
    E_lane = E;
    nocc_lane = nocc;
}

// component and local index of parameter p; comp = 2 for alpha
__device__ __forceinline__ int combo_split(const GenCombo& cb, int p, int& local) {
    if (p < cb.g1.P) {
        local = p;
        return 0;
    }
    if (p < cb.g1.P + cb.g2.P) {
        local = p - cb.g1.P;
        return 1;
    }
    local = 0;
    return 2;
}
__device__ __forceinline__ double combo_grad_ratio(int comp, double F, const ComboW& cw) {
    return comp == 0 ? -F * (cw.a1 / cw.den) : (comp == 1 ? -F * (cw.a2 / cw.den) : cw.ga);
}
__device__ __forceinline__ double combo_feature(const GenCombo& cb, const GenWarp& ws, const GenWarp& ws2, int comp, int local,
                                                long long reg, long long ext, double lgsum) {
    return comp == 0 ? gen_feature(cb.g1, ws, local, reg, ext, lgsum)
                     : (comp == 1 ? gen_feature(cb.g2, ws2, local, reg, ext, lgsum) : 0.0);
}
__device__ __forceinline__ double combo_dfeature(const GenCombo& cb, const GenWarp& ws, const GenWarp& ws2, int comp,
                                                 int local, int i, int j) {
    return comp == 0 ? gen_dfeature(cb.g1, ws, local, i, j) : (comp == 1 ? gen_dfeature(cb.g2, ws2, local, i, j) : 0.0);
}

// from-scratch state: potentials, totals, lr = phi_1 - phi_2, log|psi_1| scale
__device__ __forceinline__ void combo_state(const GenCombo& cb, const GenWarp& ws, const GenWarp& ws2, int lane,
                                            long long& reg, long long& ext, double& lgsum, double& phi1, double& lr) {
    gen_potential(cb.g1, ws, lane);
    gen_potential(cb.g2, ws2, lane);
    gen_totals(cb.g1, ws, lane, reg, ext, lgsum);
    phi1 = gen_phi(cb.g1, ws, lane, lgsum);
    lr = phi1 - gen_phi(cb.g2, ws2, lane, lgsum);
}

#define GEN_MPL_SWITCH(mpl, CALL)       \
    switch (mpl) {                      \
        case 1: { constexpr int MPL_ = 1; CALL; } break; \
        case 2: { constexpr int MPL_ = 2; CALL; } break; \
        case 4: { constexpr int MPL_ = 4; CALL; } break; \
        default: { constexpr int MPL_ = 8; CALL; } break; \
    }

// ------------------------------------------------------------------------------------------
// probe / evaluation
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEN_BLOCK)
combo_probe_kernel(const __grid_constant__ GenCombo cb, const uint64_t* __restrict__ addr, size_t n,
                   const double* __restrict__ u01, double* __restrict__ val_out, double* __restrict__ grad_out,
                   double* __restrict__ e_loc, double* __restrict__ tau_out, double* __restrict__ grad_ratio,
                   int* __restrict__ n_hops, double* __restrict__ rates, int* __restrict__ chosen,
                   uint64_t* __restrict__ next_addr, int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const GenModel& gm = cb.g1;
    const GenWarp ws(smem_raw, gm.MP, GenWarp::ComboTag{});
    const GenWarp ws2 = ws.second();
    const int M = gm.M, P = cb.P;
    for (size_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
        gen_decode(gm, addr + idx * (size_t)gm.W, ws, lane);
        long long reg, ext;
        double lgsum, phi1, lr;
        combo_state(cb, ws, ws2, lane, reg, ext, lgsum, phi1, lr);
        const ComboW cw = combo_weights(cb.alpha, lr);
        const double m = fmax(0.0, lr);
        const double val = exp(m - phi1) * cw.den;
        if (val_out && lane == 0) val_out[idx] = val;
        if (grad_out || grad_ratio)
            for (int p = lane; p < P; p += 32) {
                int local;
                const int comp = combo_split(cb, p, local);
                const double G = combo_grad_ratio(comp, combo_feature(cb, ws, ws2, comp, local, reg, ext, lgsum), cw);
                if (grad_out) grad_out[idx * (size_t)P + p] = G * val;
                if (grad_ratio) grad_ratio[idx * (size_t)P + p] = G;
            }
        if (!(e_loc || tau_out || n_hops || rates || chosen || next_addr)) continue;
        double S_lane, E_lane;
        int nocc_lane;
        combo_rates(cb, ws, ws2, cw, lane, S_lane, E_lane, nocc_lane);
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
        if (rates) {
            double* row = rates + idx * (size_t)(2 * M);
            for (int h = lane; h < 2 * M; h += 32) row[h] = 0.0;
            __syncwarp();
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
// walkers
// ------------------------------------------------------------------------------------------
template <int PPL>
__global__ void __launch_bounds__(GEN_BLOCK)
combo_walker_kernel(const __grid_constant__ GenCombo cb, const __grid_constant__ PhiloxKeys pk, uint8_t* __restrict__ occ_g,
                    int Mpad, double* __restrict__ acc_g, size_t n, uint64_t goff, uint64_t step0, uint64_t steps,
                    int accumulate, int reset_acc, const double* __restrict__ shift, double* __restrict__ partials,
                    double* __restrict__ rec_tau, double* __restrict__ rec_eloc, uint64_t* __restrict__ rec_addr,
                    int* __restrict__ err_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const GenModel& gm = cb.g1;
    const GenWarp ws(smem_raw, gm.MP, GenWarp::ComboTag{});
    const GenWarp ws2 = ws.second();
    const size_t gwarp = blockIdx.x, nwarps = gridDim.x;
    const int M = gm.M, P = cb.P;
    const int AW = 3 + 2 * P;
    const double E0 = (accumulate && shift) ? shift[0] : 0.0;
    double G0[PPL];
    int comp[PPL], local[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
        const int p = lane + 32 * q;
        G0[q] = (accumulate && shift && p < P) ? shift[1 + p] : 0.0;
        comp[q] = (p < P) ? combo_split(cb, p, local[q]) : 3;
        if (p >= P) local[q] = 0;
    }
    double ps[GEN_HEAD];
    double pv[4][PPL];
#pragma unroll
    for (int i = 0; i < GEN_HEAD; ++i) ps[i] = 0.0;
#pragma unroll
    for (int q = 0; q < PPL; ++q) pv[0][q] = pv[1][q] = pv[2][q] = pv[3][q] = 0.0;
    bool bad = false;

    for (size_t w = gwarp; w < n; w += nwarps) {
        for (int k = lane; k < gm.MP; k += 32) ws.occ[k] = (k < M) ? (int)occ_g[w * (size_t)Mpad + k] : 0;
        __syncwarp();
        long long reg, ext;
        double lgsum, phi1, lr;
        combo_state(cb, ws, ws2, lane, reg, ext, lgsum, phi1, lr);
        double F[PPL], aG[PPL], aGE[PPL];
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
            F[q] = (comp[q] < 2) ? combo_feature(cb, ws, ws2, comp[q], local[q], reg, ext, lgsum) : 0.0;
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
                const int p = lane + 32 * q;
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

        for (uint64_t k = 0; k < steps; ++k) {
            const uint64_t step = step0 + k;
            if (k != 0 && (k % GEN_REFRESH) == 0) {
                combo_state(cb, ws, ws2, lane, reg, ext, lgsum, phi1, lr);
#pragma unroll
                for (int q = 0; q < PPL; ++q)
                    if (comp[q] < 2) F[q] = combo_feature(cb, ws, ws2, comp[q], local[q], reg, ext, lgsum);
            }
            // the 32 lanes evaluate 32 consecutive Philox blocks at once (64 steps); see gen_walker_kernel
            const uint64_t blk = step >> 1;
            if (k == 0 || ((step & 1ull) == 0ull && (blk & 31ull) == 0ull)) {
                const uint64_t mine = (blk & ~31ull) + (uint64_t)lane;
                philox4x32_10_keyed((uint32_t)mine, (uint32_t)(mine >> 32), c2, c3, pk, rnd);
            }
            const int rsrc = (int)(blk & 31ull);
            const uint32_t wlo = __shfl_sync(0xffffffffu, (step & 1ull) ? rnd[2] : rnd[0], rsrc);
            const uint32_t whi = __shfl_sync(0xffffffffu, (step & 1ull) ? rnd[3] : rnd[1], rsrc);
            const double u01 = u01_from_words(wlo, whi);
            const ComboW cw = combo_weights(cb.alpha, lr);
            double S_lane, E_lane;
            int nocc_lane;
            combo_rates(cb, ws, ws2, cw, lane, S_lane, E_lane, nocc_lane);
            const double cum = warp_scan_incl(S_lane, lane);
            const double S = __shfl_sync(0xffffffffu, cum, 31);
            const double Etot = warp_sum_all(E_lane);
            const int nocc = warp_sum_all(nocc_lane);
            const double tau = 1.0 / S;
            const double eloc = gm.u * (double)reg + gm.v * (double)ext - gm.t * Etot;
            bad |= !(tau < 1.0e300);
            if (rec_tau && lane == 0) {
                rec_tau[k * n + w] = tau;
                rec_eloc[k * n + w] = eloc;
            }
            if (rec_addr) gen_encode(gm, ws, lane, rec_addr + (k * n + w) * (size_t)gm.W);
            const double dE = eloc - E0;
            const double tE = tau * dE;
            st += tau;
            ste += tE;
            stee = fma(tE, dE, stee);
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const double dG = combo_grad_ratio(comp[q] < 3 ? comp[q] : 0, F[q], cw) - G0[q];
                aG[q] = fma(tau, dG, aG[q]);
                aGE[q] = fma(tE, dG, aGE[q]);
            }
            hops += (unsigned)(2 * nocc);
            double T = u01 * S;
            if (!(T < S)) T = __longlong_as_double(__double_as_longlong(S) - 1);
            int src, right;
            GEN_MPL_SWITCH(gm.MPL, (gen_select<MPL_>(gm, ws, lane, S_lane, cum, T, src, right)));
            const int dst = right ? (src + 1 == M ? 0 : src + 1) : (src == 0 ? M - 1 : src - 1);
            __syncwarp();  // the owner's exponents are visible
            const int h = 2 * src + (right ? 0 : 1);
            lr += ws.e1[h] - ws.e2[h];
#pragma unroll
            for (int q = 0; q < PPL; ++q)
                if (comp[q] < 2) F[q] += combo_dfeature(cb, ws, ws2, comp[q], local[q], src, dst);
            reg += ws.occ[dst] - ws.occ[src] + 1;
            ext += gen_dext(gm, ws, src, dst);
            __syncwarp();
            for (int mo = lane; mo < M; mo += 32) {
                ws.V[mo] += __ldg(cb.g1.Qt + (size_t)dst * M + mo) - __ldg(cb.g1.Qt + (size_t)src * M + mo);
                ws.V2[mo] += __ldg(cb.g2.Qt + (size_t)dst * M + mo) - __ldg(cb.g2.Qt + (size_t)src * M + mo);
            }
            if (lane == 0) {
                ws.occ[src] -= 1;
                ws.occ[dst] += 1;
            }
            __syncwarp();
        }

        for (int k = lane; k < M; k += 32) occ_g[w * (size_t)Mpad + k] = (uint8_t)ws.occ[k];
        if (accumulate) {
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
                const int p = lane + 32 * q;
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
        const int p = lane + 32 * q;
        if (p < P) {
            out[GEN_HEAD + p] = pv[0][q];
            out[GEN_HEAD + P + p] = pv[1][q];
            out[GEN_HEAD + 2 * P + p] = pv[2][q];
            out[GEN_HEAD + 3 * P + p] = pv[3][q];
        }
    }
}

// ------------------------------------------------------------------------------------------
// LocalEnergyEvaluator
// ------------------------------------------------------------------------------------------
template <int PPL>
__global__ void __launch_bounds__(GEN_BLOCK)
combo_sweep_kernel(const __grid_constant__ GenCombo cb, const uint64_t* __restrict__ basis, uint64_t dim,
                   uint64_t rank_begin, const uint64_t* __restrict__ binom, uint64_t chunk, int with_grad,
                   double* __restrict__ partials, double* __restrict__ terms_out, uint64_t* __restrict__ states_out,
                   uint64_t first, uint64_t count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const GenModel& gm = cb.g1;
    const GenWarp ws(smem_raw, gm.MP, GenWarp::ComboTag{});
    const GenWarp ws2 = ws.second();
    uint64_t* aw = reinterpret_cast<uint64_t*>(smem_raw + GenWarp::bytes_combo(gm.MP));
    const uint64_t gwarp = blockIdx.x;
    const int M = gm.M, P = cb.P;
    const bool per_state = terms_out != nullptr || states_out != nullptr;
    const uint64_t lo = per_state ? first + gwarp * chunk : gwarp * chunk;
    const uint64_t end = per_state ? first + count : dim;
    const uint64_t hi = (lo + chunk < end) ? lo + chunk : end;
    double num = 0.0, den = 0.0, ng[PPL], dg[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) ng[q] = dg[q] = 0.0;
    uint64_t x = 0;
    for (uint64_t sidx = lo; sidx < hi; ++sidx) {
        if (basis) {
            if (lane < gm.W) aw[lane] = basis[(uint64_t)lane * dim + sidx];
        } else {
            x = (sidx == lo) ? unrank_colex(rank_begin + sidx, gm.N, gm.B, binom) : next_same_popcount(x);
            if (lane == 0) aw[0] = x;
        }
        __syncwarp();
        if (states_out && lane < gm.W) states_out[(sidx - first) * (uint64_t)gm.W + lane] = aw[lane];
        gen_decode(gm, aw, ws, lane);
        long long reg, ext;
        double lgsum, phi1, lr;
        combo_state(cb, ws, ws2, lane, reg, ext, lgsum, phi1, lr);
        const ComboW cw = combo_weights(cb.alpha, lr);
        double S_lane, E_lane;
        int nocc_lane;
        combo_rates(cb, ws, ws2, cw, lane, S_lane, E_lane, nocc_lane);
        __syncwarp();
        const double Etot = warp_sum_all(E_lane);
        const double D = gm.u * (double)reg + gm.v * (double)ext;
        const double eloc = D - gm.t * Etot;
        const double psi = exp(fmax(0.0, lr) - phi1) * cw.den;
        const double psi2 = psi * psi;
        if (per_state) {
            if (terms_out && lane == 0) {
                terms_out[(sidx - first) * (uint64_t)(2 + 2 * P)] = psi2 * eloc;
                terms_out[(sidx - first) * (uint64_t)(2 + 2 * P) + 1] = psi2;
            }
        } else {
            num += psi2 * eloc;
            den += psi2;
        }
        if (with_grad) {
            // per-hop weights of the neighbour: overwrite the exponents with (a1/den, a2/den, ga) of lr' = lr + e1 - e2
            for (int h = lane; h < 2 * M; h += 32) {
                const ComboW nw = combo_weights(cb.alpha, lr + ws.e1[h] - ws.e2[h]);
                ws.e1[h] = nw.a1 / nw.den;
                ws.e2[h] = nw.a2 / nw.den;
                ws.ga[h] = nw.ga;
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int p = lane + 32 * q;
                if (p < P) {
                    int local;
                    const int comp = combo_split(cb, p, local);
                    const double Fk = combo_feature(cb, ws, ws2, comp, local, reg, ext, lgsum);
                    const double Gk = combo_grad_ratio(comp, Fk, cw);
                    double cross = 0.0;  // sum over hops of melem * ratio * (G_k + G_l)
                    for (int k = 0; k < M; ++k) {
                        if (ws.occ[k] > 0) {
                            const int kr = k + 1 == M ? 0 : k + 1, kl = k == 0 ? M - 1 : k - 1;
#pragma unroll
                            for (int dir = 0; dir < 2; ++dir) {
                                const int h = 2 * k + dir;
                                const double Fl = Fk + combo_dfeature(cb, ws, ws2, comp, local, k, dir == 0 ? kr : kl);
                                const double Gl = comp == 0 ? -Fl * ws.e1[h] : (comp == 1 ? -Fl * ws.e2[h] : ws.ga[h]);
                                cross = fma(ws.mr[h], Gk + Gl, cross);
                            }
                        }
                    }
                    // localenergy.jl:109-118 with k_grad = G_k psi_k, l_grad = G_l psi_l
                    const double s_ng = psi2 * (2.0 * Gk * D + cross);
                    const double s_dg = 2.0 * Gk * psi2;
                    if (per_state) {
                        if (terms_out) {
                            terms_out[(sidx - first) * (uint64_t)(2 + 2 * P) + 2 + p] = s_ng;
                            terms_out[(sidx - first) * (uint64_t)(2 + 2 * P) + 2 + P + p] = s_dg;
                        }
                    } else {
                        ng[q] += s_ng;
                        dg[q] += s_dg;
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
        const int p = lane + 32 * q;
        if (p < P) {
            out[2 + p] = ng[q];
            out[2 + P + p] = dg[q];
        }
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static size_t combo_smem(const GenCombo& cb) { return GenWarp::bytes_combo(cb.g1.MP) + 4 * sizeof(uint64_t); }
template <typename K>
static cudaError_t combo_prepare(K kernel, size_t smem) {
    if (smem <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
template <typename K>
static cudaError_t combo_grid_for(K kernel, const GenCombo& cb, size_t n, int sm_count, int* grid) {
    const size_t smem = combo_smem(cb);
    if (cudaError_t e = combo_prepare(kernel, smem)) return e;
    int per_sm = 0;
    if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, GEN_BLOCK, smem)) return e;
    if (per_sm < 1) per_sm = 1;
    const size_t resident = (size_t)per_sm * sm_count;
    *grid = (int)(n < resident ? n : resident);
    if (*grid < 1) *grid = 1;
    return cudaSuccess;
}
cudaError_t combo_walker_grid(const GenCombo& cb, size_t n, int sm_count, int* grid) {
    switch (gen_ppl(cb.P)) {
        case 1: return combo_grid_for(combo_walker_kernel<1>, cb, n, sm_count, grid);
        case 2: return combo_grid_for(combo_walker_kernel<2>, cb, n, sm_count, grid);
        case 4: return combo_grid_for(combo_walker_kernel<4>, cb, n, sm_count, grid);
        case 8: return combo_grid_for(combo_walker_kernel<8>, cb, n, sm_count, grid);
        default: return cudaErrorInvalidValue;
    }
}
cudaError_t combo_sweep_grid(const GenCombo& cb, size_t n, int sm_count, int* grid) {
    switch (gen_ppl(cb.P)) {
        case 1: return combo_grid_for(combo_sweep_kernel<1>, cb, n, sm_count, grid);
        case 2: return combo_grid_for(combo_sweep_kernel<2>, cb, n, sm_count, grid);
        case 4: return combo_grid_for(combo_sweep_kernel<4>, cb, n, sm_count, grid);
        case 8: return combo_grid_for(combo_sweep_kernel<8>, cb, n, sm_count, grid);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_combo_probe(const GenProbeLaunch& L, const GenCombo& cb, cudaStream_t s) {
    const size_t smem = combo_smem(cb);
    if (cudaError_t e = combo_prepare(combo_probe_kernel, smem)) return e;
    size_t grid = L.n;
    if (grid > 65535u * 16u) grid = 65535u * 16u;
    if (grid < 1) grid = 1;
    combo_probe_kernel<<<(unsigned)grid, GEN_BLOCK, smem, s>>>(cb, L.addr, L.n, L.u01, L.val, L.grad, L.e_loc, L.tau,
                                                              L.grad_ratio, L.n_hops, L.rates, L.chosen, L.next_addr, L.err_flag);
    return cudaGetLastError();
}
template <int PPL>
static cudaError_t combo_walker_launch(const GenWalkerLaunch& L, const GenCombo& cb, cudaStream_t s) {
    const size_t smem = combo_smem(cb);
    if (cudaError_t e = combo_prepare(combo_walker_kernel<PPL>, smem)) return e;
    combo_walker_kernel<PPL><<<L.grid, GEN_BLOCK, smem, s>>>(cb, philox_keys(L.seed), L.occ, L.Mpad, L.acc, L.n, L.global_offset,
                                                           L.step0, L.steps, L.accumulate, L.reset_acc, L.shift, L.partials,
                                                           L.rec_tau, L.rec_eloc, L.rec_addr, L.err_flag);
    return cudaGetLastError();
}
cudaError_t launch_combo_walkers(const GenWalkerLaunch& L, const GenCombo& cb, cudaStream_t s) {
    switch (gen_ppl(cb.P)) {
        case 1: return combo_walker_launch<1>(L, cb, s);
        case 2: return combo_walker_launch<2>(L, cb, s);
        case 4: return combo_walker_launch<4>(L, cb, s);
        case 8: return combo_walker_launch<8>(L, cb, s);
        default: return cudaErrorInvalidValue;
    }
}
template <int PPL>
static cudaError_t combo_sweep_launch(const GenSweepLaunch& L, const GenCombo& cb, cudaStream_t s) {
    const size_t smem = combo_smem(cb);
    if (cudaError_t e = combo_prepare(combo_sweep_kernel<PPL>, smem)) return e;
    combo_sweep_kernel<PPL><<<L.grid, GEN_BLOCK, smem, s>>>(cb, L.basis, L.dim, L.rank_begin, L.binom, L.chunk, L.with_grad,
                                                          L.partials, L.terms_out, L.states_out, L.first, L.count);
    return cudaGetLastError();
}
cudaError_t launch_combo_sweep(const GenSweepLaunch& L, const GenCombo& cb, cudaStream_t s) {
    switch (gen_ppl(cb.P)) {
        case 1: return combo_sweep_launch<1>(L, cb, s);
        case 2: return combo_sweep_launch<2>(L, cb, s);
        case 4: return combo_sweep_launch<4>(L, cb, s);
        case 8: return combo_sweep_launch<8>(L, cb, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace gwz
