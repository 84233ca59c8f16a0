// gwz_sparse.cu -- the hot path over a STORED OPERATOR: any Hamiltonian, any ansatz.
//
// The reference's kinetic_sample! (src/kinetic-vqmc/qmc.jl:14-44) and LocalEnergyEvaluator (src/localenergy.jl:94-155)
// are generic in the Hamiltonian: they only call offdiagonals(H, k), diagonal_element(H, k) and ansatz(k, p)
// (test/qmc.jl:10-14 runs them on HubbardReal1D, HubbardMom1D and a two-component HubbardRealSpace).  The bit-level
// kernels of this library cover the Hubbard chain; every other operator reaches the GPU materialised over its
// basis: build_basis(H) enumerates the addresses, row k holds offdiagonals(H, basis[k]) in the reference's order as
// (column index, matrix element), diag[k] = diagonal_element, and the ansatz is a table psi[k] with
// grad_ratio[k][p] = (d psi/d p_p)/psi (uploaded, or filled on the device for GutzwillerAnsatz, gutzwiller.jl:25-36).
//
// HBM layout: row_ptr u64[dim+1], col u32[nnz], melem f64[nnz], rpsi f64[nnz] (the amplitude of every row entry's
// target, gathered once per table so that a walker step STREAMS its row -- 20 B per hop, no gather), diag f64[dim],
// amplitudes f64[dim], a 64 B record per state (row bounds, diag, psi, gradient ratios) and the ansatz table
// f64[dim][1+P] (psi, then grad_ratio: one state's data in one sector, what the gradient sweep gathers).  When the
// rows are regular enough each state also gets a ROW BLOCK at a fixed stride (head + the row in chunks of G entries,
// col | melem | rpsi, 128 B aligned): the block address is k * stride, so a walker step is one dependent HBM access
// (index -> block) instead of three (index -> record -> row).
// Walkers: index u32[n], accumulators f64[4+2P][n].  One walker per GROUP of G = 4/8/16/32 lanes (G from the mean row
// length): the lanes read a row chunk by chunk (coalesced), an inclusive scan inside each chunk gives the reference's
// cumulative buffer in row order and the move is the first entry with u*S < cumsum (pick_random_from_cumsum,
// qmc.jl:52-61); the first R chunks stay in registers for the pick, longer rows are re-read.  Bound: HBM (random
// 128 B lines) and issue about evenly; occupancy (64 registers, 8 CTAs/SM) and bytes per hop are what matter.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "gwz_generic.cuh"
#include "gwz_handles.h"

using namespace gwz;

namespace gwz {

struct SparseDev {
    const unsigned long long* row_ptr;
    const uint32_t* col;
    const double* melem;
    const double* diag;
    const double* rpsi;  // [nnz] psi of each row entry's target, gathered once per table (sparse_expand_kernel): a walker
                         // step streams (col, melem, rpsi) of its row instead of gathering an amplitude per hop
    const double* rec;  // [dim][RS] what a walker needs of its OWN state in one aligned 64 B block (RS = 8 for P <= 4):
                        // row begin, row end (as bits), diag, psi, grad_ratio[0..P)
    int RS;
    const double* blk;  // [dim][BS] row BLOCKS (null: rows too ragged to store at a fixed stride): per state one 128 B
                        // aligned block = head (row length, diag, psi, grad_ratio[0..min(P,5))) + the row in chunks of G
                        // entries (col u32[G], melem[G], rpsi[G]), zero padded.  Block address = k * BS: a walker
                        // step is ONE dependent HBM access (index -> block) instead of index -> record -> row.
    int BS;
    const double* amp;  // [dim] psi_k alone (the expansion and the value-only sweep gather from it)
    const double* tab;  // [dim][1 + P]: psi_k, then grad_ratio[k][0..P) -- one state's ansatz data shares a sector
    unsigned long long dim;
    int P;
    __device__ __forceinline__ double psi(unsigned long long k) const { return amp[k]; }
    __device__ __forceinline__ double psi_tab(unsigned long long k) const { return tab[k * (unsigned)(1 + P)]; }
    __device__ __forceinline__ double gratio(unsigned long long k, int p) const { return tab[k * (unsigned)(1 + P) + 1 + p]; }
};

#ifndef GWZ_SP_MINB
#define GWZ_SP_MINB 8
#endif
#ifndef GWZ_SP_SWEEP_MINB
#define GWZ_SP_SWEEP_MINB 8
#endif
#ifndef GWZ_SP_R
#define GWZ_SP_R 2
#endif
constexpr int SP_BLOCK = 128;
constexpr int SP_R = GWZ_SP_R;  // row chunks whose cumulative sums stay in registers

template <int G>
__device__ __forceinline__ unsigned sp_mask() {
    if (G == 32) return 0xffffffffu;
    return ((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1));
}
template <int G>
__device__ __forceinline__ double sp_scan(double v, int lane, unsigned mask) {
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const double t = __shfl_up_sync(mask, v, o, G);
        if (lane >= o) v += t;
    }
    return v;
}
template <int G>
__device__ __forceinline__ double sp_sum(double v, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
    return v;
}

struct SpStep {
    double tau, eloc, S;
    int nhops, chosen;  // chosen: 0-based position in the row, -1 = none (non-finite sums)
    uint32_t next;
};

// kinetic_sample! at index k for one group of G lanes
template <int G, bool RATES>
__device__ __forceinline__ SpStep sp_step(const SparseDev& sd, uint32_t k, int lane, unsigned mask, double u01,
                                          double* rates /* RATES: |psi_l / psi_k| in row order */) {
    const double* rec = sd.rec + (size_t)k * sd.RS;
    const ulonglong2 be = *reinterpret_cast<const ulonglong2*>(rec);
    const double2 dv = *reinterpret_cast<const double2*>(rec + 2);
    const unsigned long long b = be.x, e = be.y;
    const double v1 = dv.y;
    double num_l = 0.0, base = 0.0;
    double cc[SP_R];
    uint32_t cn[SP_R];
#pragma unroll
    for (int c = 0; c < SP_R; ++c) {
        cc[c] = 0.0;
        cn[c] = 0u;
        if (b + (unsigned long long)(c * G) < e) {  // uniform in the group
            const unsigned long long i = b + (unsigned long long)(c * G + lane);
            double r = 0.0;
            if (i < e) {
                cn[c] = sd.col[i];
                const double v2 = sd.rpsi[i];
                r = fabs(v2);                       // qmc.jl:27
                num_l = fma(sd.melem[i], v2, num_l);  // qmc.jl:28
                if (RATES) rates[i - b] = r / fabs(v1);
            }
            const double sc = sp_scan<G>(r, lane, mask);
            cc[c] = base + sc;  // qmc.jl:30: the cumulative buffer, in row order
            base += __shfl_sync(mask, sc, G - 1, G);
        }
    }
    const double baseR = base;
    for (unsigned long long i0 = b + (unsigned long long)(SP_R * G); i0 < e; i0 += G) {
        const unsigned long long i = i0 + lane;
        double r = 0.0;
        if (i < e) {
            const double v2 = sd.rpsi[i];
            r = fabs(v2);
            num_l = fma(sd.melem[i], v2, num_l);
            if (RATES) rates[i - b] = r / fabs(v1);
        }
        const double sc = sp_scan<G>(r, lane, mask);
        base += __shfl_sync(mask, sc, G - 1, G);
    }
    SpStep o;
    o.S = base;
    o.tau = fabs(v1) / base;                                        // qmc.jl:33
    o.eloc = (dv.x * v1 + sp_sum<G>(num_l, mask)) / v1;              // qmc.jl:21,37
    o.nhops = (int)(e - b);
    o.chosen = -1;
    o.next = k;
    const double T = u01 * base;  // pick_random_from_cumsum (qmc.jl:52-61)
    bool found = false;
#pragma unroll
    for (int c = 0; c < SP_R; ++c) {
        if (!found && b + (unsigned long long)(c * G) < e) {
            const unsigned hit = __ballot_sync(mask, T < cc[c]) & mask;
            if (hit) {
                const int src = (__ffs(hit) - 1) & (G - 1);
                o.next = __shfl_sync(mask, cn[c], src, G);
                o.chosen = c * G + src;
                found = true;
            }
        }
    }
    if (!found) {
        double run = baseR;
        for (unsigned long long i0 = b + (unsigned long long)(SP_R * G); i0 < e; i0 += G) {
            const unsigned long long i = i0 + lane;
            double r = 0.0;
            uint32_t cl = 0u;
            if (i < e) {
                cl = sd.col[i];
                r = fabs(sd.rpsi[i]);
            }
            const double sc = sp_scan<G>(r, lane, mask);
            const unsigned hit = __ballot_sync(mask, T < run + sc) & mask;
            if (hit) {
                const int src = (__ffs(hit) - 1) & (G - 1);
                o.next = __shfl_sync(mask, cl, src, G);
                o.chosen = (int)(i0 - b) + src;
                break;
            }
            run += __shfl_sync(mask, sc, G - 1, G);
        }
    }
    return o;
}

// doubles per chunk of G row entries in a block: col u32[G] | melem[G] | rpsi[G]
__host__ __device__ constexpr int sp_chunk_doubles(int G) { return G / 2 + 2 * G; }
constexpr int SP_HEAD = 8;  // doubles in the head of a block

// kinetic_sample! at index k from the row block (same arithmetic, same order as sp_step)
template <int G, bool RATES>
__device__ __forceinline__ SpStep sp_step_blk(const SparseDev& sd, uint32_t k, int lane, unsigned mask, double u01,
                                              double* rates) {
    constexpr int CH = sp_chunk_doubles(G);
    const double* blk = sd.blk + (size_t)k * sd.BS;
    // the head and the first SP_R chunks are loaded together: none of their addresses depends on the row length
    // (blocks hold at least SP_R chunks; entries past the row end are zeros: rate 0, never picked)
    const double2 h01 = *reinterpret_cast<const double2*>(blk);
    const double v1 = blk[2];
    double cc[SP_R], rr[SP_R], mm[SP_R];
    uint32_t cn[SP_R];
#pragma unroll
    for (int c = 0; c < SP_R; ++c) {
        const double* ch = blk + SP_HEAD + c * CH;
        cn[c] = reinterpret_cast<const uint32_t*>(ch)[lane];
        mm[c] = ch[G / 2 + lane];
        rr[c] = ch[G / 2 + G + lane];
    }
    const int len = (int)__double_as_longlong(h01.x);
    double num_l = 0.0, base = 0.0;
#pragma unroll
    for (int c = 0; c < SP_R; ++c) {
        cc[c] = 0.0;
        if (c * G < len) {  // uniform in the group
            const double r = fabs(rr[c]);                // qmc.jl:27
            num_l = fma(mm[c], rr[c], num_l);            // qmc.jl:28
            if (RATES && c * G + lane < len) rates[c * G + lane] = r / fabs(v1);
            const double sc = sp_scan<G>(r, lane, mask);
            cc[c] = base + sc;  // qmc.jl:30
            base += __shfl_sync(mask, sc, G - 1, G);
        }
    }
    const double baseR = base;
    for (int j0 = SP_R * G; j0 < len; j0 += G) {
        const double* ch = blk + SP_HEAD + (j0 / G) * CH;
        const double v2 = ch[G / 2 + G + lane];
        const double r = fabs(v2);
        num_l = fma(ch[G / 2 + lane], v2, num_l);
        if (RATES && j0 + lane < len) rates[j0 + lane] = r / fabs(v1);
        const double sc = sp_scan<G>(r, lane, mask);
        base += __shfl_sync(mask, sc, G - 1, G);
    }
    SpStep o;
    o.S = base;
    o.tau = fabs(v1) / base;                                  // qmc.jl:33
    o.eloc = (h01.y * v1 + sp_sum<G>(num_l, mask)) / v1;      // qmc.jl:21,37
    o.nhops = len;
    o.chosen = -1;
    o.next = k;
    const double T = u01 * base;  // pick_random_from_cumsum (qmc.jl:52-61)
    bool found = false;
#pragma unroll
    for (int c = 0; c < SP_R; ++c) {
        if (!found && c * G < len) {
            const unsigned hit = __ballot_sync(mask, T < cc[c]) & mask;
            if (hit) {
                const int src = (__ffs(hit) - 1) & (G - 1);
                o.next = __shfl_sync(mask, cn[c], src, G);
                o.chosen = c * G + src;
                found = true;
            }
        }
    }
    if (!found) {
        double run = baseR;
        for (int j0 = SP_R * G; j0 < len; j0 += G) {
            const double* ch = blk + SP_HEAD + (j0 / G) * CH;
            const uint32_t cl = reinterpret_cast<const uint32_t*>(ch)[lane];
            const double sc = sp_scan<G>(fabs(ch[G / 2 + G + lane]), lane, mask);
            const unsigned hit = __ballot_sync(mask, T < run + sc) & mask;
            if (hit) {
                const int src = (__ffs(hit) - 1) & (G - 1);
                o.next = __shfl_sync(mask, cl, src, G);
                o.chosen = j0 + src;
                break;
            }
            run += __shfl_sync(mask, sc, G - 1, G);
        }
    }
    return o;
}

// accumulator rows per walker: 0 sum tau, 1 sum tau dE, 2 sum tau dE^2, 3 hops, 4+p sum tau G_p, 4+P+p sum tau G_p dE
__host__ __device__ inline int sp_acc_rows(int P) { return 4 + 2 * P; }

template <int G, int PPL, bool BLK>
__global__ void __launch_bounds__(SP_BLOCK, PPL == 1 ? GWZ_SP_MINB : (PPL == 2 ? (GWZ_SP_MINB < 6 ? GWZ_SP_MINB : 6) : 1))
sparse_walker_kernel(const __grid_constant__ SparseDev sd, const __grid_constant__ PhiloxKeys pk, uint32_t* __restrict__ idx,
                     double* __restrict__ acc, size_t n, uint64_t goff, uint64_t step0, uint64_t steps, int accumulate,
                     int reset_acc, double E0, double* __restrict__ rec_tau, double* __restrict__ rec_eloc,
                     uint32_t* __restrict__ rec_idx, double* __restrict__ hist, int* __restrict__ err_flag) {
    const int lane = threadIdx.x & (G - 1);
    const unsigned mask = sp_mask<G>();
    const size_t group = ((size_t)blockIdx.x * SP_BLOCK + threadIdx.x) / G, ngroups = (size_t)gridDim.x * SP_BLOCK / G;
    const int P = sd.P;
    bool bad = false;
    for (size_t w = group; w < n; w += ngroups) {
        uint32_t k = idx[w];
        double st = 0.0, ste = 0.0, stee = 0.0, hops = 0.0;
        double sg[PPL], sge[PPL];
#pragma unroll
        for (int q = 0; q < PPL; ++q) sg[q] = sge[q] = 0.0;
        if (accumulate && !reset_acc) {
            st = acc[w];
            ste = acc[n + w];
            stee = acc[2 * n + w];
            hops = acc[3 * n + w];
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int p = lane + G * q;
                if (p < P) {
                    sg[q] = acc[(size_t)(4 + p) * n + w];
                    sge[q] = acc[(size_t)(4 + P + p) * n + w];
                }
            }
        }
        const uint64_t wid = goff + (uint64_t)w;
        const uint32_t c2 = (uint32_t)wid, c3 = (uint32_t)(wid >> 32);
        uint32_t rnd[4] = {0u, 0u, 0u, 0u};
        for (uint64_t s = 0; s < steps; ++s) {
            const uint64_t step = step0 + s;
            // one Philox block serves two steps; with G >= 8 the lanes of the group evaluate G consecutive blocks at
            // once (lane l: block (blk & ~(G-1)) + l) and hand the words to the group when their step comes
            // (measured: +3 % at G = 8, -4 % at G = 4, where every lane keeps evaluating its walker's block)
            constexpr int SH = G >= 8 ? G : 1;
            const uint64_t blk = step >> 1;
            if (s == 0 || ((step & 1ull) == 0ull && (blk & (uint64_t)(SH - 1)) == 0ull)) {
                const uint64_t mine = (blk & ~(uint64_t)(SH - 1)) + (uint64_t)(SH > 1 ? lane : 0);
                philox4x32_10_keyed((uint32_t)mine, (uint32_t)(mine >> 32), c2, c3, pk, rnd);
            }
            uint32_t wlo = (step & 1ull) ? rnd[2] : rnd[0], whi = (step & 1ull) ? rnd[3] : rnd[1];
            if (SH > 1) {
                const int src = (int)(blk & (uint64_t)(SH - 1));
                wlo = __shfl_sync(mask, wlo, src, G);
                whi = __shfl_sync(mask, whi, src, G);
            }
            const double u01 = u01_from_words(wlo, whi);
            const SpStep o = BLK ? sp_step_blk<G, false>(sd, k, lane, mask, u01, nullptr)
                                 : sp_step<G, false>(sd, k, lane, mask, u01, nullptr);
            if (o.chosen < 0 || !(o.tau < 1.0e300)) {  // qmc.jl:34-36
                bad = true;
                break;
            }
            if (BLK) {  // the move is known: start the next state's row block on its way (one 128-byte line per lane)
                const double* nb = sd.blk + (size_t)o.next * sd.BS;
                if (lane * 16 < sd.BS && lane * 16 < SP_HEAD + SP_R * sp_chunk_doubles(G))
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + lane * 16));
            }
            if (lane == 0) {
                if (rec_tau) {
                    rec_tau[s * n + w] = o.tau;
                    rec_eloc[s * n + w] = o.eloc;
                    if (rec_idx) rec_idx[s * n + w] = k;  // the address sampled from (qmc.jl:89)
                }
                if (hist) atomicAdd(hist + k, o.tau);
            }
            const double dE = o.eloc - E0, tE = o.tau * dE;
            st += o.tau;
            ste += tE;
            stee = fma(tE, dE, stee);
            hops += (double)o.nhops;
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int p = lane + G * q;
                if (p < P) {
                    // qmc.jl:38; the block head holds the first five ratios (same 128 B line as the row's start)
                    const double g = (BLK && P <= 5) ? sd.blk[(size_t)k * sd.BS + 3 + p] : sd.rec[(size_t)k * sd.RS + 4 + p];
                    sg[q] = fma(o.tau, g, sg[q]);
                    sge[q] = fma(tE, g, sge[q]);
                }
            }
            k = o.next;
        }
        if (lane == 0) idx[w] = k;
        if (accumulate) {
            if (lane == 0) {
                acc[w] = st;
                acc[n + w] = ste;
                acc[2 * n + w] = stee;
                acc[3 * n + w] = hops;
            }
#pragma unroll
            for (int q = 0; q < PPL; ++q) {
                const int p = lane + G * q;
                if (p < P) {
                    acc[(size_t)(4 + p) * n + w] = sg[q];
                    acc[(size_t)(4 + P + p) * n + w] = sge[q];
                }
            }
        }
    }
    if (bad) atomicOr(err_flag, 1);
}

__device__ __forceinline__ double sp_block_sum(double v, double* sm /* [SP_BLOCK/32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < SP_BLOCK / 32; ++i) t += sm[i];
    return t;
}

// state.jl:42-49 per walker, then the sums over walkers of state.jl:137-149 and of the pooled estimator:
// partials[block][GEN_HEAD + 4P] in the layout of the ansatz-generic engine (gwz_generic.cuh)
__global__ void __launch_bounds__(SP_BLOCK)
sparse_finalize_kernel(const double* __restrict__ acc, size_t n, int P, double E0, double steps_per_walker,
                       double* __restrict__ partials) {
    __shared__ double sm[SP_BLOCK / 32];
    const int L = gen_vec_len(P);
    double* out = partials + (size_t)blockIdx.x * L;
    const size_t t0 = (size_t)blockIdx.x * SP_BLOCK + threadIdx.x, nt = (size_t)gridDim.x * SP_BLOCK;
    double h[GEN_HEAD];
#pragma unroll
    for (int i = 0; i < GEN_HEAD; ++i) h[i] = 0.0;
    for (size_t w = t0; w < n; w += nt) {
        const double st = acc[w], ste = acc[n + w], stee = acc[2 * n + w];
        const double val_w = E0 + ste / st;
        h[GS_WALKERS] += 1.0;
        h[GS_VAL] += val_w;
        h[GS_VAL2] = fma(val_w, val_w, h[GS_VAL2]);
        h[GS_T] += st;
        h[GS_TE] += st * val_w;
        h[GS_TEE] += st * E0 * E0 + 2.0 * E0 * ste + stee;
        h[GS_HOPS] += acc[3 * n + w];
        h[GS_STEPS] += steps_per_walker;
    }
#pragma unroll
    for (int i = 0; i < GEN_HEAD; ++i) {
        const double v = sp_block_sum(h[i], sm);
        if (threadIdx.x == 0) out[i] = v;
    }
    for (int p = 0; p < P; ++p) {
        double g1 = 0.0, g2 = 0.0, tg = 0.0, tge = 0.0;
        for (size_t w = t0; w < n; w += nt) {
            const double st = acc[w], ste = acc[n + w];
            const double sg = acc[(size_t)(4 + p) * n + w], sge = acc[(size_t)(4 + P + p) * n + w];
            const double gw = 2.0 * (sge / st - (ste / st) * (sg / st));  // state.jl:46-48 (shift-invariant form)
            g1 += gw;
            g2 = fma(gw, gw, g2);
            tg += sg;
            tge += sge + E0 * sg;
        }
        g1 = sp_block_sum(g1, sm);
        g2 = sp_block_sum(g2, sm);
        tg = sp_block_sum(tg, sm);
        tge = sp_block_sum(tge, sm);
        if (threadIdx.x == 0) {
            out[GEN_HEAD + p] = g1;
            out[GEN_HEAD + P + p] = g2;
            out[GEN_HEAD + 2 * P + p] = tg;
            out[GEN_HEAD + 3 * P + p] = tge;
        }
    }
}

// LocalEnergyEvaluator (localenergy.jl:94-155): one row per group; per-group sums of (num, den, num_grad[P],
// den_grad[P]) to partials[group][2 + 2P]; optional per-row terms
template <int G, int PPL>
__global__ void __launch_bounds__(SP_BLOCK, PPL == 1 ? GWZ_SP_SWEEP_MINB : 1)
sparse_sweep_kernel(const __grid_constant__ SparseDev sd, unsigned long long row_begin, unsigned long long row_end,
                    int with_grad, double* __restrict__ partials, double* __restrict__ terms) {
    constexpr int PB = 4;  // parameters per pass over a row
    const int lane = threadIdx.x & (G - 1);
    const unsigned mask = sp_mask<G>();
    const size_t group = ((size_t)blockIdx.x * SP_BLOCK + threadIdx.x) / G, ngroups = (size_t)gridDim.x * SP_BLOCK / G;
    const int P = with_grad ? sd.P : 0, L = 2 + 2 * P;
    double num = 0.0, den = 0.0;
    double ng[PPL], dg[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) ng[q] = dg[q] = 0.0;
    for (unsigned long long k = row_begin + group; k < row_end; k += ngroups) {
        const unsigned long long b = sd.row_ptr[k], e = sd.row_ptr[k + 1];
        const double kv = sd.psi(k), dgl = sd.diag[k];
        double off = 0.0;
        for (int p0 = 0; p0 == 0 || p0 < P; p0 += PB) {
            double off_l = 0.0, og_l[PB];  // sum_l melem * l_val;  sum_l melem * l_grad, l_grad = grad_ratio[l] * l_val
#pragma unroll
            for (int j = 0; j < PB; ++j) og_l[j] = 0.0;
            for (unsigned long long i = b + lane; i < e; i += G) {
                const uint32_t c = sd.col[i];
                const double m = sd.melem[i], lv = P ? sd.psi_tab(c) : sd.psi(c);
                off_l = fma(m, lv, off_l);
                const double mv = m * lv;
#pragma unroll
                for (int j = 0; j < PB; ++j)
                    if (p0 + j < P) og_l[j] = fma(mv, sd.gratio(c, p0 + j), og_l[j]);
            }
            if (p0 == 0) {
                off = sp_sum<G>(off_l, mask);
                const double num_k = kv * kv * dgl + off * kv, den_k = kv * kv;  // localenergy.jl:108,111,116
                num += num_k;
                den += den_k;
                if (terms && lane == 0) {
                    terms[(k - row_begin) * L] = num_k;
                    terms[(k - row_begin) * L + 1] = den_k;
                }
            }
#pragma unroll
            for (int j = 0; j < PB; ++j) {
                const int p = p0 + j;
                if (p < P) {
                    const double og = sp_sum<G>(og_l[j], mask);
                    const double kg = sd.gratio(k, p) * kv;            // k_grad
                    const double ngk = 2.0 * kg * kv * dgl + (off * kg + og * kv);  // localenergy.jl:109,117
                    const double dgk = 2.0 * kv * kg;                                // localenergy.jl:112
#pragma unroll
                    for (int q = 0; q < PPL; ++q)
                        if (p == lane + G * q) {
                            ng[q] += ngk;
                            dg[q] += dgk;
                        }
                    if (terms && lane == 0) {
                        terms[(k - row_begin) * L + 2 + p] = ngk;
                        terms[(k - row_begin) * L + 2 + P + p] = dgk;
                    }
                }
            }
        }
    }
    double* out = partials + group * L;
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

// per-index kinetic_sample! outputs (parity surface)
template <int G, bool BLK>
__global__ void __launch_bounds__(SP_BLOCK)
sparse_probe_kernel(const __grid_constant__ SparseDev sd, const uint32_t* __restrict__ idx, size_t n,
                    const double* __restrict__ u01, double* __restrict__ e_loc, double* __restrict__ tau,
                    double* __restrict__ grad_ratio, int32_t* __restrict__ n_hops, double* __restrict__ rates,
                    size_t rates_stride, int32_t* __restrict__ chosen, uint32_t* __restrict__ next_idx,
                    int* __restrict__ err_flag) {
    const int lane = threadIdx.x & (G - 1);
    const unsigned mask = sp_mask<G>();
    const size_t group = ((size_t)blockIdx.x * SP_BLOCK + threadIdx.x) / G, ngroups = (size_t)gridDim.x * SP_BLOCK / G;
    for (size_t i = group; i < n; i += ngroups) {
        const uint32_t k = idx[i];
        const double u = u01 ? u01[i] : 0.5;
        const SpStep o = BLK ? (rates ? sp_step_blk<G, true>(sd, k, lane, mask, u, rates + i * rates_stride)
                                      : sp_step_blk<G, false>(sd, k, lane, mask, u, nullptr))
                             : (rates ? sp_step<G, true>(sd, k, lane, mask, u, rates + i * rates_stride)
                                      : sp_step<G, false>(sd, k, lane, mask, u, nullptr));
        if (o.chosen < 0 && lane == 0) atomicOr(err_flag, 1);
        if (lane == 0) {
            if (e_loc) e_loc[i] = o.eloc;
            if (tau) tau[i] = o.tau;
            if (n_hops) n_hops[i] = o.nhops;
            if (chosen) chosen[i] = o.chosen + 1;  // 1-based like the reference
            if (next_idx) next_idx[i] = o.next;
        }
        if (grad_ratio)
            for (int p = lane; p < sd.P; p += G) grad_ratio[i * sd.P + p] = sd.gratio(k, p);
    }
}

// GutzwillerAnsatz over the stored diagonal (gutzwiller.jl:25-36): psi = exp(-g D), grad psi / psi = -D
__global__ void sparse_fill_gutzwiller_kernel(const double* __restrict__ diag, unsigned long long dim, double g,
                                              double* __restrict__ amp, double2* __restrict__ tab) {
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < dim;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const double d = diag[k], v = exp(-g * d);
        amp[k] = v;
        tab[k] = make_double2(v, -d);
    }
}

// rpsi[i] = psi[col[i]]: one gather per stored entry and table, instead of one per hop of every walker step
__global__ void sparse_expand_kernel(const uint32_t* __restrict__ col, const double* __restrict__ amp,
                                     unsigned long long nnz, double* __restrict__ rpsi) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
         i += (unsigned long long)gridDim.x * blockDim.x)
        rpsi[i] = amp[col[i]];
}

// the walkers' per-state record (see SparseDev::rec)
__global__ void sparse_build_rec_kernel(const unsigned long long* __restrict__ row_ptr, const double* __restrict__ diag,
                                        const double* __restrict__ amp, const double* __restrict__ tab, int P, int RS,
                                        unsigned long long dim, double* __restrict__ rec) {
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < dim;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        double* r = rec + k * (unsigned)RS;
        r[0] = __longlong_as_double((long long)row_ptr[k]);
        r[1] = __longlong_as_double((long long)row_ptr[k + 1]);
        r[2] = diag[k];
        r[3] = amp[k];
        for (int p = 0; p < P; ++p) r[4 + p] = tab[k * (unsigned)(1 + P) + 1 + p];
    }
}

// row blocks (see SparseDev::blk): one thread per state copies its row; the block memory was zeroed before
template <int G>
__global__ void sparse_fill_blocks_kernel(const __grid_constant__ SparseDev sd, double* __restrict__ blk) {
    constexpr int CH = sp_chunk_doubles(G);
    const int P = sd.P;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < sd.dim;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        double* b = blk + k * (unsigned)sd.BS;
        const unsigned long long r0 = sd.row_ptr[k], r1 = sd.row_ptr[k + 1];
        b[0] = __longlong_as_double((long long)(r1 - r0));
        b[1] = sd.diag[k];
        b[2] = sd.amp[k];
        for (int p = 0; p < P && p < 5; ++p) b[3 + p] = sd.gratio(k, p);
        for (unsigned long long i = r0; i < r1; ++i) {
            const int j = (int)(i - r0), c = j / G, l = j % G;
            double* ch = b + SP_HEAD + c * CH;
            const uint32_t cl = sd.col[i];
            reinterpret_cast<uint32_t*>(ch)[l] = cl;
            ch[G / 2 + l] = sd.melem[i];
            ch[G / 2 + G + l] = sd.amp[cl];
        }
    }
}

__global__ void sparse_fill_idx_kernel(uint32_t* idx, size_t n, uint32_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) idx[i] = v;
}

}  // namespace gwz

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
struct gwz_sparse {
    uint32_t magic = gwz::GWZ_TAG_SPARSE;  // handle type tag, checked by every entry point
    int device = 0;  // destroy must not dereference parent handles (a host GC may release them in any order)
    gwz_context* ctx = nullptr;
    uint64_t dim = 0, nnz = 0, max_row = 0;
    int P = 0, G = 8, G_base = 8;  // G_base: lanes per walker from the row lengths; G: raised so that P fits 8 per lane
    uint64_t table_version = 0;  // 0 = no ansatz table yet
    unsigned long long* d_row_ptr = nullptr;
    uint32_t* d_col = nullptr;
    double* d_melem = nullptr;
    double* d_diag = nullptr;
    double* d_rpsi = nullptr;  // [nnz]
    double* d_rec = nullptr;   // [dim][RS]
    size_t rec_cap = 0;
    double* d_blk = nullptr;   // [dim][BS] row blocks, or null
    size_t blk_cap = 0;
    double* d_amp = nullptr;  // [dim]
    double* d_tab = nullptr;  // [dim][1 + P]
    size_t tab_cap = 0;
    double* d_partials = nullptr;
    size_t partials_cap = 0;
    float last_ms = 0.f;
    SparseDev sd{};
};
struct gwz_swalkers {
    uint32_t magic = gwz::GWZ_TAG_SWALKERS;  // handle type tag, checked by every entry point
    int device = 0;
    gwz_sparse* h = nullptr;
    size_t n = 0;
    uint64_t offset = 0, seed = 0, steps_done = 0, acc_steps = 0, table_seen = 0;
    int P_acc = -1;
    uint32_t* d_idx = nullptr;
    double* d_acc = nullptr;
    double* d_partials = nullptr;
    size_t partials_cap = 0;
    double* d_vec = nullptr;
    double* h_vec = nullptr;
    size_t vec_cap = 0;
    double shift_e = 0.0;
};

static int sp_ppl(int P, int G) {
    const int need = (P + G - 1) / G;
    return need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : 8;
}

#define GWZ_SP_DISPATCH_G(G_, CALL)                                \
    switch (G_) {                                                  \
        case 4: { constexpr int G = 4; CALL; } break;              \
        case 8: { constexpr int G = 8; CALL; } break;              \
        case 16: { constexpr int G = 16; CALL; } break;            \
        default: { constexpr int G = 32; CALL; } break;            \
    }
#define GWZ_SP_DISPATCH(G_, PPL_, CALL)                                                   \
    GWZ_SP_DISPATCH_G(G_, switch (PPL_) {                                                 \
        case 1: { constexpr int PPL = 1; CALL; } break;                                   \
        case 2: { constexpr int PPL = 2; CALL; } break;                                   \
        case 4: { constexpr int PPL = 4; CALL; } break;                                   \
        default: { constexpr int PPL = 8; CALL; } break;                                  \
    })

#define GWZ_SP_DISPATCH_BLK(B_, CALL)             \
    if (B_) { constexpr bool BLK = true; CALL; }  \
    else { constexpr bool BLK = false; CALL; }

// lanes per walker / row: the smallest power of two that covers a quarter of a mean row -- the path is bound by the
// latency of its dependent gathers, so more walkers in flight beat fewer chunks per row (measured: L = 24 runs
// fastest with G = 8, L = 8 with G = 4) -- raised until the parameters fit 8 per lane.  GWZ_SPARSE_GROUP overrides.
static int sp_group_base(double mean_row) {
    int G = 4;
    while (G < 32 && 4.0 * G < mean_row) G <<= 1;
    if (const char* e = std::getenv("GWZ_SPARSE_GROUP")) {
        const int v = std::atoi(e);
        if (v == 4 || v == 8 || v == 16 || v == 32) G = v;
    }
    return G;
}
static int sp_group(int G, int P) {
    while (G < 32 && P > 8 * G) G <<= 1;
    return G;
}

extern "C" int gwz_sparse_destroy(gwz_sparse* h) {
    if (h && h->magic != gwz::GWZ_TAG_SPARSE) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_sparse_destroy: not a gwz_sparse handle");
    if (!h) return GWZ_OK;
    if (cudaSetDevice(h->device) == cudaSuccess) {
        cudaFree(h->d_row_ptr);
        cudaFree(h->d_col);
        cudaFree(h->d_melem);
        cudaFree(h->d_diag);
        cudaFree(h->d_amp);
        cudaFree(h->d_rpsi);
        cudaFree(h->d_rec);
        cudaFree(h->d_blk);
        cudaFree(h->d_tab);
        cudaFree(h->d_partials);
    }
    h->magic = 0;
    delete h;
    return GWZ_OK;
}

extern "C" int gwz_sparse_create(gwz_context* ctx, uint64_t dim, const uint64_t* row_ptr, const uint32_t* col,
                                 const double* melem, const double* diag, gwz_sparse** out) {
    GWZ_HANDLE(ctx, gwz::GWZ_TAG_CONTEXT, "gwz_sparse_create: NULL or foreign handle (expected a gwz_context)");
    GWZ_REQUIRE(ctx && row_ptr && diag && out, "gwz_sparse_create: NULL argument");
    GWZ_REQUIRE(dim >= 1 && dim <= 0xffffffffull, "gwz_sparse_create: dimension must be in [1, 2^32)");
    GWZ_REQUIRE(row_ptr[0] == 0, "gwz_sparse_create: row_ptr[0] must be 0");
    uint64_t max_row = 0;
    for (uint64_t k = 0; k < dim; ++k) {
        GWZ_REQUIRE(row_ptr[k + 1] >= row_ptr[k], "gwz_sparse_create: row_ptr must be non-decreasing");
        const uint64_t len = row_ptr[k + 1] - row_ptr[k];
        GWZ_REQUIRE(len < (1ull << 31), "gwz_sparse_create: row too long");
        if (len > max_row) max_row = len;
    }
    const uint64_t nnz = row_ptr[dim];
    GWZ_REQUIRE(nnz == 0 || (col && melem), "gwz_sparse_create: NULL col/melem");
    for (uint64_t i = 0; i < nnz; ++i) {
        GWZ_REQUIRE(col[i] < dim, "gwz_sparse_create: column index outside the basis");
        GWZ_REQUIRE(std::isfinite(melem[i]), "gwz_sparse_create: non-finite matrix element");
    }
    for (uint64_t k = 0; k < dim; ++k) GWZ_REQUIRE(std::isfinite(diag[k]), "gwz_sparse_create: non-finite diagonal element");
    if (int rc = use_device(ctx)) return rc;
    gwz_sparse* h = new gwz_sparse();
    HandleGuard<gwz_sparse, gwz_sparse_destroy> guard(h);
    h->ctx = ctx;
    h->device = ctx->device;
    h->dim = dim;
    h->nnz = nnz;
    h->max_row = max_row;
    GWZ_CUDA(cudaMalloc(&h->d_row_ptr, sizeof(uint64_t) * (dim + 1)));
    GWZ_CUDA(cudaMalloc(&h->d_col, sizeof(uint32_t) * (nnz ? nnz : 1)));
    GWZ_CUDA(cudaMalloc(&h->d_melem, sizeof(double) * (nnz ? nnz : 1)));
    GWZ_CUDA(cudaMalloc(&h->d_diag, sizeof(double) * dim));
    GWZ_CUDA(cudaMemcpy(h->d_row_ptr, row_ptr, sizeof(uint64_t) * (dim + 1), cudaMemcpyHostToDevice));
    if (nnz) {
        GWZ_CUDA(cudaMemcpy(h->d_col, col, sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice));
        GWZ_CUDA(cudaMemcpy(h->d_melem, melem, sizeof(double) * nnz, cudaMemcpyHostToDevice));
    }
    GWZ_CUDA(cudaMemcpy(h->d_diag, diag, sizeof(double) * dim, cudaMemcpyHostToDevice));
    h->sd.row_ptr = h->d_row_ptr;
    h->sd.col = h->d_col;
    h->sd.melem = h->d_melem;
    h->sd.diag = h->d_diag;
    GWZ_CUDA(cudaMalloc(&h->d_amp, sizeof(double) * dim));
    h->sd.amp = h->d_amp;
    GWZ_CUDA(cudaMalloc(&h->d_rpsi, sizeof(double) * (nnz ? nnz : 1)));
    h->sd.rpsi = h->d_rpsi;
    h->sd.tab = nullptr;
    h->sd.dim = dim;
    h->sd.P = 0;
    h->G_base = sp_group_base((double)nnz / (double)dim);
    h->G = h->G_base;
    *out = guard.release();
    return GWZ_OK;
}

extern "C" int gwz_sparse_info(const gwz_sparse* h, uint64_t* dim, uint64_t* nnz, uint64_t* max_row, int32_t* num_parameters,
                               int32_t* lanes_per_walker) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_sparse_info: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h != nullptr, "gwz_sparse_info: NULL handle");
    if (dim) *dim = h->dim;
    if (nnz) *nnz = h->nnz;
    if (max_row) *max_row = h->max_row;
    if (num_parameters) *num_parameters = h->P;
    if (lanes_per_walker) *lanes_per_walker = h->G;
    return GWZ_OK;
}

extern "C" int gwz_sparse_layout(const gwz_sparse* h, int32_t* row_blocks, uint64_t* block_bytes) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_sparse_layout: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h != nullptr, "gwz_sparse_layout: NULL handle");
    if (row_blocks) *row_blocks = h->sd.blk != nullptr;
    if (block_bytes) *block_bytes = h->sd.blk ? (uint64_t)h->sd.BS * 8 : 0;
    return GWZ_OK;
}

static int sp_reserve_table(gwz_sparse* h, int P) {
    const size_t need = (size_t)h->dim * (size_t)(1 + P);
    if (need > h->tab_cap) {
        if (h->d_tab) cudaFree(h->d_tab);
        h->d_tab = nullptr;
        h->tab_cap = 0;
        GWZ_CUDA(cudaMalloc(&h->d_tab, sizeof(double) * need));
        h->tab_cap = need;
    }
    const int RS = (4 + P + 3) / 4 * 4 <= 8 ? 8 : (4 + P + 3) / 4 * 4;  // whole 32 B sectors; one 64 B block for P <= 4
    if ((size_t)h->dim * RS > h->rec_cap) {
        if (h->d_rec) cudaFree(h->d_rec);
        h->d_rec = nullptr;
        h->rec_cap = 0;
        GWZ_CUDA(cudaMalloc(&h->d_rec, sizeof(double) * h->dim * RS));
        h->rec_cap = (size_t)h->dim * RS;
    }
    h->sd.rec = h->d_rec;
    h->sd.RS = RS;
    h->P = P;
    h->sd.P = P;
    h->sd.tab = h->d_tab;
    h->G = sp_group(h->G_base, P);
    return GWZ_OK;
}

// Row blocks are used when the rows are regular enough: at most twice the bytes of the row-pointer layout (or
// 64 MiB) and at most a third of the free device memory.  GWZ_SPARSE_BLOCKS=0 keeps the row-pointer layout.
static int sp_plan_blocks(gwz_sparse* h) {
    h->sd.blk = nullptr;
    h->sd.BS = 0;
    if (const char* e = std::getenv("GWZ_SPARSE_BLOCKS"))
        if (std::atoi(e) == 0) return GWZ_OK;
    const int G = h->G;
    const uint64_t chunks = std::max<uint64_t>((h->max_row + G - 1) / G, SP_R);
    const uint64_t BS = (SP_HEAD + chunks * sp_chunk_doubles(G) + 15) / 16 * 16;
    if (BS > (1u << 30)) return GWZ_OK;
    const uint64_t blk_bytes = h->dim * BS * 8, csr_bytes = h->nnz * 20 + h->dim * 72;
    size_t free_b = 0, total_b = 0;
    GWZ_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t have = free_b + h->blk_cap * 8;
    if (blk_bytes > std::max<uint64_t>(2 * csr_bytes, 64ull << 20) || blk_bytes > have / 3) return GWZ_OK;
    if (h->dim * BS > h->blk_cap) {
        if (h->d_blk) cudaFree(h->d_blk);
        h->d_blk = nullptr;
        h->blk_cap = 0;
        GWZ_CUDA(cudaMalloc(&h->d_blk, blk_bytes));
        h->blk_cap = h->dim * BS;
    }
    h->sd.blk = h->d_blk;
    h->sd.BS = (int)BS;
    return GWZ_OK;
}

static int sp_expand(gwz_sparse* h) {
    gwz_context* c = h->ctx;
    const unsigned rgrid = (unsigned)std::min<uint64_t>((h->dim + 255) / 256, (uint64_t)c->prop.multiProcessorCount * 16);
    api_count_launch();
    sparse_build_rec_kernel<<<rgrid, 256, 0, c->stream>>>(h->d_row_ptr, h->d_diag, h->d_amp, h->d_tab, h->P, h->sd.RS, h->dim,
                                                          h->d_rec);
    GWZ_CUDA(cudaGetLastError());
    if (h->nnz == 0) return GWZ_OK;
    const unsigned grid = (unsigned)std::min<uint64_t>((h->nnz + 255) / 256, (uint64_t)c->prop.multiProcessorCount * 16);
    api_count_launch();
    sparse_expand_kernel<<<grid, 256, 0, c->stream>>>(h->d_col, h->d_amp, h->nnz, h->d_rpsi);
    GWZ_CUDA(cudaGetLastError());
    if (int rc = sp_plan_blocks(h)) return rc;
    if (h->sd.blk) {
        GWZ_CUDA(cudaMemsetAsync(h->d_blk, 0, sizeof(double) * h->dim * h->sd.BS, c->stream));
        api_count_launch();
        GWZ_SP_DISPATCH_G(h->G, (sparse_fill_blocks_kernel<G><<<rgrid, 256, 0, c->stream>>>(h->sd, h->d_blk)));
        GWZ_CUDA(cudaGetLastError());
    }
    return GWZ_OK;
}

// any ansatz as a table: psi[dim] = ansatz(basis[k], params), grad_ratio[dim][P] = grad psi / psi (NULL when P = 0,
// e.g. VectorAnsatz, vector.jl:6-12)
extern "C" int gwz_sparse_set_table(gwz_sparse* h, const double* psi, const double* grad_ratio, int32_t P) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_sparse_set_table: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h && psi, "gwz_sparse_set_table: NULL argument");
    GWZ_REQUIRE(P >= 0 && P <= 256, "gwz_sparse_set_table: number of parameters must be in [0, 256]");
    GWZ_REQUIRE(P == 0 || grad_ratio, "gwz_sparse_set_table: NULL grad_ratio with P > 0");
    for (uint64_t k = 0; k < h->dim; ++k) GWZ_REQUIRE(std::isfinite(psi[k]), "gwz_sparse_set_table: non-finite amplitude");
    for (uint64_t k = 0; k < h->dim * (uint64_t)P; ++k)
        GWZ_REQUIRE(std::isfinite(grad_ratio[k]), "gwz_sparse_set_table: non-finite gradient ratio");
    gwz_context* c = h->ctx;
    if (int rc = use_device(c)) return rc;
    if (int rc = sp_reserve_table(h, P)) return rc;
    std::vector<double> tab((size_t)h->dim * (1 + P));
    for (uint64_t k = 0; k < h->dim; ++k) {
        tab[k * (1 + P)] = psi[k];
        for (int p = 0; p < P; ++p) tab[k * (1 + P) + 1 + p] = grad_ratio[k * P + p];
    }
    GWZ_CUDA(cudaMemcpyAsync(h->d_tab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(h->d_amp, psi, sizeof(double) * h->dim, cudaMemcpyHostToDevice, c->stream));
    if (int rc = sp_expand(h)) return rc;
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    ++h->table_version;
    return GWZ_OK;
}

extern "C" int gwz_sparse_set_gutzwiller(gwz_sparse* h, double g) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_sparse_set_gutzwiller: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h != nullptr, "gwz_sparse_set_gutzwiller: NULL handle");
    GWZ_REQUIRE(std::isfinite(g), "gwz_sparse_set_gutzwiller: non-finite parameter");
    gwz_context* c = h->ctx;
    if (int rc = use_device(c)) return rc;
    if (int rc = sp_reserve_table(h, 1)) return rc;
    const unsigned grid = (unsigned)std::min<uint64_t>((h->dim + 255) / 256, (uint64_t)c->prop.multiProcessorCount * 8);
    api_count_launch();
    sparse_fill_gutzwiller_kernel<<<grid, 256, 0, c->stream>>>(h->d_diag, h->dim, g, h->d_amp, reinterpret_cast<double2*>(h->d_tab));
    GWZ_CUDA(cudaGetLastError());
    if (int rc = sp_expand(h)) return rc;
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    ++h->table_version;
    return GWZ_OK;
}

extern "C" int gwz_sparse_get_table(gwz_sparse* h, double* psi, double* grad_ratio) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_sparse_get_table: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h != nullptr, "gwz_sparse_get_table: NULL handle");
    GWZ_REQUIRE(h->table_version > 0, "gwz_sparse_get_table: no ansatz table set");
    gwz_context* c = h->ctx;
    if (int rc = use_device(c)) return rc;
    const int P = h->P;
    std::vector<double> tab((size_t)h->dim * (1 + P));
    GWZ_CUDA(cudaMemcpy(tab.data(), h->d_tab, sizeof(double) * tab.size(), cudaMemcpyDeviceToHost));
    for (uint64_t k = 0; k < h->dim; ++k) {
        if (psi) psi[k] = tab[k * (1 + P)];
        if (grad_ratio)
            for (int p = 0; p < P; ++p) grad_ratio[k * P + p] = tab[k * (1 + P) + 1 + p];
    }
    return GWZ_OK;
}

static int sp_grid(const gwz_sparse* h, size_t units) {
    const size_t per_cta = SP_BLOCK / h->G;
    const size_t need = (units + per_cta - 1) / per_cta, resident = (size_t)h->ctx->prop.multiProcessorCount * 16;
    return (int)std::max<size_t>(1, std::min(need, resident));
}

// val_and_grad(le, params) / le(params) over rows [row_begin, row_end) (all rows: 0, dim); with a communicator the
// sums are all-reduced, so ranks may each take a slice of the rows
static int sp_sweep(gwz_sparse* h, uint64_t row_begin, uint64_t row_end, bool with_grad, bool allreduce,
                    double* sums /* host, 2+2P */, double* terms_host) {
    gwz_context* c = h->ctx;
    const int P = with_grad ? h->P : 0, L = 2 + 2 * P, ppl = sp_ppl(P, h->G);
    const uint64_t count = row_end - row_begin;
    const int grid = sp_grid(h, count);
    const size_t rows = (size_t)grid * (SP_BLOCK / h->G);
    if (rows * L > h->partials_cap) {
        if (h->d_partials) cudaFree(h->d_partials);
        h->d_partials = nullptr;
        h->partials_cap = 0;
        GWZ_CUDA(cudaMalloc(&h->d_partials, sizeof(double) * rows * L));
        h->partials_cap = rows * L;
    }
    DevBuf<double> d_out, d_terms;
    GWZ_CUDA(d_out.alloc(L));
    if (terms_host) GWZ_CUDA(d_terms.alloc(count * L));
    GWZ_CUDA(cudaEventRecord(c->ev0, c->stream));
    api_count_launch();
    GWZ_SP_DISPATCH(h->G, ppl, (sparse_sweep_kernel<G, PPL><<<grid, SP_BLOCK, 0, c->stream>>>(
                                   h->sd, row_begin, row_end, with_grad ? 1 : 0, h->d_partials, d_terms.p)));
    GWZ_CUDA(cudaGetLastError());
    GWZ_CUDA(cudaEventRecord(c->ev1, c->stream));
    GWZ_LAUNCH(launch_reduce_partials(h->d_partials, (int)rows, L, d_out.p, c->stream));
    if (c->comm && allreduce) {
        const NcclApi* api = nccl_api(nullptr);
        if (!api) return api_fail(GWZ_ERR_NCCL, "NCCL unavailable");
        GWZ_NCCL(api, api->AllReduce(d_out.p, d_out.p, L, kNcclFloat64, kNcclSum, c->comm, c->stream));
    }
    GWZ_CUDA(cudaMemcpyAsync(sums, d_out.p, sizeof(double) * L, cudaMemcpyDeviceToHost, c->stream));
    if (terms_host) GWZ_CUDA(cudaMemcpyAsync(terms_host, d_terms.p, sizeof(double) * count * L, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    GWZ_CUDA(cudaEventElapsedTime(&h->last_ms, c->ev0, c->ev1));
    return GWZ_OK;
}

extern "C" int gwz_sparse_sweep(gwz_sparse* h, uint64_t row_begin, uint64_t row_end, double* val, double* grad) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_sparse_sweep: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h && val, "gwz_sparse_sweep: NULL argument");
    GWZ_REQUIRE(h->table_version > 0, "gwz_sparse_sweep: no ansatz table set");
    GWZ_REQUIRE(row_begin < row_end && row_end <= h->dim, "gwz_sparse_sweep: bad row range");
    if (int rc = use_device(h->ctx)) return rc;
    const bool with_grad = grad != nullptr && h->P > 0;
    std::vector<double> s(2 + 2 * (size_t)h->P, 0.0);
    if (int rc = sp_sweep(h, row_begin, row_end, with_grad, true, s.data(), nullptr)) return rc;
    const double num = s[0], den = s[1];
    *val = num / den;  // localenergy.jl:124
    if (with_grad)
        for (int p = 0; p < h->P; ++p) grad[p] = (s[2 + p] * den - num * s[2 + h->P + p]) / (den * den);  // :125-128
    return GWZ_OK;
}

extern "C" int gwz_sparse_sweep_terms(gwz_sparse* h, uint64_t first, uint64_t count, double* out) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_sparse_sweep_terms: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h && out, "gwz_sparse_sweep_terms: NULL argument");
    GWZ_REQUIRE(h->table_version > 0, "gwz_sparse_sweep_terms: no ansatz table set");
    GWZ_REQUIRE(count >= 1 && first + count <= h->dim, "gwz_sparse_sweep_terms: bad row range");
    if (int rc = use_device(h->ctx)) return rc;
    std::vector<double> s(2 + 2 * (size_t)h->P, 0.0);
    return sp_sweep(h, first, first + count, true, false, s.data(), out);
}

extern "C" int gwz_sparse_last_kernel_ms(const gwz_sparse* h, float* ms) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_sparse_last_kernel_ms: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h && ms, "gwz_sparse_last_kernel_ms: NULL argument");
    *ms = h->last_ms;
    return GWZ_OK;
}

extern "C" int gwz_sparse_probe(gwz_sparse* h, const uint32_t* idx, size_t n, const double* u01, double* e_loc,
                                double* residence_time, double* grad_ratio, int32_t* n_hops, double* rates,
                                int32_t* chosen, uint32_t* next_idx) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_sparse_probe: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h && (n == 0 || idx), "gwz_sparse_probe: NULL argument");
    GWZ_REQUIRE(h->table_version > 0, "gwz_sparse_probe: no ansatz table set");
    if (n == 0) return GWZ_OK;
    for (size_t i = 0; i < n; ++i) GWZ_REQUIRE(idx[i] < h->dim, "gwz_sparse_probe: index outside the basis");
    gwz_context* c = h->ctx;
    if (int rc = use_device(c)) return rc;
    const size_t stride = (size_t)(h->max_row ? h->max_row : 1);
    DevBuf<uint32_t> d_idx, d_next;
    DevBuf<double> d_u, d_e, d_t, d_g, d_r;
    DevBuf<int32_t> d_nh, d_ch;
    GWZ_CUDA(d_idx.alloc(n));
    GWZ_CUDA(d_next.alloc(n));
    GWZ_CUDA(d_u.alloc(n));
    GWZ_CUDA(d_e.alloc(n));
    GWZ_CUDA(d_t.alloc(n));
    GWZ_CUDA(d_g.alloc(n * (size_t)(h->P ? h->P : 1)));
    GWZ_CUDA(d_nh.alloc(n));
    GWZ_CUDA(d_ch.alloc(n));
    if (rates) {
        GWZ_CUDA(d_r.alloc(n * stride));
        GWZ_CUDA(cudaMemsetAsync(d_r.p, 0, sizeof(double) * n * stride, c->stream));
    }
    GWZ_CUDA(cudaMemcpyAsync(d_idx.p, idx, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
    if (u01) GWZ_CUDA(cudaMemcpyAsync(d_u.p, u01, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    const int grid = sp_grid(h, n);
    api_count_launch();
    GWZ_SP_DISPATCH_BLK(h->sd.blk != nullptr, GWZ_SP_DISPATCH_G(h->G, (sparse_probe_kernel<G, BLK><<<grid, SP_BLOCK, 0, c->stream>>>(
                                h->sd, d_idx.p, n, u01 ? d_u.p : nullptr, d_e.p, d_t.p, d_g.p, d_nh.p, rates ? d_r.p : nullptr,
                                stride, d_ch.p, d_next.p, c->d_err))));
    GWZ_CUDA(cudaGetLastError());
    if (e_loc) GWZ_CUDA(cudaMemcpyAsync(e_loc, d_e.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (residence_time) GWZ_CUDA(cudaMemcpyAsync(residence_time, d_t.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (grad_ratio && h->P) GWZ_CUDA(cudaMemcpyAsync(grad_ratio, d_g.p, sizeof(double) * n * h->P, cudaMemcpyDeviceToHost, c->stream));
    if (n_hops) GWZ_CUDA(cudaMemcpyAsync(n_hops, d_nh.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    if (rates) GWZ_CUDA(cudaMemcpyAsync(rates, d_r.p, sizeof(double) * n * stride, cudaMemcpyDeviceToHost, c->stream));
    if (chosen) GWZ_CUDA(cudaMemcpyAsync(chosen, d_ch.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    if (next_idx) GWZ_CUDA(cudaMemcpyAsync(next_idx, d_next.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return api_check_err_flag(c);
}

// ---- walkers ---------------------------------------------------------------------------------
extern "C" int gwz_swalkers_destroy(gwz_swalkers* w) {
    if (w && w->magic != gwz::GWZ_TAG_SWALKERS) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_swalkers_destroy: not a gwz_swalkers handle");
    if (!w) return GWZ_OK;
    if (cudaSetDevice(w->device) == cudaSuccess) {
        cudaFree(w->d_idx);
        cudaFree(w->d_acc);
        cudaFree(w->d_partials);
        cudaFree(w->d_vec);
        if (w->h_vec) cudaFreeHost(w->h_vec);
    }
    w->magic = 0;
    delete w;
    return GWZ_OK;
}

extern "C" int gwz_swalkers_create(gwz_sparse* h, uint64_t n_walkers, uint64_t global_offset, uint32_t start_index,
                                   uint64_t seed, gwz_swalkers** out) {
    GWZ_HANDLE(h, gwz::GWZ_TAG_SPARSE, "gwz_swalkers_create: NULL or foreign handle (expected a gwz_sparse)");
    GWZ_REQUIRE(h && out, "gwz_swalkers_create: NULL argument");
    GWZ_REQUIRE(n_walkers >= 1, "gwz_swalkers_create: need at least one walker");
    GWZ_REQUIRE(start_index < h->dim, "gwz_swalkers_create: starting index outside the basis");
    gwz_context* c = h->ctx;
    if (int rc = use_device(c)) return rc;
    gwz_swalkers* w = new gwz_swalkers();
    HandleGuard<gwz_swalkers, gwz_swalkers_destroy> guard(w);
    w->h = h;
    w->device = h->device;
    w->n = (size_t)n_walkers;
    w->offset = global_offset;
    w->seed = seed;
    GWZ_CUDA(cudaMalloc(&w->d_idx, sizeof(uint32_t) * w->n));
    api_count_launch();
    sparse_fill_idx_kernel<<<(unsigned)std::min<size_t>((w->n + 255) / 256, 4096), 256, 0, c->stream>>>(w->d_idx, w->n, start_index);
    GWZ_CUDA(cudaGetLastError());
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    *out = guard.release();
    return GWZ_OK;
}

extern "C" int gwz_swalkers_get_indices(gwz_swalkers* w, uint32_t* idx) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_SWALKERS, "gwz_swalkers_get_indices: NULL or foreign handle (expected a gwz_swalkers)");
    GWZ_REQUIRE(w && idx, "gwz_swalkers_get_indices: NULL argument");
    if (int rc = use_device(w->h->ctx)) return rc;
    GWZ_CUDA(cudaMemcpy(idx, w->d_idx, sizeof(uint32_t) * w->n, cudaMemcpyDeviceToHost));
    return GWZ_OK;
}
extern "C" int gwz_swalkers_set_indices(gwz_swalkers* w, const uint32_t* idx) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_SWALKERS, "gwz_swalkers_set_indices: NULL or foreign handle (expected a gwz_swalkers)");
    GWZ_REQUIRE(w && idx, "gwz_swalkers_set_indices: NULL argument");
    for (size_t i = 0; i < w->n; ++i) GWZ_REQUIRE(idx[i] < w->h->dim, "gwz_swalkers_set_indices: index outside the basis");
    if (int rc = use_device(w->h->ctx)) return rc;
    GWZ_CUDA(cudaMemcpy(w->d_idx, idx, sizeof(uint32_t) * w->n, cudaMemcpyHostToDevice));
    return GWZ_OK;
}
extern "C" int gwz_swalkers_set_step_counter(gwz_swalkers* w, uint64_t steps_done) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_SWALKERS, "gwz_swalkers_set_step_counter: NULL or foreign handle (expected a gwz_swalkers)");
    GWZ_REQUIRE(w != nullptr, "gwz_swalkers_set_step_counter: NULL handle");
    w->steps_done = steps_done;
    return GWZ_OK;
}

static int sp_walkers_launch(gwz_swalkers* w, uint64_t steps, int accumulate, int reset_acc, double* rec_tau, double* rec_eloc,
                             uint32_t* rec_idx, double* hist) {
    gwz_sparse* h = w->h;
    gwz_context* c = h->ctx;
    const int grid = sp_grid(h, w->n), ppl = sp_ppl(h->P, h->G);
    api_count_launch();
    GWZ_SP_DISPATCH_BLK(h->sd.blk != nullptr, GWZ_SP_DISPATCH(h->G, ppl, (sparse_walker_kernel<G, PPL, BLK><<<grid, SP_BLOCK, 0, c->stream>>>(
                                   h->sd, philox_keys(w->seed), w->d_idx, w->d_acc, w->n, w->offset, w->steps_done, steps,
                                   accumulate, reset_acc, accumulate ? w->shift_e : 0.0, rec_tau, rec_eloc, rec_idx, hist,
                                   c->d_err))));
    GWZ_CUDA(cudaGetLastError());
    w->steps_done += steps;
    return GWZ_OK;
}

static int sp_run(gwz_swalkers* w, uint64_t steps, uint64_t burn_in, unsigned flags, double* rec_tau, double* rec_eloc,
                  uint32_t* rec_idx, double* hist_host, gwz_gen_result* out, double* grad, double* grad_err,
                  double* pooled_grad) {
    GWZ_REQUIRE(w != nullptr, "gwz_swalkers_run: NULL handle");
    GWZ_REQUIRE(steps >= 1, "gwz_swalkers_run: steps_per_walker must be >= 1");
    gwz_sparse* h = w->h;
    GWZ_REQUIRE(h->table_version > 0, "gwz_swalkers_run: no ansatz table set");
    gwz_context* c = h->ctx;
    if (int rc = use_device(c)) return rc;
    const int P = h->P, L = gen_vec_len(P);
    bool keep = (flags & GWZ_RUN_KEEP_ACC) != 0 && w->acc_steps > 0;
    if (w->P_acc != P) {
        GWZ_REQUIRE(!keep, "gwz_swalkers_run: GWZ_RUN_KEEP_ACC after the number of parameters changed");
        if (w->d_acc) cudaFree(w->d_acc);
        w->d_acc = nullptr;
        GWZ_CUDA(cudaMalloc(&w->d_acc, sizeof(double) * w->n * sp_acc_rows(P)));
        w->P_acc = P;
    }
    if ((size_t)L > w->vec_cap) {
        if (w->d_vec) cudaFree(w->d_vec);
        if (w->h_vec) cudaFreeHost(w->h_vec);
        w->d_vec = w->h_vec = nullptr;
        GWZ_CUDA(cudaMalloc(&w->d_vec, sizeof(double) * L));
        GWZ_CUDA(cudaMallocHost(&w->h_vec, sizeof(double) * L));
        w->vec_cap = L;
    }
    const int fgrid = (int)std::max<size_t>(1, std::min<size_t>((w->n + SP_BLOCK - 1) / SP_BLOCK, (size_t)c->prop.multiProcessorCount * 4));
    if ((size_t)fgrid * L > w->partials_cap) {
        if (w->d_partials) cudaFree(w->d_partials);
        w->d_partials = nullptr;
        GWZ_CUDA(cudaMalloc(&w->d_partials, sizeof(double) * fgrid * L));
        w->partials_cap = (size_t)fgrid * L;
    }
    if (!keep && w->table_seen != h->table_version) {
        // reference point of the shifted sums: the Rayleigh quotient of the current table
        std::vector<double> s(2, 0.0);
        if (int rc = sp_sweep(h, 0, h->dim, false, false, s.data(), nullptr)) return rc;  // every rank sweeps all rows
        w->shift_e = s[0] / s[1];
        if (!std::isfinite(w->shift_e)) w->shift_e = 0.0;
        w->table_seen = h->table_version;
    }
    GWZ_CUDA(cudaEventRecord(c->ev0, c->stream));
    if (burn_in > 0)
        if (int rc = sp_walkers_launch(w, burn_in, 0, 0, nullptr, nullptr, nullptr, nullptr)) return rc;
    const size_t cells = (size_t)steps * w->n;
    DevBuf<double> d_tau, d_el, d_hist;
    DevBuf<uint32_t> d_ri;
    if (rec_tau) {
        GWZ_CUDA(d_tau.alloc(cells));
        GWZ_CUDA(d_el.alloc(cells));
        if (rec_idx) GWZ_CUDA(d_ri.alloc(cells));
    }
    if (hist_host) {
        GWZ_CUDA(d_hist.alloc(h->dim));
        GWZ_CUDA(cudaMemsetAsync(d_hist.p, 0, sizeof(double) * h->dim, c->stream));
    }
    if (int rc = sp_walkers_launch(w, steps, 1, keep ? 0 : 1, rec_tau ? d_tau.p : nullptr, rec_tau ? d_el.p : nullptr,
                                   rec_idx ? d_ri.p : nullptr, hist_host ? d_hist.p : nullptr))
        return rc;
    GWZ_CUDA(cudaEventRecord(c->ev1, c->stream));
    w->acc_steps = keep ? w->acc_steps + steps : steps;
    api_count_launch();
    sparse_finalize_kernel<<<fgrid, SP_BLOCK, 0, c->stream>>>(w->d_acc, w->n, P, w->shift_e, (double)w->acc_steps, w->d_partials);
    GWZ_CUDA(cudaGetLastError());
    GWZ_LAUNCH(launch_reduce_partials(w->d_partials, fgrid, L, w->d_vec, c->stream));
    if (c->comm) {
        const NcclApi* api = nccl_api(nullptr);
        if (!api) return api_fail(GWZ_ERR_NCCL, "NCCL unavailable");
        // one ncclAllReduce(sum, float64, 8 + 4P) per run, on the sampling stream
        GWZ_NCCL(api, api->AllReduce(w->d_vec, w->d_vec, L, kNcclFloat64, kNcclSum, c->comm, c->stream));
        if (hist_host) GWZ_NCCL(api, api->AllReduce(d_hist.p, d_hist.p, h->dim, kNcclFloat64, kNcclSum, c->comm, c->stream));
    }
    GWZ_CUDA(cudaMemcpyAsync(w->h_vec, w->d_vec, sizeof(double) * L, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (rec_tau) {
        GWZ_CUDA(cudaMemcpyAsync(rec_tau, d_tau.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
        GWZ_CUDA(cudaMemcpyAsync(rec_eloc, d_el.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, c->stream));
        if (rec_idx) GWZ_CUDA(cudaMemcpyAsync(rec_idx, d_ri.p, sizeof(uint32_t) * cells, cudaMemcpyDeviceToHost, c->stream));
    }
    if (hist_host) GWZ_CUDA(cudaMemcpyAsync(hist_host, d_hist.p, sizeof(double) * h->dim, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    if (int rc = api_check_err_flag(c)) return rc;
    float ms = 0.f;
    GWZ_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double* v = w->h_vec;
    const double nw = v[GS_WALKERS], T = v[GS_T];
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const double val = v[GS_VAL] / nw, pe = v[GS_TE] / T;
    const bool pooled = (flags & GWZ_RUN_POOLED) != 0;
    if (out) {
        out->val = pooled ? pe : val;
        out->err = nw > 1.5 ? std::sqrt(std::fmax((v[GS_VAL2] / nw - val * val) * nw / (nw - 1.0), 0.0) / nw) : nan;
        out->pooled_val = pe;
        out->pooled_var = std::fmax(v[GS_TEE] / T - pe * pe, 0.0);
        out->total_time = T;
        out->walkers = (uint64_t)(nw + 0.5);
        out->steps = (uint64_t)(v[GS_STEPS] + 0.5);
        out->hops = (uint64_t)(v[GS_HOPS] + 0.5);
        out->kernel_ms = ms;
        out->np = P;
    }
    for (int p = 0; p < P; ++p) {
        const double g = v[GEN_HEAD + p] / nw;
        const double pg = 2.0 * (v[GEN_HEAD + 3 * P + p] / T - pe * (v[GEN_HEAD + 2 * P + p] / T));
        if (grad) grad[p] = pooled ? pg : g;
        if (pooled_grad) pooled_grad[p] = pg;
        if (grad_err)
            grad_err[p] = nw > 1.5 ? std::sqrt(std::fmax((v[GEN_HEAD + P + p] / nw - g * g) * nw / (nw - 1.0), 0.0) / nw) : nan;
    }
    return GWZ_OK;
}

extern "C" int gwz_swalkers_run(gwz_swalkers* w, uint64_t steps_per_walker, uint64_t burn_in, unsigned flags,
                                gwz_gen_result* out, double* grad, double* grad_err, double* pooled_grad) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_SWALKERS, "gwz_swalkers_run: NULL or foreign handle (expected a gwz_swalkers)");
    return sp_run(w, steps_per_walker, burn_in, flags, nullptr, nullptr, nullptr, nullptr, out, grad, grad_err, pooled_grad);
}
extern "C" int gwz_swalkers_run_record(gwz_swalkers* w, uint64_t steps_per_walker, unsigned flags, double* tau, double* eloc,
                                       uint32_t* idx, gwz_gen_result* out, double* grad, double* grad_err,
                                       double* pooled_grad) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_SWALKERS, "gwz_swalkers_run_record: NULL or foreign handle (expected a gwz_swalkers)");
    GWZ_REQUIRE(tau && eloc, "gwz_swalkers_run_record: NULL series buffer");
    return sp_run(w, steps_per_walker, 0, flags, tau, eloc, idx, nullptr, out, grad, grad_err, pooled_grad);
}
// DVec(res) (state.jl:303-319): total residence time per basis index, hist[dim]
extern "C" int gwz_swalkers_sample_vector(gwz_swalkers* w, uint64_t steps_per_walker, uint64_t burn_in, unsigned flags,
                                          double* hist, gwz_gen_result* out) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_SWALKERS, "gwz_swalkers_sample_vector: NULL or foreign handle (expected a gwz_swalkers)");
    GWZ_REQUIRE(hist != nullptr, "gwz_swalkers_sample_vector: NULL histogram buffer");
    return sp_run(w, steps_per_walker, burn_in, flags, nullptr, nullptr, nullptr, hist, out, nullptr, nullptr, nullptr);
}
// per-walker estimates (state.jl:42-49) of the steps accumulated so far
extern "C" int gwz_swalkers_estimates(gwz_swalkers* w, double* val_w, double* grad_w) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_SWALKERS, "gwz_swalkers_estimates: NULL or foreign handle (expected a gwz_swalkers)");
    GWZ_REQUIRE(w != nullptr, "gwz_swalkers_estimates: NULL handle");
    GWZ_REQUIRE(w->acc_steps > 0 && w->d_acc, "gwz_swalkers_estimates: no steps accumulated yet");
    if (int rc = use_device(w->h->ctx)) return rc;
    const int P = w->P_acc;
    const size_t n = w->n;
    std::vector<double> a(n * sp_acc_rows(P));
    GWZ_CUDA(cudaMemcpy(a.data(), w->d_acc, sizeof(double) * a.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) {
        const double st = a[i], ste = a[n + i];
        if (val_w) val_w[i] = w->shift_e + ste / st;
        if (grad_w)
            for (int p = 0; p < P; ++p)
                grad_w[i * P + p] = 2.0 * (a[(size_t)(4 + P + p) * n + i] / st - (ste / st) * (a[(size_t)(4 + p) * n + i] / st));
    }
    return GWZ_OK;
}
