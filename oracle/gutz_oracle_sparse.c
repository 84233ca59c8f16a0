/*
 * gutz_oracle_sparse.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE), part 3.
 *
 * The hot path of the reference for ANY Hamiltonian and ANY ansatz, restated over a materialised operator:
 * the basis is `build_basis(H)` (an enumerated list of addresses), row k of the operator holds
 * `offdiagonals(H, basis[k])` in the reference's own order as (column index, matrix element) pairs, diag[k] is
 * `diagonal_element(H, basis[k])`, and the ansatz is a table psi[k] = ansatz(basis[k], params) with
 * grad_ratio[k][p] = (d psi / d p_p) / psi at basis[k].  This is how HubbardMom1D, HubbardRealSpace, ... of
 * test/qmc.jl:10-14 and test/ansatz.jl:44-49 reach the path: only Rimu's operator code
 * differs, and that is replaced by the stored rows.
 *
 * Each function follows the cited reference lines literally (sequential sums, the reference's prob_buffer and
 * pick_random_from_cumsum).  PARITY PIN: for HubbardReal1D + GutzwillerAnsatz the rows are generated from part 1 of
 * this oracle (pinned against README.md:94,104) and the two restatements must agree bit for bit
 * (tests/test_oracle_sparse.py); for other operators the reference holds no literal vectors: "parity unpinned"
 * beyond that cross-check, Hermiticity/spectrum checks and exact diagonalisation in the tests.
 *
 * All file:line citations are relative to /root/reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gutz_oracle.h"

/* kinetic_sample! (src/kinetic-vqmc/qmc.jl:14-44) at basis index k; prob_buffer holds row-length doubles.
 * Returns 0, or 1 on a non-finite residence time (qmc.jl:34-36). */
int go_sparse_kinetic_sample(const uint64_t *row_ptr, const uint32_t *col, const double *melem, const double *diag,
                             const double *psi, const double *grad_ratio, int P, uint32_t k, double u01,
                             double *prob_buffer, uint32_t *new_index, double *residence_time, double *local_energy,
                             double *gradient, int *chosen, int *n_hops) {
    const uint64_t b = row_ptr[k], e = row_ptr[k + 1];
    const int h = (int)(e - b);
    const double val1 = psi[k];
    double local_energy_num = diag[k] * val1; /* qmc.jl:20-21 */
    double residence_time_denom = 0.0;
    for (int i = 0; i < h; ++i) { /* qmc.jl:25-31 */
        const double val2 = psi[col[b + i]];
        residence_time_denom += fabs(val2);
        local_energy_num += melem[b + i] * val2;
        prob_buffer[i] = residence_time_denom;
    }
    const double tau = fabs(val1) / residence_time_denom; /* qmc.jl:33 */
    *residence_time = tau;
    *local_energy = local_energy_num / val1; /* qmc.jl:37 */
    for (int p = 0; p < P; ++p) gradient[p] = grad_ratio[(size_t)k * P + p]; /* qmc.jl:38 */
    *n_hops = h;
    if (!isfinite(tau)) return 1;
    /* pick_random_from_cumsum (qmc.jl:52-61) */
    const double target = u01 * prob_buffer[h - 1];
    int i = 0;
    while (!(target < prob_buffer[i])) ++i;
    *chosen = i + 1;
    *new_index = col[b + i];
    return 0;
}

/* kinetic_vqmc! (qmc.jl:73-106) for n_walkers walkers over the stored operator; random numbers as go_kinetic_vqmc
 * (go_u01(seed, walker0 + w, step0 + k)).  idx[n_walkers] in/out.  Series (optional) are [walker][step];
 * val_w/grad_w: per-walker estimates (state.jl:42-49).  Returns 0 or 1. */
int go_sparse_kinetic_vqmc(const uint64_t *row_ptr, const uint32_t *col, const double *melem, const double *diag,
                           const double *psi, const double *grad_ratio, int P, uint32_t *idx, int64_t n_walkers,
                           int64_t steps, uint64_t seed, uint64_t walker0, uint64_t step0, uint32_t *series_idx,
                           double *series_tau, double *series_eloc, double *val_w, double *grad_w, int64_t *total_hops,
                           uint64_t dim) {
    uint64_t maxrow = 1;
    for (uint64_t k = 0; k < dim; ++k)
        if (row_ptr[k + 1] - row_ptr[k] > maxrow) maxrow = row_ptr[k + 1] - row_ptr[k];
    int err = 0;
    int64_t hops = 0;
#pragma omp parallel for schedule(static) reduction(| : err) reduction(+ : hops)
    for (int64_t w = 0; w < n_walkers; ++w) {
        double *prob_buffer = (double *)malloc(sizeof(double) * maxrow);
        double *tau = (double *)malloc(sizeof(double) * (size_t)steps);
        double *el = (double *)malloc(sizeof(double) * (size_t)steps);
        double *gr = (double *)malloc(sizeof(double) * (size_t)steps * (P > 0 ? P : 1));
        uint32_t cur = idx[w];
        for (int64_t s = 0; s < steps; ++s) {
            const double u01 = go_u01(seed, walker0 + (uint64_t)w, step0 + (uint64_t)s);
            uint32_t nxt = cur;
            int chosen = 0, nh = 0;
            if (series_idx) series_idx[w * steps + s] = cur; /* qmc.jl:89 stores the address sampled from */
            err |= go_sparse_kinetic_sample(row_ptr, col, melem, diag, psi, grad_ratio, P, cur, u01, prob_buffer, &nxt,
                                            &tau[s], &el[s], gr + (size_t)s * P, &chosen, &nh);
            hops += nh;
            cur = nxt;
        }
        idx[w] = cur;
        if (series_tau) memcpy(series_tau + w * steps, tau, sizeof(double) * (size_t)steps);
        if (series_eloc) memcpy(series_eloc + w * steps, el, sizeof(double) * (size_t)steps);
        /* state.jl:42-49: val = sum(tau .* E) / sum(tau); grad = 2 (sum(tau G E) / T - val sum(tau G) / T) */
        double T = 0.0, TE = 0.0;
        for (int64_t s = 0; s < steps; ++s) {
            T += tau[s];
            TE += tau[s] * el[s];
        }
        const double val = TE / T;
        if (val_w) val_w[w] = val;
        for (int p = 0; p < P; ++p) {
            double TG = 0.0, TGE = 0.0;
            for (int64_t s = 0; s < steps; ++s) {
                const double g = gr[(size_t)s * P + p];
                TG += tau[s] * g;
                TGE += tau[s] * g * el[s];
            }
            if (grad_w) grad_w[(size_t)w * P + p] = 2.0 * (TGE / T - val * TG / T);
        }
        free(prob_buffer);
        free(tau);
        free(el);
        free(gr);
    }
    if (total_hops) *total_hops = hops;
    return err;
}

/* LocalEnergyEvaluator over the stored operator: per-row terms of localenergy.jl:105-120,
 * out = (num, den, num_grad[P], den_grad[P]) */
void go_sparse_lee_terms(const uint64_t *row_ptr, const uint32_t *col, const double *melem, const double *diag,
                         const double *psi, const double *grad_ratio, int P, uint64_t k, double *out) {
    const double k_val = psi[k];
    const double dg = diag[k];
    double num = k_val * k_val * dg;
    const double den = k_val * k_val;
    double *num_grad = out + 2, *den_grad = out + 2 + P;
    for (int p = 0; p < P; ++p) {
        const double k_grad = grad_ratio[(size_t)k * P + p] * k_val;
        num_grad[p] = 2.0 * k_grad * k_val * dg;
        den_grad[p] = 2.0 * k_val * k_grad;
    }
    for (uint64_t i = row_ptr[k]; i < row_ptr[k + 1]; ++i) {
        const uint32_t l = col[i];
        const double l_val = psi[l], m = melem[i];
        num += l_val * m * k_val;
        for (int p = 0; p < P; ++p) {
            const double k_grad = grad_ratio[(size_t)k * P + p] * k_val;
            const double l_grad = grad_ratio[(size_t)l * P + p] * l_val;
            num_grad[p] += m * (l_val * k_grad + l_grad * k_val);
        }
    }
    out[0] = num;
    out[1] = den;
}

/* val_and_grad(le, params) (localenergy.jl:94-129): long double accumulation of the per-row terms (the parity
 * target of the device reduction, like mode 1 of go_lee_val_and_grad) */
void go_sparse_lee_val_and_grad(const uint64_t *row_ptr, const uint32_t *col, const double *melem, const double *diag,
                                const double *psi, const double *grad_ratio, int P, uint64_t dim, double *val,
                                double *grad) {
    long double num = 0.0L, den = 0.0L;
    long double *ng = (long double *)calloc((size_t)(2 * P + 1), sizeof(long double));
    double *t = (double *)malloc(sizeof(double) * (size_t)(2 + 2 * P));
    for (uint64_t k = 0; k < dim; ++k) {
        go_sparse_lee_terms(row_ptr, col, melem, diag, psi, grad_ratio, P, k, t);
        num += t[0];
        den += t[1];
        for (int p = 0; p < 2 * P; ++p) ng[p] += t[2 + p];
    }
    *val = (double)(num / den);
    for (int p = 0; p < P; ++p) grad[p] = (double)((ng[p] * den - num * ng[P + p]) / (den * den));
    free(ng);
    free(t);
}
