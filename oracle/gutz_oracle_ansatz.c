/*
 * gutz_oracle_ansatz.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE), part 2.
 *
 * Literal C restatement of the reference's other single-component ansaetze and of the
 * ansatz-generic forms of kinetic_sample! / LocalEnergyEvaluator:
 *     src/ansatz/gutzwiller.jl:25-36      GutzwillerAnsatz       exp(-g * diagonal_element(H, addr))
 *     src/ansatz/extgutzwiller.jl:31-35   ExtendedGutzwillerAnsatz (call operator; see note)
 *     src/ansatz/multinomial.jl:33-59     MultinomialAnsatz
 *     src/ansatz/densityprofile.jl:20-28  DensityProfileAnsatz
 *     src/ansatz/jastrow.jl:26-58         JastrowAnsatz
 *     src/ansatz/jastrow.jl:84-103        RelativeJastrowAnsatz
 *     src/kinetic-vqmc/qmc.jl:14-61       kinetic_sample! (every neighbour amplitude evaluated in full)
 *     src/localenergy.jl:94-155           LocalEnergyEvaluator val_and_grad / value
 * for HubbardReal1D and Rimu's ExtendedHubbardReal1D (diagonal u*sum n(n-1)/2 + v*sum n_i n_{i+1},
 * periodic; Rimu `extended_hubbard_interaction`).
 *
 * Like the reference, nothing here uses closed-form amplitude ratios: psi(addr2) is evaluated
 * from the occupation numbers of addr2 for every hop.  That is the point -- the GPU engine works
 * with ratios and potentials, and is checked against this.
 *
 * Note on ExtendedGutzwillerAnsatz: the reference's val_and_grad (extgutzwiller.jl:22-29) refers
 * to an undefined name and cannot run (its test is commented out, test/ansatz.jl:52); the call
 * operator (extgutzwiller.jl:31-35) is well defined: exp(-g1*sum n(n-1)/2 - g2*sum n_i n_{i+1}),
 * without the factors u, v.  The oracle follows the call operator and uses its analytic gradient.
 *
 * PARITY PIN: no literal vectors exist in the reference for these ansaetze (its tests compare
 * against ForwardDiff / rayleigh_quotient live).  tests/test_oracle_ansatz.py pins this file the
 * same way: central finite differences for every gradient, a dense-matrix Rayleigh quotient for
 * every LocalEnergyEvaluator value, the u = 0 exactness of the multinomial ansatz
 * (test/ansatz.jl:66-80), and equality with part 1 of the oracle (itself pinned bit-for-bit to the
 * README) for GutzwillerAnsatz.  "parity unpinned" beyond those.
 */
#include "gutz_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GO_MAXM 256

int go_ansatz_init(go_ansatz *a, const go_model *m, double v, int kind, int normalize) {
    a->m = *m;
    a->v = v;
    a->kind = kind;
    a->kind1 = a->kind2 = -1;
    a->norm1 = a->norm2 = 0;
    const int M = m->M;
    switch (kind) {
        case GO_ANSATZ_GUTZWILLER: a->P = 1; break;
        case GO_ANSATZ_EXT_GUTZWILLER: a->P = 2; break;
        case GO_ANSATZ_MULTINOMIAL: a->P = 1; break;
        case GO_ANSATZ_DENSITY_PROFILE: a->P = M; break;              /* densityprofile.jl:15 */
        case GO_ANSATZ_RELATIVE_JASTROW: a->P = (M + 1) / 2; break;   /* jastrow.jl:78 cld(M, 2) */
        case GO_ANSATZ_JASTROW: a->P = M * (M + 1) / 2; break;        /* jastrow.jl:20 */
        default: return 1;
    }
    /* multinomial.jl:29-31: gamma(N + 1) / M^N */
    a->normalization = normalize ? tgamma((double)m->N + 1.0) / pow((double)M, (double)m->N) : 1.0;
    return 0;
}

int go_ansatz_init_combination(go_ansatz *a, const go_model *m, double v, int kind1, int normalize1, int kind2,
                               int normalize2) {
    go_ansatz c1, c2;
    if (go_ansatz_init(&c1, m, v, kind1, normalize1) || go_ansatz_init(&c2, m, v, kind2, normalize2)) return 1;
    a->m = *m;
    a->v = v;
    a->kind = GO_ANSATZ_COMBINATION;
    a->P = c1.P + c2.P + 1; /* combination.jl:14 */
    a->normalization = 1.0;
    a->kind1 = kind1;
    a->kind2 = kind2;
    a->norm1 = normalize1;
    a->norm2 = normalize2;
    return 0;
}

/* Rimu extended_hubbard_interaction: (sum_i n_i n_{i+1} periodic, sum_i n_i (n_i - 1) / 2) */
static void ext_interaction(const int *o, int M, long *ext, long *reg) {
    long e = 0, r = 0;
    for (int i = 0; i < M; ++i) {
        e += (long)o[i] * o[(i + 1) % M];
        r += (long)o[i] * (o[i] - 1);
    }
    *ext = e;
    *reg = r / 2;
}

double go_gen_diagonal(const go_ansatz *a, const go_addr *addr) {
    int o[GO_MAXM];
    go_to_onr(&a->m, addr, o);
    long ext, reg;
    ext_interaction(o, a->m.M, &ext, &reg);
    return a->m.u * (double)reg + a->v * (double)ext;
}

/* val_and_grad(ansatz, addr, params); grad may be NULL (the call operator) */
void go_gen_val_and_grad(const go_ansatz *a, const go_addr *addr, const double *p, double *val, double *grad) {
    const int M = a->m.M;
    int o[GO_MAXM];
    go_to_onr(&a->m, addr, o);
    switch (a->kind) {
        case GO_ANSATZ_GUTZWILLER: { /* gutzwiller.jl:25-36 */
            const double diag = go_gen_diagonal(a, addr);
            const double v = exp(-p[0] * diag);
            *val = v;
            if (grad) grad[0] = -diag * v;
            break;
        }
        case GO_ANSATZ_EXT_GUTZWILLER: { /* extgutzwiller.jl:31-35 */
            long ext, reg;
            ext_interaction(o, M, &ext, &reg);
            const double v = exp(-p[0] * (double)reg + -p[1] * (double)ext);
            *val = v;
            if (grad) {
                grad[0] = -(double)reg * v;
                grad[1] = -(double)ext * v;
            }
            break;
        }
        case GO_ANSATZ_MULTINOMIAL: { /* multinomial.jl:39-59 */
            double w = 1.0;
            for (int i = 0; i < M; ++i)
                if (o[i] > 0) w *= 1.0 / tgamma((double)o[i] + 1.0);
            const double y = w * a->normalization;
            const double v = pow(y, p[0]);
            *val = v;
            if (grad) grad[0] = v * log(y);
            break;
        }
        case GO_ANSATZ_DENSITY_PROFILE: { /* densityprofile.jl:20-28 */
            double e = 0.0;
            for (int i = 0; i < M; ++i) e += (double)o[i] * p[i];
            const double v = exp(-e);
            *val = v;
            if (grad)
                for (int i = 0; i < M; ++i) grad[i] = -v * (double)o[i];
            break;
        }
        case GO_ANSATZ_RELATIVE_JASTROW: { /* jastrow.jl:84-103; circshift_dot(o, o, d) */
            double e = 0.0;
            double prod[GO_MAXM];
            for (int d = 0; d < a->P; ++d) {
                long s = 0;
                for (int i = 0; i < M; ++i) s += (long)o[i] * o[(i + d) % M];
                prod[d] = (double)s;
                e += p[d] * prod[d];
            }
            const double v = exp(-e);
            *val = v;
            if (grad)
                for (int d = 0; d < a->P; ++d) grad[d] = -(prod[d] * v);
            break;
        }
        case GO_ANSATZ_JASTROW: { /* jastrow.jl:26-58 */
            double e = 0.0;
            int k = 0;
            for (int i = 0; i < M; ++i)
                for (int j = i; j < M; ++j, ++k) e += p[k] * (double)(o[i] * o[j]);
            const double v = exp(-e);
            *val = v;
            if (grad) {
                k = 0;
                for (int i = 0; i < M; ++i)
                    for (int j = i; j < M; ++j, ++k) grad[k] = -(double)(o[i] * o[j]) * v;
            }
            break;
        }
        case GO_ANSATZ_COMBINATION: { /* combination.jl:17-50 */
            go_ansatz c1, c2;
            go_ansatz_init(&c1, &a->m, a->v, a->kind1, a->norm1);
            go_ansatz_init(&c2, &a->m, a->v, a->kind2, a->norm2);
            const double alpha = p[a->P - 1];
            double v1, v2;
            double *g1 = grad ? grad : NULL, *g2 = grad ? grad + c1.P : NULL;
            go_gen_val_and_grad(&c1, addr, p, &v1, g1);
            go_gen_val_and_grad(&c2, addr, p + c1.P, &v2, g2);
            *val = alpha * v1 + (1 - alpha) * v2;
            if (grad) {
                for (int k = 0; k < c1.P; ++k) g1[k] *= alpha;
                for (int k = 0; k < c2.P; ++k) g2[k] *= (1 - alpha);
                grad[a->P - 1] = v1 - v2;
            }
            break;
        }
        default: *val = 0.0;
    }
}

/* qmc.jl:14-61 */
int go_gen_kinetic_sample(const go_ansatz *a, const double *params, const go_addr *a1, double u01,
                          double *prob_buffer, go_addr *new_addr, double *residence_time, double *local_energy,
                          double *grad_ratio, int *chosen_out, int *n_hops) {
    const go_model *m = &a->m;
    const int nh = go_num_offdiagonals(m, a1);
    double val1;
    double *grad1 = (double *)malloc(sizeof(double) * (size_t)a->P);
    go_gen_val_and_grad(a, a1, params, &val1, grad1);
    const double diag = go_gen_diagonal(a, a1);
    double num = diag * val1;
    double denom = 0.0;
    for (int i = 1; i <= nh; ++i) {
        go_addr a2;
        const double melem = go_get_offdiagonal(m, a1, i, &a2);
        double val2;
        go_gen_val_and_grad(a, &a2, params, &val2, NULL);
        denom += fabs(val2);
        num += melem * val2;
        prob_buffer[i - 1] = denom;
    }
    const double tau = fabs(val1) / denom;
    *residence_time = tau;
    *local_energy = num / val1;
    for (int k = 0; k < a->P; ++k) grad_ratio[k] = grad1[k] / val1;
    free(grad1);
    if (n_hops) *n_hops = nh;
    if (!isfinite(tau)) return 1;
    /* pick_random_from_cumsum (clamped to the last hop) */
    const double c = u01 * prob_buffer[nh - 1];
    int i = 1;
    while (i < nh && !(c < prob_buffer[i - 1])) ++i;
    if (chosen_out) *chosen_out = i;
    go_get_offdiagonal(m, a1, i, new_addr);
    return 0;
}

int go_gen_hop_rates(const go_ansatz *a, const double *params, const go_addr *a1, double *rates) {
    const go_model *m = &a->m;
    const int nh = go_num_offdiagonals(m, a1);
    double val1;
    go_gen_val_and_grad(a, a1, params, &val1, NULL);
    for (int i = 1; i <= nh; ++i) {
        go_addr a2;
        go_get_offdiagonal(m, a1, i, &a2);
        double val2;
        go_gen_val_and_grad(a, &a2, params, &val2, NULL);
        rates[i - 1] = fabs(val2) / fabs(val1);
    }
    return nh;
}

/* localenergy.jl:105-120 for one state: out = (num, den, num_grad[P], den_grad[P]) */
void go_gen_lee_terms(const go_ansatz *a, const go_addr *k, const double *params, double *out) {
    const go_model *m = &a->m;
    const int P = a->P;
    double *k_grad = (double *)malloc(sizeof(double) * 2 * (size_t)P);
    double *l_grad = k_grad + P;
    double k_val;
    go_gen_val_and_grad(a, k, params, &k_val, k_grad);
    const double diag = go_gen_diagonal(a, k);
    double num = (k_val * k_val) * diag;
    double *num_grad = out + 2, *den_grad = out + 2 + P;
    for (int q = 0; q < P; ++q) {
        num_grad[q] = 2 * k_grad[q] * k_val * diag;
        den_grad[q] = 2 * k_val * k_grad[q];
    }
    const int nh = go_num_offdiagonals(m, k);
    for (int i = 1; i <= nh; ++i) {
        go_addr l;
        const double melem = go_get_offdiagonal(m, k, i, &l);
        double l_val;
        go_gen_val_and_grad(a, &l, params, &l_val, l_grad);
        num += l_val * melem * k_val;
        for (int q = 0; q < P; ++q) num_grad[q] += melem * (l_val * k_grad[q] + l_grad[q] * k_val);
    }
    out[0] = num;
    out[1] = k_val * k_val;
    free(k_grad);
}

/* localenergy.jl:94-129; sums in long double (exactly-rounded-ish: the parity target) */
void go_gen_lee_val_and_grad(const go_ansatz *a, const go_addr *basis, int64_t dim, const double *params,
                             double *val, double *grad) {
    const int P = a->P;
    long double num = 0, den = 0;
    long double *ng = (long double *)calloc(2 * (size_t)P, sizeof(long double));
    long double *dg = ng + P;
    double *t = (double *)malloc(sizeof(double) * (2 + 2 * (size_t)P));
    for (int64_t i = 0; i < dim; ++i) {
        go_gen_lee_terms(a, &basis[i], params, t);
        num += t[0];
        den += t[1];
        for (int q = 0; q < P; ++q) {
            ng[q] += t[2 + q];
            dg[q] += t[2 + P + q];
        }
    }
    *val = (double)(num / den);
    if (grad)
        for (int q = 0; q < P; ++q) grad[q] = (double)((ng[q] * den - num * dg[q]) / (den * den));
    free(ng);
    free(t);
}

/* kinetic_vqmc! (qmc.jl:73-106) + per-walker estimates (state.jl:42-49) with the Philox stream of
 * part 1 (go_u01).  val_w[n_walkers], grad_w[n_walkers x P]; series (optional) [walker][step]. */
int go_gen_kinetic_vqmc(const go_ansatz *a, const double *params, go_addr *addrs, int64_t n_walkers, int64_t steps,
                        uint64_t seed, uint64_t walker0, uint64_t step0, int n_threads, double *series_tau,
                        double *series_eloc, go_addr *series_addr, double *val_w, double *grad_w,
                        int64_t *total_hops) {
    const int P = a->P;
    int rc = 0;
    int64_t hops = 0;
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : rc) reduction(+ : hops) num_threads(n_threads > 0 ? n_threads : 1)
    for (int64_t w = 0; w < n_walkers; ++w) {
        double *tau = (double *)malloc(sizeof(double) * (size_t)steps * (2 + (size_t)P));
        double *el = tau + steps;
        double *gr = el + steps;
        double *buf = (double *)malloc(sizeof(double) * 2 * (size_t)a->m.M);
        go_addr cur = addrs[w];
        for (int64_t k = 0; k < steps; ++k) {
            go_addr nxt;
            int nh = 0;
            const double u01 = go_u01(seed, walker0 + (uint64_t)w, step0 + (uint64_t)k);
            rc |= go_gen_kinetic_sample(a, params, &cur, u01, buf, &nxt, &tau[k], &el[k], &gr[k * P], NULL, &nh);
            hops += nh;
            if (series_addr) series_addr[w * steps + k] = cur;
            cur = nxt;
        }
        addrs[w] = cur;
        /* state.jl:42-49 */
        double st = 0, ste = 0;
        for (int64_t k = 0; k < steps; ++k) {
            st += tau[k];
            ste += tau[k] * el[k];
        }
        const double val = ste / st;
        if (val_w) val_w[w] = val;
        if (grad_w)
            for (int q = 0; q < P; ++q) {
                double s = 0;
                for (int64_t k = 0; k < steps; ++k) s += tau[k] * gr[k * P + q] * (el[k] - val);
                grad_w[w * P + q] = 2 * s / st;
            }
        if (series_tau) memcpy(series_tau + w * steps, tau, sizeof(double) * (size_t)steps);
        if (series_eloc) memcpy(series_eloc + w * steps, el, sizeof(double) * (size_t)steps);
        free(tau);
        free(buf);
    }
    if (total_hops) *total_hops = hops;
    return rc;
}

/* ---------------------------------------------------------------------------------------------
 * "Optimised CPU" bound for the Gutzwiller path: NOT a restatement of the reference but
 * what a tuned CPU code could do -- occupation arrays, closed-form amplitude ratios exp(-g u Delta)
 * from a table, sqrt table, incrementally carried diagonal.  Same hop order, same Philox stream and
 * the same selection rule as go_kinetic_vqmc, so both produce the same trajectories (up to the last
 * bits of the rates); it is timed next to the faithful port as an honest upper bound of the CPU side
 * and doubles as a cross-check of the port.
 * ------------------------------------------------------------------------------------------- */
int go_kinetic_vqmc_fast(const go_model *m, double g, go_addr *addrs, int64_t n_walkers, int64_t steps, uint64_t seed,
                         uint64_t walker0, uint64_t step0, int n_threads, double *val_w, double *grad_w,
                         int64_t *total_hops) {
    const int M = m->M, N = m->N;
    double *exptab = (double *)malloc(sizeof(double) * (size_t)(2 * N + 3));
    double *sqrtab = (double *)malloc(sizeof(double) * (size_t)(N + 2));
    for (int i = 0; i < 2 * N + 3; ++i) exptab[i] = exp(-g * m->u * (double)(i - N));
    for (int i = 0; i < N + 2; ++i) sqrtab[i] = sqrt((double)i);
    int rc = 0;
    int64_t hops = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(| : rc) reduction(+ : hops) num_threads(n_threads > 0 ? n_threads : 1)
    for (int64_t w = 0; w < n_walkers; ++w) {
        int n[GO_MAXM];
        double cum[2 * GO_MAXM];
        int src[2 * GO_MAXM];
        go_to_onr(m, &addrs[w], n);
        long d = 0;
        for (int k = 0; k < M; ++k) d += (long)n[k] * (n[k] - 1) / 2;
        double st = 0, ste = 0, stg = 0, stge = 0;
        for (int64_t k = 0; k < steps; ++k) {
            double S = 0.0, E = 0.0;
            int h = 0;
            for (int i = 0; i < M; ++i) {
                const int ni = n[i];
                if (ni == 0) continue;
                const int r = (i + 1 == M) ? 0 : i + 1, l = (i == 0) ? M - 1 : i - 1;
                const double rr = exptab[n[r] - ni + 1 + N], rl = exptab[n[l] - ni + 1 + N];
                S += rr;
                cum[h] = S;
                src[h++] = 2 * i;
                S += rl;
                cum[h] = S;
                src[h++] = 2 * i + 1;
                E += sqrtab[ni] * (sqrtab[n[r] + 1] * rr + sqrtab[n[l] + 1] * rl);
            }
            const double D = m->u * (double)d;
            const double tau = 1.0 / S, eloc = D - m->t * E, gr = -D;
            if (!isfinite(tau)) rc |= 1;
            st += tau;
            ste += tau * eloc;
            stg += tau * gr;
            stge += tau * gr * eloc;
            hops += h;
            const double c = go_u01(seed, walker0 + (uint64_t)w, step0 + (uint64_t)k) * S;
            int q = 0;
            while (q < h - 1 && !(c < cum[q])) ++q;
            const int i = src[q] >> 1;
            const int j = (src[q] & 1) ? ((i == 0) ? M - 1 : i - 1) : ((i + 1 == M) ? 0 : i + 1);
            d += n[j] - n[i] + 1;
            n[i] -= 1;
            n[j] += 1;
        }
        go_from_onr(m, n, &addrs[w]);
        const double val = ste / st;
        if (val_w) val_w[w] = val;
        if (grad_w) grad_w[w] = 2.0 * (stge / st - val * (stg / st));
    }
    free(exptab);
    free(sqrtab);
    if (total_hops) *total_hops = hops;
    return rc;
}
