/*
 * gutz_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See gutz_oracle.h.
 *
 * Literal restatement of (paths relative to /root/reference):
 *   src/kinetic-vqmc/qmc.jl:14-61,73-106     kinetic_sample!, pick_random_from_cumsum, kinetic_vqmc!
 *   src/kinetic-vqmc/state.jl:42-60,137-164,194-231  estimators, resample!
 *   src/ansatz/gutzwiller.jl:25-36           GutzwillerAnsatz
 *   src/localenergy.jl:94-155                LocalEnergyEvaluator
 *   src/amsgrad.jl:117-226,240-263           adaptive_gradient_descent, gradient_descent, amsgrad
 * and of the Rimu.jl boundary those lines call: BoseFS unary bitstring,
 * HubbardReal1D diagonal_element / offdiagonals (hopnextneighbour order), near_uniform,
 * build_basis (breadth-first closure), StatsTools.blocking_analysis.
 *
 * Compile with -ffp-contract=off: Julia does not fuse a*b+c, and the README pins are bit-exact.
 */
#include "gutz_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ======================================================================================
 * Multi-word bitstring helpers (Rimu BitString{B,C,T}; here always 64-bit chunks, w[0] LSB)
 * ==================================================================================== */
static inline int a_is_zero(const go_addr *a, int W) {
    for (int i = 0; i < W; i++)
        if (a->w[i]) return 0;
    return 1;
}
static inline void a_shr(go_addr *a, int W, int k) {
    int ws = k >> 6, bs = k & 63;
    for (int i = 0; i < W; i++) {
        uint64_t lo = (i + ws < W) ? a->w[i + ws] : 0;
        uint64_t hi = (i + ws + 1 < W) ? a->w[i + ws + 1] : 0;
        a->w[i] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
    }
}
static inline void a_shl1(go_addr *a, int W) {
    for (int i = W - 1; i >= 0; i--) a->w[i] = (a->w[i] << 1) | (i ? (a->w[i - 1] >> 63) : 0);
}
static inline int a_trailing_zeros(const go_addr *a, int W) {
    int c = 0;
    for (int i = 0; i < W; i++) {
        if (a->w[i]) return c + __builtin_ctzll(a->w[i]);
        c += 64;
    }
    return c;
}
static inline int a_trailing_ones(const go_addr *a, int W) {
    int c = 0;
    for (int i = 0; i < W; i++) {
        uint64_t inv = ~a->w[i];
        if (inv) return c + __builtin_ctzll(inv);
        c += 64;
    }
    return c;
}
static inline int a_get(const go_addr *a, int pos) { return (int)((a->w[pos >> 6] >> (pos & 63)) & 1u); }
static inline void a_flip(go_addr *a, int pos) { a->w[pos >> 6] ^= (uint64_t)1 << (pos & 63); }
static inline void a_mask_to(go_addr *a, int W, int B) {
    for (int i = 0; i < GO_MAXW; i++) {
        int lo = 64 * i;
        if (i >= W || lo >= B) a->w[i] = 0;
        else if (B - lo < 64) a->w[i] &= (((uint64_t)1 << (B - lo)) - 1);
    }
}
static inline int a_equal(const go_addr *a, const go_addr *b) {
    for (int i = 0; i < GO_MAXW; i++)
        if (a->w[i] != b->w[i]) return 0;
    return 1;
}

/* occupied-mode iterator (Rimu `occupied_modes(::BoseFS)`): yields (mode 1-based, occnum, bit
 * position of the bottom of the run) in ascending mode order */
typedef struct {
    go_addr bits;
    int W;
    int mode; /* zeros passed so far = 0-based mode index */
    int pos;
} occ_iter;
static inline void occ_begin(occ_iter *it, const go_model *m, const go_addr *a) {
    it->bits = *a;
    it->W = m->W;
    it->mode = 0;
    it->pos = 0;
}
static inline int occ_next(occ_iter *it, int *mode, int *occ, int *pos) {
    if (a_is_zero(&it->bits, it->W)) return 0;
    int z = a_trailing_zeros(&it->bits, it->W);
    a_shr(&it->bits, it->W, z);
    it->mode += z;
    it->pos += z;
    int n = a_trailing_ones(&it->bits, it->W);
    *mode = it->mode + 1;
    *occ = n;
    *pos = it->pos;
    a_shr(&it->bits, it->W, n);
    it->pos += n;
    return 1;
}

/* ======================================================================================
 * Rimu boundary
 * ==================================================================================== */
int go_model_init(go_model *m, int N, int M, double u, double t) {
    if (N < 1 || M < 2) return 1;
    int B = N + M - 1;
    if (B > 64 * GO_MAXW) return 2;
    m->N = N;
    m->M = M;
    m->u = u;
    m->t = t;
    m->B = B;
    m->W = (B + 63) / 64;
    return 0;
}

/* Appendix A.1: mode 1 at the LSB end, n_i ones followed by a single 0 separator */
void go_from_onr(const go_model *m, const int *onr, go_addr *out) {
    memset(out, 0, sizeof(*out));
    int pos = 0;
    for (int i = 0; i < m->M; i++) {
        for (int k = 0; k < onr[i]; k++) a_flip(out, pos++);
        pos++; /* separator */
    }
}
void go_to_onr(const go_model *m, const go_addr *a, int *onr) {
    for (int i = 0; i < m->M; i++) onr[i] = 0;
    occ_iter it;
    int mode, occ, pos;
    occ_begin(&it, m, a);
    while (occ_next(&it, &mode, &occ, &pos)) onr[mode - 1] = occ;
}
/* near_uniform_onr: fillingfactor everywhere, first `extras` modes get one more */
void go_near_uniform(const go_model *m, go_addr *out) {
    int *onr = (int *)malloc(sizeof(int) * (size_t)m->M);
    int ff = m->N / m->M, extras = m->N % m->M;
    for (int i = 0; i < m->M; i++) onr[i] = ff + (i < extras ? 1 : 0);
    go_from_onr(m, onr, out);
    free(onr);
}

/* Appendix A.3: diagonal_element(H, a) = u * sum_i n_i (n_i - 1) / 2 */
double go_diagonal_element(const go_model *m, const go_addr *a) {
    occ_iter it;
    int mode, occ, pos;
    int64_t s = 0;
    occ_begin(&it, m, a);
    while (occ_next(&it, &mode, &occ, &pos)) s += (int64_t)occ * (occ - 1);
    return m->u * (double)s / 2;
}

int go_num_offdiagonals(const go_model *m, const go_addr *a) {
    occ_iter it;
    int mode, occ, pos, c = 0;
    occ_begin(&it, m, a);
    while (occ_next(&it, &mode, &occ, &pos)) c++;
    return 2 * c;
}

/* number of ones directly below bit B (occupation of mode M) */
static int top_run(const go_model *m, const go_addr *a) {
    int c = 0;
    for (int p = m->B - 1; p >= 0 && a_get(a, p); p--) c++;
    return c;
}

/* Appendix A.4: entry `chosen` (1-based): source = ((chosen+1)>>1)-th occupied mode; odd hops
 * right (wrapping M->1), even hops left (wrapping 1->M); melem = -t*sqrt(n_src*(n_dst+1)).
 * The scan restarts from the first mode on every call, as Rimu's find_occupied_mode does. */
double go_get_offdiagonal(const go_model *m, const go_addr *a, int chosen, go_addr *out) {
    occ_iter it;
    int mode = 0, occ = 0, pos = 0;
    int pmode = 0, pocc = 0; /* previous occupied mode */
    int k = (chosen + 1) >> 1;
    occ_begin(&it, m, a);
    for (int c = 1;; c++) {
        int mo, oc, po;
        if (!occ_next(&it, &mo, &oc, &po)) return NAN; /* chosen out of range */
        if (c == k) {
            mode = mo;
            occ = oc;
            pos = po;
            break;
        }
        pmode = mo;
        pocc = oc;
    }
    *out = *a;
    int ndst;
    if (chosen & 1) { /* right: dst = mode + 1 */
        if (mode < m->M) {
            /* occupation of the next mode = ones after this run's separator */
            go_addr rest = it.bits; /* iterator sits on the separator after the run */
            a_shr(&rest, m->W, 1);
            ndst = a_trailing_ones(&rest, m->W);
            a_flip(out, pos + occ - 1);
            a_flip(out, pos + occ);
        } else { /* M -> 1 */
            ndst = a_trailing_ones(a, m->W);
            if (ndst > m->B) ndst = m->B;
            a_shl1(out, m->W);
            out->w[0] |= 1;
            a_mask_to(out, m->W, m->B);
        }
    } else { /* left: dst = mode - 1 */
        if (mode > 1) {
            ndst = (pmode == mode - 1) ? pocc : 0;
            a_flip(out, pos - 1);
            a_flip(out, pos);
        } else { /* 1 -> M */
            ndst = top_run(m, a);
            a_shr(out, m->W, 1);
            a_flip(out, m->B - 1); /* bit B-1 is 0 after the shift */
        }
    }
    return -m->t * sqrt((double)((int64_t)occ * (ndst + 1)));
}

double go_dimension(const go_model *m) {
    /* C(N+M-1, N) */
    double r = 1.0;
    int n = m->B, k = m->N < m->M - 1 ? m->N : m->M - 1;
    for (int i = 1; i <= k; i++) r = r * (double)(n - k + i) / (double)i;
    return floor(r + 0.5);
}

/* --- open-addressing set of addresses for build_basis --- */
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
static inline uint64_t a_hash(const go_addr *a) {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (int i = 0; i < GO_MAXW; i++) h = mix64(h ^ a->w[i]) + 0x9E3779B97F4A7C15ULL;
    return h;
}

/* Appendix A.5: breadth-first closure from the start address, offdiagonals in order,
 * append unseen targets; NOT sorted. (localenergy.jl:86 `basis=build_basis(ham)`) */
int64_t go_build_basis(const go_model *m, const go_addr *start, go_addr *basis, int64_t cap) {
    uint64_t tcap = 16;
    while ((int64_t)tcap < 2 * cap + 16) tcap <<= 1;
    int64_t *table = (int64_t *)malloc(sizeof(int64_t) * tcap);
    if (!table) return -2;
    for (uint64_t i = 0; i < tcap; i++) table[i] = -1;
    int64_t dim = 0, head = 0;
    basis[dim] = *start;
    table[a_hash(start) & (tcap - 1)] = dim;
    dim++;
    while (head < dim) {
        go_addr cur = basis[head++];
        int h = go_num_offdiagonals(m, &cur);
        for (int c = 1; c <= h; c++) {
            go_addr nb;
            go_get_offdiagonal(m, &cur, c, &nb);
            uint64_t slot = a_hash(&nb) & (tcap - 1);
            int found = 0;
            while (table[slot] >= 0) {
                if (a_equal(&basis[table[slot]], &nb)) {
                    found = 1;
                    break;
                }
                slot = (slot + 1) & (tcap - 1);
            }
            if (!found) {
                if (dim >= cap) {
                    free(table);
                    return -1;
                }
                basis[dim] = nb;
                table[slot] = dim;
                dim++;
            }
        }
    }
    free(table);
    return dim;
}

/* ======================================================================================
 * GutzwillerAnsatz  (src/ansatz/gutzwiller.jl:25-36)
 * ==================================================================================== */
double go_ansatz_val(const go_model *m, const go_addr *a, double g) {
    return exp(-g * go_diagonal_element(m, a)); /* gutzwiller.jl:34-36 */
}
void go_ansatz_val_and_grad(const go_model *m, const go_addr *a, double g, double *val, double *der) {
    double diag = go_diagonal_element(m, a); /* gutzwiller.jl:27 */
    *val = exp(-g * diag);                   /* :29 */
    *der = -diag * (*val);                   /* :30 (signed zero at diag = 0, README.md:65) */
}

/* ======================================================================================
 * kinetic_sample!  (src/kinetic-vqmc/qmc.jl:14-44) + pick_random_from_cumsum (:52-61)
 * ==================================================================================== */
int go_kinetic_sample(const go_model *m, double g, const go_addr *a1, double u01,
                      double *prob_buffer, go_addr *new_addr, double *residence_time,
                      double *local_energy, double *grad_ratio, int *chosen_out, int *n_hops) {
    int h = go_num_offdiagonals(m, a1); /* :15 length(offdiags) */
    double val1, grad1;
    go_ansatz_val_and_grad(m, a1, g, &val1, &grad1); /* :19 */
    double diag = go_diagonal_element(m, a1);        /* :20 */
    double local_energy_num = diag * val1;           /* :21 */
    double residence_time_denom = 0.0;               /* :23 */
    for (int i = 1; i <= h; i++) {                   /* :25-31 */
        go_addr addr2;
        double melem = go_get_offdiagonal(m, a1, i, &addr2);
        double val2 = go_ansatz_val(m, &addr2, g);
        residence_time_denom += fabs(val2);
        local_energy_num += melem * val2;
        prob_buffer[i - 1] = residence_time_denom;
    }
    *residence_time = fabs(val1) / residence_time_denom; /* :33 */
    if (n_hops) *n_hops = h;
    if (!isfinite(*residence_time)) return 1;            /* :34-36 */
    *local_energy = local_energy_num / val1;             /* :37 */
    *grad_ratio = grad1 / val1;                          /* :38 */
    /* pick_random_from_cumsum (:52-61); the reference has no upper bound on i under
     * @inbounds -- clamp to the last hop instead of reading out of bounds (Appendix B.7) */
    double chosen = u01 * prob_buffer[h - 1];
    int i = 1;
    while (i < h && !(chosen < prob_buffer[i - 1])) i++;
    if (chosen_out) *chosen_out = i;
    go_get_offdiagonal(m, a1, i, new_addr); /* :41 */
    return 0;
}

int go_hop_rates(const go_model *m, double g, const go_addr *a, double *rates, double *melems) {
    int h = go_num_offdiagonals(m, a);
    double val1 = go_ansatz_val(m, a, g);
    for (int i = 1; i <= h; i++) {
        go_addr addr2;
        double melem = go_get_offdiagonal(m, a, i, &addr2);
        rates[i - 1] = fabs(go_ansatz_val(m, &addr2, g)) / fabs(val1);
        if (melems) melems[i - 1] = melem;
    }
    return h;
}

/* ======================================================================================
 * Philox4x32-10
 * ==================================================================================== */
void go_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}
double go_u01(uint64_t seed, uint64_t walker, uint64_t step) {
    uint64_t blk = step >> 1;
    uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)walker, (uint32_t)(walker >> 32)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t x[4];
    go_philox4x32_10(ctr, key, x);
    int o = (int)(step & 1) * 2;
    uint64_t r = ((uint64_t)x[o + 1] << 32) | x[o];
    return (double)(r >> 11) * 0x1.0p-53;
}

/* ======================================================================================
 * Estimators (src/kinetic-vqmc/state.jl)
 * ==================================================================================== */
/* state.jl:42-49: val = mean(E, weights tau); grad = 2*mean(G.*(E.-val), weights tau) */
void go_state_val_and_grad(const double *tau, const double *eloc, const double *gr, int64_t n,
                           double *val, double *grad) {
    long double st = 0, ste = 0;
    for (int64_t k = 0; k < n; k++) {
        st += tau[k];
        ste += (long double)tau[k] * eloc[k];
    }
    double v = (double)(ste / st);
    long double sg = 0;
    for (int64_t k = 0; k < n; k++) sg += (long double)tau[k] * (gr[k] * (eloc[k] - v));
    *val = v;
    *grad = (double)(2 * (sg / st));
}
/* state.jl:137-149 */
void go_result_val_and_grad(const double *val_w, const double *grad_w, int64_t n_walkers,
                            double *val, double *grad) {
    double vals = 0, grads = 0;
    for (int64_t w = 0; w < n_walkers; w++) {
        vals += val_w[w];
        grads += grad_w[w];
    }
    *val = vals / (double)n_walkers;
    *grad = grads / (double)n_walkers;
}

/* qmc.jl:73-97 per walker, qmc.jl:99-106 threaded over walkers */
int go_kinetic_vqmc(const go_model *m, double g, go_addr *addrs, int64_t n_walkers, int64_t steps,
                    uint64_t seed, uint64_t walker0, uint64_t step0, int n_threads,
                    go_addr *series_addr, double *series_tau, double *series_eloc,
                    double *series_grad, double *val_w, double *grad_w, int64_t *total_hops) {
    int err = 0;
    int64_t hops_total = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(| : err) reduction(+ : hops_total)
    for (int64_t w = 0; w < n_walkers; w++) {
        /* resize!(addresses...) etc. (qmc.jl:80-83): per-walker per-step arrays */
        go_addr *sa = series_addr ? series_addr + w * steps : (go_addr *)malloc(sizeof(go_addr) * (size_t)steps);
        double *st = series_tau ? series_tau + w * steps : (double *)malloc(sizeof(double) * (size_t)steps);
        double *se = series_eloc ? series_eloc + w * steps : (double *)malloc(sizeof(double) * (size_t)steps);
        double *sg = series_grad ? series_grad + w * steps : (double *)malloc(sizeof(double) * (size_t)steps);
        double *prob_buffer = (double *)malloc(sizeof(double) * 2 * (size_t)m->M);
        go_addr curr = addrs[w];
        for (int64_t k = 0; k < steps; k++) { /* :85-94 */
            go_addr next;
            double tau, eloc, gr;
            int chosen, nh;
            double u01 = go_u01(seed, walker0 + (uint64_t)w, step0 + (uint64_t)k);
            err |= go_kinetic_sample(m, g, &curr, u01, prob_buffer, &next, &tau, &eloc, &gr, &chosen, &nh);
            sa[k] = curr;
            st[k] = tau;
            se[k] = eloc;
            sg[k] = gr;
            curr = next;
            hops_total += nh;
        }
        addrs[w] = curr; /* :95 */
        if (val_w && grad_w) go_state_val_and_grad(st, se, sg, steps, &val_w[w], &grad_w[w]);
        if (!series_addr) free(sa);
        if (!series_tau) free(st);
        if (!series_eloc) free(se);
        if (!series_grad) free(sg);
        free(prob_buffer);
    }
    if (total_hops) *total_hops = hops_total;
    return err;
}

/* ======================================================================================
 * LocalEnergyEvaluator (src/localenergy.jl:94-155)
 * ==================================================================================== */
void go_lee_terms(const go_model *m, const go_addr *k, double g, double out[4]) {
    double k_val, k_grad;
    go_ansatz_val_and_grad(m, k, g, &k_val, &k_grad); /* :105 */
    double diag = go_diagonal_element(m, k);          /* :106 */
    double num = (k_val * k_val) * diag;              /* :108 abs2(k_val)*diag */
    double num_grad = ((2 * k_grad) * k_val) * diag;  /* :109 */
    double den = k_val * k_val;                       /* :111 */
    double den_grad = (2 * k_val) * k_grad;           /* :112 */
    int h = go_num_offdiagonals(m, k);
    for (int c = 1; c <= h; c++) {                    /* :114-118 */
        go_addr l;
        double melem = go_get_offdiagonal(m, k, c, &l);
        double l_val, l_grad;
        go_ansatz_val_and_grad(m, &l, g, &l_val, &l_grad);
        num += (l_val * melem) * k_val;
        num_grad += melem * (l_val * k_grad + l_grad * k_val);
    }
    out[0] = num;
    out[1] = den;
    out[2] = num_grad;
    out[3] = den_grad;
}

typedef struct {
    long double s, c;
} neumaier;
static inline void nm_add(neumaier *a, long double x) {
    long double t = a->s + x;
    if (fabsl(a->s) >= fabsl(x)) a->c += (a->s - t) + x;
    else a->c += (x - t) + a->s;
    a->s = t;
}

static void lee_reduce(const go_model *m, const go_addr *basis, int64_t dim, double g, int mode,
                       int n_threads, double acc[4]) {
    if (mode == 0) { /* sequential left fold: result = add(result, f(k)) */
        acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
        for (int64_t i = 0; i < dim; i++) {
            double tmp[4];
            go_lee_terms(m, &basis[i], g, tmp);
            for (int j = 0; j < 4; j++) acc[j] = acc[j] + tmp[j];
        }
    } else if (mode == 1) { /* compensated long-double sums: order-independent to ~1e-19 */
        neumaier s[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int64_t i = 0; i < dim; i++) {
            double tmp[4];
            go_lee_terms(m, &basis[i], g, tmp);
            for (int j = 0; j < 4; j++) nm_add(&s[j], tmp[j]);
        }
        for (int j = 0; j < 4; j++) acc[j] = (double)(s[j].s + s[j].c);
    } else { /* threaded chunked fold (Folds.mapreduce with threads) */
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#ifdef _OPENMP
        if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(static) reduction(+ : a0, a1, a2, a3)
        for (int64_t i = 0; i < dim; i++) {
            double tmp[4];
            go_lee_terms(m, &basis[i], g, tmp);
            a0 += tmp[0];
            a1 += tmp[1];
            a2 += tmp[2];
            a3 += tmp[3];
        }
        acc[0] = a0;
        acc[1] = a1;
        acc[2] = a2;
        acc[3] = a3;
    }
}

void go_lee_val_and_grad(const go_model *m, const go_addr *basis, int64_t dim, double g, int mode,
                         int n_threads, double *val, double *grad) {
    double a[4];
    lee_reduce(m, basis, dim, g, mode, n_threads, a);
    double numerator = a[0], denominator = a[1], numerator_grad = a[2], denominator_grad = a[3];
    *val = numerator / denominator; /* :126 */
    *grad = (numerator_grad * denominator - numerator * denominator_grad) / (denominator * denominator); /* :127 */
}

/* value-only variant localenergy.jl:131-155 (same per-state num/den; ansatz(k) == val_and_grad value) */
double go_lee_val(const go_model *m, const go_addr *basis, int64_t dim, double g, int mode, int n_threads) {
    if (mode == 0) {
        double num_acc = 0.0, den_acc = 0.0;
        for (int64_t i = 0; i < dim; i++) {
            const go_addr *k = &basis[i];
            double k_val = go_ansatz_val(m, k, g);
            double diag = go_diagonal_element(m, k);
            double num = (k_val * k_val) * diag;
            double den = k_val * k_val;
            int h = go_num_offdiagonals(m, k);
            for (int c = 1; c <= h; c++) {
                go_addr l;
                double melem = go_get_offdiagonal(m, k, c, &l);
                double l_val = go_ansatz_val(m, &l, g);
                num += (l_val * melem) * k_val;
            }
            num_acc = num_acc + num;
            den_acc = den_acc + den;
        }
        return num_acc / den_acc;
    }
    double a[4];
    lee_reduce(m, basis, dim, g, mode, n_threads, a);
    return a[0] / a[1];
}

/* ======================================================================================
 * resample!  (src/kinetic-vqmc/state.jl:194-231), literal
 * ==================================================================================== */
int go_resample(double *result, int64_t len, const double *times, const double *values, int64_t n) {
    double total_time = 0.0;
    for (int64_t j = 0; j < n; j++) total_time += times[j];
    double t = 0.0, mu = 0.0, val = 0.0;
    int64_t k = 1, i = 1;
    for (;;) {
        while (t >= (double)i && i <= len) { /* :211-215 */
            result[i - 1] = val;
            i += 1;
            mu -= val;
        }
        while (t < (double)i && k <= n) { /* :217-223 */
            val = values[k - 1];
            double step = (double)len * times[k - 1] / total_time;
            double room = (double)i - t;
            mu += val * (step < room ? step : room);
            t += step;
            k += 1;
        }
        if (i > len) break; /* :225 */
        result[i - 1] = mu;
        mu = (t - (double)i) * val;
        i += 1;
    }
    return 0;
}

/* ======================================================================================
 * blocking_analysis (Rimu.StatsTools: blocks_with_m + mtest; Flyvbjerg-Petersen pair averaging +
 * the M test of Jonsson, PRE 98, 043304).  PARITY UNPINNED: Rimu's source is not available; restated
 * from the paper's listing (quantile indexed by the level, M[k] < q[k], the top level never accepted)
 * with the finite-n M_j of the paper's theorem and corrected (n-1) variance / lag-1 autocovariance,
 * which is how Rimu's blocks_with_m evaluates them by default.  Two-pass moments in long double; the
 * product (csrc/gwz_blocking.cuh) uses centred one-pass moments, tests/blocking_checker.py is a third,
 * independent numpy version written from the published listing.
 * ==================================================================================== */
static const double CHI2_Q99[] = {
    /* the table printed in the listing: chi-square 0.99 quantiles, 1..30 degrees of freedom */
    6.634897,  9.210340,  11.344867, 13.276704, 15.086272, 16.811894, 18.475307, 20.090235, 21.665994, 23.209251,
    24.724970, 26.216967, 27.688250, 29.141238, 30.577914, 31.999927, 33.408664, 34.805306, 36.190869, 37.566235,
    38.932173, 40.289360, 41.638398, 42.979820, 44.314105, 45.641683, 46.962942, 48.278236, 49.587884, 50.892181};

void go_blocking_analysis(const double *v_in, int64_t n, double alpha, go_blocking_result *out) {
    (void)alpha; /* only alpha = 0.01 tabulated */
    out->mean = 0;
    out->err = out->err_err = out->p_cov = NAN;
    out->k = -1;
    out->blocks = 0;
    if (n < 1) return;
    if (n == 1) {
        out->mean = v_in[0];
        out->k = 0;
        out->blocks = 1;
        return;
    }
    int n_steps = 0;
    while (((int64_t)1 << (n_steps + 1)) <= n) n_steps++; /* floor(log2(n)) */
    double *v = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(v, v_in, sizeof(double) * (size_t)n);
    double *se = (double *)malloc(sizeof(double) * (size_t)n_steps);
    double *see = (double *)malloc(sizeof(double) * (size_t)n_steps);
    double *mj = (double *)malloc(sizeof(double) * (size_t)n_steps);
    int64_t *blocks = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_steps);
    double mean0 = 0;
    int64_t len = n;
    for (int s = 0; s < n_steps; s++) {
        long double sum = 0;
        for (int64_t i = 0; i < len; i++) sum += v[i];
        double mean = (double)(sum / len);
        if (s == 0) mean0 = mean;
        long double ss = 0, gam = 0;
        for (int64_t i = 0; i < len; i++) ss += ((long double)v[i] - mean) * ((long double)v[i] - mean);
        for (int64_t i = 0; i + 1 < len; i++) gam += ((long double)v[i] - mean) * ((long double)v[i + 1] - mean);
        double var = (double)(ss / (len - 1));     /* var(v; corrected = true) */
        double gamma = (double)(gam / (len - 1));  /* autocovariance(v, 1; corrected = true) */
        blocks[s] = len;
        se[s] = sqrt(var / (double)len);
        see[s] = se[s] / sqrt(2.0 * (double)(len - 1));
        /* M_j = n ((n-1) var / n^2 + gamma)^2 / var^2 ; 0/0 = NaN for a constant series */
        double q = ((double)(len - 1) * var / ((double)len * (double)len) + gamma);
        mj[s] = (double)len * q * q / (var * var);
        int64_t half = len / 2;
        for (int64_t i = 0; i < half; i++) v[i] = 0.5 * (v[2 * i] + v[2 * i + 1]);
        len = half;
    }
    /* M_k = sum_{j >= k} mj[j]; first k = 1 .. n_steps-1 with M_k < q[k] */
    int kk = -1;
    double tail = 0;
    double *M = (double *)malloc(sizeof(double) * (size_t)n_steps);
    for (int s = n_steps - 1; s >= 0; s--) {
        tail += mj[s];
        M[s] = tail;
    }
    for (int s = 0; s + 1 < n_steps; s++) {
        if (M[s] < CHI2_Q99[s < 30 ? s : 29]) {
            kk = s;
            break;
        }
    }
    out->mean = mean0;
    if (kk >= 0) {
        out->err = se[kk];
        out->err_err = see[kk];
        out->p_cov = se[kk] * se[kk];
        out->k = kk + 1; /* 1-based level like Julia */
        out->blocks = blocks[kk];
    }
    free(v);
    free(se);
    free(see);
    free(mj);
    free(blocks);
    free(M);
}

/* state.jl:50-60 */
void go_state_val_err_and_grad(const double *tau, const double *eloc, const double *gr, int64_t n,
                               double *val, double *err, double *grad) {
    double *rs = (double *)malloc(sizeof(double) * (size_t)n);
    go_resample(rs, n, tau, eloc, n);
    go_blocking_result b;
    go_blocking_analysis(rs, n, 0.01, &b);
    free(rs);
    *val = b.mean;
    *err = b.err;
    long double st = 0, sg = 0;
    for (int64_t k = 0; k < n; k++) {
        st += tau[k];
        sg += (long double)tau[k] * (gr[k] * (eloc[k] - b.mean));
    }
    *grad = (double)(2 * (sg / st));
}

/* ======================================================================================
 * adaptive_gradient_descent (src/amsgrad.jl:117-226); phi/psi pairs :240-244 and :259-263
 * ==================================================================================== */
int go_adaptive_gradient_descent(go_objective val_and_grad, go_objective val_err_and_grad, void *user,
                                 const double *p0, int np, const go_gd_opts *o, double *tr_param,
                                 double *tr_value, double *tr_error, double *tr_grad, double *tr_m1,
                                 double *tr_m2, double *tr_dp, int *converged, int *reason) {
    double *p = (double *)malloc(sizeof(double) * (size_t)np * 5);
    double *grad = p + np, *m1 = p + 2 * np, *m2 = p + 3 * np, *dp = p + 4 * np;
    double val, err = 0.0, old_val = INFINITY;
    memcpy(p, p0, sizeof(double) * (size_t)np);
    val_and_grad(user, p, np, &val, &err, grad); /* :136 */
    for (int i = 0; i < np; i++) {
        m1[i] = o->first_moment_init ? o->first_moment_init[i] : grad[i];  /* :138-142 */
        m2[i] = o->second_moment_init ? o->second_moment_init[i] : 0.0;    /* :143-147 */
    }
    int iter = 0, rows = 0;
    *converged = 0;
    *reason = 0;
    while (iter <= o->maxiter) { /* :166 -> maxiter+1 iterations */
        iter += 1;
        val_err_and_grad(user, p, np, &val, &err, grad); /* :168 */
        for (int i = 0; i < np; i++) {
            if (o->use_amsgrad) {
                m1[i] = (1 - o->beta1) * m1[i] + o->beta1 * grad[i];               /* :260 */
                double cand = (1 - o->beta2) * m2[i] + o->beta2 * (grad[i] * grad[i]); /* :261 */
                m2[i] = cand > m2[i] ? cand : m2[i];
            } else {
                m1[i] = grad[i]; /* :241 */
                m2[i] = 1.0;     /* :242 */
            }
        }
        double ndp = 0, ngrad = 0;
        for (int i = 0; i < np; i++) {
            int nf = !(o->fix_params && o->fix_params[i]);
            grad[i] = grad[i] * (double)nf;                            /* :171 */
            dp[i] = (double)nf * -(o->alpha * m1[i] / sqrt(m2[i]));    /* :173 */
            ndp += dp[i] * dp[i];
            ngrad += grad[i] * grad[i];
        }
        double dval = old_val - val; /* :174 */
        old_val = val;
        memcpy(tr_param + (size_t)rows * np, p, sizeof(double) * (size_t)np); /* :177-183 */
        tr_value[rows] = val;
        tr_error[rows] = err;
        memcpy(tr_grad + (size_t)rows * np, grad, sizeof(double) * (size_t)np);
        memcpy(tr_m1 + (size_t)rows * np, m1, sizeof(double) * (size_t)np);
        memcpy(tr_m2 + (size_t)rows * np, m2, sizeof(double) * (size_t)np);
        memcpy(tr_dp + (size_t)rows * np, dp, sizeof(double) * (size_t)np);
        rows++;
        if (o->early_stop) { /* :185-210 */
            if (sqrt(ndp) < o->param_tol) { *reason = 1; *converged = 1; break; }
            if (sqrt(ngrad) < o->grad_tol) { *reason = 2; *converged = 1; break; }
            if (fabs(dval) < o->val_tol) { *reason = 3; *converged = 1; break; }
            if (fabs(dval) < err * o->err_tol) { *reason = 4; *converged = 1; break; }
        }
        for (int i = 0; i < np; i++) p[i] = p[i] + dp[i]; /* :212 */
    }
    free(p);
    return rows;
}

int go_lee_objective_val_and_grad(void *user, const double *p, int np, double *val, double *err,
                                  double *grad) {
    (void)np;
    const go_lee_objective *o = (const go_lee_objective *)user;
    go_lee_val_and_grad(o->m, o->basis, o->dim, p[0], o->mode, 0, val, grad);
    *err = 0.0; /* abstract.jl:59-62 fallback */
    return 0;
}
