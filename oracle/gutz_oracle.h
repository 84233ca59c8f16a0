/*
 * gutz_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A literal C restatement of the kinetic-VQMC hot path of mtsch/Gutzwiller.jl and of the
 * Rimu.jl semantics it calls into (Rimu is an un-vendored, un-pinned dependency of the
 * reference: Project.toml:16, no [compat]).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product
 * (libgutzwiller_b200.so) never links, loads or calls it.
 *
 * PARITY PIN: the deterministic part (addresses, diagonal/offdiagonal elements, ansatz,
 * build_basis order, LocalEnergyEvaluator left fold) is pinned bit-for-bit against the
 * literal outputs printed in the reference README (README.md:54,65,94,104,186) -- see
 * tests/test_oracle_golden.py.  The error-bar part (resample! is literal; blocking_analysis is
 * Rimu.StatsTools, restated from the published Flyvbjerg-Petersen / Jonsson M-test algorithm)
 * has no literal vector in the reference: "parity unpinned" for go_blocking_analysis.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef GUTZ_ORACLE_H
#define GUTZ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GO_MAXW 4 /* 64-bit words per address: BoseFS{100,100} needs 199 bits */

typedef struct {
    uint64_t w[GO_MAXW]; /* w[0] = least significant word */
} go_addr;

typedef struct {
    int N, M;    /* particles, modes */
    double u, t; /* HubbardReal1D(addr; u, t) */
    int B;       /* N + M - 1 bits */
    int W;       /* ceil(B / 64) */
} go_model;

/* ---- Rimu boundary ------------------------------------------- */
int go_model_init(go_model *m, int N, int M, double u, double t);
void go_from_onr(const go_model *m, const int *onr, go_addr *out);
void go_to_onr(const go_model *m, const go_addr *a, int *onr);
void go_near_uniform(const go_model *m, go_addr *out);
double go_diagonal_element(const go_model *m, const go_addr *a);
int go_num_offdiagonals(const go_model *m, const go_addr *a);
/* chosen is 1-based like Julia; returns matrix element, writes new address */
double go_get_offdiagonal(const go_model *m, const go_addr *a, int chosen, go_addr *out);
/* BFS closure from `start`; returns dim (or -1 if cap exceeded); basis must hold cap entries */
int64_t go_build_basis(const go_model *m, const go_addr *start, go_addr *basis, int64_t cap);
double go_dimension(const go_model *m);

/* ---- GutzwillerAnsatz (src/ansatz/gutzwiller.jl:25-36) ------------------------------- */
double go_ansatz_val(const go_model *m, const go_addr *a, double g);
void go_ansatz_val_and_grad(const go_model *m, const go_addr *a, double g, double *val, double *der);

/* ---- kinetic_sample! (src/kinetic-vqmc/qmc.jl:14-61) --------------------------------- */
/* u01 replaces rand(); prob_buffer must hold 2*M doubles (filled with the cumulative sums,
 * exactly as the reference's prob_buffer).  Returns 0, or 1 on non-finite residence time. */
int go_kinetic_sample(const go_model *m, double g, const go_addr *a1, double u01,
                      double *prob_buffer, go_addr *new_addr, double *residence_time,
                      double *local_energy, double *grad_ratio, int *chosen, int *n_hops);

/* Per-hop absolute rates |psi(addr2)| / |psi(addr1)| in reference hop order (for parity of
 * "transition rates"); rates must hold 2*M doubles. Returns the number of hops. */
int go_hop_rates(const go_model *m, double g, const go_addr *a, double *rates, double *melems);

/* ---- Philox4x32-10 counter-based stream (Salmon et al., SC'11; Random123) ------------- */
void go_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* uniform in [0,1) for (seed, walker, step): key=(seed lo,hi), ctr=(step>>1 lo,hi, walker lo,hi),
 * words (0,1) for even steps, (2,3) for odd steps; u = (x_hi:x_lo >> 11) * 2^-53 */
double go_u01(uint64_t seed, uint64_t walker, uint64_t step);

/* ---- kinetic_vqmc! (qmc.jl:73-106), estimators (state.jl:42-49,137-149,287-298) ------- */
/* Runs `steps` steps for each of n_walkers walkers (OpenMP over walkers = Threads.@threads,
 * qmc.jl:102).  addrs[n_walkers] in/out (curr_address); step0 = steps already done (stream
 * offset); walker0 = global id of walker 0.  Per-step series are stored like the reference
 * does (qmc.jl:89-92): if the series pointers are non-NULL they receive [walker][step]
 * row-major arrays, else scratch is allocated internally (and still written, to stay
 * faithful for timing).  Per-walker estimates (state.jl:42-49) go to val_w/grad_w if non-NULL.
 * Returns 0 or 1 (infinite residence time, qmc.jl:34-36). */
int go_kinetic_vqmc(const go_model *m, double g, go_addr *addrs, int64_t n_walkers, int64_t steps,
                    uint64_t seed, uint64_t walker0, uint64_t step0, int n_threads,
                    go_addr *series_addr, double *series_tau, double *series_eloc,
                    double *series_grad, double *val_w, double *grad_w, int64_t *total_hops);

/* state.jl:42-49 for one walker's series */
void go_state_val_and_grad(const double *tau, const double *eloc, const double *gr, int64_t n,
                           double *val, double *grad);
/* state.jl:137-149: mean over walkers */
void go_result_val_and_grad(const double *val_w, const double *grad_w, int64_t n_walkers,
                            double *val, double *grad);

/* ---- LocalEnergyEvaluator (src/localenergy.jl:94-155) -------------------------------- */
/* mode 0: sequential left fold in basis order (bit-for-bit pin, README.md:94,104);
 * mode 1: exactly-rounded-ish (long double Neumaier) sums -- the parity target for the GPU;
 * mode 2: OpenMP chunked fold (what Folds.mapreduce does with threads) -- for CPU timing. */
void go_lee_val_and_grad(const go_model *m, const go_addr *basis, int64_t dim, double g, int mode,
                         int n_threads, double *val, double *grad);
double go_lee_val(const go_model *m, const go_addr *basis, int64_t dim, double g, int mode,
                  int n_threads);
/* per-state (num, den, num_grad, den_grad) of localenergy.jl:105-120 */
void go_lee_terms(const go_model *m, const go_addr *k, double g, double out[4]);

/* ---- resample! (state.jl:194-231) and blocking (Rimu.StatsTools; parity unpinned) ------ */
int go_resample(double *result, int64_t len, const double *times, const double *values, int64_t n);
typedef struct {
    double mean, err, err_err, p_cov;
    int k;
    int64_t blocks;
} go_blocking_result;
void go_blocking_analysis(const double *v, int64_t n, double alpha, go_blocking_result *out);
/* state.jl:50-60 */
void go_state_val_err_and_grad(const double *tau, const double *eloc, const double *gr, int64_t n,
                               double *val, double *err, double *grad);

/* ---- adaptive_gradient_descent / amsgrad (src/amsgrad.jl:117-226,240-263) ------------- */
typedef int (*go_objective)(void *user, const double *p, int np, double *val, double *err,
                            double *grad);
typedef struct {
    int maxiter;
    double alpha, beta1, beta2;
    int use_amsgrad; /* 1 = amsgrad, 0 = gradient_descent */
    int early_stop;
    double grad_tol, param_tol, val_tol, err_tol;
    const double *first_moment_init;  /* NULL = gradient at p0 */
    const double *second_moment_init; /* NULL = zeros */
    const uint8_t *fix_params;        /* NULL = none fixed */
} go_gd_opts;
/* trace arrays hold (maxiter+1) rows of np doubles (value/error: maxiter+1 doubles).
 * reason: 0 iterations, 1 params, 2 gradient, 3 value, 4 errorbars. Returns #iterations. */
int go_adaptive_gradient_descent(go_objective val_and_grad, go_objective val_err_and_grad, void *user,
                                 const double *p0, int np, const go_gd_opts *o, double *tr_param,
                                 double *tr_value, double *tr_error, double *tr_grad, double *tr_m1,
                                 double *tr_m2, double *tr_dp, int *converged, int *reason);
/* built-in objective: LocalEnergyEvaluator over a stored basis (mode 0 fold) */
typedef struct {
    const go_model *m;
    const go_addr *basis;
    int64_t dim;
    int mode;
} go_lee_objective;
int go_lee_objective_val_and_grad(void *user, const double *p, int np, double *val, double *err,
                                  double *grad);

/* ---- part 2 (gutz_oracle_ansatz.c): the other single-component ansaetze, generic in P -------- */
enum go_ansatz_kind {
    GO_ANSATZ_GUTZWILLER = 0,       /* src/ansatz/gutzwiller.jl */
    GO_ANSATZ_EXT_GUTZWILLER = 1,   /* src/ansatz/extgutzwiller.jl (call operator) */
    GO_ANSATZ_MULTINOMIAL = 2,      /* src/ansatz/multinomial.jl */
    GO_ANSATZ_DENSITY_PROFILE = 3,  /* src/ansatz/densityprofile.jl */
    GO_ANSATZ_RELATIVE_JASTROW = 4, /* src/ansatz/jastrow.jl:60-103 */
    GO_ANSATZ_JASTROW = 5,          /* src/ansatz/jastrow.jl:4-58 */
    GO_ANSATZ_COMBINATION = 6       /* src/ansatz/combination.jl: alpha*a1 + (1-alpha)*a2, params (p1..., p2..., alpha) */
};
typedef struct {
    go_model m;
    double v;             /* ExtendedHubbardReal1D nearest-neighbour interaction (0: HubbardReal1D) */
    int kind, P;          /* ansatz kind, num_parameters */
    double normalization; /* MultinomialAnsatz(H; normalize) */
    int kind1, kind2;     /* CombinationAnsatz: kinds of the two components */
    int norm1, norm2;     /* ... and their normalize flags */
} go_ansatz;
/* a1 + a2 (combination.jl:13-15) */
int go_ansatz_init_combination(go_ansatz *a, const go_model *m, double v, int kind1, int normalize1, int kind2,
                               int normalize2);
int go_ansatz_init(go_ansatz *a, const go_model *m, double v, int kind, int normalize);
double go_gen_diagonal(const go_ansatz *a, const go_addr *addr);
/* grad (P doubles) may be NULL: the ansatz call operator */
void go_gen_val_and_grad(const go_ansatz *a, const go_addr *addr, const double *params, double *val, double *grad);
int go_gen_kinetic_sample(const go_ansatz *a, const double *params, const go_addr *a1, double u01,
                          double *prob_buffer, go_addr *new_addr, double *residence_time, double *local_energy,
                          double *grad_ratio, int *chosen, int *n_hops);
int go_gen_hop_rates(const go_ansatz *a, const double *params, const go_addr *a1, double *rates);
/* out = (num, den, num_grad[P], den_grad[P]) */
void go_gen_lee_terms(const go_ansatz *a, const go_addr *k, const double *params, double *out);
void go_gen_lee_val_and_grad(const go_ansatz *a, const go_addr *basis, int64_t dim, const double *params,
                             double *val, double *grad);
int go_gen_kinetic_vqmc(const go_ansatz *a, const double *params, go_addr *addrs, int64_t n_walkers, int64_t steps,
                        uint64_t seed, uint64_t walker0, uint64_t step0, int n_threads, double *series_tau,
                        double *series_eloc, go_addr *series_addr, double *val_w, double *grad_w,
                        int64_t *total_hops);

/* tuned CPU code for the Gutzwiller path (closed-form ratios, tables): an upper bound of the CPU side, not a
 * restatement of the reference */
int go_kinetic_vqmc_fast(const go_model *m, double g, go_addr *addrs, int64_t n_walkers, int64_t steps, uint64_t seed,
                         uint64_t walker0, uint64_t step0, int n_threads, double *val_w, double *grad_w,
                         int64_t *total_hops);

/* ---- part 3 (gutz_oracle_sparse.c): the same path over a materialised operator (any Hamiltonian, any ansatz) ----
 * rows = offdiagonals(H, basis[k]) in reference order as (col, melem); diag[k]; psi[k]; grad_ratio[k][P] */
int go_sparse_kinetic_sample(const uint64_t *row_ptr, const uint32_t *col, const double *melem, const double *diag,
                             const double *psi, const double *grad_ratio, int P, uint32_t k, double u01,
                             double *prob_buffer, uint32_t *new_index, double *residence_time, double *local_energy,
                             double *gradient, int *chosen, int *n_hops);
int go_sparse_kinetic_vqmc(const uint64_t *row_ptr, const uint32_t *col, const double *melem, const double *diag,
                           const double *psi, const double *grad_ratio, int P, uint32_t *idx, int64_t n_walkers,
                           int64_t steps, uint64_t seed, uint64_t walker0, uint64_t step0, uint32_t *series_idx,
                           double *series_tau, double *series_eloc, double *val_w, double *grad_w, int64_t *total_hops,
                           uint64_t dim);
void go_sparse_lee_terms(const uint64_t *row_ptr, const uint32_t *col, const double *melem, const double *diag,
                         const double *psi, const double *grad_ratio, int P, uint64_t k, double *out);
void go_sparse_lee_val_and_grad(const uint64_t *row_ptr, const uint32_t *col, const double *melem, const double *diag,
                                const double *psi, const double *grad_ratio, int P, uint64_t dim, double *val,
                                double *grad);

#ifdef __cplusplus
}
#endif
#endif
