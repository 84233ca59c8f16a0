/*
 * gutzwiller_b200.h -- C ABI of libgutzwiller_b200.so
 *
 * B200 (sm_100a) implementation of the kinetic variational-QMC hot path of mtsch/Gutzwiller.jl
 * for HubbardReal1D + GutzwillerAnsatz.  The reference has no FFI layer (it is pure Julia); the
 * drop-in boundary is its exported Julia API, and every entry point below names the reference
 * definition it stands behind (paths relative to the reference repository root).  A Julia host
 * binds these with `ccall` (see INTEGRATION.md and gutzwiller.jl_b200/julia/GutzwillerB200.jl); the
 * Python mirror in gutzwiller.jl_b200/ binds the same symbols with ctypes.
 *
 * Conventions
 *  - every function returns an int status (GWZ_OK = 0); gwz_last_error() gives the message of
 *    the last failure on the calling thread.  No C++ exception crosses this boundary.
 *  - the caller owns all in/out buffers (plain host pointers); the library owns the opaque
 *    handles.  Calls block until the device work they issued is complete.
 *  - a handle may be used by one host thread at a time; distinct handles concurrently.
 *  - addresses are BoseFS{N,M} unary bitstrings (Rimu BitString semantics): W = ceil((N+M-1)/64)
 *    64-bit words per address, word 0 = least significant, mode 1 at bit 0, n_i ones followed
 *    by one 0 separator per mode.
 *  - all floating point is FP64 (eltype(H) = Float64, src/ansatz/gutzwiller.jl:18).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    GWZ_ERR_CUDA.
 */
#ifndef GUTZWILLER_B200_H
#define GUTZWILLER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWZ_VERSION 200
#define GWZ_MAX_WORDS 4       /* 64-bit words per address (N+M-1 <= 251) */
#define GWZ_NUM_PARAMS 1      /* GutzwillerAnsatz has one parameter g (gutzwiller.jl:14) */
#define GWZ_NCCL_ID_BYTES 128 /* sizeof(ncclUniqueId) */
#define GWZ_NUM_ACC 13        /* doubles in the cross-device accumulator vector */
#define GWZ_NUM_BACC 38       /* doubles in the cross-device vector of the blocking analyses */

enum gwz_status {
    GWZ_OK = 0,
    GWZ_ERR_INVALID = 1,   /* bad argument (Julia: ArgumentError) */
    GWZ_ERR_CUDA = 2,      /* CUDA runtime failure / no device */
    GWZ_ERR_NCCL = 3,      /* NCCL failure / libnccl not loadable */
    GWZ_ERR_NONFINITE = 4, /* "Infinite residence time" (src/kinetic-vqmc/qmc.jl:34-36) */
    GWZ_ERR_NOMEM = 5
};

/* which device code evaluates a step */
enum gwz_kernel_kind {
    GWZ_KERNEL_ENUMERATE = 0, /* walks the hops in the reference's offdiagonals order (qmc.jl:25-31);
                                 move = first i with u*S < cumsum_i (qmc.jl:52-61) */
    GWZ_KERNEL_BITPARALLEL = 1 /* production path: bond classes by popcount, class-ordered move
                                  selection (same distribution, different u -> hop map) */
};

/* flags of gwz_vqmc_run */
#define GWZ_RUN_KEEP_ACC 1u /* continue the estimators (kinetic_vqmc!(res), qmc.jl:108-111) instead of
                               clearing them first (KineticVQMC `_reset!`, wrapper.jl:74-79) */
#define GWZ_RUN_POOLED 2u   /* report the pooled ratio estimator (sum over ALL walkers and steps) in val/grad
                               instead of the reference's mean of per-walker estimates (state.jl:137-149);
                               the latter carries an O(tau_corr/steps) bias per walker that does not average
                               out over walkers, which matters for many short walkers */

#define GWZ_RUN_SERIES 4u   /* keep the (residence_time, local_energy) series of the accumulated steps in HBM, like
                               the reference's KineticVQMCWalkerState does on the host (state.jl:15-17, qmc.jl:89-92);
                               gwz_walkers_blocking analyses them.  16 B of HBM per walker-step. */
#define GWZ_RUN_BLOCKING 8u /* GWZ_RUN_SERIES + run val_err_and_grad's per-walker resample + blocking_analysis
                               (state.jl:50-60,150-164) on the device in the same call: fills out->blocking */

typedef struct gwz_context gwz_context;
typedef struct gwz_model gwz_model;
typedef struct gwz_walkers gwz_walkers;
typedef struct gwz_sweep gwz_sweep;

const char *gwz_last_error(void);
int gwz_version(void);
/* number of CUDA kernels launched by this library so far (process-wide; for benchmarks) */
int gwz_launch_count(uint64_t *n);

/* ---- context: one CUDA device + stream (+ optional NCCL communicator) --------------------- */
int gwz_context_create(int device, gwz_context **out);
int gwz_context_destroy(gwz_context *ctx);
int gwz_context_info(const gwz_context *ctx, int *device, int *sm_count, int *sm_clock_khz,
                     int *cc_major, int *cc_minor, size_t *global_mem_bytes);
/* raw cudaStream_t of the context, for timing with CUDA events on the launching stream */
int gwz_context_stream(const gwz_context *ctx, void **cuda_stream);
int gwz_context_synchronize(gwz_context *ctx);
/* Multi-GPU (replaces Threads.@threads over walkers, qmc.jl:99-106, across devices).
 * One process per GPU: rank 0 calls gwz_nccl_unique_id, ships the bytes to the other ranks,
 * every rank calls gwz_context_comm_init.  One process driving several GPUs (a Julia host):
 * create one context per device and call gwz_context_comm_init_all once. */
int gwz_nccl_unique_id(uint8_t id[GWZ_NCCL_ID_BYTES]);
int gwz_context_comm_init(gwz_context *ctx, int world_size, int rank, const uint8_t id[GWZ_NCCL_ID_BYTES]);
int gwz_context_comm_init_all(gwz_context *const *ctxs, int n);
int gwz_context_comm_info(const gwz_context *ctx, int *world_size, int *rank);

/* ---- model: HubbardReal1D(BoseFS{N,M}; u, t) + GutzwillerAnsatz(H) ------------------------- */
/* replaces `H = HubbardReal1D(addr; u, t); ansatz = GutzwillerAnsatz(H)` (gutzwiller.jl:14-21) */
int gwz_model_create_hubbard1d_gutzwiller(gwz_context *ctx, int N, int M, double u, double t, gwz_model **out);
int gwz_model_destroy(gwz_model *m);
int gwz_model_info(const gwz_model *m, int *N, int *M, double *u, double *t, int *bits, int *words);
/* Rimu `near_uniform(BoseFS{N,M})`, `BoseFS(onr)`, `onr(addr)` -- host helpers, W words per address */
int gwz_model_near_uniform(const gwz_model *m, uint64_t *addr_words);
int gwz_model_from_onr(const gwz_model *m, const int32_t *onr, uint64_t *addr_words);
int gwz_model_to_onr(const gwz_model *m, const uint64_t *addr_words, int32_t *onr);
/* number of Fock states C(N+M-1, N) as a double, and exactly when it fits in 64 bits (else 0) */
int gwz_model_dimension(const gwz_model *m, double *dim, uint64_t *dim_exact);

/* ---- per-address probe: one kinetic_sample! per address (qmc.jl:14-44) --------------------- */
/* addr_words: n x W.  params: GWZ_NUM_PARAMS doubles.  u01: n uniforms in [0,1) or NULL (0.5).
 * Outputs (each may be NULL): e_loc[n] (qmc.jl:37), residence_time[n] (:33), grad_ratio[n x P]
 * (:38), n_hops[n] (:15), rates[n x 2M] = |psi(addr2)|/|psi(addr1)| in reference hop order,
 * zero padded (ENUMERATE only), chosen[n] 1-based reference hop index of the selected move,
 * next_addr[n x W] (:40-41). */
int gwz_probe_addresses(gwz_model *m, const uint64_t *addr_words, size_t n, const double *params,
                        const double *u01, int kernel_kind, double *e_loc, double *residence_time,
                        double *grad_ratio, int32_t *n_hops, double *rates, int32_t *chosen,
                        uint64_t *next_addr);
/* GutzwillerAnsatz value and gradient per address (gutzwiller.jl:25-36): val[n], grad[n x P] */
int gwz_ansatz_val_and_grad(gwz_model *m, const uint64_t *addr_words, size_t n, const double *params,
                            double *val, double *grad);

/* ---- walkers: KineticVQMCWalkerState x n (state.jl:10-28), resident in HBM ------------------ */
/* n_walkers local walkers whose global ids are global_offset .. global_offset+n_walkers-1 (the
 * Philox stream of a walker depends only on (seed, global id, step), so results do not depend on
 * how walkers are sharded).  start_addr: W words, broadcast to all walkers (state.jl:26). */
int gwz_walkers_create(gwz_model *m, uint64_t n_walkers, uint64_t global_offset, const uint64_t *start_addr,
                       uint64_t seed, gwz_walkers **out);
int gwz_walkers_destroy(gwz_walkers *w);
int gwz_walkers_count(const gwz_walkers *w, uint64_t *n_walkers, uint64_t *steps_done);
/* curr_address of every walker (n x W) -- continuation / checkpoint (state.jl:14, qmc.jl:95) */
int gwz_walkers_get_addresses(gwz_walkers *w, uint64_t *addr_words);
int gwz_walkers_set_addresses(gwz_walkers *w, const uint64_t *addr_words);
int gwz_walkers_set_step_counter(gwz_walkers *w, uint64_t steps_done);

/* val_err_and_grad(::KineticVQMCResult) / local_energy_estimator (state.jl:150-164, 265-285): per walker
 * blocking_analysis(resample(residence_times, local_energies)), combined over ALL walkers of all ranks */
typedef struct gwz_blocking_summary {
    double mean;                 /* mean over walkers of the blocking means (state.jl:163 `vals / walkers`) */
    double err;                  /* sqrt(sum_w err_w^2) / walkers (state.jl:163 = :279); NaN if any walker's M test
                                    failed, exactly as the reference's sum would be */
    double err_err;              /* sqrt(sum_w err_err_w^2) / walkers (state.jl:280) */
    double err_ok;               /* err over the walkers whose M test succeeded (= err when failed == 0) */
    double grad[GWZ_NUM_PARAMS]; /* mean over walkers of 2 mean_tau(G (E - b_w.mean)) (state.jl:57-58) */
    uint64_t walkers;            /* walkers analysed (all ranks); 0 = no analysis was run */
    uint64_t failed;             /* walkers with k = -1 ("M test failed, more data needed") */
    int32_t k_min, k_max;        /* accepted blocking levels over the successful walkers (-1 if none) */
    uint64_t series_steps;       /* length of each walker's series */
    uint64_t resample_len;       /* length it was resampled to (state.jl:183 `len`) */
    float kernel_ms;             /* device time of the blocking kernel (this rank) */
    float reserved;
} gwz_blocking_summary;

typedef struct gwz_vqmc_result {
    double val;                      /* mean over walkers of per-walker ratio estimates (state.jl:137-149) */
    double err;                      /* standard error of that mean across walkers (NaN for 1 walker) */
    double grad[GWZ_NUM_PARAMS];     /* state.jl:42-49 then :137-149 */
    double grad_err[GWZ_NUM_PARAMS]; /* cross-walker standard error of grad */
    double pooled_val;               /* sum tau E / sum tau over all walkers and steps */
    double pooled_grad[GWZ_NUM_PARAMS];
    double pooled_var;               /* time-weighted variance of E_loc over all walkers and steps */
    double walker_var;               /* var(res) (state.jl:293-298): mean over walkers of var(E_loc, FrequencyWeights(tau)),
                                        uncorrected (StatsBase default) */
    double total_time;               /* sum of residence times */
    uint64_t walkers;                /* global walker count (all ranks) */
    uint64_t steps;                  /* global walker-steps sampled by THIS call (with GWZ_RUN_KEEP_ACC the estimators
                                        also hold the earlier calls' steps) */
    uint64_t hops;                   /* global hop evaluations of this call = sum over its steps of length(offdiagonals) */
    double acc[GWZ_NUM_ACC];         /* raw all-reduced accumulator vector */
    float kernel_ms;                 /* device time of the sampling kernel(s) of this call (this rank) */
    float reserved;
    gwz_blocking_summary blocking;   /* filled by GWZ_RUN_BLOCKING, else zero (walkers = 0) */
} gwz_vqmc_result;

/* kinetic_vqmc!(states; steps) (qmc.jl:73-106) + val_and_grad(result) (state.jl:137-149) + the
 * cross-device all-reduce.  burn_in steps are sampled first without accumulating (0 = reference
 * behaviour).  Collective when the context has a communicator: every rank must call it. */
int gwz_vqmc_run(gwz_walkers *w, const double *params, uint64_t steps_per_walker, uint64_t burn_in,
                 int kernel_kind, unsigned flags, gwz_vqmc_result *out);
/* same, for several walker sets living on different devices of ONE process (contexts joined by
 * gwz_context_comm_init_all); out receives the global result */
int gwz_vqmc_run_group(gwz_walkers *const *ws, int n, const double *params, uint64_t steps_per_walker,
                       uint64_t burn_in, int kernel_kind, unsigned flags, gwz_vqmc_result *out);
/* per-walker estimates of the estimators accumulated so far: val_w[n], grad_w[n x P] (state.jl:42-49) */
int gwz_walkers_estimates(gwz_walkers *w, double *val_w, double *grad_w);
/* Recording run: like gwz_vqmc_run but also returns the per-step series the reference stores
 * (qmc.jl:89-92): tau[steps x n], eloc[steps x n] (step-major), addr[steps x n x W] (may be NULL). */
int gwz_vqmc_run_record(gwz_walkers *w, const double *params, uint64_t steps_per_walker, int kernel_kind,
                        unsigned flags, double *tau, double *eloc, uint64_t *addr, gwz_vqmc_result *out);

/* Parity surface of the PRODUCTION kernel (GWZ_KERNEL_BITPARALLEL, 4 <= M <= 250): steps_per_walker recorded steps; per
 * step the residence time, the local energy and the selected hop, hop = 2 z + right with z the 0-based mode on the
 * left of the bond (z, z+1 mod M), right = 1: a particle moves z -> z+1, 0: z+1 -> z; -1: no move.  A trajectory is
 * its start address plus its hops (qmc.jl:40-41, 89-92).  u01 (may be NULL): steps x n uniforms in [0,1) used instead
 * of the walkers' Philox stream, so that one kinetic_sample! per address can be probed with a prescribed u.
 * tau, eloc, hop: steps x n, step-major. */
int gwz_vqmc_run_trace(gwz_walkers *w, const double *params, uint64_t steps_per_walker, unsigned flags, const double *u01,
                       double *tau, double *eloc, int32_t *hop, gwz_vqmc_result *out);

/* Device-side error bars.  Analyses the series kept by the last GWZ_RUN_SERIES run(s) of this walker set (they must
 * cover the accumulated steps): per walker resample to resample_len values (0 = the series length, the default of
 * state.jl:183) and blocking_analysis, then the cross-walker combination.  Collective when the context has a
 * communicator.  Per-walker outputs (each may be NULL, n local walkers): mean_w, err_w, err_err_w (NaN when the M
 * test failed), k_w (-1 failed), blocks_w.  One thread per walker, series read from HBM once. */
int gwz_walkers_blocking(gwz_walkers *w, uint64_t resample_len, gwz_blocking_summary *out, double *mean_w,
                         double *err_w, double *err_err_w, int32_t *k_w, int64_t *blocks_w);
/* the series themselves (host copies; tau, eloc: steps x n, step-major; either may be NULL); *steps = series length */
int gwz_walkers_series(gwz_walkers *w, uint64_t *steps, double *tau, double *eloc);
/* The same analysis for caller-supplied series (host pointers, step-major steps x n): n independent series through
 * the same kernel; never collective.  This is the parity surface of the blocking kernel.  resampled (may be NULL;
 * resample_len x n, step-major, needs resample_len >= 2): the equal-time series the kernel analysed, i.e.
 * resample(residence_times, values; len) of state.jl:183-231 per column. */
int gwz_series_blocking(gwz_context *ctx, const double *tau, const double *eloc, uint64_t steps, uint64_t n,
                        uint64_t resample_len, gwz_blocking_summary *out, double *mean_w, double *err_w,
                        double *err_err_w, int32_t *k_w, int64_t *blocks_w, double *resampled);

/* DVec(res) (state.jl:303-319): runs steps_per_walker recorded steps and deposits every step's residence time on its
 * address in an HBM hash table; returns the distinct addresses in ascending order with the 2-norm-normalised
 * values (normalize!(dv)).  cap = capacity of addr_out / val_out; *count = number of distinct addresses (if it exceeds
 * cap the call fails with GWZ_ERR_INVALID after setting *count).  Single-word addresses only (N+M-1 <= 64). */
int gwz_vqmc_sample_vector(gwz_walkers *w, const double *params, uint64_t steps_per_walker, int kernel_kind,
                           unsigned flags, uint64_t cap, uint64_t *addr_out, double *val_out, uint64_t *count,
                           gwz_vqmc_result *out);

/* ---- LocalEnergyEvaluator (src/localenergy.jl:81-173) --------------------------------------- */
/* basis_words: dim x W stored basis (reference layout `basis::Vector{A}`, localenergy.jl:84), copied
 * to HBM once; or NULL => all C(N+M-1,N) states enumerated on the fly by rank in [rank_begin,
 * rank_end) (rank_end = 0 => dimension), requires N+M-1 <= 64. */
int gwz_sweep_create(gwz_model *m, const uint64_t *basis_words, uint64_t dim, uint64_t rank_begin,
                     uint64_t rank_end, gwz_sweep **out);
int gwz_sweep_destroy(gwz_sweep *s);
int gwz_sweep_dim(const gwz_sweep *s, uint64_t *dim);
/* le(p) (localenergy.jl:131-155) */
int gwz_sweep_val(gwz_sweep *s, const double *params, double *val);
/* val_and_grad(le, p) (localenergy.jl:94-129); acc4 (may be NULL) = (num, den, num_grad, den_grad)
 * after the cross-device all-reduce */
int gwz_sweep_val_and_grad(gwz_sweep *s, const double *params, double *val, double *grad, double *acc4);
/* per-state (num, den, num_grad, den_grad) of localenergy.jl:105-120 for states first..first+count-1
 * of this sweep: out[count x 4] */
int gwz_sweep_terms(gwz_sweep *s, const double *params, uint64_t first, uint64_t count, double *out);
/* the states themselves (enumeration order of this sweep): out[count x W] */
int gwz_sweep_states(gwz_sweep *s, uint64_t first, uint64_t count, uint64_t *out);
/* device time (ms) of the last sweep kernel */
int gwz_sweep_last_kernel_ms(const gwz_sweep *s, float *ms);
/* Host-only helper (no device work): a ranked sweep of HubbardReal1D + GutzwillerAnsatz evaluates one representative
 * per orbit of the chain's translations and reflection (the largest rotation of the unary string and of its mirror
 * image), weighted with the orbit size.  Representatives carry the largest occupation in mode M, so only the ranks
 * [*first_rank, *first_rank + *count) of [rank_begin, rank_begin + dim) are visited; at commensurate filling the uniform
 * state (*uniform_state, if *has_uniform) is evaluated on its own. */
int gwz_sweep_representative_range(int N, int M, uint64_t rank_begin, uint64_t dim, uint64_t *first_rank,
                                   uint64_t *count, int32_t *has_uniform, uint64_t *uniform_state);

/* ---- error bars: resample! (state.jl:194-231) + Rimu.StatsTools.blocking_analysis ---------- */
int gwz_resample(double *result, size_t len, const double *residence_times, const double *values, size_t n);
typedef struct gwz_blocking_result {
    double mean, err, err_err, p_cov;
    int32_t k; /* -1: blocking unsuccessful (state.jl:250-253) */
    int64_t blocks;
} gwz_blocking_result;
int gwz_blocking_analysis(const double *v, size_t n, gwz_blocking_result *out);

/* ---- adaptive_gradient_descent / gradient_descent / amsgrad (src/amsgrad.jl:105-263) -------- */
typedef int (*gwz_objective_fn)(void *user, const double *p, int np, double *val, double *err, double *grad);
typedef struct gwz_gd_opts {
    int32_t maxiter;     /* 100 */
    int32_t use_amsgrad; /* 1: amsgrad (amsgrad.jl:259-263), 0: gradient_descent (:240-244) */
    int32_t early_stop;  /* 1 */
    int32_t reserved;
    double alpha, beta1, beta2;                   /* 0.01, 0.1, 0.01 */
    double grad_tol, param_tol, val_tol, err_tol; /* sqrt(eps) x3, 0.0 */
    const double *first_moment_init;              /* NULL: gradient at p0 (amsgrad.jl:138-142) */
    const double *second_moment_init;             /* NULL: zeros (:143-147) */
    const uint8_t *fix_params;                    /* NULL: none */
} gwz_gd_opts;
int gwz_gd_opts_default(gwz_gd_opts *o);
/* rows = maxiter + 1; vector traces hold rows x np doubles, scalar traces rows doubles */
typedef struct gwz_gd_trace {
    int32_t iterations, converged, reason; /* reason: 0 iterations 1 params 2 gradient 3 value 4 errorbars */
    int32_t np;
    double *param, *value, *error, *gradient, *first_moment, *second_moment, *param_delta;
} gwz_gd_trace;
/* generic driver: any objective (amsgrad.jl:117-226) */
int gwz_adaptive_gradient_descent(gwz_objective_fn val_and_grad, gwz_objective_fn val_err_and_grad, void *user,
                                  const double *p0, int np, const gwz_gd_opts *opts, gwz_gd_trace *trace);
/* built-in objectives: amsgrad(qmc::KineticVQMC, p0) and amsgrad(le::LocalEnergyEvaluator, p0).
 * gwz_amsgrad_vqmc: every evaluation is val_err_and_grad(qmc, p) (wrapper.jl:98-105): burn_in unaccumulated steps,
 * then steps_per_walker steps from the current walker positions.  flags | GWZ_RUN_BLOCKING: value, error bar and
 * gradient of state.jl:150-164 (device-side blocking; the error feeds the err_tol stop, amsgrad.jl:204); without it
 * the error column is the cross-walker standard error.  last (may be NULL): result of the final evaluation, what the
 * reference leaves in qmc.result. */
int gwz_amsgrad_vqmc(gwz_walkers *w, const double *p0, const gwz_gd_opts *opts, uint64_t steps_per_walker,
                     uint64_t burn_in, int kernel_kind, unsigned flags, gwz_gd_trace *trace, gwz_vqmc_result *last);
int gwz_amsgrad_sweep(gwz_sweep *s, const double *p0, const gwz_gd_opts *opts, gwz_gd_trace *trace);

/* ---- ansatz-generic engine ---------------------------------------------
 * The reference's other single-component ansaetze, and Rimu's ExtendedHubbardReal1D, through the same
 * three operations (probe / walkers / sweep) with P = num_parameters(ansatz) parameters.  One warp per
 * walker; hops are visited in the reference's offdiagonals order and the move is the reference's
 * pick_random_from_cumsum (qmc.jl:52-61), so a walker follows the reference trajectory for the same
 * uniform stream.  Parameter vectors are P doubles in the reference's order. */
enum gwz_ansatz_kind {
    GWZ_ANSATZ_GUTZWILLER = 0,       /* src/ansatz/gutzwiller.jl:14-36:   P = 1 */
    GWZ_ANSATZ_EXT_GUTZWILLER = 1,   /* src/ansatz/extgutzwiller.jl:14-35 (call operator): P = 2 */
    GWZ_ANSATZ_MULTINOMIAL = 2,      /* src/ansatz/multinomial.jl:17-59:  P = 1 */
    GWZ_ANSATZ_DENSITY_PROFILE = 3,  /* src/ansatz/densityprofile.jl:8-28: P = M */
    GWZ_ANSATZ_RELATIVE_JASTROW = 4, /* src/ansatz/jastrow.jl:71-103:     P = cld(M, 2) */
    GWZ_ANSATZ_JASTROW = 5,          /* src/ansatz/jastrow.jl:13-58:      P = M (M + 1) / 2 */
    GWZ_ANSATZ_COMBINATION = 6       /* src/ansatz/combination.jl: a1 + a2, P = P1 + P2 + 1 (gwz_ansatz_create_combination) */
};
typedef struct gwz_ansatz gwz_ansatz;
typedef struct gwz_gwalkers gwz_gwalkers;
typedef struct gwz_gsweep gwz_gsweep;

/* Rimu ExtendedHubbardReal1D(addr; u, v, t): diagonal u sum n(n-1)/2 + v sum n_i n_{i+1} (periodic).
 * Such a model is accepted by the gwz_gen_* / gwz_gwalkers_* / gwz_gsweep_* entry points only. */
int gwz_model_create_ext_hubbard1d(gwz_context *ctx, int N, int M, double u, double v, double t, gwz_model **out);
int gwz_model_info_ext(const gwz_model *m, double *v);
/* `ansatz = XAnsatz(H)`; normalize: MultinomialAnsatz(H; normalize) (multinomial.jl:22-31) */
int gwz_ansatz_create(gwz_model *m, int kind, int normalize, gwz_ansatz **out);
/* `a1 + a2` (combination.jl:13-15): alpha a1 + (1 - alpha) a2 with parameters (p1..., p2..., alpha); both components
 * from the kinds above */
int gwz_ansatz_create_combination(gwz_model *m, int kind1, int normalize1, int kind2, int normalize2, gwz_ansatz **out);
int gwz_ansatz_destroy(gwz_ansatz *a);
int gwz_ansatz_info(const gwz_ansatz *a, int *kind, int *num_parameters);
/* ansatz(addr, p) and val_and_grad(ansatz, addr, p): val[n], grad[n x P] (either may be NULL) */
int gwz_gen_ansatz_eval(gwz_ansatz *a, const uint64_t *addr_words, size_t n, const double *params, double *val,
                        double *grad);
/* one kinetic_sample! per address (qmc.jl:14-44); outputs as gwz_probe_addresses, grad_ratio[n x P] */
int gwz_gen_probe_addresses(gwz_ansatz *a, const uint64_t *addr_words, size_t n, const double *params,
                            const double *u01, double *e_loc, double *residence_time, double *grad_ratio,
                            int32_t *n_hops, double *rates, int32_t *chosen, uint64_t *next_addr);

int gwz_gwalkers_create(gwz_ansatz *a, uint64_t n_walkers, uint64_t global_offset, const uint64_t *start_addr,
                        uint64_t seed, gwz_gwalkers **out);
int gwz_gwalkers_destroy(gwz_gwalkers *w);
int gwz_gwalkers_count(const gwz_gwalkers *w, uint64_t *n_walkers, uint64_t *steps_done);
int gwz_gwalkers_get_addresses(gwz_gwalkers *w, uint64_t *addr_words);
int gwz_gwalkers_set_addresses(gwz_gwalkers *w, const uint64_t *addr_words);
int gwz_gwalkers_set_step_counter(gwz_gwalkers *w, uint64_t steps_done);
typedef struct gwz_gen_result {
    double val, err;        /* mean over walkers of the per-walker estimates, its cross-walker standard error */
    double pooled_val, pooled_var, total_time;
    uint64_t walkers, steps, hops;
    float kernel_ms;
    int32_t np;
} gwz_gen_result;
/* as gwz_vqmc_run; grad, grad_err, pooled_grad: P doubles each (any may be NULL) */
int gwz_gen_vqmc_run(gwz_gwalkers *w, const double *params, uint64_t steps_per_walker, uint64_t burn_in,
                     unsigned flags, gwz_gen_result *out, double *grad, double *grad_err, double *pooled_grad);
int gwz_gen_vqmc_run_record(gwz_gwalkers *w, const double *params, uint64_t steps_per_walker, unsigned flags,
                            double *tau, double *eloc, uint64_t *addr, gwz_gen_result *out, double *grad,
                            double *grad_err, double *pooled_grad);
int gwz_gwalkers_estimates(gwz_gwalkers *w, double *val_w, double *grad_w);

/* LocalEnergyEvaluator(H, ansatz; basis) (localenergy.jl:81-173) for any of the ansaetze above */
int gwz_gsweep_create(gwz_ansatz *a, const uint64_t *basis_words, uint64_t dim, uint64_t rank_begin,
                      uint64_t rank_end, gwz_gsweep **out);
int gwz_gsweep_destroy(gwz_gsweep *s);
int gwz_gsweep_dim(const gwz_gsweep *s, uint64_t *dim);
/* grad == NULL: le(p) (localenergy.jl:131-155); else val_and_grad(le, p) (:94-129), grad[P] */
int gwz_gsweep_val_and_grad(gwz_gsweep *s, const double *params, double *val, double *grad);
/* per-state (num, den, num_grad[P], den_grad[P]): out[count x (2 + 2P)] */
int gwz_gsweep_terms(gwz_gsweep *s, const double *params, uint64_t first, uint64_t count, double *out);
int gwz_gsweep_states(gwz_gsweep *s, uint64_t first, uint64_t count, uint64_t *out);
int gwz_gsweep_last_kernel_ms(const gwz_gsweep *s, float *ms);

/* ---- VectorAnsatz (src/ansatz/vector.jl:1-12): a stored vector as a 0-parameter ansatz --------
 * The (address, amplitude) pairs are kept in an HBM hash table; addresses the vector does not hold have amplitude 0
 * (Rimu DVec semantics).  HubbardReal1D only. */
typedef struct gwz_vansatz gwz_vansatz;
typedef struct gwz_vwalkers gwz_vwalkers;
/* `VectorAnsatz(vector)`: addr_words n x W, values n */
int gwz_vector_ansatz_create(gwz_model *m, const uint64_t *addr_words, const double *values, uint64_t n,
                             gwz_vansatz **out);
int gwz_vector_ansatz_destroy(gwz_vansatz *a);
/* va(addr, _) (vector.jl:12), batched */
int gwz_vector_ansatz_lookup(gwz_vansatz *a, const uint64_t *addr_words, size_t n, double *val);
/* LocalEnergyEvaluator(H, VectorAnsatz(v))(): the Rayleigh quotient over basis = keys(v) (vector.jl:10) */
int gwz_vector_ansatz_rayleigh(gwz_vansatz *a, double *val);
/* kinetic_vqmc(H, VectorAnsatz(v), Float64[]; ...) (test/qmc.jl:22-23): as gwz_gwalkers_* with P = 0 */
int gwz_vwalkers_create(gwz_vansatz *a, uint64_t n_walkers, uint64_t global_offset, const uint64_t *start_addr,
                        uint64_t seed, gwz_vwalkers **out);
int gwz_vwalkers_destroy(gwz_vwalkers *w);
int gwz_vwalkers_get_addresses(gwz_vwalkers *w, uint64_t *addr_words);
int gwz_vwalkers_run(gwz_vwalkers *w, uint64_t steps_per_walker, uint64_t burn_in, unsigned flags, gwz_gen_result *out);
int gwz_vwalkers_run_record(gwz_vwalkers *w, uint64_t steps_per_walker, unsigned flags, double *tau, double *eloc,
                            uint64_t *addr, gwz_gen_result *out);

/* ---- Stored operators: the path for ANY Hamiltonian and ANY ansatz --------
 * kinetic_sample! (src/kinetic-vqmc/qmc.jl:14-44) and LocalEnergyEvaluator (src/localenergy.jl:94-155) only call
 * offdiagonals(H, k), diagonal_element(H, k) and ansatz(k, params).  For operators other than the Hubbard chain
 * (HubbardMom1D, HubbardRealSpace, ... of test/qmc.jl:10-14) the caller materialises them over basis =
 * build_basis(H): row k = offdiagonals(H, basis[k]) in the reference's order as (column index, matrix element),
 * diag[k] = diagonal_element(H, basis[k]); the ansatz is a table psi[k] = ansatz(basis[k], params) with
 * grad_ratio[k][p] = (d psi / d p_p) / psi.  Walker state is the basis index.  Results use gwz_gen_result. */
typedef struct gwz_sparse gwz_sparse;
typedef struct gwz_swalkers gwz_swalkers;
/* row_ptr[dim + 1] (row_ptr[0] = 0), col[nnz] < dim, melem[nnz], diag[dim]; dim < 2^32.  Duplicate targets inside a
 * row are allowed (Rimu lists both directions of a length-2 periodic axis). */
int gwz_sparse_create(gwz_context *ctx, uint64_t dim, const uint64_t *row_ptr, const uint32_t *col,
                      const double *melem, const double *diag, gwz_sparse **out);
int gwz_sparse_destroy(gwz_sparse *h);
int gwz_sparse_info(const gwz_sparse *h, uint64_t *dim, uint64_t *nnz, uint64_t *max_row, int32_t *num_parameters,
                    int32_t *lanes_per_walker);
/* layout the walkers read after the last table was set: row_blocks = 1 when every state's row sits in a fixed-stride
 * block (one dependent HBM access per step; chosen when the rows are regular enough), block_bytes = its stride */
int gwz_sparse_layout(const gwz_sparse *h, int32_t *row_blocks, uint64_t *block_bytes);
/* any ansatz as a table (P = num_parameters; grad_ratio may be NULL when P = 0, e.g. VectorAnsatz, vector.jl:6-12) */
int gwz_sparse_set_table(gwz_sparse *h, const double *psi, const double *grad_ratio, int32_t P);
/* GutzwillerAnsatz(H) (gutzwiller.jl:25-36) filled on the device: psi = exp(-g diag), grad psi / psi = -diag */
int gwz_sparse_set_gutzwiller(gwz_sparse *h, double g);
int gwz_sparse_get_table(gwz_sparse *h, double *psi, double *grad_ratio);
/* per-index kinetic_sample! outputs; rates: n x max_row (|psi_l / psi_k| in row order, 0 padded), chosen 1-based */
int gwz_sparse_probe(gwz_sparse *h, const uint32_t *idx, size_t n, const double *u01, double *e_loc,
                     double *residence_time, double *grad_ratio, int32_t *n_hops, double *rates, int32_t *chosen,
                     uint32_t *next_idx);
/* val_and_grad(le, params) (localenergy.jl:94-129; grad == NULL: le(params), :131-155) over rows
 * [row_begin, row_end); with a communicator the sums are all-reduced, so each rank may pass its slice of the rows */
int gwz_sparse_sweep(gwz_sparse *h, uint64_t row_begin, uint64_t row_end, double *val, double *grad);
/* per-row (num, den, num_grad[P], den_grad[P]): out[count x (2 + 2P)] */
int gwz_sparse_sweep_terms(gwz_sparse *h, uint64_t first, uint64_t count, double *out);
int gwz_sparse_last_kernel_ms(const gwz_sparse *h, float *ms);
/* walkers: global ids [global_offset, global_offset + n) key the random streams as everywhere else */
int gwz_swalkers_create(gwz_sparse *h, uint64_t n_walkers, uint64_t global_offset, uint32_t start_index,
                        uint64_t seed, gwz_swalkers **out);
int gwz_swalkers_destroy(gwz_swalkers *w);
int gwz_swalkers_get_indices(gwz_swalkers *w, uint32_t *idx);
int gwz_swalkers_set_indices(gwz_swalkers *w, const uint32_t *idx);
int gwz_swalkers_set_step_counter(gwz_swalkers *w, uint64_t steps_done);
/* as gwz_gen_vqmc_run, with the table currently set on the operator */
int gwz_swalkers_run(gwz_swalkers *w, uint64_t steps_per_walker, uint64_t burn_in, unsigned flags,
                     gwz_gen_result *out, double *grad, double *grad_err, double *pooled_grad);
/* series are [step][walker]; idx (optional) = the index each step was sampled from (qmc.jl:89) */
int gwz_swalkers_run_record(gwz_swalkers *w, uint64_t steps_per_walker, unsigned flags, double *tau, double *eloc,
                            uint32_t *idx, gwz_gen_result *out, double *grad, double *grad_err, double *pooled_grad);
/* DVec(res) (state.jl:303-319): total residence time per basis index, hist[dim] (all-reduced with a communicator) */
int gwz_swalkers_sample_vector(gwz_swalkers *w, uint64_t steps_per_walker, uint64_t burn_in, unsigned flags,
                               double *hist, gwz_gen_result *out);
/* per-walker estimates (state.jl:42-49): val_w[n], grad_w[n x P] */
int gwz_swalkers_estimates(gwz_swalkers *w, double *val_w, double *grad_w);

#ifdef __cplusplus
}
#endif
#endif /* GUTZWILLER_B200_H */
