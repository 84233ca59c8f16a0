/* c_consumer.c -- a plain C99 consumer of include/gutzwiller_b200.h, compiled with gcc and linked against
 * libgutzwiller_b200.so by tests/test_capi_and_host.py.  It does what a foreign-language host (the Julia `ccall`
 * shim) does: fills gwz_gd_opts / gwz_gd_trace / gwz_vqmc_result through the header and runs BASELINE config 1
 * (HubbardReal1D BoseFS{10,10} u = 2, GutzwillerAnsatz g = 1), one LocalEnergyEvaluator sweep and a short AMSGrad
 * run.  Struct layout and calling convention are therefore checked by a compiler, not by a regular expression.
 *
 * Exit status: 0 = all checks passed; 77 = no CUDA device (the library refused, as it must: it has no CPU path);
 * anything else = failure (message on stderr). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gutzwiller_b200.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int rc_ = (call);                                                                    \
        if (rc_ != GWZ_OK) {                                                                 \
            fprintf(stderr, "%s:%d: %s -> %d: %s\n", __FILE__, __LINE__, #call, rc_, gwz_last_error()); \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)
#define REQUIRE(cond)                                                        \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "%s:%d: check failed: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(void) {
    if (gwz_version() != GWZ_VERSION) {
        fprintf(stderr, "header %d vs library %d\n", GWZ_VERSION, gwz_version());
        return 1;
    }
    gwz_context *ctx = NULL;
    int rc = gwz_context_create(0, &ctx);
    if (rc == GWZ_ERR_CUDA) {
        printf("no CUDA device: %s\n", gwz_last_error());
        return 77;
    }
    REQUIRE(rc == GWZ_OK && ctx != NULL);

    /* H = HubbardReal1D(near_uniform(BoseFS{10,10}); u = 2.0); ansatz = GutzwillerAnsatz(H)  (README.md:33-40) */
    gwz_model *model = NULL;
    CHECK(gwz_model_create_hubbard1d_gutzwiller(ctx, 10, 10, 2.0, 1.0, &model));
    int N, M, bits, words;
    double u, t;
    CHECK(gwz_model_info(model, &N, &M, &u, &t, &bits, &words));
    REQUIRE(N == 10 && M == 10 && bits == 19 && words == 1 && u == 2.0 && t == 1.0);
    uint64_t start[GWZ_MAX_WORDS] = {0};
    CHECK(gwz_model_near_uniform(model, start));
    REQUIRE(start[0] == 0x55555ull); /* 0b1010101010101010101, README.md:43 */

    /* a handle of the wrong type is an error message, not a crash */
    gwz_vqmc_result junk;
    double p1[GWZ_NUM_PARAMS] = {1.0};
    REQUIRE(gwz_vqmc_run((gwz_walkers *)model, p1, 10, 0, GWZ_KERNEL_BITPARALLEL, 0u, &junk) == GWZ_ERR_INVALID);
    REQUIRE(strstr(gwz_last_error(), "foreign handle") != NULL);

    /* le = LocalEnergyEvaluator(H, ansatz); val_and_grad(le, [1.0])  (README.md:94,104) */
    gwz_sweep *sweep = NULL;
    CHECK(gwz_sweep_create(model, NULL, 0, 0, 0, &sweep));
    uint64_t dim = 0;
    CHECK(gwz_sweep_dim(sweep, &dim));
    REQUIRE(dim == 92378ull);
    double val = 0.0, grad[GWZ_NUM_PARAMS] = {0.0}, acc4[4];
    CHECK(gwz_sweep_val_and_grad(sweep, p1, &val, grad, acc4));
    REQUIRE(fabs(val - (-7.825819465047312)) < 1e-11 && fabs(grad[0] - 10.614147776143358) < 1e-10);
    REQUIRE(fabs(acc4[0] / acc4[1] - val) < 1e-13);

    /* kinetic_vqmc(H, ansatz, [1.0]; walkers = 1, steps = 1e5): BASELINE config 1, error bars from the device-side
     * blocking analysis (README.md:218 prints -7.8369 +- 0.016555 for its own random stream) */
    gwz_walkers *one = NULL;
    CHECK(gwz_walkers_create(model, 1, 0, start, 12345ull, &one));
    gwz_vqmc_result r;
    memset(&r, 0xAB, sizeof r);
    CHECK(gwz_vqmc_run(one, p1, 100000, 0, GWZ_KERNEL_BITPARALLEL, GWZ_RUN_BLOCKING, &r));
    REQUIRE(r.walkers == 1 && r.steps == 100000 && r.hops > 1000000 && r.total_time > 0.0);
    REQUIRE(r.blocking.walkers == 1 && r.blocking.failed == 0 && r.blocking.series_steps == 100000);
    REQUIRE(r.blocking.err > 0.005 && r.blocking.err < 0.05 && r.blocking.k_min >= 3);
    REQUIRE(fabs(r.blocking.mean - r.val) < 1e-9 && fabs(r.val - val) < 5.0 * r.blocking.err);
    REQUIRE(fabs(r.acc[0] - 1.0) < 1e-15 && r.walker_var > 0.0 && r.reserved == 0.0f);
    /* kinetic_vqmc!(res; steps): the estimators and the series continue */
    CHECK(gwz_vqmc_run(one, p1, 50000, 0, GWZ_KERNEL_BITPARALLEL, GWZ_RUN_KEEP_ACC | GWZ_RUN_BLOCKING, &r));
    REQUIRE(r.steps == 50000 && r.blocking.series_steps == 150000);
    uint64_t nw = 0, done = 0;
    CHECK(gwz_walkers_count(one, &nw, &done));
    REQUIRE(nw == 1 && done == 150000);

    /* amsgrad(KineticVQMC(H, ansatz; samples = 2^22, walkers = 2^14), [1.0]; maxiter = 20) and the same on le */
    gwz_walkers *many = NULL;
    CHECK(gwz_walkers_create(model, 1u << 14, 0, start, 777ull, &many));
    gwz_gd_opts opts;
    CHECK(gwz_gd_opts_default(&opts));
    REQUIRE(opts.maxiter == 100 && opts.use_amsgrad == 1 && opts.alpha == 0.01 && opts.beta1 == 0.1 && opts.beta2 == 0.01);
    REQUIRE(opts.first_moment_init == NULL && opts.fix_params == NULL && opts.err_tol == 0.0);
    opts.maxiter = 20;
    enum { ROWS = 21 };
    double param[ROWS], value[ROWS], error[ROWS], gradient[ROWS], m1[ROWS], m2[ROWS], dp[ROWS];
    gwz_gd_trace tr;
    memset(&tr, 0, sizeof tr);
    tr.param = param;
    tr.value = value;
    tr.error = error;
    tr.gradient = gradient;
    tr.first_moment = m1;
    tr.second_moment = m2;
    tr.param_delta = dp;
    gwz_vqmc_result last;
    CHECK(gwz_amsgrad_vqmc(many, p1, &opts, 256, 50, GWZ_KERNEL_BITPARALLEL, GWZ_RUN_BLOCKING, &tr, &last));
    REQUIRE(tr.iterations == ROWS && tr.np == 1 && tr.converged == 0 && tr.reason == 0);
    REQUIRE(param[0] == 1.0 && param[ROWS - 1] < 0.9 && value[ROWS - 1] < value[0]);
    REQUIRE(error[0] > 0.0 && error[0] < 0.05 && last.blocking.walkers == (1u << 14) && last.val == last.val);
    REQUIRE(fabs(m2[0] - 0.01 * gradient[0] * gradient[0]) < 1e-12 * m2[0]); /* amsgrad.jl:261 from m2 = 0 */
    REQUIRE(fabs(dp[0] + 0.01 * m1[0] / sqrt(m2[0])) < 1e-15);               /* amsgrad.jl:173 */
    /* fix_params freezes the parameter; moment initialisers are honoured */
    unsigned char fixed[1] = {1};
    double m1_init[1] = {3.0}, m2_init[1] = {4.0};
    opts.fix_params = fixed;
    opts.first_moment_init = m1_init;
    opts.second_moment_init = m2_init;
    opts.maxiter = 2;
    opts.early_stop = 0;
    CHECK(gwz_amsgrad_sweep(sweep, p1, &opts, &tr));
    REQUIRE(tr.iterations == 3 && param[2] == 1.0 && dp[0] == 0.0 && gradient[0] == 0.0);
    REQUIRE(fabs(m1[0] - (0.9 * 3.0 + 0.1 * 10.614147776143358)) < 1e-9 && m2[0] >= 4.0);
    REQUIRE(fabs(value[0] - val) < 1e-12 && error[0] == 0.0);

    CHECK(gwz_walkers_destroy(many));
    CHECK(gwz_walkers_destroy(one));
    CHECK(gwz_sweep_destroy(sweep));
    CHECK(gwz_model_destroy(model));
    CHECK(gwz_context_destroy(ctx));
    printf("c consumer ok: le = %.15f grad = %.15f; vqmc = %.5f +- %.5f (k = %d); amsgrad p -> %.4f\n", val, grad[0],
           r.val, r.blocking.err, (int)r.blocking.k_min, param[ROWS - 1]);
    return 0;
}
