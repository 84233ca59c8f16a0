// gwz_api_generic.cu -- C ABI of the ansatz-generic engine (include/gutzwiller_b200.h, second part).
//
// Host-side orchestration only: the parameter-dependent tables of the log-linear form
// (Q, b, hop constants; O(M^2) doubles per call), HBM buffers, launches, the NCCL all-reduce of the
// accumulator vector.  No sampled or swept quantity is computed on the host.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "gwz_handles.h"

using namespace gwz;

struct gwz_ansatz {
    uint32_t magic = gwz::GWZ_TAG_ANSATZ;  // handle type tag, checked by every entry point
    int device = 0;  // destroy must not dereference parent handles (a host GC may release them in any order)
    gwz_model* m;
    int kind, P, normalize;
    double logK;
    int Mpad;
    GenModel gm;            // device pointers filled in; Qt/bvec/cR/cL/s updated per call
    double* d_static = nullptr;  // logtab, lgam, sqrtab
    uint16_t* d_pairs = nullptr;
    double* d_par = nullptr;  // Qt[M*M], bvec[M], cR[M], cL[M]
    double* h_par = nullptr;  // pinned staging of d_par
    // CombinationAnsatz (combination.jl): two owned component handles; kind = GWZ_ANSATZ_COMBINATION, P = P1 + P2 + 1
    gwz_ansatz* c1 = nullptr;
    gwz_ansatz* c2 = nullptr;
};

struct gwz_gwalkers {
    uint32_t magic = gwz::GWZ_TAG_GWALKERS;  // handle type tag, checked by every entry point
    int device = 0;
    gwz_ansatz* a;
    size_t n;
    uint64_t offset, seed, steps_done, acc_steps;
    uint8_t* d_occ = nullptr;
    double* d_acc = nullptr;
    double* d_partials = nullptr;
    size_t partials_rows = 0;
    double* d_vec = nullptr;
    double* h_vec = nullptr;    // pinned
    double* d_shift = nullptr;  // [1 + P]
    double* h_shift = nullptr;  // pinned
    bool have_shift = false;
    uint64_t* d_addr1 = nullptr;  // W words scratch
};

struct gwz_gsweep {
    uint32_t magic = gwz::GWZ_TAG_GSWEEP;  // handle type tag, checked by every entry point
    int device = 0;
    gwz_ansatz* a;
    uint64_t* d_basis = nullptr;  // SoA [W][dim]
    uint64_t dim = 0, rank_begin = 0;
    uint64_t* d_binom = nullptr;
    double* d_partials = nullptr;
    size_t partials_rows = 0;
    double* d_vec = nullptr;
    double* h_vec = nullptr;
    float last_ms = 0.f;
    // ranked sweeps of an ansatz that is invariant under translations and reflection of the chain: the orbit
    // representatives (state, orbit size) are stored at the second sweep of a handle and evaluated from then on
    std::vector<uint64_t> h_binom;
    uint64_t* d_rep = nullptr;
    uint32_t* d_mult = nullptr;
    uint64_t nrep = 0, calls = 0, build_at = 3;
    float t_enum = 0.f;
    int rep_state = 0;  // 0 not built, 1 built, -1 not applicable / not cacheable
};

// ------------------------------------------------------------------------------------------
// model / ansatz
// ------------------------------------------------------------------------------------------
extern "C" int gwz_model_create_ext_hubbard1d(gwz_context* ctx, int N, int M, double u, double v, double t,
                                              gwz_model** out) {
    GWZ_HANDLE(ctx, gwz::GWZ_TAG_CONTEXT, "gwz_model_create_ext_hubbard1d: NULL or foreign handle (expected a gwz_context)");
    if (int rc = gwz_model_create_hubbard1d_gutzwiller(ctx, N, M, u, t, out)) return rc;
    (*out)->v = v;
    return GWZ_OK;
}
extern "C" int gwz_model_info_ext(const gwz_model* m, double* v) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_model_info_ext: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && v, "gwz_model_info_ext: NULL argument");
    *v = m->v;
    return GWZ_OK;
}

static int ansatz_num_parameters(int kind, int M) {
    switch (kind) {
        case GWZ_ANSATZ_GUTZWILLER: return 1;
        case GWZ_ANSATZ_EXT_GUTZWILLER: return 2;
        case GWZ_ANSATZ_MULTINOMIAL: return 1;
        case GWZ_ANSATZ_DENSITY_PROFILE: return M;           // densityprofile.jl:15
        case GWZ_ANSATZ_RELATIVE_JASTROW: return (M + 1) / 2;  // jastrow.jl:78
        case GWZ_ANSATZ_JASTROW: return M * (M + 1) / 2;       // jastrow.jl:20
        default: return -1;
    }
}

extern "C" int gwz_ansatz_create(gwz_model* m, int kind, int normalize, gwz_ansatz** out) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_ansatz_create: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && out, "gwz_ansatz_create: NULL argument");
    const int M = m->M, N = m->N;
    const int P = ansatz_num_parameters(kind, M);
    GWZ_REQUIRE(P > 0, "gwz_ansatz_create: unknown ansatz kind");
    GWZ_REQUIRE(gen_ppl(P) > 0, "gwz_ansatz_create: more than 256 parameters are not supported by the walker kernels");
    GWZ_REQUIRE(N <= 255, "gwz_ansatz_create: N must be <= 255");
    if (int rc = use_device(m->ctx)) return rc;
    gwz_ansatz* a = new gwz_ansatz();
    HandleGuard<gwz_ansatz, gwz_ansatz_destroy> guard(a);
    a->m = m;
    a->device = m->ctx->device;
    a->kind = kind;
    a->P = P;
    a->normalize = normalize;
    // multinomial.jl:29-31: K = gamma(N + 1) / M^N
    a->logK = (kind == GWZ_ANSATZ_MULTINOMIAL && normalize) ? std::lgamma((double)N + 1.0) - (double)N * std::log((double)M) : 0.0;
    a->Mpad = (M + 3) & ~3;
    const int T = N + 2;
    std::vector<double> st(3 * (size_t)T);
    for (int i = 0; i < T; ++i) {
        st[i] = i > 0 ? std::log((double)i) : 0.0;
        st[T + i] = std::lgamma((double)i + 1.0);
        st[2 * T + i] = std::sqrt((double)i);
    }
    GWZ_CUDA(cudaMalloc(&a->d_static, sizeof(double) * st.size()));
    GWZ_CUDA(cudaMemcpy(a->d_static, st.data(), sizeof(double) * st.size(), cudaMemcpyHostToDevice));
    std::vector<uint16_t> pairs(2 * (size_t)P, 0);
    if (kind == GWZ_ANSATZ_JASTROW) {
        int p = 0;
        for (int i = 0; i < M; ++i)
            for (int j = i; j < M; ++j, ++p) {  // jastrow.jl:34-40
                pairs[p] = (uint16_t)i;
                pairs[P + p] = (uint16_t)j;
            }
    }
    GWZ_CUDA(cudaMalloc(&a->d_pairs, sizeof(uint16_t) * pairs.size()));
    GWZ_CUDA(cudaMemcpy(a->d_pairs, pairs.data(), sizeof(uint16_t) * pairs.size(), cudaMemcpyHostToDevice));
    const size_t npar = (size_t)M * M + 5 * (size_t)M + 2 * GEN_FAC_PAD;  // Qt, bvec, cR, cL, facX, facY
    GWZ_CUDA(cudaMalloc(&a->d_par, sizeof(double) * npar));
    GWZ_CUDA(cudaMallocHost(&a->h_par, sizeof(double) * npar));
    GenModel& gm = a->gm;
    std::memset(&gm, 0, sizeof(gm));
    gm.N = N;
    gm.M = M;
    gm.P = P;
    gm.kind = kind;
    gm.B = m->B;
    gm.W = m->W;
    {
        const int need = (M + 31) / 32;  // modes per lane, rounded to the instantiated 1, 2, 4, 8
        gm.MPL = need <= 1 ? 1 : (need <= 2 ? 2 : (need <= 4 ? 4 : 8));
    }
    gm.MP = 32 * gm.MPL;
    gm.u = m->u;
    gm.v = m->v;
    gm.t = m->t;
    gm.logK = a->logK;
    gm.Qt = a->d_par;
    gm.bvec = a->d_par + (size_t)M * M;
    gm.cR = gm.bvec + M;
    gm.cL = gm.cR + M;
    gm.facX = gm.cL + M;
    gm.facY = gm.facX + M + GEN_FAC_PAD;
    gm.logtab = a->d_static;
    gm.lgam = a->d_static + T;
    gm.sqrtab = a->d_static + 2 * T;
    gm.pair_i = a->d_pairs;
    gm.pair_j = a->d_pairs + P;
    *out = guard.release();
    return GWZ_OK;
}
extern "C" int gwz_ansatz_create_combination(gwz_model* m, int kind1, int normalize1, int kind2, int normalize2,
                                             gwz_ansatz** out) {
    GWZ_HANDLE(m, gwz::GWZ_TAG_MODEL, "gwz_ansatz_create_combination: NULL or foreign handle (expected a gwz_model)");
    GWZ_REQUIRE(m && out, "gwz_ansatz_create_combination: NULL argument");
    gwz_ansatz *c1 = nullptr, *c2 = nullptr;
    if (int rc = gwz_ansatz_create(m, kind1, normalize1, &c1)) return rc;
    if (int rc = gwz_ansatz_create(m, kind2, normalize2, &c2)) {
        gwz_ansatz_destroy(c1);
        return rc;
    }
    const int P = c1->P + c2->P + 1;  // combination.jl:14
    if (gen_ppl(P) <= 0) {
        gwz_ansatz_destroy(c1);
        gwz_ansatz_destroy(c2);
        return api_fail(GWZ_ERR_INVALID, "gwz_ansatz_create_combination: more than 256 parameters are not supported");
    }
    gwz_ansatz* a = new gwz_ansatz();
    a->m = m;
    a->device = m->ctx->device;
    a->kind = GWZ_ANSATZ_COMBINATION;
    a->P = P;
    a->normalize = 0;
    a->logK = 0.0;
    a->Mpad = c1->Mpad;
    a->gm = c1->gm;  // geometry (N, M, B, W, MP, MPL, u, v, t); the per-component tables are taken from c1 / c2
    a->gm.P = P;
    a->c1 = c1;
    a->c2 = c2;
    *out = a;
    return GWZ_OK;
}

extern "C" int gwz_ansatz_destroy(gwz_ansatz* a) {
    if (a && a->magic != gwz::GWZ_TAG_ANSATZ) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_ansatz_destroy: not a gwz_ansatz handle");
    if (!a) return GWZ_OK;
    if (a->c1) gwz_ansatz_destroy(a->c1);
    if (a->c2) gwz_ansatz_destroy(a->c2);
    if (cudaSetDevice(a->device) == cudaSuccess) {
        if (a->d_static) cudaFree(a->d_static);
        if (a->d_pairs) cudaFree(a->d_pairs);
        if (a->d_par) cudaFree(a->d_par);
        if (a->h_par) cudaFreeHost(a->h_par);
    }
    a->magic = 0;
    delete a;
    return GWZ_OK;
}
extern "C" int gwz_ansatz_info(const gwz_ansatz* a, int* kind, int* num_parameters) {
    GWZ_HANDLE(a, gwz::GWZ_TAG_ANSATZ, "gwz_ansatz_info: NULL or foreign handle (expected a gwz_ansatz)");
    GWZ_REQUIRE(a != nullptr, "gwz_ansatz_info: NULL handle");
    if (kind) *kind = a->kind;
    if (num_parameters) *num_parameters = a->P;
    return GWZ_OK;
}

// phi(n) = 1/2 n^T Q n + b^T n + s (sum log n! - log K) for the given parameters; enqueues the upload
static int upload_params(gwz_ansatz* a, const double* p, GenModel* gm_out) {
    gwz_model* m = a->m;
    gwz_context* c = m->ctx;
    const int M = m->M;
    double* Q = a->h_par;
    double* b = Q + (size_t)M * M;
    double* cR = b + M;
    double* cL = cR + M;
    double* facX = cL + M;
    double* facY = facX + M + GEN_FAC_PAD;
    const size_t npar = (size_t)M * M + 5 * (size_t)M + 2 * GEN_FAC_PAD;
    // the previous upload from this staging buffer must have completed
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    std::memset(Q, 0, sizeof(double) * npar);
    double s = 0.0;
    auto ring = [&](double w) {  // w * (S + S^-1): sum_k n_k n_{k+1} = 1/2 n^T (S + S^-1) n
        for (int k = 0; k < M; ++k) {
            const int r = (k + 1) % M;
            Q[(size_t)k * M + r] += w;
            Q[(size_t)r * M + k] += w;
        }
    };
    switch (a->kind) {
        case GWZ_ANSATZ_GUTZWILLER:  // exp(-g * diagonal_element(H, addr)), gutzwiller.jl:25-36
            for (int k = 0; k < M; ++k) {
                Q[(size_t)k * M + k] = p[0] * m->u;
                b[k] = -0.5 * p[0] * m->u;
            }
            if (m->v != 0.0) ring(p[0] * m->v);
            break;
        case GWZ_ANSATZ_EXT_GUTZWILLER:  // exp(-g1 sum n(n-1)/2 - g2 sum n_i n_{i+1}), extgutzwiller.jl:31-35
            for (int k = 0; k < M; ++k) {
                Q[(size_t)k * M + k] = p[0];
                b[k] = -0.5 * p[0];
            }
            ring(p[1]);
            break;
        case GWZ_ANSATZ_MULTINOMIAL:  // (K prod 1/n_i!)^p, multinomial.jl:53-59
            s = p[0];
            break;
        case GWZ_ANSATZ_DENSITY_PROFILE:  // exp(-sum p_i n_i), densityprofile.jl:20-28
            for (int k = 0; k < M; ++k) b[k] = p[k];
            break;
        case GWZ_ANSATZ_RELATIVE_JASTROW:  // exp(-sum_d p_d sum_k n_k n_{k+d}), jastrow.jl:84-103
            for (int d = 0; d < a->P; ++d)
                for (int k = 0; k < M; ++k) {
                    const int r = (k + d) % M;
                    Q[(size_t)k * M + r] += p[d];
                    Q[(size_t)r * M + k] += p[d];
                }
            break;
        case GWZ_ANSATZ_JASTROW: {  // exp(-sum_{i<=j} p_ij n_i n_j), jastrow.jl:26-58
            int q = 0;
            for (int i = 0; i < M; ++i)
                for (int j = i; j < M; ++j, ++q) {
                    if (i == j) {
                        Q[(size_t)i * M + i] = 2.0 * p[q];
                    } else {
                        Q[(size_t)i * M + j] = p[q];
                        Q[(size_t)j * M + i] = p[q];
                    }
                }
            break;
        }
        default: return api_fail(GWZ_ERR_INVALID, "unknown ansatz kind");
    }
    for (int k = 0; k < M; ++k) {
        const int r = (k + 1) % M, l = (k + M - 1) % M;
        cR[k] = 0.5 * Q[(size_t)r * M + r] + 0.5 * Q[(size_t)k * M + k] - Q[(size_t)k * M + r];
        cL[k] = 0.5 * Q[(size_t)l * M + l] + 0.5 * Q[(size_t)k * M + k] - Q[(size_t)k * M + l];
    }
    for (int k = 0; k < a->P; ++k) GWZ_REQUIRE(std::isfinite(p[k]), "parameters must be finite");
    // translation-invariant couplings (Q[k][l] = q[(l - k) mod M] exactly, no multinomial term): the walker kernel
    // carries the hop rates and multiplies them by exp(+-D2[distance to the hop]), D2 = second difference of q
    bool circ = s == 0.0 && M >= 4;
    for (int k = 1; k < M && circ; ++k)
        for (int l = 0; l < M; ++l)
            if (Q[(size_t)k * M + l] != Q[(size_t)((l - k + M) % M)]) {
                circ = false;
                break;
            }
    if (circ)
        for (int i = 0; i < M + GEN_FAC_PAD; ++i) {  // entry i: distance (i - 1) mod M
            const int x = (i - 1 + M) % M;
            const double d2 = Q[(x + 1) % M] - 2.0 * Q[x] + Q[(x + M - 1) % M];
            facX[i] = std::exp(d2);
            facY[i] = std::exp(-d2);
        }
    GWZ_CUDA(cudaMemcpyAsync(a->d_par, a->h_par, sizeof(double) * npar, cudaMemcpyHostToDevice, c->stream));
    *gm_out = a->gm;
    gm_out->s = s;
    gm_out->circ = circ ? 1 : 0;
    return GWZ_OK;
}

// parameters -> device tables for a plain ansatz (gm) or a combination (cb); *is_combo tells which
static int prepare_params(gwz_ansatz* a, const double* p, GenModel* gm, GenCombo* cb, bool* is_combo) {
    *is_combo = a->c1 != nullptr;
    if (!*is_combo) return upload_params(a, p, gm);
    for (int k = 0; k < a->P; ++k) GWZ_REQUIRE(std::isfinite(p[k]), "parameters must be finite");
    if (int rc = upload_params(a->c1, p, &cb->g1)) return rc;
    if (int rc = upload_params(a->c2, p + a->c1->P, &cb->g2)) return rc;
    cb->alpha = p[a->P - 1];
    cb->P = a->P;
    *gm = a->gm;
    return GWZ_OK;
}

// ------------------------------------------------------------------------------------------
// probe / ansatz evaluation
// ------------------------------------------------------------------------------------------
static int probe_impl(gwz_ansatz* a, const uint64_t* addr_words, size_t n, const double* params, const double* u01,
                      double* val, double* grad, double* e_loc, double* tau, double* grad_ratio, int32_t* n_hops,
                      double* rates, int32_t* chosen, uint64_t* next_addr) {
    GWZ_REQUIRE(a && params, "gwz_gen_probe: NULL argument");
    GWZ_REQUIRE(n == 0 || addr_words, "gwz_gen_probe: addr_words is NULL");
    if (n == 0) return GWZ_OK;
    gwz_context* c = a->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int W = a->m->W, M = a->m->M, P = a->P;
    GenProbeLaunch L;
    std::memset(&L, 0, sizeof(L));
    GenCombo cb;
    bool combo = false;
    if (int rc = prepare_params(a, params, &L.gm, &cb, &combo)) return rc;
    DevBuf<uint64_t> d_addr, d_next;
    DevBuf<double> d_u, d_val, d_grad, d_e, d_t, d_g, d_r;
    DevBuf<int32_t> d_h, d_c;
    GWZ_CUDA(d_addr.alloc(n * W));
    GWZ_CUDA(cudaMemcpyAsync(d_addr.p, addr_words, sizeof(uint64_t) * n * W, cudaMemcpyHostToDevice, c->stream));
    if (u01) {
        GWZ_CUDA(d_u.alloc(n));
        GWZ_CUDA(cudaMemcpyAsync(d_u.p, u01, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    }
    if (val) GWZ_CUDA(d_val.alloc(n));
    if (grad) GWZ_CUDA(d_grad.alloc(n * P));
    if (e_loc) GWZ_CUDA(d_e.alloc(n));
    if (tau) GWZ_CUDA(d_t.alloc(n));
    if (grad_ratio) GWZ_CUDA(d_g.alloc(n * P));
    if (n_hops) GWZ_CUDA(d_h.alloc(n));
    if (rates) GWZ_CUDA(d_r.alloc(n * 2 * M));
    if (chosen) GWZ_CUDA(d_c.alloc(n));
    if (next_addr) GWZ_CUDA(d_next.alloc(n * W));
    L.addr = d_addr.p;
    L.n = n;
    L.u01 = d_u.p;
    L.val = d_val.p;
    L.grad = d_grad.p;
    L.e_loc = d_e.p;
    L.tau = d_t.p;
    L.grad_ratio = d_g.p;
    L.n_hops = d_h.p;
    L.rates = d_r.p;
    L.chosen = d_c.p;
    L.next_addr = d_next.p;
    L.err_flag = c->d_err;
    if (combo) GWZ_LAUNCH(launch_combo_probe(L, cb, c->stream));
    else GWZ_LAUNCH(launch_gen_probe(L, c->stream));
    auto back = [&](void* dst, const void* src, size_t bytes) {
        return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
    };
    GWZ_CUDA(back(val, d_val.p, sizeof(double) * n));
    GWZ_CUDA(back(grad, d_grad.p, sizeof(double) * n * P));
    GWZ_CUDA(back(e_loc, d_e.p, sizeof(double) * n));
    GWZ_CUDA(back(tau, d_t.p, sizeof(double) * n));
    GWZ_CUDA(back(grad_ratio, d_g.p, sizeof(double) * n * P));
    GWZ_CUDA(back(n_hops, d_h.p, sizeof(int32_t) * n));
    GWZ_CUDA(back(rates, d_r.p, sizeof(double) * n * 2 * M));
    GWZ_CUDA(back(chosen, d_c.p, sizeof(int32_t) * n));
    GWZ_CUDA(back(next_addr, d_next.p, sizeof(uint64_t) * n * W));
    GWZ_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return api_check_err_flag(c);
}

extern "C" int gwz_gen_ansatz_eval(gwz_ansatz* a, const uint64_t* addr_words, size_t n, const double* params, double* val,
                                   double* grad) {
    GWZ_HANDLE(a, gwz::GWZ_TAG_ANSATZ, "gwz_gen_ansatz_eval: NULL or foreign handle (expected a gwz_ansatz)");
    return probe_impl(a, addr_words, n, params, nullptr, val, grad, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                      nullptr);
}
extern "C" int gwz_gen_probe_addresses(gwz_ansatz* a, const uint64_t* addr_words, size_t n, const double* params,
                                       const double* u01, double* e_loc, double* residence_time, double* grad_ratio,
                                       int32_t* n_hops, double* rates, int32_t* chosen, uint64_t* next_addr) {
    GWZ_HANDLE(a, gwz::GWZ_TAG_ANSATZ, "gwz_gen_probe_addresses: NULL or foreign handle (expected a gwz_ansatz)");
    return probe_impl(a, addr_words, n, params, u01, nullptr, nullptr, e_loc, residence_time, grad_ratio, n_hops, rates,
                      chosen, next_addr);
}

// ------------------------------------------------------------------------------------------
// walkers
// ------------------------------------------------------------------------------------------
extern "C" int gwz_gwalkers_create(gwz_ansatz* a, uint64_t n_walkers, uint64_t global_offset, const uint64_t* start_addr,
                                   uint64_t seed, gwz_gwalkers** out) {
    GWZ_HANDLE(a, gwz::GWZ_TAG_ANSATZ, "gwz_gwalkers_create: NULL or foreign handle (expected a gwz_ansatz)");
    GWZ_REQUIRE(a && start_addr && out, "gwz_gwalkers_create: NULL argument");
    GWZ_REQUIRE(n_walkers >= 1, "gwz_gwalkers_create: need at least one walker");
    gwz_context* c = a->m->ctx;
    if (int rc = use_device(c)) return rc;
    // the start address must be a valid BoseFS{N,M} string
    {
        int ones = 0;
        for (int p = 0; p < a->m->B; ++p) ones += (int)((start_addr[p >> 6] >> (p & 63)) & 1ull);
        GWZ_REQUIRE(ones == a->m->N, "gwz_gwalkers_create: start address does not hold N particles");
    }
    gwz_gwalkers* w = new gwz_gwalkers();
    HandleGuard<gwz_gwalkers, gwz_gwalkers_destroy> guard(w);
    w->a = a;
    w->device = a->device;
    w->n = (size_t)n_walkers;
    w->offset = global_offset;
    w->seed = seed;
    w->steps_done = 0;
    w->acc_steps = 0;
    const int P = a->P, L = gen_vec_len(P);
    GWZ_CUDA(cudaMalloc(&w->d_occ, (size_t)a->Mpad * w->n));
    GWZ_CUDA(cudaMalloc(&w->d_acc, sizeof(double) * (size_t)(3 + 2 * P) * w->n));
    GWZ_CUDA(cudaMemsetAsync(w->d_acc, 0, sizeof(double) * (size_t)(3 + 2 * P) * w->n, c->stream));
    GWZ_CUDA(cudaMalloc(&w->d_vec, sizeof(double) * L));
    GWZ_CUDA(cudaMallocHost(&w->h_vec, sizeof(double) * L));
    GWZ_CUDA(cudaMalloc(&w->d_shift, sizeof(double) * (1 + P)));
    GWZ_CUDA(cudaMallocHost(&w->h_shift, sizeof(double) * (1 + P)));
    GWZ_CUDA(cudaMalloc(&w->d_addr1, sizeof(uint64_t) * GWZ_MAX_WORDS));
    GWZ_CUDA(cudaMemcpyAsync(w->d_addr1, start_addr, sizeof(uint64_t) * a->m->W, cudaMemcpyHostToDevice, c->stream));
    GWZ_LAUNCH(launch_gen_unary_to_occ(w->d_addr1, w->n, a->gm, a->Mpad, w->d_occ, true, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    *out = guard.release();
    return GWZ_OK;
}
extern "C" int gwz_gwalkers_destroy(gwz_gwalkers* w) {
    if (w && w->magic != gwz::GWZ_TAG_GWALKERS) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_gwalkers_destroy: not a gwz_gwalkers handle");
    if (!w) return GWZ_OK;
    if (cudaSetDevice(w->device) == cudaSuccess) {
        cudaFree(w->d_occ);
        cudaFree(w->d_acc);
        cudaFree(w->d_partials);
        cudaFree(w->d_vec);
        if (w->h_vec) cudaFreeHost(w->h_vec);
        cudaFree(w->d_shift);
        if (w->h_shift) cudaFreeHost(w->h_shift);
        cudaFree(w->d_addr1);
    }
    w->magic = 0;
    delete w;
    return GWZ_OK;
}
extern "C" int gwz_gwalkers_count(const gwz_gwalkers* w, uint64_t* n_walkers, uint64_t* steps_done) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_GWALKERS, "gwz_gwalkers_count: NULL or foreign handle (expected a gwz_gwalkers)");
    GWZ_REQUIRE(w != nullptr, "gwz_gwalkers_count: NULL handle");
    if (n_walkers) *n_walkers = w->n;
    if (steps_done) *steps_done = w->steps_done;
    return GWZ_OK;
}
extern "C" int gwz_gwalkers_get_addresses(gwz_gwalkers* w, uint64_t* addr_words) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_GWALKERS, "gwz_gwalkers_get_addresses: NULL or foreign handle (expected a gwz_gwalkers)");
    GWZ_REQUIRE(w && addr_words, "gwz_gwalkers_get_addresses: NULL argument");
    gwz_context* c = w->a->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int W = w->a->m->W;
    DevBuf<uint64_t> tmp;
    GWZ_CUDA(tmp.alloc(w->n * W));
    GWZ_LAUNCH(launch_gen_occ_to_unary(w->d_occ, w->n, w->a->gm, w->a->Mpad, tmp.p, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(addr_words, tmp.p, sizeof(uint64_t) * w->n * W, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}
extern "C" int gwz_gwalkers_set_addresses(gwz_gwalkers* w, const uint64_t* addr_words) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_GWALKERS, "gwz_gwalkers_set_addresses: NULL or foreign handle (expected a gwz_gwalkers)");
    GWZ_REQUIRE(w && addr_words, "gwz_gwalkers_set_addresses: NULL argument");
    gwz_context* c = w->a->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int W = w->a->m->W;
    DevBuf<uint64_t> tmp;
    GWZ_CUDA(tmp.alloc(w->n * W));
    GWZ_CUDA(cudaMemcpyAsync(tmp.p, addr_words, sizeof(uint64_t) * w->n * W, cudaMemcpyHostToDevice, c->stream));
    GWZ_LAUNCH(launch_gen_unary_to_occ(tmp.p, w->n, w->a->gm, w->a->Mpad, w->d_occ, false, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}
extern "C" int gwz_gwalkers_set_step_counter(gwz_gwalkers* w, uint64_t steps_done) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_GWALKERS, "gwz_gwalkers_set_step_counter: NULL or foreign handle (expected a gwz_gwalkers)");
    GWZ_REQUIRE(w != nullptr, "gwz_gwalkers_set_step_counter: NULL handle");
    w->steps_done = steps_done;
    return GWZ_OK;
}

static int gen_run(gwz_gwalkers* w, const double* params, uint64_t steps, uint64_t burn_in, unsigned flags, double* rec_tau,
                   double* rec_eloc, uint64_t* rec_addr, gwz_gen_result* out, double* grad, double* grad_err,
                   double* pooled_grad) {
    GWZ_REQUIRE(w && params, "gwz_gen_vqmc_run: NULL argument");
    GWZ_REQUIRE(steps >= 1, "gwz_gen_vqmc_run: steps_per_walker must be >= 1");
    gwz_ansatz* a = w->a;
    gwz_context* c = a->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int P = a->P, L = gen_vec_len(P), W = a->m->W;
    GenWalkerLaunch K;
    std::memset(&K, 0, sizeof(K));
    GenCombo cb;
    bool combo = false;
    if (int rc = prepare_params(a, params, &K.gm, &cb, &combo)) return rc;
    int grid = 0, per_cta = 1;
    if (combo) GWZ_CUDA(combo_walker_grid(cb, w->n, c->prop.multiProcessorCount, &grid));
    else GWZ_CUDA(gen_walker_grid(K.gm, w->n, c->prop.multiProcessorCount, &grid, &per_cta));
    const size_t rows = (size_t)grid * per_cta;
    if (rows > w->partials_rows) {
        if (w->d_partials) cudaFree(w->d_partials);
        w->d_partials = nullptr;
        GWZ_CUDA(cudaMalloc(&w->d_partials, sizeof(double) * rows * L));
        w->partials_rows = rows;
    }
    K.occ = w->d_occ;
    K.Mpad = a->Mpad;
    K.acc = w->d_acc;
    K.n = w->n;
    K.global_offset = w->offset;
    K.seed = w->seed;
    K.partials = w->d_partials;
    K.grid = grid;
    K.err_flag = c->d_err;
    K.shift = w->d_shift;
    GWZ_CUDA(cudaEventRecord(c->ev0, c->stream));
    if (burn_in > 0) {
        K.step0 = w->steps_done;
        K.steps = burn_in;
        K.accumulate = 0;
        K.reset_acc = 0;
        if (combo) GWZ_LAUNCH(launch_combo_walkers(K, cb, c->stream));
        else GWZ_LAUNCH(launch_gen_walkers(K, c->stream));
        w->steps_done += burn_in;
    }
    const bool keep = (flags & GWZ_RUN_KEEP_ACC) != 0 && w->acc_steps > 0;
    if (!keep && !w->have_shift) {
        // reference point of the shifted sums: E_loc and the gradient ratios of walker 0's current address
        DevBuf<double> d_e, d_g;
        GWZ_CUDA(d_e.alloc(1));
        GWZ_CUDA(d_g.alloc(P));
        GWZ_LAUNCH(launch_gen_occ_to_unary(w->d_occ, 1, K.gm, a->Mpad, w->d_addr1, c->stream));
        GenProbeLaunch Pr;
        std::memset(&Pr, 0, sizeof(Pr));
        Pr.gm = K.gm;
        Pr.addr = w->d_addr1;
        Pr.n = 1;
        Pr.e_loc = d_e.p;
        Pr.grad_ratio = d_g.p;
        Pr.err_flag = c->d_err;
        if (combo) GWZ_LAUNCH(launch_combo_probe(Pr, cb, c->stream));
        else GWZ_LAUNCH(launch_gen_probe(Pr, c->stream));
        GWZ_CUDA(cudaMemcpyAsync(w->d_shift, d_e.p, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        GWZ_CUDA(cudaMemcpyAsync(w->d_shift + 1, d_g.p, sizeof(double) * P, cudaMemcpyDeviceToDevice, c->stream));
        GWZ_CUDA(cudaStreamSynchronize(c->stream));
        w->have_shift = true;
        GWZ_CUDA(cudaEventRecord(c->ev0, c->stream));
    }
    K.step0 = w->steps_done;
    K.steps = steps;
    K.accumulate = 1;
    K.reset_acc = keep ? 0 : 1;
    DevBuf<double> d_tau, d_el;
    DevBuf<uint64_t> d_ra;
    if (rec_tau) {
        GWZ_REQUIRE(rec_eloc != nullptr, "gwz_gen_vqmc_run_record: tau and eloc must both be given");
        GWZ_CUDA(d_tau.alloc(steps * w->n));
        GWZ_CUDA(d_el.alloc(steps * w->n));
        K.rec_tau = d_tau.p;
        K.rec_eloc = d_el.p;
        if (rec_addr) {
            GWZ_CUDA(d_ra.alloc(steps * w->n * W));
            K.rec_addr = d_ra.p;
        }
    }
    if (combo) GWZ_LAUNCH(launch_combo_walkers(K, cb, c->stream));
    else GWZ_LAUNCH(launch_gen_walkers(K, c->stream));
    GWZ_CUDA(cudaEventRecord(c->ev1, c->stream));
    w->steps_done += steps;
    w->acc_steps = keep ? w->acc_steps + steps : steps;
    GWZ_LAUNCH(launch_reduce_partials(w->d_partials, (int)rows, L, w->d_vec, c->stream));
    if (c->comm) {
        const NcclApi* api = nccl_api(nullptr);
        if (!api) return api_fail(GWZ_ERR_NCCL, "NCCL unavailable");
        // one ncclAllReduce(sum, float64, 8 + 4P) per run, on the sampling stream
        GWZ_NCCL(api, api->AllReduce(w->d_vec, w->d_vec, L, kNcclFloat64, kNcclSum, c->comm, c->stream));
    }
    GWZ_CUDA(cudaMemcpyAsync(w->h_vec, w->d_vec, sizeof(double) * L, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (rec_tau) {
        GWZ_CUDA(cudaMemcpyAsync(rec_tau, d_tau.p, sizeof(double) * steps * w->n, cudaMemcpyDeviceToHost, c->stream));
        GWZ_CUDA(cudaMemcpyAsync(rec_eloc, d_el.p, sizeof(double) * steps * w->n, cudaMemcpyDeviceToHost, c->stream));
        if (rec_addr)
            GWZ_CUDA(cudaMemcpyAsync(rec_addr, d_ra.p, sizeof(uint64_t) * steps * w->n * W, cudaMemcpyDeviceToHost, c->stream));
    }
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

extern "C" int gwz_gen_vqmc_run(gwz_gwalkers* w, const double* params, uint64_t steps_per_walker, uint64_t burn_in,
                                unsigned flags, gwz_gen_result* out, double* grad, double* grad_err, double* pooled_grad) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_GWALKERS, "gwz_gen_vqmc_run: NULL or foreign handle (expected a gwz_gwalkers)");
    return gen_run(w, params, steps_per_walker, burn_in, flags, nullptr, nullptr, nullptr, out, grad, grad_err, pooled_grad);
}
extern "C" int gwz_gen_vqmc_run_record(gwz_gwalkers* w, const double* params, uint64_t steps_per_walker, unsigned flags,
                                       double* tau, double* eloc, uint64_t* addr, gwz_gen_result* out, double* grad,
                                       double* grad_err, double* pooled_grad) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_GWALKERS, "gwz_gen_vqmc_run_record: NULL or foreign handle (expected a gwz_gwalkers)");
    GWZ_REQUIRE(tau && eloc, "gwz_gen_vqmc_run_record: NULL series buffer");
    return gen_run(w, params, steps_per_walker, 0, flags, tau, eloc, addr, out, grad, grad_err, pooled_grad);
}
extern "C" int gwz_gwalkers_estimates(gwz_gwalkers* w, double* val_w, double* grad_w) {
    GWZ_HANDLE(w, gwz::GWZ_TAG_GWALKERS, "gwz_gwalkers_estimates: NULL or foreign handle (expected a gwz_gwalkers)");
    GWZ_REQUIRE(w != nullptr, "gwz_gwalkers_estimates: NULL handle");
    GWZ_REQUIRE(w->acc_steps > 0, "gwz_gwalkers_estimates: no steps accumulated yet");
    gwz_context* c = w->a->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int P = w->a->P;
    DevBuf<double> dv, dg;
    GWZ_CUDA(dv.alloc(w->n));
    GWZ_CUDA(dg.alloc(w->n * P));
    GWZ_LAUNCH(launch_gen_estimates(w->d_acc, w->n, P, w->d_shift, dv.p, dg.p, c->stream));
    if (val_w) GWZ_CUDA(cudaMemcpyAsync(val_w, dv.p, sizeof(double) * w->n, cudaMemcpyDeviceToHost, c->stream));
    if (grad_w) GWZ_CUDA(cudaMemcpyAsync(grad_w, dg.p, sizeof(double) * w->n * P, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}

// ------------------------------------------------------------------------------------------
// LocalEnergyEvaluator
// ------------------------------------------------------------------------------------------
extern "C" int gwz_gsweep_create(gwz_ansatz* a, const uint64_t* basis_words, uint64_t dim, uint64_t rank_begin,
                                 uint64_t rank_end, gwz_gsweep** out) {
    GWZ_HANDLE(a, gwz::GWZ_TAG_ANSATZ, "gwz_gsweep_create: NULL or foreign handle (expected a gwz_ansatz)");
    GWZ_REQUIRE(a && out, "gwz_gsweep_create: NULL argument");
    gwz_model* m = a->m;
    gwz_context* c = m->ctx;
    if (int rc = use_device(c)) return rc;
    gwz_gsweep* s = new gwz_gsweep();
    HandleGuard<gwz_gsweep, gwz_gsweep_destroy> guard(s);
    s->a = a;
    s->device = a->device;
    if (basis_words) {
        GWZ_REQUIRE(dim >= 1, "gwz_gsweep_create: empty basis");
        const int W = m->W;
        DevBuf<uint64_t> tmp;
        GWZ_CUDA(tmp.alloc(dim * W));
        GWZ_CUDA(cudaMalloc(&s->d_basis, sizeof(uint64_t) * dim * W));
        GWZ_CUDA(cudaMemcpyAsync(tmp.p, basis_words, sizeof(uint64_t) * dim * W, cudaMemcpyHostToDevice, c->stream));
        GWZ_LAUNCH(launch_transpose_addr(tmp.p, s->d_basis, dim, W, true, c->stream));
        GWZ_CUDA(cudaStreamSynchronize(c->stream));
        s->dim = dim;
    } else {
        GWZ_REQUIRE(m->B <= 64, "gwz_gsweep_create: ranked enumeration needs N+M-1 <= 64");
        std::vector<uint64_t> tab;
        binomial_table(tab);
        const uint64_t total = tab[(size_t)m->B * 65 + m->N];
        GWZ_REQUIRE(total != UINT64_MAX, "gwz_gsweep_create: dimension overflows 64 bits");
        if (rank_end == 0) rank_end = total;
        GWZ_REQUIRE(rank_begin < rank_end && rank_end <= total, "gwz_gsweep_create: bad rank range");
        GWZ_CUDA(cudaMalloc(&s->d_binom, sizeof(uint64_t) * tab.size()));
        GWZ_CUDA(cudaMemcpy(s->d_binom, tab.data(), sizeof(uint64_t) * tab.size(), cudaMemcpyHostToDevice));
        s->dim = rank_end - rank_begin;
        s->rank_begin = rank_begin;
        s->h_binom = std::move(tab);
    }
    const int L = 2 + 2 * a->P;
    GWZ_CUDA(cudaMalloc(&s->d_vec, sizeof(double) * L));
    GWZ_CUDA(cudaMallocHost(&s->h_vec, sizeof(double) * L));
    *out = guard.release();
    return GWZ_OK;
}
extern "C" int gwz_gsweep_destroy(gwz_gsweep* s) {
    if (s && s->magic != gwz::GWZ_TAG_GSWEEP) return gwz::api_fail(GWZ_ERR_INVALID, "gwz_gsweep_destroy: not a gwz_gsweep handle");
    if (!s) return GWZ_OK;
    if (cudaSetDevice(s->device) == cudaSuccess) {
        if (s->d_basis) cudaFree(s->d_basis);
        if (s->d_binom) cudaFree(s->d_binom);
        if (s->d_partials) cudaFree(s->d_partials);
        cudaFree(s->d_rep);
        cudaFree(s->d_mult);
        cudaFree(s->d_vec);
        if (s->h_vec) cudaFreeHost(s->h_vec);
    }
    s->magic = 0;
    delete s;
    return GWZ_OK;
}
extern "C" int gwz_gsweep_dim(const gwz_gsweep* s, uint64_t* dim) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_GSWEEP, "gwz_gsweep_dim: NULL or foreign handle (expected a gwz_gsweep)");
    GWZ_REQUIRE(s && dim, "gwz_gsweep_dim: NULL argument");
    *dim = s->dim;
    return GWZ_OK;
}

extern "C" int gwz_gsweep_val_and_grad(gwz_gsweep* s, const double* params, double* val, double* grad) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_GSWEEP, "gwz_gsweep_val_and_grad: NULL or foreign handle (expected a gwz_gsweep)");
    GWZ_REQUIRE(s && params && val, "gwz_gsweep_val_and_grad: NULL argument");
    gwz_ansatz* a = s->a;
    gwz_context* c = a->m->ctx;
    if (int rc = use_device(c)) return rc;
    const int P = a->P, L = 2 + 2 * P;
    GenSweepLaunch K;
    std::memset(&K, 0, sizeof(K));
    GenCombo cb;
    bool combo = false;
    if (int rc = prepare_params(a, params, &K.gm, &cb, &combo)) return rc;
    // Gutzwiller, ExtendedGutzwiller, Multinomial and RelativeJastrow amplitudes and features only depend on the
    // occupations up to translations and reflection of the periodic chain: one representative per orbit suffices
    gwz_model* m = a->m;
    // (a handle that sweeps part of the rank range on its own keeps counting single states: its partial sums must
    // mean the same at every call; with a communicator only the all-reduced total is visible)
    const bool whole = c->comm != nullptr ||
                       (s->rank_begin == 0 && !s->h_binom.empty() && s->dim == s->h_binom[(size_t)m->B * 65 + m->N]);
    const bool invariant = whole && !combo && s->d_basis == nullptr && m->N + m->M <= 64 &&
                           (K.gm.kind == GEN_GUTZWILLER || K.gm.kind == GEN_EXT_GUTZWILLER ||
                            K.gm.kind == GEN_MULTINOMIAL || K.gm.kind == GEN_RELATIVE_JASTROW);
    // Unlike the Gutzwiller sweep (which counts orbits in every mode), this sweep counts single states until the
    // representatives are stored: ranks sharing a communicator must switch together -- at a fixed call, and only if
    // every one of them could store its share.
    if (c->comm) s->build_at = 3;
    if (invariant && s->rep_state == 0 && ++s->calls >= s->build_at && s->calls >= 2) {
        s->rep_state = -1;
        const RepRange rr = representative_range(s->h_binom, m->N, m->M, m->B, s->rank_begin, s->dim);
        int stored = 0;
        if (int rc = build_representative_basis(c, m->N, m->M, m->B, s->d_binom, rr, &s->d_rep, &s->d_mult, &s->nrep, &stored))
            return rc;
        double failed = stored ? 0.0 : 1.0;  // summed over the ranks
        if (c->comm) {
            const NcclApi* api = nccl_api(nullptr);
            if (!api) return api_fail(GWZ_ERR_NCCL, "NCCL unavailable");
            GWZ_CUDA(cudaMemcpyAsync(s->d_vec, &failed, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            GWZ_NCCL(api, api->AllReduce(s->d_vec, s->d_vec, 1, kNcclFloat64, kNcclSum, c->comm, c->stream));
            GWZ_CUDA(cudaMemcpyAsync(&failed, s->d_vec, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            GWZ_CUDA(cudaStreamSynchronize(c->stream));
        }
        if (failed < 0.5) {
            s->rep_state = 1;  // possibly with nrep == 0: this rank's range holds no representative
        } else {
            cudaFree(s->d_rep);
            cudaFree(s->d_mult);
            s->d_rep = nullptr;
            s->d_mult = nullptr;
            s->nrep = 0;
        }
    }
    const bool reps = invariant && s->rep_state == 1;
    const uint64_t ndim = reps ? s->nrep : s->dim;
    if (reps && ndim == 0) {  // this rank's range holds no representative: it contributes zeros to the all-reduce
        GWZ_CUDA(cudaMemsetAsync(s->d_vec, 0, sizeof(double) * L, c->stream));
        if (c->comm) {
            const NcclApi* api = nccl_api(nullptr);
            if (!api) return api_fail(GWZ_ERR_NCCL, "NCCL unavailable");
            GWZ_NCCL(api, api->AllReduce(s->d_vec, s->d_vec, L, kNcclFloat64, kNcclSum, c->comm, c->stream));
        }
        GWZ_CUDA(cudaMemcpyAsync(s->h_vec, s->d_vec, sizeof(double) * L, cudaMemcpyDeviceToHost, c->stream));
        GWZ_CUDA(cudaStreamSynchronize(c->stream));
        s->last_ms = 0.f;
        const double num0 = s->h_vec[0], den0 = s->h_vec[1];
        *val = num0 / den0;
        if (grad)
            for (int p = 0; p < P; ++p) grad[p] = (s->h_vec[2 + p] * den0 - num0 * s->h_vec[2 + P + p]) / (den0 * den0);
        return GWZ_OK;
    }
    int grid = 0, per_cta = 1;
    if (combo) GWZ_CUDA(combo_sweep_grid(cb, (size_t)ndim, c->prop.multiProcessorCount, &grid));
    else GWZ_CUDA(gen_sweep_grid(K.gm, (size_t)ndim, c->prop.multiProcessorCount, &grid, &per_cta));
    const size_t rows = (size_t)grid * per_cta;
    if (rows > s->partials_rows) {
        if (s->d_partials) cudaFree(s->d_partials);
        s->d_partials = nullptr;
        GWZ_CUDA(cudaMalloc(&s->d_partials, sizeof(double) * rows * L));
        s->partials_rows = rows;
    }
    K.basis = reps ? s->d_rep : s->d_basis;
    K.dim = ndim;
    K.rank_begin = reps ? 0 : s->rank_begin;
    K.binom = s->d_binom;
    K.weight = reps ? s->d_mult : nullptr;
    K.chunk = (ndim + rows - 1) / rows;
    K.with_grad = grad ? 1 : 0;
    K.partials = s->d_partials;
    K.grid = grid;
    if (!grad) GWZ_CUDA(cudaMemsetAsync(s->d_partials, 0, sizeof(double) * rows * L, c->stream));
    GWZ_CUDA(cudaEventRecord(c->ev0, c->stream));
    if (combo) GWZ_LAUNCH(launch_combo_sweep(K, cb, c->stream));
    else GWZ_LAUNCH(launch_gen_sweep(K, c->stream));
    GWZ_CUDA(cudaEventRecord(c->ev1, c->stream));
    GWZ_LAUNCH(launch_reduce_partials(s->d_partials, (int)rows, L, s->d_vec, c->stream));
    if (c->comm) {
        const NcclApi* api = nccl_api(nullptr);
        if (!api) return api_fail(GWZ_ERR_NCCL, "NCCL unavailable");
        GWZ_NCCL(api, api->AllReduce(s->d_vec, s->d_vec, L, kNcclFloat64, kNcclSum, c->comm, c->stream));
    }
    GWZ_CUDA(cudaMemcpyAsync(s->h_vec, s->d_vec, sizeof(double) * L, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    GWZ_CUDA(cudaEventElapsedTime(&s->last_ms, c->ev0, c->ev1));
    if (invariant && s->rep_state == 0 && s->calls <= 2 && !c->comm) {  // rent or buy, as in gwz_sweep (gwz_api.cu)
        s->t_enum = s->calls == 1 ? s->last_ms : std::min(s->t_enum, s->last_ms);
        const double T = std::max((double)s->t_enum, 1e-3);
        s->build_at = s->calls == 1 ? 3 : std::max<uint64_t>(3, (uint64_t)std::ceil((1.0 + 0.3 * T) / (0.7 * T)));
        if (const char* e = std::getenv("GWZ_SWEEP_CACHE_AT")) s->build_at = (uint64_t)std::max(2, std::atoi(e));
    }
    const double num = s->h_vec[0], den = s->h_vec[1];
    *val = num / den;  // localenergy.jl:124-128
    if (grad)
        for (int p = 0; p < P; ++p) grad[p] = (s->h_vec[2 + p] * den - num * s->h_vec[2 + P + p]) / (den * den);
    return GWZ_OK;
}

static int gsweep_terms_impl(gwz_gsweep* s, const double* params, uint64_t first, uint64_t count, double* out,
                             uint64_t* states) {
    gwz_ansatz* a = s->a;
    gwz_context* c = a->m->ctx;
    if (int rc = use_device(c)) return rc;
    GWZ_REQUIRE(first + count <= s->dim, "gwz_gsweep_terms: range exceeds the basis");
    if (count == 0) return GWZ_OK;
    const int P = a->P, L = 2 + 2 * P, W = a->m->W;
    GenSweepLaunch K;
    std::memset(&K, 0, sizeof(K));
    GenCombo cb;
    bool combo = false;
    if (params) {
        if (int rc = prepare_params(a, params, &K.gm, &cb, &combo)) return rc;
    } else {
        std::vector<double> zero((size_t)P, 0.0);
        if (int rc = prepare_params(a, zero.data(), &K.gm, &cb, &combo)) return rc;
    }
    DevBuf<double> d_out;
    DevBuf<uint64_t> d_st;
    if (out) GWZ_CUDA(d_out.alloc(count * L));
    if (states) GWZ_CUDA(d_st.alloc(count * W));
    int grid = 0, per_cta = 1;
    if (combo) GWZ_CUDA(combo_sweep_grid(cb, (size_t)count, c->prop.multiProcessorCount, &grid));
    else GWZ_CUDA(gen_sweep_grid(K.gm, (size_t)count, c->prop.multiProcessorCount, &grid, &per_cta));
    const size_t rows = (size_t)grid * per_cta;
    K.basis = s->d_basis;
    K.dim = s->dim;
    K.rank_begin = s->rank_begin;
    K.binom = s->d_binom;
    K.chunk = (count + rows - 1) / rows;
    K.with_grad = out ? 1 : 0;
    K.terms_out = d_out.p;
    K.states_out = d_st.p;
    K.first = first;
    K.count = count;
    K.grid = grid;
    if (combo) GWZ_LAUNCH(launch_combo_sweep(K, cb, c->stream));
    else GWZ_LAUNCH(launch_gen_sweep(K, c->stream));
    if (out) GWZ_CUDA(cudaMemcpyAsync(out, d_out.p, sizeof(double) * count * L, cudaMemcpyDeviceToHost, c->stream));
    if (states) GWZ_CUDA(cudaMemcpyAsync(states, d_st.p, sizeof(uint64_t) * count * W, cudaMemcpyDeviceToHost, c->stream));
    GWZ_CUDA(cudaStreamSynchronize(c->stream));
    return GWZ_OK;
}
extern "C" int gwz_gsweep_terms(gwz_gsweep* s, const double* params, uint64_t first, uint64_t count, double* out) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_GSWEEP, "gwz_gsweep_terms: NULL or foreign handle (expected a gwz_gsweep)");
    GWZ_REQUIRE(s && params && out, "gwz_gsweep_terms: NULL argument");
    return gsweep_terms_impl(s, params, first, count, out, nullptr);
}
extern "C" int gwz_gsweep_states(gwz_gsweep* s, uint64_t first, uint64_t count, uint64_t* out) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_GSWEEP, "gwz_gsweep_states: NULL or foreign handle (expected a gwz_gsweep)");
    GWZ_REQUIRE(s && out, "gwz_gsweep_states: NULL argument");
    return gsweep_terms_impl(s, nullptr, first, count, nullptr, out);
}
extern "C" int gwz_gsweep_last_kernel_ms(const gwz_gsweep* s, float* ms) {
    GWZ_HANDLE(s, gwz::GWZ_TAG_GSWEEP, "gwz_gsweep_last_kernel_ms: NULL or foreign handle (expected a gwz_gsweep)");
    GWZ_REQUIRE(s && ms, "gwz_gsweep_last_kernel_ms: NULL argument");
    *ms = s->last_ms;
    return GWZ_OK;
}
