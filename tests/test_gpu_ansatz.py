"""GPU parity of the ansatz-generic engine (ExtendedGutzwiller, Multinomial, DensityProfile, RelativeJastrow,
Jastrow, Gutzwiller on ExtendedHubbardReal1D) against part 2 of the CPU oracle, through the C ABI:
per-address amplitudes / local energies / rates / moves at 1e-12 relative, step-for-step trajectories,
LocalEnergyEvaluator values and gradients at 1e-12, statistical containment, AMSGrad (test/amsgrad.jl)."""
import numpy as np
import pytest

from conftest import edge_onrs, random_onr

pytestmark = pytest.mark.gpu
RTOL = 1e-12

KINDS = {"gutzwiller": "GutzwillerAnsatz", "ext_gutzwiller": "ExtendedGutzwillerAnsatz",
         "multinomial": "MultinomialAnsatz", "density_profile": "DensityProfileAnsatz",
         "relative_jastrow": "RelativeJastrowAnsatz", "jastrow": "JastrowAnsatz",
         # CombinationAnsatz: GutzwillerAnsatz(H) + MultinomialAnsatz(H) is the reference's own test (test/ansatz.jl:58)
         ("gutzwiller", "multinomial"): None, ("relative_jastrow", "density_profile"): None}
SHAPES = [  # N, M, u, v, t
    (5, 5, 1.0, 0.0, 0.1),     # test/ansatz.jl:45
    (5, 5, 1.0, 1.0, 0.1),     # test/ansatz.jl:46 (ExtendedHubbardReal1D)
    (4, 6, 2.0, 0.5, 1.0),
    (7, 3, 0.7, 0.3, 1.0),
    (3, 2, 1.0, 0.4, 1.0),     # M = 2: both neighbours are the same mode
    (12, 12, 2.0, 0.0, 1.0),
    (40, 37, 2.0, 0.25, 1.0),  # two modes per lane, two 64-bit words
    (120, 131, 2.0, 0.1, 0.5),  # eight modes per lane, four words, up to 131 parameters
]


def _make(gw, oracle_mod, kind, N, M, u, v, t, normalize=False):
    o = oracle_mod.Oracle(N, M, u, t)
    ao = oracle_mod.AnsatzOracle(o, kind, v=v, normalize=normalize)
    addr = gw.near_uniform(N, M)
    H = gw.ExtendedHubbardReal1D(addr, u=u, v=v, t=t) if v else gw.HubbardReal1D(addr, u=u, t=t)
    if isinstance(kind, tuple):
        parts = []
        for k in kind:
            c = gw.GenericGutzwillerAnsatz if (k == "gutzwiller" and v) else getattr(gw, KINDS[k])
            parts.append(c(H))
        return o, ao, H, parts[0] + parts[1]
    cls = getattr(gw, KINDS[kind])
    if kind == "gutzwiller":
        cls = gw.GenericGutzwillerAnsatz  # the generic engine also for plain HubbardReal1D (cross-check of the fast path)
    ans = cls(H, normalize=True) if (kind == "multinomial" and normalize) else cls(H)
    return o, ao, H, ans


def _params(P, rng, scale=0.3):
    return rng.random(P) * scale


def _addresses(o, N, M, seed, n_random=30):
    rng = np.random.default_rng(seed)
    onrs = edge_onrs(N, M)
    for conc in (0.3, 1.0, 5.0, 50.0):
        onrs += [random_onr(rng, N, M, conc) for _ in range(n_random // 4)]
    return np.stack([o.from_onr(x) for x in onrs])


def _close(a, b, rtol=RTOL, atol=1e-13):
    return np.all(np.abs(np.asarray(a) - np.asarray(b)) <= atol + rtol * np.abs(np.asarray(b)))


@pytest.mark.parametrize("kind", list(KINDS))
@pytest.mark.parametrize("N,M,u,v,t", SHAPES)
def test_eval_and_probe_match_oracle(gw, oracle_mod, kind, N, M, u, v, t):
    if kind == "jastrow" and M * (M + 1) // 2 > 256:
        pytest.skip("more than 256 parameters")
    o, ao, H, ans = _make(gw, oracle_mod, kind, N, M, u, v, t)
    assert ans.N_PARAMS == ao.P == gw.num_parameters(ans)
    rng = np.random.default_rng(N * 17 + M)
    p = _params(ao.P, rng, 0.05 if N > 20 else 0.3)
    addrs = _addresses(o, N, M, seed=M * 31 + N)
    h = ans.handle()
    val, grad = h.eval(addrs[:, :o.W], p)
    u01 = rng.random(len(addrs))
    out = h.probe(addrs[:, :o.W], p, u01, want_rates=True)
    for i, a in enumerate(addrs):
        v1, g1 = ao.val_and_grad(a, p)
        if not (1e-280 < abs(v1) < 1e280):
            continue  # the reference's absolute amplitude under/overflows here
        assert _close(val[i], v1) and _close(grad[i], g1, atol=1e-13 * abs(v1) * max(1.0, np.abs(g1 / v1).max()))
        ks = ao.kinetic_sample(a, p, float(u01[i]))
        assert out["n_hops"][i] == ks["n_hops"]
        # E_loc = D - t * sum(...) cancels for clumped states: the tolerance is relative to the terms, and the
        # reference's own psi2/psi1 through absolute amplitudes carries ~|log psi| * eps
        scale = max(1.0, abs(ao.diagonal_element(a)), abs(np.log(abs(v1))))
        assert _close(out["e_loc"][i], ks["eloc"], atol=1e-12 * scale)
        assert _close(out["residence_time"][i], ks["tau"], rtol=1e-12 * scale)
        assert _close(out["grad_ratio"][i], ks["grad"], atol=1e-12)
        rates = ao.hop_rates(a, p)
        assert _close(out["rates"][i][: len(rates)], rates, rtol=1e-12 * scale) and np.all(out["rates"][i][len(rates):] == 0)
        # the move: same hop unless u01*S falls within rounding distance of a boundary of the cumulative sums
        cs = ks["cumsum"]
        T = u01[i] * cs[-1]
        if np.min(np.abs(cs - T)) > 1e-9 * cs[-1]:
            assert out["chosen"][i] == ks["chosen"]
            assert np.array_equal(out["next_addr"][i], ks["new_addr"][: o.W])


@pytest.mark.parametrize("kind,N,M,u,v,t", [("relative_jastrow", 5, 5, 1.0, 0.0, 0.1), ("density_profile", 4, 6, 2.0, 0.5, 1.0),
                                            ("jastrow", 5, 5, 1.0, 1.0, 0.1), ("multinomial", 7, 3, 0.7, 0.3, 1.0),
                                            ("ext_gutzwiller", 12, 12, 2.0, 0.3, 1.0), ("gutzwiller", 3, 2, 1.0, 0.4, 1.0),
                                            ("relative_jastrow", 40, 37, 2.0, 0.25, 1.0),
                                            (("gutzwiller", "multinomial"), 5, 5, 1.0, 0.0, 0.1),
                                            (("relative_jastrow", "density_profile"), 7, 6, 2.0, 0.5, 1.0)])
def test_walkers_follow_the_oracle_step_for_step(gw, oracle_mod, kind, N, M, u, v, t):
    """Same Philox stream, same hop order, same selection rule => identical trajectories and estimates."""
    o, ao, H, ans = _make(gw, oracle_mod, kind, N, M, u, v, t)
    rng = np.random.default_rng(11)
    p = _params(ao.P, rng, 0.05 if N > 20 else 0.3)
    nw, steps, seed, off = 9, 300, 12345, 5
    w = gw.GenWalkers(ans.handle(), nw, H.address.words, seed=seed, global_offset=off)
    r, tau, eloc, addr = w.run_record(p, steps)
    ref = ao.kinetic_vqmc(p, np.tile(np.pad(H.address.words, (0, 4 - o.W)), (nw, 1)), steps, seed=seed, walker0=off,
                          series=True)
    assert np.array_equal(addr.transpose(1, 0, 2), ref["series_addr"][:, :, : o.W])
    assert _close(tau.T, ref["series_tau"]) and _close(eloc.T, ref["series_eloc"], atol=1e-11)
    assert np.array_equal(w.addresses(), ref["addrs"][:, : o.W])
    vw, gwv = w.estimates()
    assert _close(vw, ref["val_w"], rtol=1e-10) 
    assert _close(gwv, ref["grad_w"], rtol=1e-8, atol=1e-9 * max(1.0, np.abs(ref["grad_w"]).max()))
    assert _close(r.val, ref["val"], rtol=1e-10) and r.walkers == nw and r.steps == nw * steps
    # continuation (kinetic_vqmc!(res; steps)) keeps the stream and the estimators
    r2 = w.run(p, 50, keep_acc=True)
    ref2 = ao.kinetic_vqmc(p, ref["addrs"].copy(), 50, seed=seed, walker0=off, step0=steps)
    assert np.array_equal(w.addresses(), ref2["addrs"][:, : o.W]) and r2.steps == nw * 50 and w.steps_done() == steps + 50


@pytest.mark.parametrize("kind", list(KINDS))
@pytest.mark.parametrize("N,M,u,v,t", [(5, 5, 1.0, 0.0, 0.1), (5, 5, 1.0, 1.0, 0.1), (4, 6, 2.0, 0.5, 1.0), (3, 2, 1.0, 0.4, 1.0)])
def test_local_energy_evaluator_matches_oracle(gw, oracle_mod, kind, N, M, u, v, t):
    o, ao, H, ans = _make(gw, oracle_mod, kind, N, M, u, v, t)
    rng = np.random.default_rng(5)
    p = _params(ao.P, rng, 0.5)
    basis = o.build_basis()
    ref_val, ref_grad = ao.lee_val_and_grad(basis, p)
    for le in (gw.LocalEnergyEvaluator(H, ans, basis=basis[:, : o.W]), gw.LocalEnergyEvaluator(H, ans)):
        assert le.dim == len(basis)
        val, grad = le.val_and_grad(p)
        assert _close(val, ref_val) and _close(le(p), ref_val)
        assert _close(grad, ref_grad, atol=1e-12 * max(1.0, np.abs(ref_grad).max()))
    le = gw.LocalEnergyEvaluator(H, ans, basis=basis[:, : o.W])
    terms = le.terms(p)
    for i in rng.integers(0, len(basis), 20):
        ref = ao.lee_terms(basis[i], p)
        assert _close(terms[i], ref, atol=1e-13 * np.abs(ref).max())
    assert np.array_equal(le.states(), basis[:, : o.W])


def test_ranked_sweep_bose12_relative_jastrow(gw, oracle_mod):
    """BASELINE config 2 shape ({12,12}, 1,352,078 states) with a 6-parameter ansatz: ranked enumeration vs
    the stored-basis sweep and the oracle on a subsample of per-state terms."""
    N = M = 12
    o, ao, H, ans = _make(gw, oracle_mod, "relative_jastrow", N, M, 2.0, 0.0, 1.0)
    p = np.array([0.3, 0.05, 0.02, 0.01, 0.0, -0.01])
    le = gw.LocalEnergyEvaluator(H, ans)
    assert le.dim == 1352078
    val, grad = le.val_and_grad(p)
    rng = np.random.default_rng(2)
    idx = rng.integers(0, le.dim, 40)
    num = den = 0.0
    for i in idx:
        st = le.states(int(i), 1)[0]
        t = le.terms(p, int(i), 1)[0]
        ref = ao.lee_terms(np.pad(st, (0, 4 - o.W)), p)
        assert _close(t, ref, atol=1e-13 * np.abs(ref).max())
    # Gutzwiller limit: p = (g u / 2, 0, ...) reproduces the fast-path sweep exactly (C_0 = sum n^2 = 2 D/u + N)
    g = 0.3331889106855026
    pg = np.zeros(6)
    pg[0] = g * 2.0 / 2
    fast = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    assert abs(le(pg) - fast(g)) < 1e-10 * abs(fast(g))
    assert np.isfinite(val) and np.all(np.isfinite(grad))


def test_vqmc_energy_within_error_bars_and_sharding(gw, oracle_mod):
    o, ao, H, ans = _make(gw, oracle_mod, "relative_jastrow", 6, 6, 2.0, 0.0, 1.0)
    p = np.array([0.25, 0.05, 0.02])
    le = gw.LocalEnergyEvaluator(H, ans)
    exact, egrad = le.val_and_grad(p)
    qmc = gw.KineticVQMC(H, ans, samples=4096 * 2000, walkers=4096, record=False, burn_in=200, estimator="pooled")
    val, grad = qmc.val_and_grad(p)
    r = qmc.result.last
    assert abs(val - exact) < 5 * r.err + 1e-3
    assert np.all(np.abs(grad - egrad) < 0.05 * max(1.0, np.abs(egrad).max()))
    # walker trajectories depend on the global id only: two shards == one set
    h = ans.handle()
    a = gw.GenWalkers(h, 64, H.address.words, seed=9, global_offset=0)
    b1 = gw.GenWalkers(h, 40, H.address.words, seed=9, global_offset=0)
    b2 = gw.GenWalkers(h, 24, H.address.words, seed=9, global_offset=40)
    a.run(p, 100); b1.run(p, 100); b2.run(p, 100)
    assert np.array_equal(a.addresses(), np.concatenate([b1.addresses(), b2.addresses()]))
    va, ga = a.estimates()
    vb = np.concatenate([b1.estimates()[0], b2.estimates()[0]])
    assert np.array_equal(va, vb)


@pytest.mark.parametrize("onr", [(1, 1, 1, 1, 1), (1, 2, 1, 1)])
def test_multinomial_is_exact_for_u0(gw, oracle_mod, onr):
    """test/ansatz.jl:66-80"""
    H = gw.HubbardReal1D(gw.BoseFS(onr), u=0.0)
    o = oracle_mod.Oracle(sum(onr), len(onr), 0.0, 1.0)
    ao = oracle_mod.AnsatzOracle(o, "multinomial", normalize=True)
    basis = o.build_basis(o.from_onr(onr))
    Hm = np.zeros((len(basis), len(basis)))
    index = {int(b[0]): i for i, b in enumerate(basis)}
    for i, b in enumerate(basis):
        for new, me in o.offdiagonals(b):
            Hm[index[int(new[0])], i] += me
    e0 = np.linalg.eigvalsh(Hm)[0]
    le = gw.LocalEnergyEvaluator(H, gw.MultinomialAnsatz(H, normalize=True), basis=basis[:, : o.W])
    assert abs(le([0.5]) - e0) < 1e-10


def test_amsgrad_relative_jastrow_and_fixed_params(gw):
    """test/amsgrad.jl:23-57: KineticVQMC + RelativeJastrow reaches the LocalEnergyEvaluator optimum; DensityProfile with a
    fixed parameter converges to the fixed value everywhere."""
    H = gw.HubbardReal1D(gw.BoseFS(1, 1, 1, 1))
    ans = gw.RelativeJastrowAnsatz(H)
    le = gw.LocalEnergyEvaluator(H, ans)
    ref = gw.amsgrad(le, [0.5, 0.5], maxiter=2000, α=0.02)
    assert ref.converged
    qmc = gw.KineticVQMC(H, ans, samples=2e6, walkers=1024, record=False, burn_in=100, estimator="pooled")
    ams = gw.amsgrad(qmc, [0.5, 0.5], maxiter=100)
    gd = gw.gradient_descent(qmc, [0.5, 0.5], maxiter=100)
    assert abs(ams.value[-1] - ref.value[-1]) < 2e-2 and abs(gd.value[-1] - ref.value[-1]) < 2e-2
    assert ams.param.shape == (101, 2) and ams.gradient.shape == (101, 2)
    dp = gw.DensityProfileAnsatz(H)
    led = gw.LocalEnergyEvaluator(H, dp)
    rng = np.random.default_rng(1)
    tr = gw.amsgrad(led, [rng.random(), rng.random(), 0.75, rng.random()], fix_params=(False, False, True, False),
                    maxiter=1000)
    assert np.all(np.abs(tr.param[-1] - 0.75) < 1e-1) and tr.param[-1][2] == 0.75


def test_gen_errors(gw):
    H = gw.HubbardReal1D(gw.near_uniform(4, 4))
    ans = gw.RelativeJastrowAnsatz(H)
    with pytest.raises(ValueError):
        ans(H.address, [0.1])  # wrong number of parameters
    big = gw.HubbardReal1D(gw.near_uniform(30, 30))
    with pytest.raises(ValueError, match="256 parameters"):
        gw.JastrowAnsatz(big).handle()
    He = gw.ExtendedHubbardReal1D(gw.near_uniform(4, 4), v=0.5)
    g = gw.GutzwillerAnsatz(He)
    assert getattr(g, "generic", False) and isinstance(g, gw.GutzwillerAnsatz)
    with pytest.raises(ValueError, match="gwz_gen"):
        gw.Walkers(He.model(), 4, He.address.words)


def test_generic_engine_against_committed_oracle_vectors(gw):
    """tests/golden/ansatz_vectors.json: per-address and LocalEnergyEvaluator parity without the oracle library."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ansatz_vectors.json")
    for c in json.load(open(path))["cases"]:
        addr = gw.near_uniform(c["N"], c["M"])
        H = gw.ExtendedHubbardReal1D(addr, u=c["u"], v=c["v"], t=c["t"]) if c["v"] else gw.HubbardReal1D(addr, u=c["u"], t=c["t"])

        def make(kind):
            if kind == "gutzwiller":
                return gw.GenericGutzwillerAnsatz(H)
            cls = getattr(gw, KINDS[kind])
            return cls(H, normalize=True) if (kind == "multinomial" and c["normalize"]) else cls(H)

        ans = make(c["kind"]) if isinstance(c["kind"], str) else make(c["kind"][0]) + make(c["kind"][1])
        p = np.array(c["params"])
        h = ans.handle()
        addrs = np.array([r["addr"] for r in c["addresses"]], dtype=np.uint64)
        val, grad = h.eval(addrs, p)
        out = h.probe(addrs, p, np.full(len(addrs), 0.37), want_rates=True)
        for i, r in enumerate(c["addresses"]):
            assert _close(val[i], r["val"]) and _close(grad[i], r["grad"], atol=1e-13 * abs(r["val"]) * max(1.0, np.abs(np.array(r["grad"]) / r["val"]).max()))
            assert _close(out["e_loc"][i], r["e_loc"], atol=1e-12) and _close(out["residence_time"][i], r["tau"])
            assert _close(out["grad_ratio"][i], r["grad_ratio"], atol=1e-12) and out["n_hops"][i] == r["n_hops"]
            assert out["chosen"][i] == r["chosen"] and _close(out["rates"][i][: len(r["rates"])], r["rates"])
        v, g = gw.LocalEnergyEvaluator(H, ans).val_and_grad(p)
        assert _close(v, c["le_val"]) and _close(g, c["le_grad"], atol=1e-12 * max(1.0, np.abs(c["le_grad"]).max()))


def test_vector_ansatz_samples_an_eigenvector(gw, oracle_mod):
    """test/qmc.jl:15-51: with the exact ground state as VectorAnsatz every local energy equals E0 (zero variance), the
    residence-time histogram converges to |v|^2, the gradient is empty, and a continuation doubles the rows."""
    N = M = 5
    o = oracle_mod.Oracle(N, M, 1.0, 0.1)  # HubbardReal1D(near_uniform(BoseFS{5,5}); t=0.1) (test/qmc.jl:11)
    basis = o.build_basis()
    index = {int(b[0]): i for i, b in enumerate(basis)}
    Hm = np.zeros((len(basis), len(basis)))
    for i, b in enumerate(basis):
        Hm[i, i] = o.diagonal_element(b)
        for new, me in o.offdiagonals(b):
            Hm[index[int(new[0])], i] += me
    w, vecs = np.linalg.eigh(Hm)
    E0, gs = w[0], vecs[:, 0] * np.sign(vecs[0, 0])
    H = gw.HubbardReal1D(gw.near_uniform(N, M), t=0.1)
    va = gw.VectorAnsatz((basis[:, :1], gs))
    assert gw.num_parameters(va) == 0
    res = gw.kinetic_vqmc(H, va, [], samples=1e6, walkers=256, record=True)
    val, grad = gw.val_and_grad(res)
    assert len(grad) == 0 and abs(val - E0) < 1e-9
    el = res.local_energies()
    assert np.max(np.abs(el - E0)) < 1e-8  # zero-variance property of an eigenvector
    addrs, vals = gw.sampled_vector(res)
    sq = gs ** 2 / np.linalg.norm(gs ** 2)
    lookup = {int(b[0]): s for b, s in zip(basis, sq)}
    dot = sum(v * lookup[int(a)] for a, v in zip(addrs[:, 0], vals))
    assert abs(dot - 1.0) < 2e-4
    n0 = len(res)
    gw.kinetic_vqmc_continue(res)
    assert len(res) == 2 * n0
    # LocalEnergyEvaluator over keys(v) is the Rayleigh quotient; lookups follow DVec semantics (0 for missing keys)
    le = gw.LocalEnergyEvaluator(H, va)
    assert abs(le([]) - E0) < 1e-10 and le.dim == len(basis)
    h = va.handle(H)
    assert np.allclose(h.lookup(basis[:7, :1]), gs[:7], rtol=0, atol=0)
    part = gw.VectorAnsatz((basis[:50, :1], gs[:50]))
    assert part.handle(H).lookup(basis[60:61, :1])[0] == 0.0
    # a walker on a partial vector never leaves its support
    if abs(gs[0]) > 0:
        r2 = gw.kinetic_vqmc(H, part, [], walkers=8, steps=200, record=True)
        seen = set(int(a) for a in r2.addresses()[:, :, 0].ravel())
        assert seen <= set(int(b[0]) for b in basis[:50])


def test_collect_to_vec_and_vector_ansatz_round_trip(gw):
    """DVec(ansatz, params) (abstract.jl:24-41) materialised on the device, fed back as a VectorAnsatz: its Rayleigh
    quotient equals the LocalEnergyEvaluator value of the original ansatz (test/ansatz.jl:37-38)."""
    H = gw.HubbardReal1D(gw.near_uniform(6, 6), u=2.0)
    for ans, p in ((gw.GutzwillerAnsatz(H), [0.4]), (gw.RelativeJastrowAnsatz(H), [0.3, 0.05, 0.02])):
        addrs, vals = gw.collect_to_vec(ans, p)
        assert addrs.shape == (462, 1) and np.all(vals > 0)
        rq = gw.LocalEnergyEvaluator(H, gw.VectorAnsatz((addrs, vals)))([])
        le = gw.LocalEnergyEvaluator(H, ans)(p)
        assert abs(rq - le) < 1e-11 * abs(le)


@pytest.mark.parametrize("cls,N,M,v", [("RelativeJastrowAnsatz", 8, 8, 0.0), ("GenericGutzwillerAnsatz", 6, 9, 0.5),
                                       ("ExtendedGutzwillerAnsatz", 7, 5, 0.4), ("MultinomialAnsatz", 10, 4, 0.0),
                                       ("DensityProfileAnsatz", 6, 6, 0.0), ("JastrowAnsatz", 5, 5, 0.0)])
def test_generic_sweep_representative_basis(gw, cls, N, M, v, monkeypatch):
    """Ansaetze that are invariant under translations and reflection of the chain sweep a stored representative basis
    (orbit-size weights) from the second call of a handle on; the others keep the state-by-state sweep.  Values and
    gradients must not depend on it."""
    H = gw.ExtendedHubbardReal1D(gw.near_uniform(N, M), u=2.0, v=v, t=0.8) if v else gw.HubbardReal1D(gw.near_uniform(N, M), u=2.0, t=0.8)
    ans = getattr(gw, cls)(H)
    rng = np.random.default_rng(4)
    p = 0.2 * rng.random(ans.N_PARAMS) + 0.05
    monkeypatch.setenv("GWZ_SWEEP_CACHE_AT", "2")
    monkeypatch.setenv("GWZ_SWEEP_CACHE", "0")
    ref = gw.LocalEnergyEvaluator(H, ans)
    v0, g0 = [ref.val_and_grad(p) for _ in range(2)][-1]
    monkeypatch.delenv("GWZ_SWEEP_CACHE")
    le = gw.LocalEnergyEvaluator(H, ans)
    outs = [le.val_and_grad(p) for _ in range(4)]
    for v1, g1 in outs:
        assert abs(v1 - v0) <= 1e-12 * abs(v0) and np.allclose(g1, g0, rtol=1e-10, atol=1e-12 * max(1.0, np.abs(g0).max()))
    assert outs[2][0] == outs[3][0] and np.array_equal(outs[2][1], outs[3][1])
    q = p * 1.3
    a, b = le.val_and_grad(q), ref.val_and_grad(q)
    assert abs(a[0] - b[0]) <= 1e-12 * abs(b[0]) and np.allclose(a[1], b[1], rtol=1e-10, atol=1e-12 * max(1.0, np.abs(b[1]).max()))
    assert abs(le(q) - b[0]) <= 1e-12 * abs(b[0])


@pytest.mark.parametrize("cls,N,M,v", [("RelativeJastrowAnsatz", 50, 50, 0.0), ("ExtendedGutzwillerAnsatz", 30, 24, 0.0),
                                       ("DensityProfileAnsatz", 20, 33, 0.0), ("GenericGutzwillerAnsatz", 24, 24, 0.4)])
def test_carried_hop_rates_follow_the_re_evaluated_ones(gw, cls, N, M, v, monkeypatch):
    """Translation-invariant couplings: the walker kernel carries exp(-e) of every hop and multiplies it by tabulated
    factors (gwz_generic.cuh: gen_rates_ti_hop), re-evaluating with exp() every GEN_REFRESH = 256 steps.  Against the
    kernel that evaluates every rate with exp() in every step (GWZ_GEN_TI=0; the one the oracle tests pin): same
    trajectories over several re-evaluation periods, residence times and local energies within 1e-12."""
    addr = gw.near_uniform(N, M)
    H = gw.ExtendedHubbardReal1D(addr, u=2.0, v=v) if v else gw.HubbardReal1D(addr, u=2.0)
    ans = getattr(gw, cls)(H)
    rng = np.random.default_rng(3)
    p = 0.02 * rng.random(ans.N_PARAMS)
    p[0] = 0.4
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("GWZ_GEN_TI", flag)
        w = gw.GenWalkers(ans.handle(), 40, H.address.words, seed=77, global_offset=3)
        r, tau, eloc, a = w.run_record(p, 700)
        out[flag] = (tau, eloc, a, w.addresses(), r.val)
    (t1, e1, a1, f1, v1), (t0, e0, a0, f0, v0) = out["1"], out["0"]
    assert np.array_equal(a1, a0) and np.array_equal(f1, f0)
    assert np.all(np.abs(t1 - t0) <= 1e-12 * np.abs(t0))
    assert np.all(np.abs(e1 - e0) <= 1e-12 * np.maximum(np.abs(e0), 1.0))
    assert abs(v1 - v0) <= 1e-12 * abs(v0)
