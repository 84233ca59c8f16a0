"""GPU parity of the walker kernel: trajectories, per-step values, estimators, statistics."""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "readme_vectors.json")))
SEED = 0x9E3779B97F4A7C15


def _words(o, a):
    return a[: o.W]


@pytest.mark.parametrize("N,M,u,g", [(10, 10, 2.0, 1.0), (20, 20, 4.0, 0.25), (50, 50, 2.0, 0.33), (100, 100, 2.0, 1.0),
                                     (7, 3, 1.0, 0.3)])
def test_enumerate_kernel_follows_oracle_trajectories(gw, oracle_mod, N, M, u, g):
    """Same Philox stream + the reference's selection rule => identical trajectories, step for step."""
    o = oracle_mod.Oracle(N, M, u, 1.0)
    model = gw.Model(N, M, u, 1.0)
    nw, steps = 37, 150
    start = o.near_uniform()
    ws = gw.Walkers(model, nw, _words(o, start), seed=SEED, global_offset=5)
    r, tau, eloc, addr = ws.run_record([g], steps, kernel=gw.KERNEL_ENUMERATE)
    ref = o.kinetic_vqmc(g, np.stack([start] * nw), steps, seed=SEED, walker0=5, series=True)
    assert np.array_equal(addr.transpose(1, 0, 2), ref["series_addr"][:, :, : o.W])
    assert np.array_equal(ws.addresses(), ref["addrs"][:, : o.W])
    assert np.allclose(tau.T, ref["series_tau"], rtol=1e-12, atol=0)
    assert np.allclose(eloc.T, ref["series_eloc"], rtol=1e-12, atol=1e-12)
    # estimators: per-walker (state.jl:42-49) and result (state.jl:137-149)
    val_w, grad_w = ws.estimates()
    assert np.allclose(val_w, ref["val_w"], rtol=1e-12, atol=0)
    assert np.allclose(grad_w[:, 0], ref["grad_w"], rtol=1e-9, atol=1e-10)
    assert abs(r.val - ref["val"]) <= 1e-12 * abs(ref["val"])
    assert abs(r.grad[0] - ref["grad"]) <= 1e-9 * abs(ref["grad"]) + 1e-10
    assert r.walkers == nw and r.steps == nw * steps and r.hops == ref["hops"]


@pytest.mark.parametrize("N,M,u,g", [(10, 10, 2.0, 1.0), (12, 12, 2.0, 0.2), (50, 50, 2.0, 0.33), (100, 100, 2.0, 0.33),
                                     (5, 7, 3.0, 0.05)])
def test_production_kernel_per_step_values(gw, oracle_mod, N, M, u, g):
    """Record the bit-parallel kernel's series and check every visited address against the oracle:
    E_loc and residence time to 1e-12, each move a genuine hop chosen by that step's uniform."""
    o = oracle_mod.Oracle(N, M, u, 1.0)
    model = gw.Model(N, M, u, 1.0)
    nw, steps = 16, 120
    start = o.near_uniform()
    ws = gw.Walkers(model, nw, _words(o, start), seed=11, global_offset=0)
    r, tau, eloc, addr = ws.run_record([g], steps, kernel=gw.KERNEL_BITPARALLEL)
    final = ws.addresses()
    full = np.zeros(4, dtype=np.uint64)
    hops = 0
    for w in range(nw):
        for k in range(steps):
            full[:] = 0
            full[: o.W] = addr[k, w]
            ks = o.kinetic_sample(full, g, 0.5)
            hops += ks["n_hops"]
            assert abs(tau[k, w] - ks["tau"]) <= 1e-12 * ks["tau"]
            assert abs(eloc[k, w] - ks["eloc"]) <= 1e-12 * abs(ks["eloc"]) + 1e-12
            nxt = addr[k + 1, w] if k + 1 < steps else final[w]
            neigh = [tuple(x[: o.W]) for x, _ in o.offdiagonals(full)]
            assert tuple(nxt) in neigh
    assert r.hops == hops
    # library estimators == reference formulas applied to the recorded series
    grads = -np.array([[o.diagonal_element(np.concatenate([addr[k, w], np.zeros(4 - o.W, dtype=np.uint64)]))
                        for k in range(steps)] for w in range(nw)])
    vals, gs = [], []
    for w in range(nw):
        st = tau[:, w].sum()
        v = (tau[:, w] * eloc[:, w]).sum() / st
        vals.append(v)
        gs.append(2 * (tau[:, w] * grads[w] * (eloc[:, w] - v)).sum() / st)
    assert abs(r.val - np.mean(vals)) <= 1e-12 * abs(np.mean(vals))
    assert abs(r.grad[0] - np.mean(gs)) <= 1e-9 * abs(np.mean(gs)) + 1e-9
    pooled = (tau * eloc).sum() / tau.sum()
    assert abs(r.pooled_val - pooled) <= 1e-12 * abs(pooled)
    assert abs(r.total_time - tau.sum()) <= 1e-12 * tau.sum()


def test_production_kernel_with_large_occupations(gw, oracle_mod):
    """Start from all bosons on one site: the first steps run through the generic (>= KCAP) bond path."""
    N, M, u, g = 9, 6, 0.5, 0.05
    o = oracle_mod.Oracle(N, M, u, 1.0)
    model = gw.Model(N, M, u, 1.0)
    onr = [0] * M
    onr[M - 1] = N
    start = o.from_onr(onr)
    ws = gw.Walkers(model, 64, _words(o, start), seed=3)
    r, tau, eloc, addr = ws.run_record([g], 60, kernel=gw.KERNEL_BITPARALLEL)
    full = np.zeros(4, dtype=np.uint64)
    big_seen = 0
    for w in range(64):
        for k in range(60):
            full[: o.W] = addr[k, w]
            big_seen += max(o.to_onr(full)) >= 4
            ks = o.kinetic_sample(full, g, 0.5)
            assert abs(tau[k, w] - ks["tau"]) <= 1e-12 * ks["tau"]
            assert abs(eloc[k, w] - ks["eloc"]) <= 1e-12 * abs(ks["eloc"]) + 1e-12
    assert big_seen > 500


def test_readme_example_one_walker(gw):
    """BASELINE config 1: HubbardReal1D BoseFS{10,10} u=2, GutzwillerAnsatz p=[1.0], steps=1e5, 1 walker
    (README.md:214-218: -7.8369 +- 0.016555).  Energy within error bars of the exact sweep value."""
    H = gw.HubbardReal1D(gw.near_uniform(10, 10), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    res = gw.kinetic_vqmc(H, ans, [1.0], steps=1e5, walkers=1)
    assert res.n_walkers == 1 and len(res) == 100000
    est = gw.local_energy_estimator(res)
    exact = GOLDEN["lee_bose10"]["val"]
    assert 0.008 < est.err < 0.035  # README: 0.0166
    assert abs(est.mean - exact) < 3.5 * est.err  # test/qmc.jl:69-70
    val, grad = gw.val_and_grad(res)
    assert abs(val - gw.mean(res)) < 1e-12
    assert abs(val - est.mean) < 1e-9  # resampling preserves the weighted mean
    v2, err, g2 = gw.val_err_and_grad(res)
    assert abs(v2 - val) < 1e-9 and abs(err - est.err) < 1e-12 and abs(g2[0] - grad[0]) < 1e-6 * abs(grad[0])
    assert gw.var(res) > 0
    # continuation doubles the length (test/qmc.jl:49-50)
    gw.kinetic_vqmc_continue(res)
    assert len(res) == 200000
    est2 = gw.local_energy_estimator(res)
    assert est2.err < est.err and abs(est2.mean - exact) < 3.5 * est2.err
    rows = res.rows()
    first = next(rows)
    assert first["walker"] == 1 and first["step"] == 1 and first["address"] == gw.near_uniform(10, 10)
    assert abs(first["local_energy"] - (-20 * math.sqrt(2) * math.exp(-2))) < 1e-13


def test_many_walkers_statistics_and_wrapper(gw):
    """test/qmc.jl:53-75 shape: Gutzwiller p=0.5; exact Rayleigh quotient within 3 sigma; wrapper agrees."""
    H = gw.HubbardReal1D(gw.near_uniform(5, 5), t=0.1)
    ans = gw.GutzwillerAnsatz(H)
    le = gw.LocalEnergyEvaluator(H, ans)
    exact, exact_grad = le.val_and_grad(0.5)
    res = gw.kinetic_vqmc(H, ans, 0.5, samples=4e6, walkers=2048, record=False)
    val1, grad1 = gw.val_and_grad(res)
    est = gw.local_energy_estimator(res)
    assert est.err > 0 and abs(est.mean - exact) < 3.5 * est.err
    qmc = gw.KineticVQMC(H, ans, samples=4e6, walkers=2048)
    val2, grad2 = gw.val_and_grad(qmc, 0.5)
    assert val1 == pytest.approx(val2, rel=1e-1) and grad1[0] == pytest.approx(grad2[0], rel=1e-1, abs=1e-2)
    v3, e3, g3 = gw.val_err_and_grad(qmc, 0.5)
    assert abs(v3 - exact) < 4 * e3 and abs(g3[0] - exact_grad[0]) < 0.05 * max(1, abs(exact_grad[0]))
    assert abs(qmc(0.5) - exact) < 4 * e3
    G = np.zeros(1)
    F = qmc(True, G, 0.5)
    assert abs(F - exact) < 4 * e3 and G[0] != 0.0
    assert qmc(None, G, 0.5) is None
    # keyword overrides: steps = samples ÷ walkers (wrapper.jl:83)
    gw.val_and_grad(qmc, 0.5, samples=2048 * 10 + 5)
    assert qmc.result.steps == 10


def test_bigger_system_energy_vs_sweep(gw):
    """BoseFS{12,12} (1,352,078 states): VQMC with burn-in vs the deterministic sweep, pooled and per-walker."""
    H = gw.HubbardReal1D(gw.near_uniform(12, 12), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    exact, egrad = gw.LocalEnergyEvaluator(H, ans).val_and_grad(0.3331889106855026)
    res = gw.kinetic_vqmc(H, ans, 0.3331889106855026, walkers=8192, steps=2000, burn_in=240, record=False)
    r = res.last
    assert abs(r.val - exact) < 4 * r.err and r.err < 0.01
    assert abs(r.pooled_val - exact) < 4 * r.err
    # The reference's estimator (mean over walkers of per-walker covariances, state.jl:42-49,137-149) is
    # biased by O(tau_corr/steps) per walker, and that does not average out over walkers (SURVEY.md 7,
    # "Estimator fidelity"); the pooled covariance is the unbiased comparison with the sweep gradient.
    assert abs(r.pooled_grad[0] - egrad[0]) < 5 * r.grad_err[0] + 0.005
    assert abs(r.grad[0] - egrad[0]) < 0.15
    assert r.walkers == 8192 and r.steps == 8192 * 2000
    pooled = gw.kinetic_vqmc(H, ans, 0.3331889106855026, walkers=8192, steps=2000, burn_in=240, estimator="pooled")
    assert pooled.last.val == pooled.last.pooled_val and pooled.last.grad[0] == pooled.last.pooled_grad[0]
    assert abs(gw.val_and_grad(pooled)[1][0] - egrad[0]) < 5 * r.grad_err[0] + 0.005


def test_sharding_invariance_logical_shards(gw, oracle_mod):
    """results do not depend on how walkers are sharded.  G logical shards on one GPU
    (global ids offset..) reduce to the single-set result up to FP reduction order."""
    N, M, u, g = 20, 20, 4.0, 0.25
    o = oracle_mod.Oracle(N, M, u, 1.0)
    model = gw.Model(N, M, u, 1.0)
    total, steps = 4096, 64
    start = o.near_uniform()[: o.W]
    single = gw.Walkers(model, total, start, seed=77).run([g], steps)
    for G in (2, 4, 8):
        sets = []
        for rnk in range(G):
            b, e = gw.shard_range(total, G, rnk)
            sets.append(gw.Walkers(model, e - b, start, seed=77, global_offset=b))
        r = gw.run_group(sets, [g], steps)
        assert r.walkers == total and r.steps == single.steps and r.hops == single.hops
        assert abs(r.val - single.val) <= 1e-13 * abs(single.val)
        assert abs(r.grad[0] - single.grad[0]) <= 1e-10 * abs(single.grad[0]) + 1e-12
        assert abs(r.pooled_val - single.pooled_val) <= 1e-13 * abs(single.pooled_val)
        addrs = np.concatenate([s.addresses() for s in sets])
        ref = gw.Walkers(model, total, start, seed=77)
        ref.run([g], steps)
        assert np.array_equal(addrs, ref.addresses())


def test_continuation_and_checkpoint(gw, oracle_mod):
    """kinetic_vqmc!(res) continues from curr_address (qmc.jl:73-97); get/set addresses + step counter
    reproduce a run exactly (checkpoint/resume)."""
    N, M = 10, 10
    o = oracle_mod.Oracle(N, M, 2.0, 1.0)
    model = gw.Model(N, M, 2.0, 1.0)
    start = o.near_uniform()[: o.W]
    a = gw.Walkers(model, 300, start, seed=9)
    ra = a.run([1.0], 100)
    ra2 = a.run([1.0], 50, keep_acc=True)
    b = gw.Walkers(model, 300, start, seed=9)
    rb = b.run([1.0], 150)
    assert np.array_equal(a.addresses(), b.addresses())
    assert abs(ra2.val - rb.val) < 1e-12 * abs(rb.val) and ra2.steps == 300 * 50 and a.steps_done() == 150
    assert abs(ra2.grad[0] - rb.grad[0]) < 1e-9 * abs(rb.grad[0])
    # checkpoint after 100 steps, resume elsewhere
    c = gw.Walkers(model, 300, start, seed=9)
    c.run([1.0], 100)
    saved = c.addresses()
    d = gw.Walkers(model, 300, start, seed=9)
    d.set_addresses(saved)
    d.set_step_counter(100)
    d.run([1.0], 50)
    assert np.array_equal(d.addresses(), b.addresses())
    assert ra.steps == 300 * 100


def test_walker_argument_errors(gw):
    model = gw.Model(4, 4, 1.0, 1.0)
    good = gw.near_uniform(4, 4).words
    with pytest.raises(ValueError):
        gw.Walkers(model, 0, good)
    with pytest.raises(ValueError):
        gw.Walkers(model, 4, np.array([0b111], dtype=np.uint64))  # 3 particles, not 4
    w = gw.Walkers(model, 4, good)
    with pytest.raises(ValueError):
        w.run([0.5], 0)
    with pytest.raises(ValueError):
        w.run([float("nan")], 10)
    with pytest.raises(ValueError):
        w.estimates()


@pytest.mark.parametrize("kernel", ["bitparallel", "enumerate"])
def test_sampler_draws_from_psi_squared(gw, kernel):
    """test/qmc.jl:53-75 ("Sampling vector squared"): the residence-time histogram of the sampled addresses
    converges to |psi|^2 = normalize(abs2(ansatz)): dot(sampled, vector_sq) ~ 1 (rtol 1e-4 at 1e6 samples)."""
    H = gw.HubbardReal1D(gw.near_uniform(5, 5), t=0.1)  # the reference's test Hamiltonian (test/qmc.jl:11)
    ans = gw.GutzwillerAnsatz(H)
    kind = gw.KERNEL_BITPARALLEL if kernel == "bitparallel" else gw.KERNEL_ENUMERATE
    res = gw.kinetic_vqmc(H, ans, 0.5, samples=2e6, walkers=256, record=True, kernel=kind)
    addrs, vals = gw.sampled_vector(res)
    le = gw.LocalEnergyEvaluator(H, ans)
    states = le.states()[:, 0]
    psi2 = le.terms(0.5)[:, 1]  # den = psi^2 per state (localenergy.jl:111)
    psi2 = psi2 / np.linalg.norm(psi2)
    assert len(states) == 126 and set(int(a) for a in addrs[:, 0]) <= set(int(s) for s in states)
    lookup = {int(s): p for s, p in zip(states, psi2)}
    dot = sum(v * lookup[int(a)] for a, v in zip(addrs[:, 0], vals))
    assert abs(dot - 1.0) < 1e-4
    assert len(addrs) == 126  # every basis state visited
    # exact energy within 3 sigma of the blocked estimate (test/qmc.jl:69-70)
    est = gw.local_energy_estimator(res)
    exact = le(0.5)
    assert est.mean - 3.5 * est.err < exact < est.mean + 3.5 * est.err


@pytest.mark.parametrize("N,M,u,g", [(10, 10, 2.0, 1.0), (50, 50, 2.0, 0.33), (100, 100, 2.0, 0.33), (9, 6, 0.5, 0.05),
                                     (7, 3, 1.0, 0.3), (33, 32, 2.0, 0.2), (64, 65, 1.0, 0.1), (120, 131, 2.0, 0.1),
                                     (2, 2, 1.0, 0.3), (17, 40, 3.0, 0.2)])
def test_planes_kernel_is_bitwise_identical_to_unary(gw, oracle_mod, N, M, u, g, monkeypatch):
    """The production kernel keeps walkers as occupation bit-planes; class order, bond order and the
    floating-point summation order are the same as on the unary bitstring, so trajectories and every
    accumulator must agree bit for bit with the unary kernel (which is checked per step against the oracle)."""
    monkeypatch.setenv("GWZ_FAST", "0")  # the round-1 kernels among themselves (round 2's has its own u -> hop map)
    o = oracle_mod.Oracle(N, M, u, 1.0)
    model = gw.Model(N, M, u, 1.0)
    onr = [0] * M
    onr[M - 1] = N  # all bosons on the last site: exercises the generic (occupation >= 4) path and the ripple
    starts = [o.near_uniform()[: o.W], o.from_onr(onr)[: o.W]]
    for start in starts:
        out = {}
        # (GWZ_PLANES, GWZ_INCR): incremental class counts (round-1 production), full re-count on the planes
        # (walker_planes_kernel, also the fallback for M < 4), unary bitstring
        for mode in (("1", "1"), ("1", "0"), ("0", "1")):
            monkeypatch.setenv("GWZ_PLANES", mode[0])
            monkeypatch.setenv("GWZ_INCR", mode[1])
            w = gw.Walkers(model, 777, start, seed=21, global_offset=3)
            r1 = w.run([g], 97, burn_in=5)
            r2 = w.run([g], 40, keep_acc=True)
            vals, grads = w.estimates()
            out[mode] = (w.addresses(), np.array(r1.acc[:]), np.array(r2.acc[:]), vals, grads, w.steps_done())
        b = out[("0", "1")]
        for mode in (("1", "1"), ("1", "0")):
            a = out[mode]
            assert np.array_equal(a[0], b[0]), mode
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), mode
            assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]) and a[5] == b[5] == 142, mode


@pytest.mark.parametrize("N,M", [(5, 5), (12, 12), (20, 20)])
def test_device_histogram_equals_host_histogram(gw, N, M):
    """gwz_vqmc_sample_vector (DVec(res), state.jl:303-319, accumulated in an HBM hash table) against the host-side
    histogram of the same recorded run."""
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    steps, walkers = 500, 128
    res = gw.kinetic_vqmc(H, ans, 0.4, walkers=walkers, steps=steps, seed=5, record=True)
    addrs, vals = gw.sampled_vector(res)
    w = gw.Walkers(H.model(), walkers, H.address.words, seed=5)
    a2, v2, r = w.sample_vector([0.4], steps)
    assert np.array_equal(addrs, a2) and np.allclose(vals, v2, rtol=1e-12, atol=0)
    assert abs(np.linalg.norm(v2) - 1.0) < 1e-12 and r.val == res.last.val
    with pytest.raises(ValueError, match="capacity"):
        w.sample_vector([0.4], steps, cap=3)
    big = gw.HubbardReal1D(gw.near_uniform(50, 50), u=2.0)
    with pytest.raises(ValueError, match="single-word"):
        gw.Walkers(big.model(), 4, big.address.words).sample_vector([0.4], 10)


@pytest.mark.parametrize("N,M,u,g", [(50, 50, 2.0, 0.5), (20, 20, 4.0, 0.25)])
def test_large_system_energy_agrees_with_the_cpu_walk_within_error_bars(gw, oracle_mod, N, M, u, g):
    """North-star criterion at the benchmark sizes, where no exact value exists: the production kernel's energy and the
    CPU walk's (oracle.kinetic_vqmc_fast: the same trajectories as the literal restatement of the reference, see
    tests/test_oracle_ansatz.py) agree within their combined cross-walker standard errors on identical (H, ansatz, p)."""
    o = oracle_mod.Oracle(N, M, u, 1.0)
    nw_cpu, burn, steps = 96, 3000, 12000
    start = np.tile(o.near_uniform(), (nw_cpu, 1))
    o.kinetic_vqmc_fast(g, start, burn, seed=11)  # thermalise in place
    ref = o.kinetic_vqmc_fast(g, start, steps, seed=11, step0=burn)
    e_cpu, s_cpu = ref["val_w"].mean(), ref["val_w"].std(ddof=1) / np.sqrt(nw_cpu)
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=u)
    res = gw.kinetic_vqmc(H, gw.GutzwillerAnsatz(H), [g], walkers=1 << 13, steps=steps, burn_in=burn, seed=99, record=False)
    e_gpu, s_gpu = res.last.val, res.last.err
    assert abs(e_gpu - e_cpu) < 4.0 * np.hypot(s_cpu, s_gpu), (e_gpu, s_gpu, e_cpu, s_cpu)
    assert s_gpu < s_cpu  # 85x more walkers
