"""GPU parity of LocalEnergyEvaluator (src/localenergy.jl:94-155) and the AMSGrad driver."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "readme_vectors.json")))
SWEEP_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sweep_vectors.json")))
RTOL = 1e-12


def test_readme_values(gw):
    """README.md:94,104: le([1.0]) = -7.825819465047312, grad 10.614147776143358 -- to 1e-12 relative
    (the reference's own left-fold carries ~1e-12 summation-order error)."""
    H = gw.HubbardReal1D(gw.near_uniform(10, 10), u=2.0)
    le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    assert le.dim == 92378
    g = GOLDEN["lee_bose10"]
    val = le([1.0])
    v2, grad = gw.val_and_grad(le, [1.0])
    assert val == v2
    assert abs(val - g["val"]) <= 2e-12 * abs(g["val"])
    assert abs(grad[0] - g["grad"]) <= 5e-12 * abs(g["grad"])
    m = GOLDEN["lbfgs_minimum"]
    vmin, gmin = le.val_and_grad(m["param"])
    assert abs(vmin - m["val"]) <= 1e-12 * abs(m["val"]) and abs(gmin[0]) < 1e-7  # README.md:186
    # (later sweeps of a handle may run over its stored representative basis: same sums in another order)
    same = lambda a, b: abs(a - b) <= 1e-14 * abs(b)  # noqa: E731
    assert same(le(1.0), val) and same(le(*[1.0]), val)  # scalar / varargs forms (localenergy.jl:171-173)
    G = np.zeros(1)
    assert same(le(True, G, [1.0]), val) and same(G[0], grad[0])  # localenergy.jl:157-169
    assert le(None, G, [1.0]) is None and same(le(True, None, [1.0]), val)


@pytest.mark.parametrize("case", SWEEP_GOLDEN["cases"], ids=lambda c: f"N{c['N']}M{c['M']}g{c['g']}")
def test_sweep_against_committed_oracle_vectors(gw, case):
    """tests/golden/sweep_vectors.json: exactly-rounded oracle sums (mode 1), incl. BASELINE config 2
    (BoseFS{12,12}, 1,352,078 states)."""
    H = gw.HubbardReal1D(gw.near_uniform(case["N"], case["M"]), u=case["u"], t=case["t"])
    le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    assert le.dim == case["dim"]
    val, grad = le.val_and_grad(case["g"])
    assert abs(val - case["val"]) <= RTOL * abs(case["val"])
    assert abs(grad[0] - case["grad"]) <= RTOL * abs(case["grad"]) + 1e-13
    acc = le.accumulators(case["g"])
    assert np.all(np.abs(acc - np.array(case["acc"])) <= RTOL * np.abs(case["acc"]) + 1e-300)


@pytest.mark.parametrize("N,M,u,t,g", [(10, 10, 2.0, 1.0, 1.0), (5, 5, 1.0, 0.1, 0.37), (6, 9, 3.0, 1.0, 0.2),
                                       (9, 4, 0.5, 1.0, 0.1)])
def test_stored_basis_and_terms_match_oracle(gw, oracle_mod, N, M, u, t, g):
    o = oracle_mod.Oracle(N, M, u, t)
    basis = o.build_basis()  # BFS order, like build_basis(H) in the reference
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=u, t=t)
    ans = gw.GutzwillerAnsatz(H)
    le_b = gw.LocalEnergyEvaluator(H, ans, basis=basis[:, : o.W])
    le_r = gw.LocalEnergyEvaluator(H, ans)
    assert le_b.dim == le_r.dim == len(basis)
    ref_val, ref_grad = o.lee_val_and_grad(basis, g, mode=1)
    for le in (le_b, le_r):
        val, grad = le.val_and_grad(g)
        assert abs(val - ref_val) <= RTOL * abs(ref_val)
        assert abs(grad[0] - ref_grad) <= RTOL * abs(ref_grad) + 1e-13
    # per-state terms (localenergy.jl:105-120), stored-basis order
    terms = le_b.terms(g)
    idx = np.random.default_rng(0).choice(len(basis), size=min(400, len(basis)), replace=False)
    for i in idx:
        ref = o.lee_terms(basis[i], g)
        assert np.all(np.abs(terms[i] - ref) <= RTOL * np.abs(ref) + 1e-300)
    # ranked enumeration visits exactly the same set of states, in increasing bitstring order
    states = le_r.states()[:, 0]
    assert np.all(np.diff(states.astype(np.int64)) > 0)
    assert set(int(x) for x in states) == set(int(x) for x in basis[:, 0])
    t_r = le_r.terms(g, first=7, count=50)
    for j in range(50):
        full = np.zeros(4, dtype=np.uint64)
        full[0] = states[7 + j]
        assert np.all(np.abs(t_r[j] - o.lee_terms(full, g)) <= RTOL * np.abs(o.lee_terms(full, g)) + 1e-300)


def test_sweep_two_word_basis(gw, oracle_mod):
    """Stored basis with W = 2 words per address (a slice of BoseFS{40,40} reached by BFS)."""
    N, M = 40, 40
    o = oracle_mod.Oracle(N, M, 2.0, 1.0)
    # first 3000 states of the BFS closure
    import ctypes as C
    om = oracle_mod
    basis = np.zeros((3000, 4), dtype=np.uint64)
    start = o._addr(o.near_uniform())
    dim = o.L.go_build_basis(C.byref(o.m), C.byref(start), basis.ctypes.data_as(C.c_void_p), 3000)
    assert dim == -1  # cap reached: the array holds the first 3000 BFS states
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=2.0)
    le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H), basis=basis[:, :2])
    ref_val, ref_grad = o.lee_val_and_grad(basis, 0.4, mode=1)
    val, grad = le.val_and_grad(0.4)
    assert abs(val - ref_val) <= RTOL * abs(ref_val) and abs(grad[0] - ref_grad) <= RTOL * abs(ref_grad)
    del om


def test_sharded_sweep_sums_to_full(gw):
    """Rank-range shards (what each GPU sweeps) add up to the full sweep."""
    import ctypes as C
    H = gw.HubbardReal1D(gw.near_uniform(10, 10), u=2.0)
    model = H.model()
    full = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H)).accumulators(0.7)
    dim = model.dimension()
    acc = np.zeros(4)
    for r in range(3):
        b, e = gw.shard_range(dim, 3, r)
        h = C.c_void_p()
        gw.capi.check(gw.lib().gwz_sweep_create(model.handle, None, 0, b, e, C.byref(h)))
        a4 = (C.c_double * 4)()
        v, g = C.c_double(), (C.c_double * 1)()
        p = (C.c_double * 1)(0.7)
        gw.capi.check(gw.lib().gwz_sweep_val_and_grad(h, p, C.byref(v), g, a4))
        acc += np.array(a4[:])
        gw.lib().gwz_sweep_destroy(h)
    assert np.all(np.abs(acc - full) <= 1e-13 * np.abs(full))


def test_amsgrad_on_sweep_matches_oracle_trace(gw, oracle_mod):
    """amsgrad(le, p0) (amsgrad.jl:259-263) is deterministic: the GPU-driven trace must equal the
    oracle's literal restatement row by row; README.md:286 row 1 pins the update rule."""
    N, M = 10, 10
    o = oracle_mod.Oracle(N, M, 2.0, 1.0)
    basis = o.build_basis()
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=2.0)
    le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    tr = gw.amsgrad(le, [1.0], maxiter=100)
    ref = o.amsgrad_lee(basis, 1.0, maxiter=100, mode=1)
    assert tr.iterations == ref["iterations"] == 101 and tr.reason == ref["reason"]
    assert np.allclose(tr.param[:, 0], ref["param"], rtol=0, atol=1e-10)
    assert np.allclose(tr.value, ref["value"], rtol=1e-11, atol=0)
    assert np.allclose(tr.gradient[:, 0], ref["gradient"], rtol=1e-9, atol=1e-10)
    assert np.allclose(tr.first_moment[:, 0], ref["first_moment"], rtol=1e-9, atol=1e-10)
    assert np.allclose(tr.second_moment[:, 0], ref["second_moment"], rtol=1e-9, atol=1e-10)
    assert np.allclose(tr.param_delta[:, 0], ref["param_delta"], rtol=1e-9, atol=1e-12)
    assert np.all(tr.error == 0.0)
    # row 1 of the README trace is the same deterministic arithmetic up to its VQMC noise
    r1 = GOLDEN["amsgrad_row1"]
    assert abs(tr.param_delta[0, 0] - r1["param_delta"]) < 2e-3 and abs(tr.gradient[0, 0] - r1["gradient"]) < 0.1
    # approaches the L-BFGS optimum of README.md:186
    assert abs(tr.param[-1, 0] - GOLDEN["lbfgs_minimum"]["param"]) < 0.05


def test_gd_vs_amsgrad_converge(gw):
    """test/amsgrad.jl:7-21: BoseFS(1,1,1,1); both reach the minimiser (atol 1e-3)."""
    H = gw.HubbardReal1D(gw.BoseFS(1, 1, 1, 1))
    le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    gd = gw.gradient_descent(le, [0.5])
    ams = gw.amsgrad(le, [0.5], maxiter=1000)
    assert gd.converged and ams.converged
    lo, hi = 0.0, 2.0
    for _ in range(60):
        m1, m2 = lo + 0.382 * (hi - lo), lo + 0.618 * (hi - lo)
        if le(m1) < le(m2):
            hi = m2
        else:
            lo = m1
    pmin = 0.5 * (lo + hi)
    assert abs(gd.param[-1, 0] - pmin) < 1e-3 and abs(ams.param[-1, 0] - pmin) < 1e-3
    # continuation forms agree exactly (test/amsgrad.jl:76)
    a = gw.amsgrad(le, [0.5], maxiter=50, early_stop=False)
    b = gw.amsgrad(le, a.param[-1], first_moment_init=a.first_moment[-1], second_moment_init=a.second_moment[-1],
                   maxiter=50, early_stop=False)
    d = gw.amsgrad(le, a, maxiter=50, early_stop=False)
    assert np.array_equal(b.param, d.param)


def test_amsgrad_driving_vqmc(gw):
    """README.md:262-271 / test/amsgrad.jl:23-35 shape: AMSGrad on the stochastic objective reaches the
    deterministic optimum (README: p = 0.33408 vs L-BFGS 0.33319, value -13.054 vs -13.005)."""
    H = gw.HubbardReal1D(gw.near_uniform(10, 10), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    qmc = gw.KineticVQMC(H, ans, samples=4e6, walkers=4096, burn_in=200)
    tr = gw.amsgrad(qmc, [1.0])
    assert tr.iterations == 101 and tr.reason == "iterations" and not tr.converged
    assert abs(tr.param[0, 0] - 1.0) < 1e-15 and abs(tr.param_delta[0, 0] + 0.1) < 0.01  # README row 1: -0.100367
    m = GOLDEN["lbfgs_minimum"]
    assert abs(tr.param[-1, 0] - m["param"]) < 0.02
    assert abs(tr.value[-1] - m["val"]) < 0.1
    assert np.all(tr.error > 0)


@pytest.mark.parametrize("N,M", [(4, 4), (6, 6), (2, 2), (3, 9), (9, 3), (12, 12), (7, 13), (3, 40), (5, 30), (20, 3), (1, 5)])
def test_orbit_reduced_sweep_equals_full_sweep(gw, N, M, monkeypatch):
    """Ranked sweeps evaluate one representative per translation orbit (weight = orbit size, including orbits with a
    non-trivial stabiliser such as |2 0 2 0>); the four sums must agree with the state-by-state sweep, also when the
    rank range is cut into shards at arbitrary places."""
    monkeypatch.setenv("GWZ_SWEEP_CACHE_AT", "2")
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=2.0, t=0.7)
    ans = gw.GutzwillerAnsatz(H)
    out = {}
    for sym in ("1", "0"):
        monkeypatch.setenv("GWZ_SWEEP_SYM", sym)
        le = gw.LocalEnergyEvaluator(H, ans)
        out[sym] = (le.accumulators(0.3), le.val_and_grad(0.3))
    a, b = out["1"], out["0"]
    assert np.all(np.abs(a[0] - b[0]) <= 1e-12 * np.abs(b[0]))
    assert abs(a[1][0] - b[1][0]) <= 1e-12 * abs(b[1][0]) and abs(a[1][1][0] - b[1][1][0]) <= 1e-11 * max(1.0, abs(b[1][1][0]))
    monkeypatch.setenv("GWZ_SWEEP_SYM", "1")
    dim = H.model().dimension()
    if dim >= 50:
        cuts = [0, dim // 7, dim // 3 + 1, dim - 5, dim]
        acc = np.zeros(4)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            h = gw.capi.C.c_void_p()
            gw.capi.check(gw.lib().gwz_sweep_create(H.model().handle, None, 0, lo, hi, gw.capi.C.byref(h)))
            v, g, a4 = gw.capi.C.c_double(), (gw.capi.C.c_double * 1)(), (gw.capi.C.c_double * 4)()
            p = (gw.capi.C.c_double * 1)(0.3)
            gw.capi.check(gw.lib().gwz_sweep_val_and_grad(h, p, gw.capi.C.byref(v), g, a4))
            gw.lib().gwz_sweep_destroy(h)
            acc += np.array(a4[:])
        assert np.all(np.abs(acc - b[0]) <= 1e-12 * np.abs(b[0]))


def test_orbit_reduced_sweep_at_the_64_bit_limit(gw, monkeypatch):
    """N + M = 64: the closed unary string fills a 64-bit word exactly (largest system the orbit test handles)."""
    H = gw.HubbardReal1D(gw.near_uniform(60, 4), u=0.5, t=1.0)
    out = {}
    for sym in ("1", "0"):
        monkeypatch.setenv("GWZ_SWEEP_SYM", sym)
        out[sym] = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H)).accumulators(0.004)
    assert np.all(np.isfinite(out["0"])) and np.all(np.abs(out["1"] - out["0"]) <= 1e-12 * np.abs(out["0"]))


@pytest.mark.parametrize("N,M", [(6, 6), (12, 12), (9, 3), (5, 30), (20, 3), (2, 2)])
def test_cached_representative_basis_gives_the_same_sums(gw, N, M, monkeypatch):
    """From its second sweep on a handle evaluates the stored orbit representatives (state, orbit size) instead of
    enumerating its rank range again: same sums as the first (enumerating) sweep, as GWZ_SWEEP_CACHE=0 and as the
    state-by-state sweep, for several parameters and across shards."""
    monkeypatch.setenv("GWZ_SWEEP_CACHE_AT", "2")  # default: when the sweeps so far would have paid for it
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=2.0, t=0.7)
    ans = gw.GutzwillerAnsatz(H)
    le = gw.LocalEnergyEvaluator(H, ans)
    first = le.accumulators(0.3)
    again = [le.accumulators(0.3) for _ in range(3)]  # built at the second call, reused afterwards
    for a in again:
        assert np.all(np.abs(a - first) <= 1e-12 * np.abs(first))
    assert np.array_equal(again[1], again[2])  # sorted representatives: a fixed summation order
    monkeypatch.setenv("GWZ_SWEEP_CACHE", "0")
    le0 = gw.LocalEnergyEvaluator(H, ans)
    monkeypatch.setenv("GWZ_SWEEP_SYM", "0")
    le_full = gw.LocalEnergyEvaluator(H, ans)
    for g in (0.0, 0.7, 1.3):
        a, b = le.accumulators(g), [le0.accumulators(g) for _ in range(2)][-1]
        monkeypatch.setenv("GWZ_SWEEP_SYM", "0")
        c = le_full.accumulators(g)
        monkeypatch.setenv("GWZ_SWEEP_SYM", "1")
        assert np.all(np.abs(a - b) <= 1e-12 * np.abs(b)) and np.all(np.abs(a - c) <= 1e-12 * np.abs(c))
    monkeypatch.delenv("GWZ_SWEEP_CACHE")
    dim = H.model().dimension()
    if dim >= 50:
        acc = np.zeros(4)
        for lo, hi in ((0, dim // 3), (dim // 3, dim - 7), (dim - 7, dim)):
            h = gw.capi.C.c_void_p()
            gw.capi.check(gw.lib().gwz_sweep_create(H.model().handle, None, 0, lo, hi, gw.capi.C.byref(h)))
            v, gr, a4 = gw.capi.C.c_double(), (gw.capi.C.c_double * 1)(), (gw.capi.C.c_double * 4)()
            p = (gw.capi.C.c_double * 1)(0.3)
            for _ in range(3):
                gw.capi.check(gw.lib().gwz_sweep_val_and_grad(h, p, gw.capi.C.byref(v), gr, a4))
            gw.lib().gwz_sweep_destroy(h)
            acc += np.array(a4[:])
        assert np.all(np.abs(acc - first) <= 1e-12 * np.abs(first))
