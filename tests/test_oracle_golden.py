"""CPU tests: pin the oracle against every literal value the reference prints for this path
(README.md:54,65,94,104,186,286; localenergy.jl:18-22) and against first-principles checks that
stand in for the reference's live cross-checks (ForwardDiff, rayleigh_quotient, eigsolve)."""
import json
import math
import os

import numpy as np
import pytest

from conftest import edge_onrs, random_onr

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "readme_vectors.json")))


@pytest.fixture(scope="module")
def o10(oracle_mod):
    return oracle_mod.Oracle(10, 10, 2.0, 1.0)


@pytest.fixture(scope="module")
def basis10(o10):
    return o10.build_basis()


def test_near_uniform_bitstring(o10):
    a = o10.near_uniform()
    assert int(a[0]) == 0b1010101010101010101  # BitString{19}: README.md:43
    assert o10.to_onr(a) == [1] * 10


def test_ansatz_at_uniform_address(o10):
    a = o10.near_uniform()
    assert o10.ansatz(a, 1.0) == GOLDEN["ansatz_uniform"]["val"]  # README.md:54
    val, der = o10.ansatz_val_and_grad(a, 1.0)
    assert val == 1.0 and der == 0.0 and math.copysign(1.0, der) == -1.0  # (1.0, [-0.0]) README.md:65


def test_basis_dimension(o10, basis10):
    assert len(basis10) == 92378 == o10.dimension()
    assert len({int(x) for x in basis10[:, 0]}) == 92378


def test_lee_readme_bit_for_bit(o10, basis10):
    g = GOLDEN["lee_bose10"]
    assert o10.lee_val(basis10, 1.0) == g["val"]  # README.md:94
    val, grad = o10.lee_val_and_grad(basis10, 1.0)
    assert val == g["val"] and grad == g["grad"]  # README.md:104


def test_lee_minimum(o10, basis10):
    g = GOLDEN["lbfgs_minimum"]
    val, grad = o10.lee_val_and_grad(basis10, g["param"])
    assert abs(val - g["val"]) < 5e-14 * abs(g["val"])  # README.md:186
    assert abs(grad) < 1e-7


def test_lee_docstring_values(oracle_mod):
    o = oracle_mod.Oracle(5, 5, 1.0, 1.0)
    b = o.build_basis()
    g = GOLDEN["lee_docstring_bose5"]
    val, grad = o.lee_val_and_grad(b, 0.5)
    assert abs(val - g["val"]) < 1e-14 * abs(val)  # localenergy.jl:18-22 (never executed by CI)
    assert abs(grad - g["grad"]) < 1e-12


def test_summation_modes_agree(o10, basis10):
    v0, g0 = o10.lee_val_and_grad(basis10, 1.0, mode=0)
    v1, g1 = o10.lee_val_and_grad(basis10, 1.0, mode=1)
    v2, g2 = o10.lee_val_and_grad(basis10, 1.0, mode=2)
    assert abs(v0 - v1) < 5e-12 * abs(v1) and abs(v2 - v1) < 5e-12 * abs(v1)
    assert abs(g0 - g1) < 1e-11 * abs(g1) and abs(g2 - g1) < 1e-11 * abs(g1)


def test_kinetic_sample_hand_check(o10):
    # SURVEY A.4 hand check: |1...1>, u=2, g=1: E = -20*sqrt(2)*e^-2, tau = e^2/20, G = -0.0
    ks = o10.kinetic_sample(o10.near_uniform(), 1.0, 0.3)
    assert ks["n_hops"] == 20
    assert abs(ks["eloc"] - (-20 * math.sqrt(2) * math.exp(-2))) < 1e-14
    assert abs(ks["tau"] - math.exp(2) / 20) < 1e-15
    assert ks["grad"] == 0.0
    assert ks["chosen"] == 7  # first i with 0.3*S < cumsum_i


def _occupation_level_hops(onr, t):
    """Independent model of offdiagonals (App. A.4) on occupation vectors."""
    M = len(onr)
    out = []
    for i in range(M):
        if onr[i] == 0:
            continue
        for d in (+1, -1):
            j = (i + d) % M
            new = list(onr)
            new[i] -= 1
            new[j] += 1
            out.append((new, -t * math.sqrt(onr[i] * (onr[j] + 1))))
    return out


@pytest.mark.parametrize("N,M", [(5, 5), (10, 10), (7, 3), (3, 9), (20, 20), (50, 50), (100, 100), (33, 32), (64, 65)])
def test_offdiagonals_match_occupation_model(oracle_mod, N, M):
    o = oracle_mod.Oracle(N, M, 1.5, 0.7)
    rng = np.random.default_rng(N * 1000 + M)
    onrs = edge_onrs(N, M) + [random_onr(rng, N, M, c) for c in (0.1, 1.0, 10.0) for _ in range(4)]
    for onr in onrs:
        a = o.from_onr(onr)
        assert o.to_onr(a) == list(onr)
        ref = _occupation_level_hops(list(onr), 0.7)
        got = o.offdiagonals(a)
        assert len(got) == len(ref) == o.num_offdiagonals(a)
        for (aw, me), (new, mref) in zip(got, ref):
            assert o.to_onr(aw) == new
            assert abs(me - mref) < 1e-15 * max(1.0, abs(mref))
        d = 1.5 * sum(n * (n - 1) for n in onr) / 2
        assert o.diagonal_element(a) == d


def test_rayleigh_quotient_dense_matrix(oracle_mod):
    # stands in for test/ansatz.jl:37 (rayleigh_quotient) and :21,32 (ForwardDiff)
    o = oracle_mod.Oracle(5, 5, 1.0, 0.1)  # HubbardReal1D(near_uniform(BoseFS{5,5}); t=0.1), test/ansatz.jl:45
    basis = o.build_basis()
    index = {int(b[0]): i for i, b in enumerate(basis)}
    dim = len(basis)
    H = np.zeros((dim, dim))
    for i, b in enumerate(basis):
        H[i, i] = o.diagonal_element(b)
        for aw, me in o.offdiagonals(b):
            H[index[int(aw[0])], i] += me
    assert np.allclose(H, H.T)
    for g in (0.0, 0.37, 1.0):
        psi = np.array([o.ansatz(b, g) for b in basis])
        rq = psi @ H @ psi / (psi @ psi)
        val, grad = o.lee_val_and_grad(basis, g)
        assert abs(val - rq) < 1e-13 * max(1, abs(rq))
        h = 1e-6
        fd = (o.lee_val(basis, g + h) - o.lee_val(basis, g - h)) / (2 * h)
        assert abs(grad - fd) < 1e-7 * max(1, abs(fd))


def test_philox_known_answers(oracle_mod):
    import ctypes as C
    L = oracle_mod.lib()
    for kat in GOLDEN["philox4x32_10"]:
        c = (C.c_uint32 * 4)(*[int(x, 16) for x in kat["ctr"]])
        k = (C.c_uint32 * 2)(*[int(x, 16) for x in kat["key"]])
        out = (C.c_uint32 * 4)()
        L.go_philox4x32_10(c, k, out)
        assert [int(x) for x in out] == [int(x, 16) for x in kat["out"]]


def test_u01_range_and_stream(o10):
    us = np.array([o10.u01(12345, 7, s) for s in range(2000)])
    assert (us >= 0).all() and (us < 1).all()
    assert abs(us.mean() - 0.5) < 0.03
    assert o10.u01(12345, 7, 10) != o10.u01(12345, 8, 10) != o10.u01(12346, 7, 10)


def test_resample_preserves_weighted_mean(oracle_mod):
    # test/runtests.jl:10-19
    o = oracle_mod.Oracle(2, 2)
    rng = np.random.default_rng(0)
    times, vals = rng.random(100), rng.random(100)
    mu = np.sum(times * vals) / np.sum(times)
    for k in (1, 2, 4, 8, 10, 16, 32, 64, 100, 200, 1000):
        assert abs(o.resample(times, vals, k).mean() - mu) < 1e-12
    with pytest.raises(ValueError):
        o.resample(times, vals[:-1])


def test_blocking_on_correlated_series(oracle_mod):
    o = oracle_mod.Oracle(2, 2)
    rng = np.random.default_rng(1)
    n = 1 << 14
    x = np.zeros(n)
    for i in range(1, n):
        x[i] = 0.9 * x[i - 1] + rng.standard_normal()
    r = o.blocking_analysis(x)
    naive = x.std(ddof=1) / math.sqrt(n)
    true = math.sqrt((1 / (1 - 0.81)) * (1 + 0.9) / (1 - 0.9) / n)
    assert r.k >= 3
    assert r.err > 2.5 * naive and abs(r.err - true) < 0.5 * true
    w = rng.standard_normal(4096)
    rw = o.blocking_analysis(w)
    # white noise: whatever level the M test accepts (its quantile is indexed by the level, so k = 1 is accepted only
    # when M_1 < 6.63), the blocked error is the naive one within its own uncertainty
    assert 1 <= rw.k <= 6 and abs(rw.err - w.std(ddof=1) / 64) < 4 * rw.err_err


def test_vqmc_statistics_vs_exact(o10, basis10):
    # README.md:218: 1 walker, 1e5 steps, g=1: -7.8369 +- 0.0166; exact -7.8258
    exact = GOLDEN["lee_bose10"]["val"]
    addrs = np.stack([o10.near_uniform()] * 8)
    out = o10.kinetic_vqmc(1.0, addrs, 20000, seed=17, series=True)
    errs = []
    for w in range(8):
        b = o10.blocking_analysis(o10.resample(out["series_tau"][w], out["series_eloc"][w]))
        errs.append(b.err)
    err = math.sqrt(sum(e * e for e in errs)) / 8
    assert 0.005 < err < 0.03
    assert abs(out["val"] - exact) < 4 * err
    # gradient: compare with the exact sweep gradient
    assert abs(out["grad"] - GOLDEN["lee_bose10"]["grad"]) < 1.5


def test_vqmc_continuation_is_deterministic(o10):
    a1 = np.stack([o10.near_uniform()] * 3)
    full = o10.kinetic_vqmc(1.0, a1.copy(), 200, seed=5, series=True)
    a2 = a1.copy()
    first = o10.kinetic_vqmc(1.0, a2, 120, seed=5, series=True)
    second = o10.kinetic_vqmc(1.0, first["addrs"], 80, seed=5, step0=120, series=True)
    assert np.array_equal(full["addrs"], second["addrs"])
    assert np.array_equal(full["series_eloc"][:, 120:], second["series_eloc"])


def test_amsgrad_update_rule_matches_readme_row1():
    # README.md:286: row 1 = [1.0] -7.86947 [10.6335] [10.6726] [1.13072] [-0.100367]
    r = GOLDEN["amsgrad_row1"]
    m2 = 0.01 * r["gradient"] ** 2
    assert abs(m2 - r["second_moment"]) < 1.2e-5  # gradient printed to 6 digits
    dp = -0.01 * r["first_moment"] / math.sqrt(m2)
    assert abs(dp - r["param_delta"]) < 1.5e-6  # 6 printed digits


def test_amsgrad_oracle_on_sweep(oracle_mod):
    # test/amsgrad.jl:7-21 shape: BoseFS(1,1,1,1), u=t=1; amsgrad and gradient_descent converge to the minimiser
    o = oracle_mod.Oracle(4, 4, 1.0, 1.0)
    basis = o.build_basis()
    gd = o.amsgrad_lee(basis, 0.5, maxiter=100, use_amsgrad=False)
    ams = o.amsgrad_lee(basis, 0.5, maxiter=1000)
    assert gd["converged"] and ams["converged"]
    # minimiser by golden-section on the oracle value
    lo, hi = 0.0, 2.0
    for _ in range(80):
        m1, m2 = lo + 0.382 * (hi - lo), lo + 0.618 * (hi - lo)
        if o.lee_val(basis, m1) < o.lee_val(basis, m2):
            hi = m2
        else:
            lo = m1
    pmin = 0.5 * (lo + hi)
    assert abs(gd["param"][-1] - pmin) < 1e-3 and abs(ams["param"][-1] - pmin) < 1e-3
    # continuation with moments follows the uninterrupted run closely (test/amsgrad.jl:52-77: atol=1e-2;
    # it is not exact because the last parameter is evaluated twice)
    a = o.amsgrad_lee(basis, 0.5, maxiter=50, early_stop=False)
    b = o.amsgrad_lee(basis, a["param"][-1], maxiter=20, early_stop=False, m1_init=a["first_moment"][-1],
                      m2_init=a["second_moment"][-1])
    full = o.amsgrad_lee(basis, 0.5, maxiter=70, early_stop=False)
    assert np.allclose(full["param"][50:], b["param"], rtol=0, atol=1e-2)
    c = o.amsgrad_lee(basis, a["param"][-1], maxiter=20, early_stop=False)  # without the moments: wobbles
    assert np.abs(full["param"][50:] - c["param"]).max() > np.abs(full["param"][50:] - b["param"]).max()
    # maxiter+1 evaluations (Appendix B.4)
    assert full["iterations"] == 71
