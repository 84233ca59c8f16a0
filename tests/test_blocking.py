"""Error bars: resample! (src/kinetic-vqmc/state.jl:194-231) + blocking_analysis
(Rimu.StatsTools, used at state.jl:53,273) -- the product's host entry points, the oracle and the DEVICE kernel
(csrc/gwz_blocking.cu) against the independent numpy checker tests/blocking_checker.py, which is written from
the published Flyvbjerg-Petersen / Jonsson papers and shares no code with either."""
import math

import numpy as np
import pytest

import blocking_checker as bc


def _series_cases(rng, count, nmax=3000):
    for _ in range(count):
        n = int(rng.integers(2, nmax))
        phi = float(rng.choice([0.0, 0.3, 0.7, 0.95]))
        yield bc.ar1(rng, n, phi) * rng.uniform(0.1, 10) + rng.uniform(-50, 50)


def _same_result(a, ref, rtol=1e-10):
    assert a.k == ref["k"] and a.blocks == ref["blocks"], (a.k, ref["k"], a.blocks, ref["blocks"])
    assert abs(a.mean - ref["mean"]) <= 1e-12 * max(1.0, abs(ref["mean"]))
    if ref["k"] > 0:
        assert abs(a.err - ref["err"]) <= rtol * ref["err"] and abs(a.err_err - ref["err_err"]) <= rtol * ref["err_err"]
    else:
        assert math.isnan(a.err) and math.isnan(a.err_err)


# ---------------------------------------------------------------------------------------------- CPU: checker itself
def test_checker_follows_the_published_listing():
    """On power-of-two lengths the checker's level table reproduces the listing's s[k] (variance) and, for long
    levels, its M statistic; the accepted level is the listing's whenever the finite-n correction cannot matter."""
    rng = np.random.default_rng(5)
    same = total = 0
    for phi in (0.0, 0.6, 0.9):
        for _ in range(40):
            x = bc.ar1(rng, 1 << 13, phi)
            mu, var_mean, k, M = bc.jonsson_block(x)
            r = bc.blocking_mtest(x, corrected=False)
            lv = r["levels"]
            # same pair-averaged series: uncorrected variance of every level equals the listing's s[i]
            assert abs(r["mean"] - mu) < 1e-12
            n = lv["n"].astype(float)
            assert np.allclose(lv["se"] ** 2 * n, [np.var(x.reshape(len(x) >> i, -1).mean(axis=1)) for i in range(13)],
                               rtol=1e-10)
            # the listing's M_j = n_j (gamma_j / s_j)^2 is the large-n limit of the theorem's M_j
            mj_listing = M - np.append(M[1:], 0.0)  # M_k - M_{k+1}
            assert np.allclose(lv["mj"][:5], mj_listing[:5], rtol=0.1, atol=0.35)
            total += 1
            same += (r["k"] == k + 1)
            if r["k"] == k + 1 and r["k"] > 0:
                assert abs(r["err"] ** 2 - var_mean) <= 1e-10 * var_mean  # s[k] / 2^(d-k)
    assert same >= 0.6 * total


def test_checker_error_bars_on_ar1_series():
    """Known integrated autocorrelation time: the blocked error approaches the exact standard error of the mean from
    below (the M test stops as soon as the remaining correlation is not detectable), the naive one does not."""
    rng = np.random.default_rng(6)
    n = 1 << 14
    for phi, lo in ((0.0, 0.97), (0.5, 0.93), (0.8, 0.85)):
        exact = bc.ar1_error_of_mean(n, phi)
        ratios, hits = [], 0
        for _ in range(60):
            r = bc.blocking_mtest(bc.ar1(rng, n, phi))
            assert r["k"] > 0
            ratios.append(r["err"] / exact)
            hits += abs(r["err"] - exact) < 3 * r["err_err"]
        assert lo < np.mean(ratios) < 1.03, (phi, np.mean(ratios))
        assert hits >= 40, (phi, hits)
        if phi > 0:
            naive = math.sqrt(1.0 / (1 - phi * phi) / n)
            assert naive < 0.75 * exact


def test_resample_definition_and_reference_property():
    """resample! against its definition (slot averages of the step function) and test/runtests.jl:10-19."""
    rng = np.random.default_rng(0)
    t, v = rng.random(100) + 1e-3, rng.random(100)
    mu = np.sum(t * v) / np.sum(t)
    for k in (1, 2, 4, 8, 10, 16, 32, 64, 100, 200, 1000):
        assert abs(bc.resample_integral(t, v, k).mean() - mu) < 1e-12


# ---------------------------------------------------------------------------------------------- CPU: oracle, host ABI
def test_oracle_blocking_matches_the_independent_checker(oracle_mod):
    o = oracle_mod.Oracle(2, 2)
    rng = np.random.default_rng(7)
    for x in _series_cases(rng, 120):
        _same_result(o.blocking_analysis(x), bc.blocking_mtest(x))
    one = o.blocking_analysis(np.array([3.0]))
    assert one.k == 0 and one.mean == 3.0 and one.blocks == 1
    const = o.blocking_analysis(np.full(64, 2.5))  # zero variance: M_j = NaN, the test cannot succeed
    assert const.k == -1 and math.isnan(const.err) and const.mean == 2.5


def test_host_entry_points_match_checker_and_oracle(gw, oracle_mod):
    o = oracle_mod.Oracle(2, 2)
    rng = np.random.default_rng(8)
    for x in _series_cases(rng, 120):
        _same_result(gw.blocking_analysis(x), bc.blocking_mtest(x))
    for n in (2, 3, 4, 5, 7, 8, 9):  # shortest series: 1 level (never accepted), 2 levels, odd tails
        x = rng.standard_normal(n)
        _same_result(gw.blocking_analysis(x), bc.blocking_mtest(x))
    t, v = rng.random(257) + 0.01, rng.standard_normal(257)
    for k in (1, 2, 7, 64, 257, 1000):
        got = gw.resample(t, v, len=k)
        assert np.array_equal(got, o.resample(t, v, k))            # the literal restatement: bit for bit
        assert np.allclose(got, bc.resample_integral(t, v, k), rtol=0, atol=1e-11)  # the definition
        assert abs(got.mean() - np.sum(t * v) / np.sum(t)) < 1e-12  # test/runtests.jl:10-19
    with pytest.raises(ValueError, match="do not match"):
        gw.resample(t, v[:-1])


# ---------------------------------------------------------------------------------------------- GPU: the device kernel
def _random_walk_series(rng, steps, n):
    """(tau, eloc) with the statistics of a kinetic walk: positive residence times with a long tail, energies that
    are AR(1)-correlated along each walker with a walker-dependent correlation time"""
    phi = rng.uniform(0.0, 0.95, size=n)
    e = np.empty((steps, n))
    e[0] = rng.standard_normal(n)
    for k in range(1, steps):
        e[k] = phi * e[k - 1] + rng.standard_normal(n)
    tau = rng.gamma(2.0, 0.5, size=(steps, n)) * np.exp(0.3 * e)  # residence time correlated with the energy
    return tau, -40.0 + 0.7 * e


@pytest.mark.gpu
@pytest.mark.parametrize("steps,n,length", [(64, 20000, None), (95, 4096, None), (100, 2048, 64), (64, 1500, 256),
                                            (1000, 300, None), (4099, 64, None), (2, 200, None), (3, 200, None),
                                            (1, 64, None), (5, 100, 1)])
def test_device_blocking_kernel_matches_checker(gw, oracle_mod, steps, n, length):
    """gwz_series_blocking (one walker per thread) against the numpy checker for every walker: k and block count
    exactly, mean / err / err_err to 1e-10; the resampled series it analyses is the oracle's (same arithmetic)."""
    o = oracle_mod.Oracle(2, 2)
    rng = np.random.default_rng(steps * 7919 + n)
    tau, eloc = _random_walk_series(rng, steps, n)
    L = steps if length is None else length
    summ, a = gw.series_blocking(tau, eloc, resample_len=length, resampled=L >= 2)
    check = range(n) if n <= 4096 else rng.choice(n, 4096, replace=False)
    near = 0
    ulp_off = 0
    for w in check:
        rs = o.resample(tau[:, w], eloc[:, w], L)
        if L >= 2:  # the device's equal-time series against the literal restatement of state.jl:194-231: bit for bit
            dev = a["resampled"][:, w]  # (its division by the total time is a Newton step, exact but for rare 1-ulp cases)
            if not np.array_equal(dev, rs):
                assert np.allclose(dev, rs, rtol=1e-13, atol=1e-13)
                ulp_off += 1
        # the definition; the algorithm's running time `t` carries ~ L eps of rounding, which the last slots see
        # multiplied by the (uncentred) values
        assert np.allclose(rs, bc.resample_integral(tau[:, w], eloc[:, w], L), rtol=0,
                           atol=1e-13 * max(L, steps) * np.abs(eloc[:, w]).max() + 1e-11)
        ref = bc.blocking_mtest(rs)
        if ref["k"] != a["k"][w]:  # an M_k within rounding of its quantile may fall on either side
            M, q = ref["M"], bc.Q99
            assert any(abs(M[j] - q[j]) < 1e-8 * q[j] for j in range(len(M) - 1)), (w, ref["k"], a["k"][w])
            near += 1
            continue
        assert a["blocks"][w] == ref["blocks"]
        assert abs(a["mean"][w] - ref["mean"]) < 1e-12 * abs(ref["mean"])
        if ref["k"] > 0:
            assert abs(a["err"][w] - ref["err"]) < 1e-10 * ref["err"]
            assert abs(a["err_err"][w] - ref["err_err"]) < 1e-10 * ref["err_err"]
        else:
            assert math.isnan(a["err"][w])
    assert near <= 2 and ulp_off <= max(1, len(check) // 100)
    # cross-walker combination (state.jl:150-164 / 265-285)
    ok = a["k"] > 0
    assert summ.walkers == n and summ.failed == int((~ok).sum())
    assert abs(summ.mean - a["mean"].mean()) < 1e-12 * abs(a["mean"].mean())
    if ok.any():
        assert abs(summ.err_ok - math.sqrt(np.sum(a["err"][ok] ** 2)) / n) < 1e-12 * summ.err_ok
        assert (summ.k_min, summ.k_max) == (a["k"][ok].min(), min(a["k"][ok].max(), 31))
    assert math.isnan(summ.err) == bool((~ok).any())


@pytest.mark.gpu
def test_device_blocking_of_the_walker_kernels_series(gw, oracle_mod):
    """End to end on the production kernel: the series kept in HBM (GWZ_RUN_SERIES), their blocking in the same call
    (GWZ_RUN_BLOCKING), against the checker run on the downloaded series; the estimator sums agree with the series."""
    H = gw.HubbardReal1D(gw.near_uniform(12, 12), u=2.0)
    model = H.model()
    n, steps = 12000, 200
    ws = gw.Walkers(model, n, H.address.words, seed=5)
    ws.run([0.5], 300)  # thermalise
    r = ws.run([0.5], steps, blocking=True)
    tau, eloc = ws.series()
    assert tau.shape == (steps, n) and np.all(tau > 0)
    b = r.blocking
    assert b.walkers == n and b.series_steps == steps and b.resample_len == steps
    val_w, grad_w = ws.estimates()
    assert np.allclose(val_w, (tau * eloc).sum(0) / tau.sum(0), rtol=1e-12)  # state.jl:42-49 from the kept series
    summ, a = ws.blocking(per_walker=True)
    assert (summ.mean, summ.err_ok, summ.failed) == (b.mean, b.err_ok, b.failed)
    o = oracle_mod.Oracle(2, 2)
    errs2, fails = 0.0, 0
    for w in range(0, n, 7):
        ref = bc.blocking_mtest(o.resample(tau[:, w], eloc[:, w], steps))
        assert ref["k"] == a["k"][w] and abs(ref["mean"] - a["mean"][w]) < 1e-12 * abs(ref["mean"])
        if ref["k"] > 0:
            assert abs(ref["err"] - a["err"][w]) < 1e-9 * ref["err"]
    ok = a["k"] > 0
    assert abs(b.mean - a["mean"].mean()) < 1e-12 * abs(b.mean)
    assert abs(b.err_ok - math.sqrt(np.sum(a["err"][ok] ** 2)) / n) < 1e-12
    # the blocking mean of a walker is its time-weighted mean (resample preserves it), so is the gradient
    assert np.allclose(a["mean"], val_w, rtol=1e-12)
    assert abs(b.grad[0] - grad_w.mean()) < 1e-9 * abs(grad_w.mean())
    # other resample lengths (local_energy_estimator(res; resample_length), state.jl:265-285)
    s64, a64 = ws.blocking(64, per_walker=True)
    for w in range(0, n, 97):
        ref = bc.blocking_mtest(o.resample(tau[:, w], eloc[:, w], 64))
        assert ref["k"] == a64["k"][w] and (ref["k"] < 0 or abs(ref["err"] - a64["err"][w]) < 1e-9 * ref["err"])
    # continuation appends to the series (kinetic_vqmc!(res; steps), qmc.jl:108-111)
    r2 = ws.run([0.5], 100, keep_acc=True, blocking=True)
    tau2, eloc2 = ws.series()
    assert tau2.shape == (300, n) and np.array_equal(tau2[:200], tau) and r2.blocking.series_steps == 300
    ref = bc.blocking_mtest(o.resample(tau2[:, 3], eloc2[:, 3], 300))
    _, a2 = ws.blocking(per_walker=True)
    assert ref["k"] == a2["k"][3] and abs(ref["err"] - a2["err"][3]) < 1e-9 * ref["err"]


@pytest.mark.gpu
def test_val_err_and_grad_uses_device_blocking_for_large_runs(gw):
    """val_err_and_grad(qmc, p) (wrapper.jl:98-105) on a run that is far too large to ship to the host returns the
    blocked error of state.jl:150-164, and amsgrad's err_tol stop (amsgrad.jl:204) acts on it."""
    H = gw.HubbardReal1D(gw.near_uniform(10, 10), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    qmc = gw.KineticVQMC(H, ans, samples=2 ** 16 * 256, walkers=2 ** 16, burn_in=100)
    val, err, grad = qmc.val_err_and_grad([0.5])
    b = qmc.result.blocking_summary
    assert b is not None and b.walkers == 2 ** 16 and b.series_steps == 256
    assert err == b.err_ok and b.failed == 0 and 1 <= b.k_min <= b.k_max <= 7
    exact, exact_grad = gw.LocalEnergyEvaluator(H, ans).val_and_grad([0.5])
    # every walker carries the O(tau_corr / steps) ratio bias of state.jl:42-49; at 256 steps it is below 1e-2
    assert abs(val - exact) < 1e-2 + 5 * err
    cross = qmc.result.last.err  # cross-walker standard error of the per-walker estimates
    assert 0.3 * cross < err < 3 * cross  # two estimates of the same uncertainty
    est = gw.local_energy_estimator(qmc.result)
    assert est.mean == val and est.err == err and est.err_err < err
    # amsgrad stops on the error bars when |dval| < err * err_tol
    res = gw.amsgrad(qmc, [0.6], maxiter=40, err_tol=1e9, val_tol=0.0, grad_tol=0.0, param_tol=0.0)
    assert res.converged and res.reason == "errorbars" and res.iterations == 2  # old_val = Inf at iteration 1
    assert res.error[0] > 0 and qmc.result.blocking_summary.walkers == 2 ** 16
    res = gw.amsgrad(qmc, [0.6], maxiter=5, err_tol=0.0)
    assert res.iterations == 6 and np.all(res.error > 0) and np.all(res.error < 1e-2)
