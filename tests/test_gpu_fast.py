"""Parity of the round-2 production walker kernel (csrc/gwz_fast.cuh, gwz_walker_fast.cu) through its own parity
surface, gwz_vqmc_run_trace: per step the residence time, the local energy and the selected hop of the kernel that
is benchmarked -- re-evaluated per address by the CPU oracle at 1e-12 (kinetic_sample!, src/kinetic-vqmc/qmc.jl:14-44)
-- and the measure of its move selection (qmc.jl:40, 52-61) with prescribed uniforms."""
import numpy as np
import pytest

from conftest import edge_onrs, random_onr

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _apply(onr, hop):
    """hop = 2 z + right on the occupation vector (in place); returns the reference hop index (1-based, Rimu
    offdiagonals order: occupied modes ascending, right hop then left hop)."""
    M = len(onr)
    z, right = hop >> 1, hop & 1
    zn = (z + 1) % M
    src, dst = (z, zn) if right else (zn, z)
    assert onr[src] > 0, "hop from an empty mode"
    rank = int(np.count_nonzero(onr[: src + 1]))  # 1-based rank of the source among the occupied modes
    onr[src] -= 1
    onr[dst] += 1
    return 2 * rank - 1 if right else 2 * rank


CASES = [(10, 10, 2.0, 1.0), (50, 50, 2.0, 1.0), (50, 50, 2.0, 0.33), (100, 100, 2.0, 0.33), (20, 20, 4.0, 1.0),
         (9, 6, 0.5, 0.05), (7, 4, 1.0, 0.3), (33, 32, 2.0, 0.2), (64, 65, 1.0, 0.1), (120, 131, 2.0, 0.1),
         (17, 40, 3.0, 0.2), (40, 8, 0.3, 0.02), (12, 5, 6.0, 1.5), (6, 240, 1.0, 0.5), (30, 200, 2.0, 0.3),
         (1, 6, 1.0, 0.5), (2, 4, 3.0, 0.7), (200, 40, 0.05, 0.01)]


@pytest.mark.parametrize("N,M,u,g", CASES)
def test_trace_of_the_production_kernel_matches_the_oracle_step_by_step(gw, oracle_mod, N, M, u, g):
    o = oracle_mod.Oracle(N, M, u, 1.0)
    model = gw.Model(N, M, u, 1.0)
    crowd = [0] * M
    crowd[M - 1] = N  # all bosons on the last site: generic bonds (occupation >= 4), the periodic bond, the ripple
    half = [0] * M
    half[0], half[M // 2] = N // 2, N - N // 2
    nw, steps = 48, 150
    for start_onr in (None, crowd, half):
        start = o.near_uniform() if start_onr is None else o.from_onr(start_onr)
        w = gw.Walkers(model, nw, start[: o.W], seed=31, global_offset=5)
        r, tau, eloc, hop = w.run_trace([g], steps)
        assert np.all(hop >= 0) and np.all(hop < 2 * M) and np.all(tau > 0)
        onr0 = np.array(o.to_onr(start))
        sum_t = np.zeros(nw)
        sum_te = np.zeros(nw)
        for k in range(nw):
            onr = onr0.copy()
            for s in range(steps):
                if k % 6 == 0 or s < 3 or s >= steps - 2:  # every step of every 6th walker, the ends of the others
                    addr = o.from_onr(onr)
                    D = o.diagonal_element(addr)
                    # the reference works with absolute amplitudes exp(-g D): beyond g D ~ 600 they are denormal or 0
                    # and the oracle is no longer accurate (the kernel's ratios have no such limit)
                    if g * D <= 600:
                        ks = o.kinetic_sample(addr, g, 0.5)
                        assert abs(tau[s, k] - ks["tau"]) <= RTOL * ks["tau"], (k, s)
                        # E_loc = D - t E_mat: 1e-12 relative to the larger of its two terms
                        scale = max(abs(ks["eloc"]), D, abs(ks["eloc"] - D))
                        assert abs(eloc[s, k] - ks["eloc"]) <= RTOL * scale + 1e-12, (k, s, eloc[s, k], ks["eloc"])
                _apply(onr, int(hop[s, k]))
            assert np.array_equal(o.from_onr(onr)[: o.W], w.addresses()[k]), k  # the trace ends where the walker is
            sum_t[k], sum_te[k] = tau[:, k].sum(), (tau[:, k] * eloc[:, k]).sum()
        val_w, _ = w.estimates()
        # the running sums are the sums of the trace.  (They are kept relative to a reference point -- E_loc of the
        # start address for a walker set's first estimator -- so only the physical start state is checked: the
        # crowded ones have |E_loc| up to e^(c N), and the sums relative to that are meaningless by design.)
        if start_onr is None:
            dev = np.abs(val_w - sum_te / sum_t) / np.abs(val_w)
            assert dev.max() < 1e-12, (dev.max(), int(dev.argmax()))
            assert abs(r.val - val_w.mean()) <= 1e-12 * abs(r.val)
        # continuation: the carried integers are rebuilt per launch; a second launch continues the same walk
        r2, tau2, eloc2, hop2 = w.run_trace([g], 20, keep_acc=True)
        onr = np.array(o.to_onr(np.pad(w.addresses()[0], (0, 4 - o.W))))
        # (replay walker 0 backwards is not possible; check its last recorded step against its final address instead)
        w_chk = gw.Walkers(model, nw, start[: o.W], seed=31, global_offset=5)
        _, tau_all, eloc_all, hop_all = w_chk.run_trace([g], steps + 20)
        assert np.array_equal(hop_all[:steps], hop) and np.array_equal(hop_all[steps:], hop2)
        assert np.array_equal(tau_all[steps:], tau2) and np.array_equal(w_chk.addresses(), w.addresses())


@pytest.mark.parametrize("N,M,u,g", [(10, 10, 2.0, 1.0), (50, 50, 2.0, 0.33), (100, 100, 2.0, 1.0), (6, 4, 1.0, 0.2),
                                     (20, 20, 4.0, 0.25), (12, 5, 6.0, 1.5)])
def test_production_selection_maps_u_onto_hops_with_measure_rate_over_S(gw, oracle_mod, N, M, u, g):
    """qmc.jl:40: the next address is drawn with probability |psi_2| / sum |psi_2|.  Stratified u: every hop owns a few
    intervals of [0, 1), so its share of n_u equally spaced uniforms is its probability to within a few cells."""
    o = oracle_mod.Oracle(N, M, u, 1.0)
    model = gw.Model(N, M, u, 1.0)
    rng = np.random.default_rng(5)
    n_u = 8192
    us = ((np.arange(n_u) + 0.5) / n_u).reshape(1, n_u)
    for onr in edge_onrs(N, M)[3:6] + [random_onr(rng, N, M, 0.2), random_onr(rng, N, M, 2.0)]:
        onr = np.array(onr)
        a = o.from_onr(onr)
        rates, _ = o.hop_rates(a, g)
        if not np.all(np.isfinite(rates)) or g * o.diagonal_element(a) > 600:
            continue
        w = gw.Walkers(model, n_u, a[: o.W], seed=1)
        _, tau, eloc, hop = w.run_trace([g], 1, u01=us)
        ref = np.array([_apply(onr.copy(), int(h)) for h in hop[0]])
        hist = np.bincount(ref, minlength=len(rates) + 1)[1:]
        assert hist.sum() == n_u
        expect = rates / rates.sum()
        assert np.max(np.abs(hist / n_u - expect)) <= 2.5 / n_u + 1e-12, (onr, hist / n_u - expect)
        ks = o.kinetic_sample(a, g, 0.5)
        assert np.all(np.abs(tau[0] - ks["tau"]) <= RTOL * ks["tau"]) and np.all(np.abs(eloc[0] - ks["eloc"]) <=
                                                                              RTOL * abs(ks["eloc"]) + 1e-12)


def test_fast_and_incremental_kernels_sample_the_same_distribution(gw, monkeypatch):
    """round-1 kernel (GWZ_FAST=0, bitwise equal to the unary kernel that follows the oracle) against the round-2
    kernel on the same system: different u -> hop maps, same energies within the cross-walker errors; and both
    against the exact sweep value."""
    H = gw.HubbardReal1D(gw.near_uniform(12, 12), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    exact, grad = gw.LocalEnergyEvaluator(H, ans).val_and_grad([0.4])
    vals = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("GWZ_FAST", flag)
        res = gw.kinetic_vqmc(H, ans, [0.4], walkers=1 << 14, steps=4000, burn_in=500, seed=77, record=False,
                              estimator="pooled")
        vals[flag] = (res.last.val, res.last.err, res.last.grad[0], res.last.grad_err[0])
    (v1, e1, g1, ge1), (v0, e0, g0, ge0) = vals["1"], vals["0"]
    assert abs(v1 - v0) < 4.5 * np.hypot(e1, e0) and abs(g1 - g0) < 4.5 * np.hypot(ge1, ge0)
    assert abs(v1 - exact) < 4.5 * e1 and abs(g1 - grad[0]) < 4.5 * ge1


@pytest.mark.parametrize("N,M,u,g", [(50, 50, 2.0, 1.0), (50, 50, 2.0, 0.33), (24, 24, 6.0, 1.2)])
def test_carried_local_energy_does_not_drift_over_long_runs(gw, oracle_mod, N, M, u, g):
    """E (the off-diagonal part of E_loc) is carried by table deltas and re-summed every FAST_REFRESH = 64 steps or when
    it has fallen 16-fold below its peak (gwz_fast.cuh): after thousands of steps every recorded E_loc and residence
    time still equals the oracle's re-evaluation of the address at 1e-12."""
    o = oracle_mod.Oracle(N, M, u, 1.0)
    model = gw.Model(N, M, u, 1.0)
    start = o.near_uniform()
    nw, steps = 8, 3000
    w = gw.Walkers(model, nw, start[: o.W], seed=5)
    _, tau, eloc, hop = w.run_trace([g], steps)
    onr0 = np.array(o.to_onr(start))
    for k in range(nw):
        onr = onr0.copy()
        for s in range(steps):
            if s % 61 == 0 or s >= steps - 3:  # every phase of the refresh period, and the end of the run
                addr = o.from_onr(onr)
                D = o.diagonal_element(addr)
                if g * D <= 600:
                    ks = o.kinetic_sample(addr, g, 0.5)
                    assert abs(tau[s, k] - ks["tau"]) <= RTOL * ks["tau"], (k, s)
                    scale = max(abs(ks["eloc"]), D, abs(ks["eloc"] - D))
                    assert abs(eloc[s, k] - ks["eloc"]) <= RTOL * scale + 1e-12, (k, s)
            _apply(onr, int(hop[s, k]))
        assert np.array_equal(o.from_onr(onr)[: o.W], w.addresses()[k]), k
