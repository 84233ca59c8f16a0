"""GPU parity: per-address local energies, residence times, gradient ratios, transition rates and
move selection against the CPU oracle (north star: 1e-12 relative in FP64)."""
import numpy as np
import pytest

from conftest import edge_onrs, random_onr

pytestmark = pytest.mark.gpu
RTOL = 1e-12  # the tolerance BASELINE.json's north_star states for per-address quantities

CASES = [  # (N, M, u, t, g): 32-bit word counts 1..8, word-boundary straddles included
    (10, 10, 2.0, 1.0, 1.0),
    (5, 5, 1.0, 0.1, 0.37),
    (12, 12, 2.0, 1.0, 0.3331889106855026),
    (20, 20, 4.0, 1.0, 0.25),
    (3, 9, 1.0, 1.0, 0.5),
    (7, 3, 1.5, 0.7, 0.4),
    (2, 2, 1.0, 1.0, 0.3),
    (33, 32, 2.0, 1.0, 0.2),
    (50, 50, 2.0, 1.0, 1.0),
    (50, 50, 2.0, 1.0, 0.33),
    (64, 65, 1.0, 1.0, 0.1),
    (80, 70, 2.0, 1.0, 0.05),
    (100, 100, 2.0, 1.0, 0.33),
    (120, 131, 2.0, 0.5, 0.02),
]


def _addresses(o, N, M, seed, n_random=60):
    rng = np.random.default_rng(seed)
    onrs = edge_onrs(N, M)
    for conc in (0.05, 0.3, 1.0, 5.0, 50.0):
        onrs += [random_onr(rng, N, M, conc) for _ in range(n_random // 5)]
    return onrs, np.stack([o.from_onr(x) for x in onrs])


def _rel(a, b):
    return abs(a - b) / max(1e-300, abs(b))


@pytest.mark.parametrize("N,M,u,t,g", CASES)
def test_probe_matches_oracle(gw, oracle_mod, N, M, u, t, g):
    o = oracle_mod.Oracle(N, M, u, t)
    model = gw.Model(N, M, u, t)
    assert model.words == o.W and model.bits == o.B
    onrs, addrs = _addresses(o, N, M, seed=N * 131 + M)
    rng = np.random.default_rng(7)
    u01 = rng.random(len(addrs))
    enum = model.probe(addrs[:, :o.W], [g], u01, kernel=gw.KERNEL_ENUMERATE, want_rates=True)
    fast = model.probe(addrs[:, :o.W], [g], u01, kernel=gw.KERNEL_BITPARALLEL)
    for i, a in enumerate(addrs):
        # overflow guard: the reference works with absolute psi = exp(-g*D); skip addresses where it underflows
        if g * o.diagonal_element(a) > 600:
            continue
        ks = o.kinetic_sample(a, g, float(u01[i]))
        rates, _ = o.hop_rates(a, g)
        for out in (enum, fast):
            assert out["n_hops"][i] == ks["n_hops"]
            assert _rel(out["e_loc"][i], ks["eloc"]) < RTOL or abs(out["e_loc"][i] - ks["eloc"]) < 1e-12
            assert _rel(out["residence_time"][i], ks["tau"]) < RTOL
            assert out["grad_ratio"][i, 0] == pytest.approx(ks["grad"], rel=RTOL, abs=0)
        # transition rates in the reference's hop order
        got = enum["rates"][i, : ks["n_hops"]]
        assert np.all(np.abs(got - rates) <= RTOL * rates)
        assert np.all(enum["rates"][i, ks["n_hops"]:] == 0)
        # reference selection rule: same hop, same next address (up to cumsum ties at 1e-16)
        cs = ks["cumsum"]
        thr = u01[i] * cs[-1]
        margin = np.min(np.abs(cs - thr)) / cs[-1]
        if margin > 1e-12:
            assert enum["chosen"][i] == ks["chosen"]
            assert np.array_equal(enum["next_addr"][i], ks["new_addr"][: o.W])
        # production selection: a valid hop of this address
        nxt = np.zeros(4, dtype=np.uint64)
        nxt[: o.W] = fast["next_addr"][i]
        c = int(fast["chosen"][i])
        assert 1 <= c <= ks["n_hops"]
        ref_next, _ = o.get_offdiagonal(a, c)
        assert np.array_equal(nxt, ref_next)


@pytest.mark.parametrize("N,M,u,g", [(10, 10, 2.0, 1.0), (50, 50, 2.0, 0.33), (100, 100, 2.0, 1.0), (6, 4, 1.0, 0.2)])
def test_production_selection_samples_the_rates(gw, oracle_mod, N, M, u, g):
    """The class-ordered move selection must map u in [0,1) onto hops with measure rate_i / S
    (qmc.jl:40: chosen ~ |psi2|).  Stratified u => deterministic check to 2/n."""
    o = oracle_mod.Oracle(N, M, u, 1.0)
    model = gw.Model(N, M, u, 1.0)
    rng = np.random.default_rng(5)
    n_u = 4096
    for onr in edge_onrs(N, M)[3:6] + [random_onr(rng, N, M, 0.2), random_onr(rng, N, M, 2.0)]:
        a = o.from_onr(onr)
        rates, _ = o.hop_rates(a, g)
        if not np.all(np.isfinite(rates)) or g * o.diagonal_element(a) > 600:
            continue
        us = (np.arange(n_u) + 0.5) / n_u
        out = model.probe(np.tile(a[: o.W], (n_u, 1)), [g], us, kernel=gw.KERNEL_BITPARALLEL)
        hist = np.bincount(out["chosen"], minlength=len(rates) + 1)[1:]
        assert hist.sum() == n_u
        expect = rates / rates.sum()
        # each hop owns one interval (or a few, never more than its class count): error <= 2 cells each
        assert np.max(np.abs(hist / n_u - expect)) <= 2.5 / n_u + 1e-12
        out_e = model.probe(np.tile(a[: o.W], (n_u, 1)), [g], us, kernel=gw.KERNEL_ENUMERATE)
        hist_e = np.bincount(out_e["chosen"], minlength=len(rates) + 1)[1:]
        assert np.max(np.abs(hist_e / n_u - expect)) <= 1.5 / n_u + 1e-12


def test_ansatz_val_and_grad(gw, oracle_mod):
    # test/ansatz.jl:16-26: value, analytic gradient; README.md:54,65
    H = gw.HubbardReal1D(gw.near_uniform(10, 10), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    assert ans(gw.near_uniform(10, 10), [1.0]) == 1.0
    val, grad = gw.val_and_grad(ans, gw.near_uniform(10, 10), [1.0])
    assert val == 1.0 and grad[0] == 0.0 and np.signbit(grad[0])  # (1.0, [-0.0])
    o = oracle_mod.Oracle(10, 10, 2.0, 1.0)
    rng = np.random.default_rng(2)
    for _ in range(10):
        onr = random_onr(rng, 10, 10, 1.0)
        v, d = o.ansatz_val_and_grad(o.from_onr(onr), 0.61)
        gv, gd = ans.val_and_grad(gw.BoseFS(onr), 0.61)
        assert abs(gv - v) <= 1e-14 * abs(v) and abs(gd[0] - d) <= 1e-14 * abs(d) + 1e-300
        assert ans(gw.BoseFS(onr), 0.61) == gv


def test_probe_edge_inputs(gw):
    model = gw.Model(4, 4, 1.0, 1.0)
    out = model.probe(np.zeros((0, 1), dtype=np.uint64), [0.5])
    assert out["e_loc"].shape == (0,)
    with pytest.raises(ValueError):
        model.probe(np.zeros((1, 1), dtype=np.uint64), [0.5, 0.1])
    with pytest.raises(ValueError):
        gw.Model(0, 4, 1.0, 1.0)
    with pytest.raises(ValueError):
        gw.Model(200, 200, 1.0, 1.0)
