"""Pins part 2 of the CPU oracle (oracle/gutz_oracle_ansatz.c: the reference's other single-component
ansaetze and the ansatz-generic kinetic_sample! / LocalEnergyEvaluator) the way the reference's own tests
pin those ansaetze (test/ansatz.jl:7-41): finite-difference gradients stand in for ForwardDiff, a dense
Hamiltonian Rayleigh quotient for Rimu's `rayleigh_quotient`, exact diagonalisation for KrylovKit
(test/ansatz.jl:66-80), and part 1 of the oracle (bit-pinned to the README) for GutzwillerAnsatz."""
import numpy as np
import pytest

KINDS = ["gutzwiller", "ext_gutzwiller", "multinomial", "density_profile", "relative_jastrow", "jastrow",
         ("gutzwiller", "multinomial"),  # GutzwillerAnsatz(H) + MultinomialAnsatz(H) (test/ansatz.jl:58)
         ("relative_jastrow", "density_profile")]


def _params(ao, rng):
    return rng.random(ao.P)  # test/ansatz.jl:55-64 uses rand(P)


def _dense_h(o, ao, basis):
    index = {tuple(int(x) for x in b): i for i, b in enumerate(basis)}
    H = np.zeros((len(basis), len(basis)))
    for i, b in enumerate(basis):
        H[i, i] = ao.diagonal_element(b)
        for new, me in o.offdiagonals(b):
            H[index[tuple(int(x) for x in new)], i] += me
    return H


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("N,M,u,v,t", [(5, 5, 1.0, 0.0, 0.1), (5, 5, 1.0, 1.0, 0.1), (4, 6, 2.0, 0.5, 1.0), (6, 4, 0.7, 0.0, 1.0)])
def test_val_and_grad_and_local_energy(oracle_mod, kind, N, M, u, v, t):
    o = oracle_mod.Oracle(N, M, u, t)
    ao = oracle_mod.AnsatzOracle(o, kind, v=v)
    rng = np.random.default_rng(7)
    p = _params(ao, rng)
    basis = o.build_basis()
    picks = [basis[0]] + [basis[i] for i in rng.integers(0, len(basis), 10)]
    for addr in picks:
        val1 = ao.val(addr, p)
        val2, grad1 = ao.val_and_grad(addr, p)
        assert val1 == val2
        for k in range(ao.P):  # central differences for ForwardDiff.gradient (test/ansatz.jl:21)
            h = 1e-6
            pp, pm = p.copy(), p.copy()
            pp[k] += h
            pm[k] -= h
            fd = (ao.val(addr, pp) - ao.val(addr, pm)) / (2 * h)
            assert abs(fd - grad1[k]) <= 1e-8 * max(1.0, abs(grad1[k]))
    # LocalEnergyEvaluator value == Rayleigh quotient of the collected vector (test/ansatz.jl:37-38)
    val, grad = ao.lee_val_and_grad(basis, p)
    H = _dense_h(o, ao, basis)
    psi = np.array([ao.val(b, p) for b in basis])
    assert abs(val - psi @ H @ psi / (psi @ psi)) <= 1e-11 * max(1.0, abs(val))
    for k in range(min(ao.P, 6)):  # ForwardDiff.gradient(le, params) (test/ansatz.jl:32)
        h = 1e-6
        pp, pm = p.copy(), p.copy()
        pp[k] += h
        pm[k] -= h
        fd = (ao.lee_val_and_grad(basis, pp)[0] - ao.lee_val_and_grad(basis, pm)[0]) / (2 * h)
        assert abs(fd - grad[k]) <= 1e-7 * max(1.0, abs(grad[k]))


def test_gutzwiller_kind_equals_part_one(oracle_mod):
    o = oracle_mod.Oracle(10, 10, 2.0, 1.0)
    ao = oracle_mod.AnsatzOracle(o, "gutzwiller")
    basis = o.build_basis()
    rng = np.random.default_rng(3)
    for i in rng.integers(0, len(basis), 25):
        a = basis[i]
        assert ao.val_and_grad(a, [0.7]) == pytest.approx(o.ansatz_val_and_grad(a, 0.7), rel=0, abs=0)
        s1, s2 = o.kinetic_sample(a, 0.7, 0.37), ao.kinetic_sample(a, [0.7], 0.37)
        assert s1["tau"] == s2["tau"] and s1["eloc"] == s2["eloc"] and s1["chosen"] == s2["chosen"]
        assert s1["grad"] == s2["grad"][0] and np.array_equal(s1["new_addr"], s2["new_addr"])
        assert np.allclose(o.lee_terms(a, 0.7), ao.lee_terms(a, [0.7]), rtol=1e-15, atol=0)
    # README.md:94,104 through the generic path (long-double sums: compare at 1e-11)
    val, grad = ao.lee_val_and_grad(basis, [1.0])
    assert abs(val - -7.825819465047312) < 1e-10 and abs(grad[0] - 10.614147776143358) < 1e-10


@pytest.mark.parametrize("onr", [(1, 1, 1, 1, 1), (1, 2, 1, 1)])
def test_multinomial_is_exact_for_u0(oracle_mod, onr):
    """test/ansatz.jl:66-80: with u = 0 the normalised multinomial ansatz at p = 1/2 is the ground state."""
    N, M = sum(onr), len(onr)
    o = oracle_mod.Oracle(N, M, 0.0, 1.0)
    ao = oracle_mod.AnsatzOracle(o, "multinomial", normalize=True)
    basis = o.build_basis(o.from_onr(onr))
    H = _dense_h(o, ao, basis)
    w, vecs = np.linalg.eigh(H)
    val, _ = ao.lee_val_and_grad(basis, [0.5])
    assert abs(val - w[0]) < 1e-10
    gs = vecs[:, 0] * np.sign(vecs[0, 0])
    psi = np.array([ao.val(b, [0.5]) for b in basis])
    assert np.allclose(psi, gs, atol=1e-10)  # DVec(bin, [0.5]) ≈ res[2][1] (normalised)


def test_generic_walk_reaches_psi_squared(oracle_mod):
    """time-weighted visit frequencies of the generic CTMC approach |psi|^2 (test/qmc.jl:38-43 style)."""
    o = oracle_mod.Oracle(3, 4, 1.5, 1.0)
    ao = oracle_mod.AnsatzOracle(o, "relative_jastrow", v=0.3)
    p = [0.3, 0.1]
    basis = o.build_basis()
    addrs = np.tile(o.near_uniform(), (4, 1))
    r = ao.kinetic_vqmc(p, addrs, 20000, seed=5, series=True)
    exact, _ = ao.lee_val_and_grad(basis, p)
    assert abs(r["val"] - exact) < 0.05


def test_tuned_cpu_walk_follows_the_faithful_port(oracle_mod):
    """oracle.kinetic_vqmc_fast (closed-form ratios, the CPU upper bound timed in bench.py) walks the same
    trajectories as the literal restatement."""
    for N, M, u, g in [(10, 10, 2.0, 1.0), (20, 20, 4.0, 0.25), (7, 3, 1.0, 0.3), (50, 50, 2.0, 0.33)]:
        o = oracle_mod.Oracle(N, M, u, 1.0)
        start = np.tile(o.near_uniform(), (4, 1))
        a = o.kinetic_vqmc(g, start.copy(), 400, seed=3, walker0=2)
        b = o.kinetic_vqmc_fast(g, start.copy(), 400, seed=3, walker0=2)
        assert np.array_equal(a["addrs"], b["addrs"]) and a["hops"] == b["hops"]
        assert np.allclose(a["val_w"], b["val_w"], rtol=1e-11) and np.allclose(a["grad_w"], b["grad_w"], rtol=1e-8, atol=1e-9)


def test_oracle_reproduces_committed_ansatz_vectors(oracle_mod):
    """tests/golden/ansatz_vectors.json (made by tests/golden/make_golden_ansatz.py) pins part 2 of the oracle against drift."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ansatz_vectors.json")
    for c in json.load(open(path))["cases"]:
        o = oracle_mod.Oracle(c["N"], c["M"], c["u"], c["t"])
        kind = tuple(c["kind"]) if isinstance(c["kind"], list) else c["kind"]
        ao = oracle_mod.AnsatzOracle(o, kind, v=c["v"], normalize=c["normalize"])
        basis = o.build_basis()
        assert len(basis) == c["dim"]
        for r in c["addresses"]:
            a = np.pad(np.array(r["addr"], dtype=np.uint64), (0, 4 - o.W))
            val, grad = ao.val_and_grad(a, c["params"])
            ks = ao.kinetic_sample(a, c["params"], 0.37)
            assert val == r["val"] and np.array_equal(grad, r["grad"]) and ks["eloc"] == r["e_loc"] and ks["tau"] == r["tau"]
            assert ks["chosen"] == r["chosen"] and np.array_equal(ao.hop_rates(a, c["params"]), r["rates"])
        lv, lg = ao.lee_val_and_grad(basis, c["params"])
        assert lv == c["le_val"] and np.array_equal(lg, c["le_grad"])
