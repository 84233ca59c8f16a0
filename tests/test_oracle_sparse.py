"""CPU tests of the stored-operator oracle (oracle/gutz_oracle_sparse.c): it must agree bit for bit with the pinned
HubbardReal1D oracle on rows generated from it, and behave like the reference's tests demand on other operators
(test/qmc.jl:16-50: eigenvector ansatz => E_loc == E0 and |psi|^2 sampling; test/ansatz.jl:37 Rayleigh quotient)."""
import numpy as np
import pytest

from oracle.oracle import Oracle, SparseOracle
import sparse_models as sm


@pytest.fixture(scope="module")
def real55():
    o = Oracle(5, 5, 1.0, 0.1)  # HubbardReal1D(near_uniform(BoseFS{5,5}); t=0.1), test/qmc.jl:11
    rp, col, mel, diag, basis = sm.hubbard_real1d(o)
    return o, SparseOracle(rp, col, mel, diag), basis


def test_rows_match_pinned_oracle_bitwise(real55):
    o, so, basis = real55
    g = 0.5
    so.set_gutzwiller(g)
    assert so.dim == 126 == len(basis)
    rng = np.random.default_rng(1)
    for k in rng.integers(0, so.dim, size=40):
        u = float(rng.random())
        a, b = o.kinetic_sample(basis[k], g, u), so.kinetic_sample(int(k), u)
        # psi_l / psi_k is exp(-g D_l) / exp(-g D_k) here and exp(-g (D_l - D_k)) nowhere: same literal formula
        assert b["n_hops"] == a["n_hops"] and b["chosen"] == a["chosen"]
        assert np.array_equal(basis[b["new_index"]], a["new_addr"])
        assert b["tau"] == a["tau"] and b["eloc"] == a["eloc"] and b["grad"][0] == a["grad"]
        assert np.array_equal(so.lee_terms(int(k)), o.lee_terms(basis[k], g))


def test_trajectories_match_pinned_oracle(real55):
    o, so, basis = real55
    g = 0.3
    so.set_gutzwiller(g)
    ref = o.kinetic_vqmc(g, np.stack([basis[0]] * 6), 300, seed=11)
    got = so.kinetic_vqmc(np.zeros(6, dtype=np.uint32), 300, seed=11)
    assert np.array_equal(basis[got["idx"]], ref["addrs"])
    assert np.allclose(got["val_w"], ref["val_w"], rtol=1e-13, atol=0)
    assert np.allclose(got["grad_w"][:, 0], ref["grad_w"], rtol=1e-11, atol=1e-13)
    v0, g0 = o.lee_val_and_grad(basis, g, mode=1)
    v1, g1 = so.lee_val_and_grad()
    assert abs(v1 - v0) <= 1e-14 * abs(v0) and abs(g1[0] - g0) <= 1e-13 * abs(g0)


def test_operators_are_hermitian_and_momentum_space_matches_real_space():
    o = Oracle(3, 5, 2.0, 1.0)
    rp, col, mel, diag, _ = sm.hubbard_real1d(o)
    Hr = sm.dense(rp, col, mel, diag)
    rpm, colm, melm, diagm, states = sm.hubbard_mom1d((3, 0, 0, 0, 0), u=2.0, t=1.0)
    Hm = sm.dense(rpm, colm, melm, diagm)
    rpf, colf, melf, diagf, fstates = sm.hubbard_2c_lattice((1, 0, 0, 1), (1, 1, 0, 0))
    Hf = sm.dense(rpf, colf, melf, diagf)
    for H in (Hr, Hm, Hf):
        assert np.allclose(H, H.T, atol=1e-14)
    assert len(fstates) == 36 and Hr.shape[0] == 35 and len(states) == 7  # total-momentum-0 sector of 35 states
    assert abs(np.linalg.eigvalsh(Hm)[0] - np.linalg.eigvalsh(Hr)[0]) < 1e-12
    # every momentum-sector eigenvalue is an eigenvalue of the full real-space operator
    er = np.linalg.eigvalsh(Hr)
    assert all(np.min(np.abs(er - e)) < 1e-11 for e in np.linalg.eigvalsh(Hm))
    # the periodic 2x2 lattice lists both directions of a length-2 axis: duplicate targets inside a row
    r0 = colf[int(rpf[0]):int(rpf[1])]
    assert len(r0) > len(set(r0.tolist()))


@pytest.mark.parametrize("name", ["real", "mom", "fermi"])
def test_eigenvector_table_samples_psi_squared(name):
    """test/qmc.jl:16-50 on the oracle: VectorAnsatz(groundstate) => E_loc = E0 at every step, and the
    residence-time histogram converges to |psi|^2."""
    if name == "real":
        rp, col, mel, diag, _ = sm.hubbard_real1d(Oracle(5, 5, 1.0, 0.1))
    elif name == "mom":
        rp, col, mel, diag, _ = sm.hubbard_mom1d((0, 0, 0, 0, 3, 0, 0, 0, 0, 0), u=4.0)  # BoseFS(10, 5 => 3), qmc.jl:12
    else:
        rp, col, mel, diag, _ = sm.hubbard_2c_lattice((1, 0, 0, 1), (1, 1, 0, 0))
    so = SparseOracle(rp, col, mel, diag)
    w, v = np.linalg.eigh(sm.dense(rp, col, mel, diag))
    so.set_table(v[:, 0])
    res = so.kinetic_vqmc(np.zeros(8, dtype=np.uint32), 20000, seed=5, series=True)
    assert np.allclose(res["series_eloc"], w[0], rtol=0, atol=1e-9 * max(1.0, abs(w[0])))
    assert abs(res["val"] - w[0]) < 1e-10 * max(1.0, abs(w[0]))
    hist = np.bincount(res["series_idx"].ravel(), weights=res["series_tau"].ravel(), minlength=so.dim)
    hist /= hist.sum()
    assert abs(np.dot(hist, v[:, 0] ** 2) / np.sqrt(np.dot(hist, hist) * np.sum(v[:, 0] ** 4)) - 1) < 2e-3


def test_lee_matches_dense_rayleigh_quotient_and_finite_differences():
    rp, col, mel, diag, _ = sm.hubbard_mom1d((0, 0, 0, 0, 3, 0, 0, 0, 0, 0), u=4.0)
    so = SparseOracle(rp, col, mel, diag)
    H = sm.dense(rp, col, mel, diag)

    def rq(g):
        psi = np.exp(-g * diag)
        return psi @ H @ psi / (psi @ psi)

    g = 0.37
    so.set_gutzwiller(g)
    val, grad = so.lee_val_and_grad()
    assert abs(val - rq(g)) < 1e-13 * abs(rq(g))
    h = 1e-6
    assert abs(grad[0] - (rq(g + h) - rq(g - h)) / (2 * h)) < 1e-7 * max(1.0, abs(grad[0]))
    # a P = 3 table: psi = exp(-p . f(k)) with arbitrary features
    rng = np.random.default_rng(3)
    F = rng.normal(size=(so.dim, 3))
    p = np.array([0.2, -0.1, 0.05])

    def rq3(p):
        psi = np.exp(-F @ p)
        return psi @ H @ psi / (psi @ psi)

    so.set_table(np.exp(-F @ p), -F)
    val, grad = so.lee_val_and_grad()
    assert abs(val - rq3(p)) < 1e-13 * abs(val)
    for i in range(3):
        e = np.zeros(3)
        e[i] = h
        assert abs(grad[i] - (rq3(p + e) - rq3(p - e)) / (2 * h)) < 1e-7 * max(1.0, abs(grad[i]))


def test_from_hubbard_reproduces_the_reference_rows(gw):
    """SparseHamiltonian.from_hubbard (host numpy) against the pinned oracle's build_basis / offdiagonals /
    diagonal_element: same rows in the same hop order, basis index = colex rank of the unary bitstring"""
    for (N, M, u, t) in ((5, 5, 1.0, 0.1), (3, 6, 2.0, 1.0), (7, 4, 4.0, 0.5)):
        o = Oracle(N, M, u, t)
        rp, col, mel, diag, basis = sm.hubbard_real1d(o)
        H = gw.HubbardReal1D(gw.BoseFS(o.to_onr(basis[0])), u=u, t=t)
        Hs = gw.SparseHamiltonian.from_hubbard(H)
        assert Hs.dim == len(basis) == o.dimension()
        index = {tuple(r): i for i, r in enumerate(Hs.addresses.tolist())}
        idx = np.array([index[tuple(o.to_onr(b))] for b in basis])
        assert idx[0] == Hs.address.index
        rpi, rps = rp.astype(np.int64), Hs.row_ptr.astype(np.int64)
        for k in range(len(basis)):
            a, b = slice(rpi[k], rpi[k + 1]), slice(rps[idx[k]], rps[idx[k] + 1])
            assert np.array_equal(idx[col[a]], Hs.col[b]) and np.array_equal(mel[a], Hs.melem[b])
            assert diag[k] == Hs.diag[idx[k]]
        # index = rank of the bitstring among all N-subsets of N+M-1 bit positions (colex)
        words = np.array([int(gw.BoseFS(r).words[0]) for r in Hs.addresses.tolist()], dtype=np.uint64)
        assert np.all(np.diff(words.astype(np.int64)) > 0)
    with pytest.raises(ValueError, match="too many"):
        gw.SparseHamiltonian.from_hubbard(gw.HubbardReal1D(gw.near_uniform(20, 20)))


def test_oracle_reproduces_committed_sparse_vectors():
    """tests/golden/sparse_vectors.json (made by tests/golden/make_golden_sparse.py) pins part 3 of the oracle"""
    import json
    import os
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_sparse", os.path.join(sys_path, "make_golden_sparse.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    golden = json.load(open(os.path.join(sys_path, "sparse_vectors.json")))
    ops = mg.operators()
    assert len(golden["cases"]) == 6
    for case in golden["cases"]:
        so = SparseOracle(*ops[case["operator"]])
        assert so.dim == case["dim"]
        if case["ansatz"] == "gutzwiller":
            so.set_gutzwiller(0.4)
        else:
            so.set_table(*mg.table(so.dim, 3, 77))
        for row in case["probes"]:
            ks = so.kinetic_sample(row["index"], 0.37)
            assert ks["eloc"] == row["e_loc"] and ks["tau"] == row["tau"] and ks["chosen"] == row["chosen"]
            assert ks["new_index"] == row["next"] and ks["n_hops"] == row["n_hops"] and list(ks["grad"]) == row["grad"]
        v, g = so.lee_val_and_grad()
        assert v == case["lee_val"] and list(g) == case["lee_grad"]
        tr = so.kinetic_vqmc(np.zeros(8, dtype=np.uint32), 120, seed=9)
        assert list(tr["idx"]) == case["final_indices"] and tr["hops"] == case["hops"]


def test_from_hubbard_extended_chain_diagonal(gw):
    """ExtendedHubbardReal1D: the nearest-neighbour term v sum_i n_i n_{i+1} enters the stored diagonal"""
    from oracle.oracle import AnsatzOracle
    o = Oracle(5, 4, 1.5, 0.3)
    ao = AnsatzOracle(o, "gutzwiller", v=0.7)
    H = gw.ExtendedHubbardReal1D(gw.near_uniform(5, 4), u=1.5, v=0.7, t=0.3)
    Hs = gw.SparseHamiltonian.from_hubbard(H)
    H0 = gw.SparseHamiltonian.from_hubbard(gw.HubbardReal1D(gw.near_uniform(5, 4), u=1.5, t=0.3))
    assert np.array_equal(Hs.col, H0.col) and np.array_equal(Hs.melem, H0.melem)  # same hops, other diagonal
    for k, n in enumerate(Hs.addresses.tolist()):
        assert abs(Hs.diag[k] - ao.diagonal_element(o.from_onr(n))) < 1e-14
    so = SparseOracle(Hs.row_ptr, Hs.col, Hs.melem, Hs.diag)
    so.set_gutzwiller(0.3)
    basis = np.stack([o.from_onr(n) for n in Hs.addresses.tolist()])
    v0, g0 = ao.lee_val_and_grad(basis, [0.3])
    v1, g1 = so.lee_val_and_grad()
    assert abs(v1 - v0) < 1e-13 * abs(v0) and abs(g1[0] - g0[0]) < 1e-12 * max(1.0, abs(g0[0]))
