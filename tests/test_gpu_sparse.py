"""GPU parity tests of the stored-operator path (csrc/gwz_sparse.cu) against oracle/gutz_oracle_sparse.c, through
the C ABI: any Hamiltonian as rows over its basis, any ansatz as a table (the reference runs
this path on HubbardReal1D, HubbardMom1D and a two-component HubbardRealSpace, test/qmc.jl:10-14)."""
import os

import numpy as np
import pytest

import sparse_models as sm

pytestmark = pytest.mark.gpu

TOL = 1e-12  # relative, FP64 (north star)


def _operators():
    from oracle.oracle import Oracle
    ops = {}
    rp, col, mel, diag, _ = sm.hubbard_real1d(Oracle(5, 5, 1.0, 0.1))  # test/qmc.jl:11
    ops["real"] = (rp, col, mel, diag)
    ops["mom"] = sm.hubbard_mom1d((0, 0, 0, 0, 3, 0, 0, 0, 0, 0), u=4.0)[:4]  # BoseFS(10, 5 => 3), test/qmc.jl:12
    ops["fermi"] = sm.hubbard_2c_lattice((1, 0, 0, 1), (1, 1, 0, 0))[:4]  # test/qmc.jl:13
    ops["ragged"] = sm.random_operator(3000, 40, seed=2, max_row=400)
    return ops


@pytest.fixture(scope="module")
def ops():
    return _operators()


@pytest.fixture(params=["blocks", "rows"])
def layout(request, monkeypatch):
    """both HBM layouts of the walk: fixed-stride row blocks (default when rows are regular enough) and row pointers"""
    if request.param == "rows":
        monkeypatch.setenv("GWZ_SPARSE_BLOCKS", "0")
    return request.param


def _pair(gw, op, group=None):
    from oracle.oracle import SparseOracle
    if group is None:
        os.environ.pop("GWZ_SPARSE_GROUP", None)
    else:
        os.environ["GWZ_SPARSE_GROUP"] = str(group)
    try:
        H = gw.SparseHamiltonian(*op)
        H.handle()
    finally:
        os.environ.pop("GWZ_SPARSE_GROUP", None)
    return H, SparseOracle(*op)


def _random_table(dim, P, seed):
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=dim) * np.exp(rng.normal(size=dim))
    psi[rng.random(dim) < 0.05] = 0.0  # neighbours outside the support of the vector (rate 0)
    psi[0] = 1.0
    return psi, rng.normal(size=(dim, P))


@pytest.mark.parametrize("name,group", [("real", None), ("mom", None), ("fermi", None), ("ragged", None), ("ragged", 4),
                                        ("ragged", 8), ("mom", 32), ("real", 16)])
def test_probe_matches_oracle(gw, ops, name, group, layout):
    H, so = _pair(gw, ops[name], group)
    h = H.handle()
    if group is not None:
        assert h.info()["lanes_per_walker"] == group
    psi, gr = _random_table(H.dim, 3, 7)
    h.set_table(psi, gr)
    assert h.info()["row_blocks"] == (layout == "blocks")
    so.set_table(psi, gr)
    rng = np.random.default_rng(5)
    idx = np.flatnonzero(psi != 0.0)
    idx = idx[rng.integers(0, idx.size, size=min(400, 4 * H.dim))].astype(np.uint32)
    u = rng.random(idx.size)
    out = h.probe(idx, u, want_rates=True)
    rp = so.row_ptr.astype(np.int64)
    for i, k in enumerate(idx):
        ref = so.kinetic_sample(int(k), float(u[i]))
        assert out["n_hops"][i] == ref["n_hops"]
        row = slice(rp[k], rp[k + 1])
        # E_loc is a signed sum: the tolerance is relative to the magnitude of its terms
        scale = (abs(so.diag[k] * psi[k]) + np.sum(np.abs(so.melem[row] * psi[so.col[row]]))) / abs(psi[k])
        assert abs(out["e_loc"][i] - ref["eloc"]) <= TOL * max(1.0, scale)
        assert abs(out["residence_time"][i] - ref["tau"]) <= TOL * ref["tau"]
        assert np.array_equal(out["grad_ratio"][i], ref["grad"])
        rates = np.abs(psi[so.col[row]]) / abs(psi[k])
        assert np.allclose(out["rates"][i, :ref["n_hops"]], rates, rtol=1e-14, atol=0)
        assert np.all(out["rates"][i, ref["n_hops"]:] == 0.0)
        # the pick: identical unless u*S falls within rounding of a cumulative sum
        c = ref["cumsum"]
        margin = np.min(np.abs(c - u[i] * c[-1])) / c[-1]
        if margin > 1e-12:
            assert out["chosen"][i] == ref["chosen"] and out["next_index"][i] == ref["new_index"]


@pytest.mark.parametrize("name", ["real", "mom", "fermi", "ragged"])
def test_trajectories_match_oracle_step_for_step(gw, ops, name, layout):
    H, so = _pair(gw, ops[name])
    if name == "ragged":
        psi, gr = _random_table(H.dim, 2, 3)
        psi = np.abs(psi) + 0.05
        ans = gw.TableAnsatz(H, lambda p: (psi, gr), 2)
        so.set_table(psi, gr)
        params = [0.0, 0.0]
    else:
        ans = gw.SparseGutzwillerAnsatz(H)
        so.set_gutzwiller(0.4)
        params = [0.4]
    n, steps = 24, 250
    ws = gw.SparseWalkers(H.handle(), ans, n, 0, seed=9)
    r, tau, eloc, addr = ws.run_record(params, steps)
    ref = so.kinetic_vqmc(np.zeros(n, dtype=np.uint32), steps, seed=9, series=True)
    assert np.array_equal(ws.indices(), ref["idx"])
    assert np.array_equal(addr[:, :, 0].T, ref["series_idx"])
    assert np.allclose(tau.T, ref["series_tau"], rtol=TOL, atol=0)
    assert np.allclose(eloc.T, ref["series_eloc"], rtol=1e-11, atol=1e-12)
    vw, gw_ = ws.estimates()
    assert np.allclose(vw, ref["val_w"], rtol=1e-10, atol=1e-12)
    assert np.allclose(gw_, ref["grad_w"], rtol=1e-8, atol=1e-10)
    assert abs(r.val - ref["val"]) <= 1e-10 * max(1.0, abs(ref["val"]))
    assert np.allclose(r.grad, ref["grad"], rtol=1e-8, atol=1e-10)
    assert r.hops == ref["hops"] and r.steps == n * steps and r.walkers == n
    # continuation: the random streams carry on at step `steps`
    ws.run(params, 50)
    ref2 = so.kinetic_vqmc(ref["idx"], 50, seed=9, step0=steps)
    assert np.array_equal(ws.indices(), ref2["idx"])


@pytest.mark.parametrize("P", [2, 7, 20, 70])
def test_many_parameter_tables_follow_the_oracle(gw, ops, P, layout):
    """gradient accumulators for P parameters (several per lane; the lane group widens when P > 8 per lane)"""
    H, so = _pair(gw, ops["real"])
    psi, gr = _random_table(H.dim, P, 21 + P)
    psi = np.abs(psi) + 0.1
    ans = gw.TableAnsatz(H, lambda p: (psi, gr), P)
    so.set_table(psi, gr)
    ws = gw.SparseWalkers(H.handle(), ans, 40, 0, seed=13)
    r = ws.run(np.zeros(P), 300)
    assert H.handle().info()["lanes_per_walker"] * 8 >= P
    ref = so.kinetic_vqmc(np.zeros(40, dtype=np.uint32), 300, seed=13)
    assert np.array_equal(ws.indices(), ref["idx"])
    vw, gw_ = ws.estimates()
    assert np.allclose(vw, ref["val_w"], rtol=1e-10, atol=1e-12)
    assert gw_.shape == (40, P) and np.allclose(gw_, ref["grad_w"], rtol=1e-7, atol=1e-9)
    assert np.allclose(r.grad, ref["grad"], rtol=1e-7, atol=1e-9)
    assert np.all(r.grad_err > 0) and np.allclose(r.pooled_grad, r.pooled_grad)


def test_shards_reproduce_the_unsharded_walk(gw, ops):
    H, _ = _pair(gw, ops["mom"])
    ans = gw.SparseGutzwillerAnsatz(H)
    one = gw.SparseWalkers(H.handle(), ans, 64, 0, seed=4)
    a = gw.SparseWalkers(H.handle(), ans, 40, 0, seed=4, global_offset=0)
    b = gw.SparseWalkers(H.handle(), ans, 24, 0, seed=4, global_offset=40)
    r1 = one.run([0.3], 200)
    ra, rb = a.run([0.3], 200), b.run([0.3], 200)
    assert np.array_equal(one.indices(), np.concatenate([a.indices(), b.indices()]))
    assert abs(r1.val - (40 * ra.val + 24 * rb.val) / 64) < 1e-12 * abs(r1.val)
    # keep_acc: two runs of 100 steps accumulate to the same estimators as one run of 200
    c = gw.SparseWalkers(H.handle(), ans, 64, 0, seed=4)
    c.run([0.3], 100)
    rc = c.run([0.3], 100, keep_acc=True)
    assert rc.val == r1.val and np.array_equal(rc.grad, r1.grad) and rc.steps == r1.steps


@pytest.mark.parametrize("name", ["real", "mom", "fermi", "ragged"])
def test_sweep_matches_oracle(gw, ops, name):
    H, so = _pair(gw, ops[name])
    h = H.handle()
    for P in (0, 1, 5, 40):
        psi, gr = _random_table(H.dim, P, 11 + P)
        h.set_table(psi, gr if P else None)
        so.set_table(psi, gr if P else None)
        v0, g0 = so.lee_val_and_grad()
        v1, g1 = h.sweep()
        assert abs(v1 - v0) <= TOL * max(1.0, abs(v0))
        assert np.allclose(g1, g0, rtol=1e-10, atol=1e-11 * max(1.0, np.abs(g0).max(initial=0.0)))
        assert abs(h.sweep(False)[0] - v0) <= TOL * max(1.0, abs(v0))
        cnt = min(20, H.dim - 3)
        terms = h.sweep_terms(3, cnt)
        for j in range(cnt):
            t = so.lee_terms(3 + j)
            assert np.allclose(terms[j], t, rtol=1e-11, atol=1e-13 * max(1.0, np.abs(t).max()))
    # row slices add up (what the ranks of a multi-GPU job each compute)
    h.set_gutzwiller(0.2)
    so.set_gutzwiller(0.2)
    v0, g0 = so.lee_val_and_grad()
    assert abs(h.sweep()[0] - v0) <= TOL * max(1.0, abs(v0)) and np.allclose(h.sweep()[1], g0, rtol=1e-10)
    psi_dev, gr_dev = h.get_table()
    assert np.allclose(psi_dev, so.psi, rtol=1e-14) and np.array_equal(gr_dev, so.gr)


@pytest.mark.parametrize("name", ["real", "mom", "fermi"])
def test_eigenvector_ansatz_samples_psi_squared(gw, ops, name):
    """test/qmc.jl:16-50: VectorAnsatz(groundstate): val == E0, DVec(res) -> |psi|^2, local_energy_estimator == E0"""
    op = ops[name]
    H = gw.SparseHamiltonian(*op)
    w, v = np.linalg.eigh(sm.dense(*op))
    E0, gs = w[0], v[:, 0]
    ans = gw.SparseVectorAnsatz(H, gs)
    res = gw.kinetic_vqmc(H, ans, [], samples=1e6, walkers=1000)
    val, grad = gw.val_and_grad(res)
    assert len(grad) == 0 and abs(val - E0) < 1e-9 * max(1.0, abs(E0))
    assert len(res) == 1_000_000
    assert abs(gw.local_energy_estimator(res).mean - E0) < 1e-6 * max(1.0, abs(E0))
    ws = gw.SparseWalkers(H.handle(), ans, 4096, 0, seed=1)
    d6 = d7 = None
    for steps in (250, 2500):
        hist, r = ws.sample_vector([], steps, burn_in=50)
        assert abs(hist.sum() - r.total_time) < 1e-9 * r.total_time
        sampled = hist / hist.sum()
        gsq = gs ** 2 / np.sum(gs ** 2)
        cos = np.dot(sampled, gsq) / np.sqrt(np.dot(sampled, sampled) * np.dot(gsq, gsq))
        assert abs(cos - 1) < (1e-4 if steps == 250 else 5e-5)
        d6, d7 = d7, np.max(np.abs(sampled - gsq))
    assert d7 < d6
    # Contination run (qmc.jl:48-49)
    n0 = len(res)
    gw.kinetic_vqmc_continue(res)
    assert len(res) == 2 * n0


@pytest.mark.parametrize("name", ["real", "mom", "fermi"])
def test_gutzwiller_on_any_operator(gw, ops, name):
    """test/qmc.jl:52-75: GutzwillerAnsatz(H) at 0.5 against the Rayleigh quotient of its vector"""
    op = ops[name]
    H = gw.SparseHamiltonian(*op)
    ans = gw.SparseGutzwillerAnsatz(H)
    Hd, diag = sm.dense(*op), op[3]
    psi = np.exp(-0.5 * diag)
    E_v = psi @ Hd @ psi / (psi @ psi)
    le = gw.LocalEnergyEvaluator(H, ans)
    assert abs(le(0.5) - E_v) < 1e-12 * max(1.0, abs(E_v))
    v, g = le.val_and_grad(0.5)
    h = 1e-6
    fd = (le(0.5 + h) - le(0.5 - h)) / (2 * h)
    assert abs(v - E_v) < 1e-12 * max(1.0, abs(E_v)) and abs(g[0] - fd) < 1e-6 * max(1.0, abs(fd))
    res = gw.kinetic_vqmc(H, ans, 0.5, samples=1e6, walkers=512, burn_in=100)
    val1, grad1 = gw.val_and_grad(res)
    qmc = gw.KineticVQMC(H, ans, samples=1e6, walkers=512, burn_in=100)
    val2, grad2 = qmc.val_and_grad(0.5)
    assert abs(val1 - val2) < 0.1 * max(1.0, abs(val1)) and np.allclose(grad1, grad2, rtol=0.1, atol=0.05)
    est = gw.local_energy_estimator(res)
    assert est.mean - 4 * est.err < E_v < est.mean + 4 * est.err
    v3, e3, g3 = qmc.val_err_and_grad(0.5)
    assert abs(v3 - E_v) < 5 * e3 + 1e-9 and abs(g3[0] - g[0]) < 0.1 * max(1.0, abs(g[0]))
    addrs, vals = gw.sampled_vector(res)
    sq = np.zeros(H.dim)
    sq[addrs[:, 0].astype(np.int64)] = vals
    target = psi ** 2 / np.linalg.norm(psi ** 2)
    assert abs(np.dot(sq, target) - 1) < 1e-3
    row = next(iter(res.rows()))
    assert row["walker"] == 1 and row["step"] == 1 and 0 <= row["address"] < H.dim


def test_gradient_of_a_table_ansatz_and_amsgrad(gw, ops):
    """a 3-parameter table ansatz on the momentum-space chain: sampled gradient against the sweep's, and
    amsgrad(le, p0) driving the stored-operator sweep to a stationary point (test/amsgrad.jl:9-21)"""
    op = ops["mom"]
    H = gw.SparseHamiltonian(*op)
    rng = np.random.default_rng(8)
    F = rng.normal(size=(H.dim, 3)) * 0.5

    def fn(p):
        return np.exp(-F @ p), -F

    ans = gw.TableAnsatz(H, fn, 3)
    p = np.array([0.3, -0.2, 0.1])
    le = gw.LocalEnergyEvaluator(H, ans)
    v, g = le.val_and_grad(p)
    qmc = gw.KineticVQMC(H, ans, samples=4e6, walkers=4096, burn_in=100)
    vq, eq, gq = qmc.val_err_and_grad(p)
    assert abs(vq - v) < 5 * eq
    r = qmc.result.last
    assert np.all(np.abs(gq - g) < 5 * r.grad_err + 1e-3)
    out = gw.amsgrad(le, p, maxiter=300, alpha=0.05)
    vmin, gmin = le.val_and_grad(out.param[-1])
    assert vmin < v and np.linalg.norm(gmin) < np.linalg.norm(g)
    assert vmin >= np.linalg.eigvalsh(sm.dense(*op))[0] - 1e-9


def test_chain_as_stored_operator_agrees_with_the_bit_level_kernels(gw):
    """HubbardReal1D{8,8} both ways: bit-level sweep kernel vs stored rows (same state order: index = rank), and a
    user-defined ansatz (the reference's JastrowAnsatz formula, jastrow.jl:4-58, written in numpy and tabulated)
    against the generic engine's JastrowAnsatz kernel"""
    H = gw.HubbardReal1D(gw.near_uniform(8, 8), u=2.0)
    Hs = gw.SparseHamiltonian.from_hubbard(H)
    le_bits = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    le_rows = gw.LocalEnergyEvaluator(Hs, gw.SparseGutzwillerAnsatz(Hs))
    assert le_rows.dim == le_bits.dim == 6435
    for g in (0.0, 0.3, 1.0):
        v0, g0 = le_bits.val_and_grad(g)
        v1, g1 = le_rows.val_and_grad(g)
        assert abs(v1 - v0) <= TOL * abs(v0) and abs(g1[0] - g0[0]) <= 1e-11 * max(1.0, abs(g0[0]))
    t0, t1 = le_bits.terms(0.3), le_rows.terms(0.3)
    assert np.allclose(t1, t0, rtol=1e-12, atol=1e-13)
    words = le_bits.states()[:, 0]
    assert np.array_equal(words, [int(gw.BoseFS(r).words[0]) for r in Hs.addresses.tolist()])
    # a user-defined ansatz over Hs.addresses
    M = 8
    iu = np.triu_indices(M)  # pairs (i <= j) in the reference's parameter order (jastrow.jl:34-40)
    feats = (Hs.addresses[:, iu[0]] * Hs.addresses[:, iu[1]]).astype(np.float64)
    custom = gw.TableAnsatz(Hs, lambda p: (np.exp(-feats @ p), -feats), feats.shape[1])
    jast = gw.JastrowAnsatz(H)
    rng = np.random.default_rng(2)
    p = 0.05 * rng.random(feats.shape[1])
    va, ga = gw.LocalEnergyEvaluator(H, jast).val_and_grad(p)
    vb, gb = gw.LocalEnergyEvaluator(Hs, custom).val_and_grad(p)
    assert abs(va - vb) <= 1e-11 * abs(va) and np.allclose(ga, gb, rtol=1e-9, atol=1e-10)
    res = gw.kinetic_vqmc(Hs, custom, p, samples=2e6, walkers=2048, burn_in=200, record=False, estimator="pooled")
    assert abs(res.last.val - vb) < 5 * res.last.err
    assert np.all(np.abs(res.last.grad - gb) < 6 * res.last.grad_err + 1e-3)


def test_against_committed_golden_vectors(gw, layout):
    """tests/golden/sparse_vectors.json: per-index outputs, LocalEnergyEvaluator values and final walker positions"""
    import importlib.util
    import json
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_sparse", os.path.join(gdir, "make_golden_sparse.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    golden = json.load(open(os.path.join(gdir, "sparse_vectors.json")))
    operators = mg.operators()
    for case in golden["cases"]:
        H = gw.SparseHamiltonian(*operators[case["operator"]])
        h = H.handle()
        if case["ansatz"] == "gutzwiller":
            ans, params = gw.SparseGutzwillerAnsatz(H), [0.4]
        else:
            psi, gr = mg.table(H.dim, 3, 77)
            ans, params = gw.TableAnsatz(H, lambda p, psi=psi, gr=gr: (psi, gr), 3), [0.0, 0.0, 0.0]
        h.load(ans, params)
        assert h.info()["row_blocks"] == (layout == "blocks")
        idx = np.array([r["index"] for r in case["probes"]], dtype=np.uint32)
        out = h.probe(idx, np.full(idx.size, 0.37))
        for i, r in enumerate(case["probes"]):
            assert abs(out["e_loc"][i] - r["e_loc"]) <= 1e-11 * max(1.0, abs(r["e_loc"]))
            assert abs(out["residence_time"][i] - r["tau"]) <= TOL * r["tau"]
            assert out["n_hops"][i] == r["n_hops"] and out["chosen"][i] == r["chosen"] and out["next_index"][i] == r["next"]
            assert np.allclose(out["grad_ratio"][i], r["grad"], rtol=1e-14, atol=0)
        v, g = h.sweep()
        assert abs(v - case["lee_val"]) <= TOL * max(1.0, abs(case["lee_val"]))
        assert np.allclose(g, case["lee_grad"], rtol=1e-10, atol=1e-12)
        ws = gw.SparseWalkers(h, ans, 8, 0, seed=9)
        r = ws.run(params, 120)
        assert list(ws.indices()) == case["final_indices"] and r.hops == case["hops"]
        assert np.allclose(ws.estimates()[0], case["val_w"], rtol=1e-10, atol=1e-12)


def test_argument_checks(gw, ops):
    rp, col, mel, diag = ops["fermi"]
    bad = col.copy()
    bad[3] = len(diag)
    with pytest.raises(ValueError, match="outside the basis"):
        gw.SparseHamiltonian(rp, bad, mel, diag).handle()
    m2 = mel.copy()
    m2[0] = np.nan
    with pytest.raises(ValueError, match="non-finite"):
        gw.SparseHamiltonian(rp, col, m2, diag).handle()
    H = gw.SparseHamiltonian(rp, col, mel, diag)
    with pytest.raises(ValueError, match="no ansatz table"):
        H.handle().sweep()
    with pytest.raises(ValueError, match="non-finite"):
        H.handle().set_table(np.full(H.dim, np.inf))
    with pytest.raises(TypeError):
        gw.kinetic_vqmc(H, gw.SparseGutzwillerAnsatz(gw.SparseHamiltonian(rp, col, mel, diag)), 0.5)
    with pytest.raises(ValueError):
        gw.SparseHamiltonian(rp, col, mel, diag, start=len(diag))
    # a walk that reaches a state whose neighbours all have amplitude 0: "Infinite residence time" (qmc.jl:34-36)
    psi = np.zeros(H.dim)
    psi[0] = 1.0
    ans = gw.SparseVectorAnsatz(H, psi)
    with pytest.raises(FloatingPointError):
        gw.SparseWalkers(H.handle(), ans, 8, 0, seed=1).run([], 10)
    # the error is reported once; the next run on the same context is clean
    ok = gw.SparseVectorAnsatz(H, np.ones(H.dim))
    assert np.isfinite(gw.SparseWalkers(H.handle(), ok, 8, 0, seed=1).run([], 10).val)


def test_amsgrad_on_one_parameter_stored_operator_objectives(gw):
    """a 1-parameter stored-operator objective must not take the in-library fast path of the bit-level
    engine (its handles are gwz_swalkers / gwz_sparse, not gwz_walkers / gwz_sweep).  Both objectives descend through
    the callback path and find the optimum the bit-level evaluator finds."""
    H = gw.HubbardReal1D(gw.near_uniform(6, 6), u=4.0)
    Hs = gw.SparseHamiltonian.from_hubbard(H)
    le_bits = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    ref = gw.amsgrad(le_bits, [0.5], maxiter=200)
    le_rows = gw.LocalEnergyEvaluator(Hs, gw.SparseGutzwillerAnsatz(Hs))
    out = gw.amsgrad(le_rows, [0.5], maxiter=200)
    assert abs(out.param[-1][0] - ref.param[-1][0]) < 1e-6 and abs(out.value[-1] - ref.value[-1]) < 1e-9
    qmc = gw.KineticVQMC(Hs, gw.SparseGutzwillerAnsatz(Hs), samples=2e6, walkers=4096, burn_in=50)
    outq = gw.amsgrad(qmc, [0.5], maxiter=40)
    assert np.isfinite(outq.value[-1]) and abs(outq.param[-1][0] - ref.param[-1][0]) < 0.1
    assert qmc.result.last is not None  # the last run stays in the objective, as in the reference (wrapper.jl:74-79)


def test_handles_of_another_kind_are_rejected(gw):
    """Handles carry a type tag: a stored-operator walker set where the bit-level entry points expect theirs (and the
    other way round) is GWZ_ERR_INVALID (ValueError), never a read through the wrong layout."""
    import ctypes as C
    from gutzwiller_b200 import capi
    L = gw.lib()
    H = gw.HubbardReal1D(gw.near_uniform(6, 6), u=2.0)
    Hs = gw.SparseHamiltonian.from_hubbard(H)
    qs = gw.KineticVQMC(Hs, gw.SparseGutzwillerAnsatz(Hs), samples=1e4, walkers=64)
    qs.val_and_grad([0.3])
    qb = gw.KineticVQMC(H, gw.GutzwillerAnsatz(H), samples=1e4, walkers=64)
    qb.val_and_grad([0.3])
    p = (C.c_double * 1)(0.3)
    r = capi.VqmcResult()
    hs = qs.result.walkers_handle.handle
    hb = qb.result.walkers_handle.handle
    assert L.gwz_vqmc_run(hb, p, 10, 0, 1, 0, C.byref(r)) == capi.OK
    rc = L.gwz_vqmc_run(hs, p, 10, 0, 1, 0, C.byref(r))
    assert rc == capi.ERR_INVALID and b"handle" in L.gwz_last_error()
    le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    v, g = C.c_double(), (C.c_double * 1)()
    assert L.gwz_sweep_val_and_grad(hb, p, C.byref(v), g, None) == capi.ERR_INVALID
    assert L.gwz_sweep_val_and_grad(le.handle, p, C.byref(v), g, None) == capi.OK
