"""CPU tests of the drop-in boundary and the host logic (no compute calls: there is no GPU here)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gutzwiller_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(gwz_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol(gw):
    import ctypes
    L = ctypes.CDLL(gw.capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/gutzwiller_b200.h but not exported"
    # and the Python binding covers exactly the declared set
    assert sorted(gw.capi.SIGNATURES) == declared


def test_no_torch_types_in_abi():
    text = open(HEADER).read()
    assert "torch" not in text.lower() and "at::" not in text
    assert 'extern "C"' in text


def test_version_and_error_channel(gw):
    L = gw.lib()
    assert L.gwz_version() == 200
    import ctypes as C
    h = C.c_void_p()
    rc = L.gwz_context_create(0, None)
    assert rc == gw.capi.ERR_INVALID and b"NULL" in L.gwz_last_error()
    del h


def test_no_cpu_fallback_without_gpu(gw):
    """Without a CUDA device every compute entry point must fail loudly (no CPU path)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(gw.GutzwillerError, match="no CUDA device"):
        gw.Context(0)
    H = gw.HubbardReal1D(gw.near_uniform(4, 4))
    with pytest.raises(gw.GutzwillerError):
        gw.kinetic_vqmc(H, gw.GutzwillerAnsatz(H), [0.5], samples=100, walkers=1)
    with pytest.raises(gw.GutzwillerError):
        gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gutzwiller.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
    assert "oracle" not in open(os.path.join(ROOT, "gutzwiller_b200.py")).read().lower()


def test_bosefs_and_near_uniform(gw, oracle_mod):
    a = gw.near_uniform(10, 10)
    assert int(a.words[0]) == 0b1010101010101010101 and a.onr() == (1,) * 10
    assert repr(a) == 'fs"|1 1 1 1 1 1 1 1 1 1⟩"'
    b = gw.BoseFS((2, 0, 1))
    assert int(b.words[0]) == 0b10011 and b.N == 3 and b.M == 3  # SURVEY App. A.1 example
    assert gw.BoseFS(1, 1, 1, 1) == gw.BoseFS((1, 1, 1, 1))
    rng = np.random.default_rng(3)
    for N, M in [(7, 3), (50, 50), (100, 100), (33, 32), (13, 40)]:
        o = oracle_mod.Oracle(N, M)
        onr = rng.multinomial(N, rng.dirichlet(np.ones(M)))
        fs = gw.BoseFS(onr)
        assert np.array_equal(fs.words, o.from_onr(onr)[: fs.words.size])
        assert fs.onr() == tuple(int(x) for x in onr)
        assert gw.BoseFS.from_words(N, M, fs.words) == fs
        nu = gw.near_uniform(N, M)
        assert np.array_equal(nu.words, o.near_uniform()[: nu.words.size])
    with pytest.raises(ValueError):
        gw.BoseFS.from_words(3, 3, [0b111111])


def test_ansatz_traits(gw):
    H = gw.HubbardReal1D(gw.near_uniform(5, 5), t=0.1)
    ans = gw.GutzwillerAnsatz(H)
    assert gw.num_parameters(ans) == 1 and gw.keytype(ans) is gw.BoseFS and gw.valtype(ans) is float
    assert gw.starting_address(H) == gw.near_uniform(5, 5)
    assert H.u == 1.0 and H.t == 0.1
    with pytest.raises(TypeError):
        gw.GutzwillerAnsatz("not a hamiltonian")


def test_shard_range_partitions(gw):
    for total in (1, 7, 2 ** 24, 1352078):
        for world in (1, 2, 3, 4, 8):
            spans = [gw.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        gw.shard_range(10, 2, 2)


def test_default_walkers_and_steps_rule(gw):
    assert gw.default_walkers(1e6) == 512 and gw.default_walkers(1e8) == 65536 and gw.default_walkers(100) == 1
    assert gw.default_walkers(1e12) == 1 << 20


def test_resample_and_blocking_host_functions(gw, oracle_mod):
    o = oracle_mod.Oracle(2, 2)
    rng = np.random.default_rng(11)
    t, v = rng.random(257) + 0.01, rng.standard_normal(257)
    for k in (1, 2, 7, 64, 257, 1000):
        got, ref = gw.resample(t, v, len=k), o.resample(t, v, k)
        assert np.allclose(got, ref, rtol=0, atol=1e-13)
        assert abs(got.mean() - np.sum(t * v) / np.sum(t)) < 1e-12  # test/runtests.jl:10-19
    with pytest.raises(ValueError, match="do not match"):
        gw.resample(t, v[:-1])
    x = np.cumsum(rng.standard_normal(5000)) * 0.01 + rng.standard_normal(5000)
    a, b = gw.blocking_analysis(x), o.blocking_analysis(x)
    assert (a.k, a.blocks) == (b.k, b.blocks) and abs(a.err - b.err) < 1e-14 and abs(a.mean - b.mean) < 1e-14
    one = gw.blocking_analysis([3.0])
    assert one.k == 0 and one.mean == 3.0


def test_generic_gradient_descent_driver_on_python_objective(gw):
    """amsgrad.jl:117-226 driven through the C callback ABI with a plain quadratic objective."""
    class Quad:
        def val_and_grad(self, p):
            return float(np.sum((p - 1.5) ** 2)), 2 * (p - 1.5)
    gd = gw.gradient_descent(Quad(), [0.0, 3.0], α=0.1, maxiter=200)
    assert gd.converged and np.allclose(gd.param[-1], 1.5, atol=1e-3)  # stops on val_tol
    ams = gw.amsgrad(Quad(), [0.0, 3.0], maxiter=30, early_stop=False)
    assert ams.iterations == 31 and not ams.converged and ams.reason == "iterations"  # maxiter+1 rows
    # first row reproduces the update rule: m1 = 0.9*g0 + 0.1*g, m2 = 0.01*g^2, dp = -0.01*m1/sqrt(m2)
    g0 = 2 * (np.array([0.0, 3.0]) - 1.5)
    assert np.allclose(ams.first_moment[0], g0) and np.allclose(ams.second_moment[0], 0.01 * g0 ** 2)
    assert np.allclose(ams.param_delta[0], -0.01 * g0 / np.sqrt(0.01 * g0 ** 2))
    # fix_params keeps a coordinate (test/amsgrad.jl:38-50)
    fx = gw.amsgrad(Quad(), [0.0, 3.0], fix_params=(False, True), maxiter=50, early_stop=False)
    assert np.all(fx.param[:, 1] == 3.0) and fx.param[-1, 0] > 0.0
    # continuation from a result == passing its last moments (test/amsgrad.jl:52-77, exact)
    a = gw.amsgrad(Quad(), [0.0, 3.0], maxiter=20, early_stop=False)
    b = gw.amsgrad(Quad(), a.param[-1], first_moment_init=a.first_moment[-1], second_moment_init=a.second_moment[-1],
                   maxiter=10, early_stop=False)
    d = gw.amsgrad(Quad(), a, maxiter=10, early_stop=False)
    assert np.array_equal(b.param, d.param)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle.oracle import Oracle
    import gutzwiller_b200 as gw
    o = Oracle(6, 6, 2.0, 1.0)
    total, steps = 10, 300
    b, e = gw.shard_range(total, world, rank)
    addrs = np.stack([o.near_uniform()] * (e - b))
    out = o.kinetic_vqmc(0.7, addrs, steps, seed=99, walker0=b)
    # the accumulator vector the GPU path all-reduces (walkers, sum val_w, sum grad_w)
    acc = torch.tensor([e - b, out["val_w"].sum(), out["grad_w"].sum()], dtype=torch.float64)
    dist.all_reduce(acc)
    # the stored-operator path shards the same way: walkers by global id, sweep rows by range
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sparse_models as sm
    from oracle.oracle import SparseOracle
    so = SparseOracle(*sm.hubbard_mom1d((0, 0, 3, 0, 0, 0), u=4.0)[:4])
    so.set_gutzwiller(0.3)
    outs = so.kinetic_vqmc(np.zeros(e - b, dtype=np.uint32), steps, seed=7, walker0=b)
    r0, r1 = gw.shard_range(so.dim, world, rank)
    terms = np.sum([so.lee_terms(k) for k in range(r0, r1)], axis=0)
    acc2 = torch.tensor([outs["val_w"].sum(), outs["grad_w"].sum(), *terms], dtype=torch.float64)
    dist.all_reduce(acc2)
    q.put((rank, acc.tolist() + acc2.tolist()))
    dist.destroy_process_group()


def test_sharding_is_invariant_under_world_size_gloo(oracle_mod):
    """World-size-2 gloo run of the N>1 host logic: contiguous shards by global walker id, Philox
    keyed by the global id, one sum all-reduce of the accumulator vector => same result as 1 rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = oracle_mod.Oracle(6, 6, 2.0, 1.0)
    single = o.kinetic_vqmc(0.7, np.stack([o.near_uniform()] * 10), 300, seed=99, walker0=0)
    for _, acc in res:
        assert acc[0] == 10
        assert abs(acc[1] / 10 - single["val"]) < 1e-12 and abs(acc[2] / 10 - single["grad"]) < 1e-11
    import sparse_models as sm
    so = oracle_mod.SparseOracle(*sm.hubbard_mom1d((0, 0, 3, 0, 0, 0), u=4.0)[:4])
    so.set_gutzwiller(0.3)
    one = so.kinetic_vqmc(np.zeros(10, dtype=np.uint32), 300, seed=7)
    v, g = so.lee_val_and_grad()
    for _, acc in res:
        assert abs(acc[3] / 10 - one["val"]) < 1e-12 * abs(one["val"]) and abs(acc[4] / 10 - one["grad"][0]) < 1e-11
        num, den, ng, dg = acc[5:9]
        assert abs(num / den - v) < 1e-12 * abs(v) and abs((ng * den - num * dg) / den ** 2 - g[0]) < 1e-11


def test_generic_ansatz_host_objects(gw):
    """Host mirrors of the other ansaetze (no device work: handles are created lazily)."""
    H = gw.HubbardReal1D(gw.BoseFS(1, 1, 1, 1, 1), t=0.1)
    He = gw.ExtendedHubbardReal1D(gw.near_uniform(5, 5), v=0.5, t=0.1)
    M = 5
    assert gw.num_parameters(gw.JastrowAnsatz(H)) == M * (M + 1) // 2          # jastrow.jl:20
    assert gw.num_parameters(gw.RelativeJastrowAnsatz(H)) == (M + 1) // 2      # cld(M, 2), jastrow.jl:78
    assert gw.num_parameters(gw.DensityProfileAnsatz(H)) == M                  # densityprofile.jl:15
    assert gw.num_parameters(gw.MultinomialAnsatz(H, normalize=True)) == 1
    assert gw.num_parameters(gw.ExtendedGutzwillerAnsatz(He)) == 2
    g = gw.GutzwillerAnsatz(H)
    ge = gw.GutzwillerAnsatz(He)
    assert not getattr(g, "generic", False) and getattr(ge, "generic", False) and isinstance(ge, gw.GutzwillerAnsatz)
    combo = g + gw.MultinomialAnsatz(H)                                       # test/ansatz.jl:58
    assert isinstance(combo, gw.CombinationAnsatz) and gw.num_parameters(combo) == 3   # combination.jl:14
    assert gw.num_parameters(gw.RelativeJastrowAnsatz(H) + gw.DensityProfileAnsatz(H)) == 3 + 5 + 1
    with pytest.raises(TypeError):
        g + gw.MultinomialAnsatz(He)      # different Hamiltonians
    with pytest.raises(TypeError):
        combo + g                         # nested combinations are not supported
    with pytest.raises(TypeError):
        gw.RelativeJastrowAnsatz("not a hamiltonian")
    assert gw.keytype(combo) is gw.BoseFS and gw.valtype(combo) is float
    assert repr(He).startswith("ExtendedHubbardReal1D(") and "v=0.5" in repr(He)


def test_graft_entry_build_is_consistent_with_the_header():
    """build() (the driver's "does it build" check) must accept the library it has just built."""
    import __graft_entry__ as g
    g.build()


def test_header_is_plain_c99(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 (what a Julia ccall / cgo / ctypes binding reads)."""
    src = tmp_path / "c.c"
    src.write_text('#include "gutzwiller_b200.h"\nint main(void) { gwz_gen_result r; gwz_vqmc_result v; (void)r; (void)v; return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                    str(src), "-o", str(tmp_path / "c.o")], check=True)


def test_sparse_host_objects(gw):
    """SparseHamiltonian / tabulated ansaetze: host-side validation and tables (no device call)"""
    import sparse_models as sm
    rp, col, mel, diag, states = sm.hubbard_2c_lattice((1, 0, 0, 1), (1, 1, 0, 0))
    H = gw.SparseHamiltonian(rp, col, mel, diag, addresses=states)
    assert H.dim == 36 and H.address.index == 0 and "dim=36" in repr(H)
    with pytest.raises(ValueError):
        gw.SparseHamiltonian(rp[:-1], col, mel, diag)
    with pytest.raises(ValueError):
        gw.SparseHamiltonian(rp, col[:-1], mel, diag)
    with pytest.raises(ValueError):
        gw.SparseHamiltonian(rp, col, mel, diag, start=36)
    a = gw.SparseGutzwillerAnsatz(H)
    assert gw.num_parameters(a) == 1
    psi, gr = a.table([0.5])
    assert np.allclose(psi, np.exp(-0.5 * diag)) and np.array_equal(gr[:, 0], -diag)
    v, g = a.val_and_grad(3, 0.5)
    assert v == psi[3] and g[0] == -diag[3] * psi[3]  # gutzwiller.jl:25-33
    va = gw.SparseVectorAnsatz(H, np.arange(36.0))
    assert gw.num_parameters(va) == 0 and va(5) == 5.0 and va.val_and_grad(5)[1].size == 0
    with pytest.raises(ValueError):
        gw.SparseVectorAnsatz(H, np.ones(35))
    with pytest.raises(TypeError):
        gw.SparseGutzwillerAnsatz(gw.HubbardReal1D(gw.near_uniform(4, 4)))
    ta = gw.TableAnsatz(H, lambda p: (np.exp(-p[0] * diag - p[1]), np.stack([-diag, -np.ones(36)], axis=1)), 2)
    assert gw.num_parameters(ta) == 2 and ta.table([0.1, 0.2])[1].shape == (36, 2)
    with pytest.raises(TypeError):  # pairs must match
        gw.kinetic_vqmc(gw.HubbardReal1D(gw.near_uniform(4, 4)), a, 0.5)


def test_sweep_representative_range_against_brute_force(gw):
    """Host-only helper behind the symmetric ranked sweep: every representative of a dihedral orbit (largest rotation of
    the unary string and of its mirror image) lies in the visited rank range, except the uniform state at commensurate
    filling, which is handed over separately -- also when the rank range is cut into shards."""
    import ctypes as C
    import itertools
    from math import comb
    L = gw.lib()

    def unary(n):
        v = pos = 0
        for k in n:
            v |= ((1 << k) - 1) << pos
            pos += k + 1
        return v

    def rank_of(x):
        r = j = 0
        for p in range(64):
            if (x >> p) & 1:
                j += 1
                r += comb(p, j)
        return r

    def query(N, M, begin, dim):
        lo, cnt, hu, ux = C.c_uint64(), C.c_uint64(), C.c_int32(), C.c_uint64()
        gw.capi.check(L.gwz_sweep_representative_range(N, M, begin, dim, C.byref(lo), C.byref(cnt), C.byref(hu), C.byref(ux)))
        return lo.value, cnt.value, hu.value, ux.value

    for N, M in ((4, 4), (6, 3), (3, 6), (5, 4), (6, 6), (2, 2), (9, 3), (1, 5), (8, 4)):
        total = comb(N + M - 1, N)
        reps = []
        for ones in itertools.combinations(range(N + M - 1), N):
            x = sum(1 << p for p in ones)
            n, run = [], 0
            for p in range(N + M - 1):
                if (x >> p) & 1:
                    run += 1
                else:
                    n.append(run)
                    run = 0
            n = tuple(n + [run])
            r = tuple(reversed(n))
            if x == max([unary(n[i:] + n[:i]) for i in range(M)] + [unary(r[i:] + r[:i]) for i in range(M)]):
                reps.append(x)
        uniform = unary((N // M,) * M) if N % M == 0 else None
        cuts = sorted({0, total // 3, (2 * total) // 3 + 1, total} - {total + 1})
        seen_uniform = 0
        for b, e in zip(cuts[:-1], cuts[1:]):
            if e <= b:
                continue
            lo, cnt, hu, ux = query(N, M, b, e - b)
            assert b <= lo and lo + cnt == e
            if hu:
                assert ux == uniform and b <= rank_of(ux) < e
                seen_uniform += 1
            for x in reps:
                if x != uniform and b <= rank_of(x) < e:
                    assert lo <= rank_of(x), (N, M, b, e, x)
        assert seen_uniform == (1 if uniform is not None else 0)
        lo, cnt, _, _ = query(N, M, 0, total)
        assert cnt < total or M == 1  # something is always skipped
    rc = L.gwz_sweep_representative_range(40, 40, 0, 1, None, None, None, None)
    assert rc == gw.capi.ERR_INVALID


def test_julia_shim_ccalls_match_the_header():
    """julia/GutzwillerB200.jl cannot run here (no Julia): at least every ccall must name a declared entry point with the
    right number of arguments of the right kind (pointer / 32- or 64-bit integer / double)."""
    jl = open(os.path.join(ROOT, "gutzwiller.jl_b200", "julia", "GutzwillerB200.jl")).read()
    calls = re.findall(r"ccall\(\(:(\w+), LIB\),\s*(\w+),\s*\(([^)]*)\)", jl, flags=re.S)
    assert len(calls) >= 40
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char \*|const char\*)\s*\*?\s*(gwz_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).replace("\n", " ")
        protos[m.group(1)] = [a.strip() for a in args.split(",")] if args.strip() not in ("", "void") else []

    def kind_c(a):
        if "*" in a:
            return "ptr"
        t = a.rsplit(" ", 1)[0].replace("const ", "").strip()
        if t == "gwz_objective_fn":  # function pointer typedef
            return "ptr"
        return {"int": "i32", "int32_t": "i32", "unsigned": "u32", "uint32_t": "u32", "uint64_t": "u64", "size_t": "u64",
                "double": "f64", "float": "f32"}[t]

    def kind_j(t):
        if t.startswith(("Ptr", "Ref")) or t == "Cstring":
            return "ptr"
        return {"Cint": "i32", "Int32": "i32", "Cuint": "u32", "UInt32": "u32", "UInt64": "u64", "Csize_t": "u64",
                "Cdouble": "f64", "Cfloat": "f32"}[t]

    for name, ret, types in calls:
        assert name in protos, f"{name} is not declared in include/gutzwiller_b200.h"
        jt = [t for t in (x.strip() for x in types.replace("\n", " ").split(",")) if t]
        assert [kind_c(a) for a in protos[name]] == [kind_j(t) for t in jt], name
        assert ret in ("Cint", "Cstring")


def test_struct_layouts_match_the_header(gw, tmp_path):
    """sizeof / offsetof of the result structs as gcc sees the header against the ctypes mirrors"""
    import ctypes as C
    src = tmp_path / "layout.c"
    src.write_text("""#include <stdio.h>
#include <stddef.h>
#include "gutzwiller_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu\\n", sizeof(gwz_vqmc_result), offsetof(gwz_vqmc_result, walkers), offsetof(gwz_vqmc_result, acc),
           offsetof(gwz_vqmc_result, kernel_ms));
    printf("%zu %zu %zu %zu\\n", sizeof(gwz_gen_result), offsetof(gwz_gen_result, walkers), offsetof(gwz_gen_result, kernel_ms),
           offsetof(gwz_gen_result, np));
    printf("%zu %zu\\n", sizeof(gwz_gd_opts), offsetof(gwz_gd_opts, first_moment_init));
    return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    rows = [[int(x) for x in line.split()] for line in subprocess.run([str(exe)], check=True, capture_output=True,
                                                                        text=True).stdout.splitlines()]
    V, G, O = gw.capi.VqmcResult, gw.capi.GenResult, gw.capi.GdOpts
    assert rows[0] == [C.sizeof(V), V.walkers.offset, V.acc.offset, V.kernel_ms.offset]
    assert rows[1] == [C.sizeof(G), G.walkers.offset, G.kernel_ms.offset, G.np.offset]
    assert rows[2] == [C.sizeof(O), O.first_moment_init.offset]


def _c_struct_fields(name):
    """[(field, kind, count)] of `typedef struct name {...} name;` in the header; kind as in the ccall test"""
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, flags=re.S).group(1)
    macros = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define (GWZ_\w+) (\d+)", hdr)}
    kinds = {"double": "f64", "float": "f32", "int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64",
             "int": "i32", "unsigned": "u32", "uint8_t": "u8"}
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(const\s+)?(\w+)\s+(.*)", decl)
        ctype, rest = m.group(2), m.group(3)
        for item in rest.split(","):
            item = item.strip()
            arr = re.match(r"(\w+)\[(\w+)\]", item)
            if item.startswith("*"):
                out.append((item.lstrip("* "), "ptr", 1))
            elif arr:
                cnt = macros[arr.group(2)] if arr.group(2) in macros else int(arr.group(2))
                out.append((arr.group(1), kinds[ctype], cnt))
            elif ctype in kinds:
                out.append((item, kinds[ctype], 1))
            else:
                out.append((item, "struct:" + ctype, 1))
    return out


def _julia_struct_fields(jl, name):
    body = re.search(r"(?:mutable )?struct %s\b[^\n]*\n(.*?)\nend" % name, jl, flags=re.S).group(1)
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"const (GWZ_\w+) = (\d+)", jl)}
    kinds = {"Cdouble": "f64", "Cfloat": "f32", "Int32": "i32", "UInt32": "u32", "Cint": "i32", "Cuint": "u32",
             "Int64": "i64", "UInt64": "u64", "UInt8": "u8"}
    out = []
    for piece in re.split(r"[;\n]", body):
        piece = piece.split("#")[0].strip()
        if not piece:
            continue
        fname, ftype = [x.strip() for x in piece.split("::")]
        tup = re.match(r"NTuple\{(\w+),(\w+)\}", ftype)
        if ftype.startswith("Ptr"):
            out.append((fname, "ptr", 1))
        elif tup:
            cnt = consts[tup.group(1)] if tup.group(1) in consts else int(tup.group(1))
            out.append((fname, kinds[tup.group(2)], cnt))
        elif ftype in kinds:
            out.append((fname, kinds[ftype], 1))
        else:
            out.append((fname, "struct:" + ftype, 1))
    return out


def test_julia_shim_structs_match_the_header():
    """Field order, types and array lengths of every struct the Julia shim mirrors, against the C header (the shim
    cannot run here; a layout slip would corrupt memory silently on the first ccall)."""
    jl = open(os.path.join(ROOT, "gutzwiller.jl_b200", "julia", "GutzwillerB200.jl")).read()
    pairs = {"VqmcResult": "gwz_vqmc_result", "BlockingSummary": "gwz_blocking_summary", "GdOpts": "gwz_gd_opts",
             "GdTrace": "gwz_gd_trace", "GenResult": "gwz_gen_result"}
    for jname, cname in pairs.items():
        c, j = _c_struct_fields(cname), _julia_struct_fields(jl, jname)
        cj = [(f, k if not k.startswith("struct:") else "struct:" + {v: kk for kk, v in pairs.items()}[k[7:]], n)
              for f, k, n in c]
        assert cj == j, (jname, cj, j)
    # the constants the shim hard-codes
    hdr = open(HEADER).read()
    for name in ("GWZ_NUM_ACC", "GWZ_NUM_PARAMS"):
        assert re.search(r"const %s = (\d+)" % name, jl).group(1) == re.search(r"#define %s (\d+)" % name, hdr).group(1)
    for name in ("GWZ_RUN_KEEP_ACC", "GWZ_RUN_POOLED", "GWZ_RUN_SERIES", "GWZ_RUN_BLOCKING"):
        assert re.search(r"const %s = Cuint\((\d+)\)" % name, jl).group(1) == \
            re.search(r"#define %s (\d+)u" % name, hdr).group(1)
    # the zero-argument constructors pass one value per field, each of the field's kind (round 1 passed Int 0 to
    # tuple-typed fields: a MethodError on the first kinetic_vqmc call)
    for jname in ("VqmcResult", "BlockingSummary"):
        start = jl.index("%s() = %s(" % (jname, jname)) + len("%s() = %s(" % (jname, jname))
        depth, args, cur = 0, [], ""
        for ch in jl[start:]:
            if ch in "([{":
                depth += 1
            if ch in ")]}":
                if depth == 0:
                    break
                depth -= 1
            if ch == "," and depth == 0:
                args.append(cur.strip())
                cur = ""
            else:
                cur += ch
        args.append(cur.strip())
        fields = _julia_struct_fields(jl, jname)
        assert len(args) == len(fields), (jname, args)
        for a, (fname, kind, cnt) in zip(args, fields):
            if cnt > 1 or a.startswith(("(", "ntuple")):
                assert cnt >= 1 and (a.startswith("ntuple") or a.startswith("(")), (jname, fname, a)
            elif kind.startswith("struct:"):
                assert a == kind[7:] + "()", (jname, fname, a)
    # API surface the reference exports for this path (state.jl:265-298, qmc.jl:108-111, amsgrad.jl:105-127)
    for needle in ("Statistics.mean(res::GPUKineticVQMCResult)", "Statistics.var(res::GPUKineticVQMCResult)",
                   "kinetic_vqmc!(res::GPUKineticVQMCResult; steps=res.steps)", "function local_energy_estimator(",
                   "fix_params", "first_moment_init", "second_moment_init", "p_init isa GPUGradientDescentResult",
                   ":gwz_vqmc_run_record", ":gwz_vqmc_sample_vector", ":gwz_walkers_blocking",
                   ":gwz_adaptive_gradient_descent"):
        assert needle in jl, needle


def _build_c_consumer(tmp_path):
    exe = tmp_path / "c_consumer"
    libdir = os.path.join(ROOT, "gutzwiller.jl_b200", "lib")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_consumer.c"), "-o", str(exe), "-L", libdir, "-lgutzwiller_b200",
                    "-Wl,-rpath," + libdir, "-lm"], check=True)
    return exe


def test_c99_consumer_compiles_links_and_refuses_without_a_gpu(tmp_path):
    """tests/c_consumer.c: the header is valid C99, every symbol it uses resolves against the shared library, and
    without a CUDA device the library refuses (exit 77) instead of computing anything on the CPU."""
    exe = _build_c_consumer(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 77 and "no CUDA device" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_c99_consumer_runs_config1_sweep_and_amsgrad(tmp_path):
    """the same program on the GPU box: BASELINE config 1 with device-side error bars, the README sweep values,
    AMSGrad through gwz_gd_opts / gwz_gd_trace, fix_params and moment initialisers"""
    exe = _build_c_consumer(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "c consumer ok" in r.stdout


def test_julia_shim_blocks_and_brackets_balance():
    """The Julia shim cannot be executed here (no Julia in the image): at least every block opener at bracket depth 0
    has its `end`, and every bracket closes -- a blatant syntax slip in the unexecuted file fails this test."""
    import re
    src = open(os.path.join(ROOT, "gutzwiller.jl_b200", "julia", "GutzwillerB200.jl")).read()
    code = re.sub(r"#=.*?=#", "", src, flags=re.S)
    code = re.sub(r'"""(?:.|\n)*?"""', '""', code)
    code = re.sub(r'"(?:\\.|[^"\\\n])*"', '""', code)
    code = re.sub(r"'(?:\\.|[^'\\\n])'", "' '", code)
    code = re.sub(r"#.*", "", code)
    tok = re.compile(r"[A-Za-z_@!][A-Za-z_0-9!]*|[\(\)\[\]\{\}]|\n|.")
    openers = {"function", "if", "for", "while", "struct", "let", "begin", "do", "try", "module", "quote", "macro"}
    pairs = {")": "(", "]": "[", "}": "{"}
    stack, brackets, line, prev = [], [], 1, None
    for m in tok.finditer(code):
        t = m.group(0)
        if t == "\n":
            line += 1
            continue
        if t in "([{":
            brackets.append((t, line))
        elif t in ")]}":
            assert brackets and brackets[-1][0] == pairs[t], f"unbalanced {t} at line {line}"
            brackets.pop()
        elif not brackets and prev not in (".", ":"):  # inside brackets: comprehensions, generators, a[end]
            if t in openers:
                stack.append((t, line))
            elif t == "end":
                assert stack, f"`end` without an opener at line {line}"
                stack.pop()
        if not t.isspace():
            prev = t
    assert not brackets, f"unclosed bracket opened at line {brackets[-1][1]}"
    assert not stack, f"unclosed `{stack[-1][0]}` opened at line {stack[-1][1]}"
