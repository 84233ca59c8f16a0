"""Multi-GPU parity script (needs >= 2 GPUs; not collected by pytest).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py            # one process per GPU, NCCL all-reduce of the accumulators
    python tests/multigpu_check.py --single-process 2   # one process, several devices (what a Julia host does)

Checks that the sharded run (global walker ids, one ncclAllReduce per run) reproduces the single-GPU run
of the same global walkers to FP reduction order, for the VQMC estimators and for the sweep.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GWZ_SWEEP_CACHE_AT", "2")  # store the representative basis at the second sweep of a handle
import gutzwiller_b200 as gw  # noqa: E402

N, M, U, G, TOTAL, STEPS = 50, 50, 2.0, 0.5, 1 << 16, 200


def reference_single_gpu(device):
    ctx = gw.Context(device)  # fresh context without communicator
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=U)
    w = gw.Walkers(H.model(ctx), TOTAL, H.address.words, seed=123)
    r = w.run([G], STEPS)
    H12 = gw.HubbardReal1D(gw.near_uniform(12, 12), u=U)
    le = gw.LocalEnergyEvaluator(H12, gw.GutzwillerAnsatz(H12), ctx=ctx)
    return r, le.accumulators(0.4)


def check(r, ref, acc, ref_acc, tag):
    assert r.walkers == TOTAL and r.steps == TOTAL * STEPS and r.hops == ref.hops, (tag, r.walkers, r.steps)
    assert abs(r.val - ref.val) <= 1e-13 * abs(ref.val), (tag, r.val, ref.val)
    assert abs(r.grad[0] - ref.grad[0]) <= 1e-10 * abs(ref.grad[0]), (tag, r.grad[0], ref.grad[0])
    assert abs(r.pooled_val - ref.pooled_val) <= 1e-13 * abs(ref.pooled_val), tag
    assert abs(r.err - ref.err) <= 1e-9 * ref.err, tag
    assert np.all(np.abs(acc - ref_acc) <= 1e-13 * np.abs(ref_acc)), (tag, acc, ref_acc)
    print(f"[{tag}] OK val={r.val:.12f} grad={r.grad[0]:.9f} pooled={r.pooled_val:.12f} sweep_num={acc[0]:.12e}", flush=True)


def check_generic(rank, world):
    """The ansatz-generic engine under the same sharding: 8 + 4P accumulators in one all-reduce per run."""
    Hg = gw.HubbardReal1D(gw.near_uniform(20, 20), u=U)
    ans = gw.RelativeJastrowAnsatz(Hg)
    p = np.linspace(0.3, 0.01, ans.N_PARAMS)
    total, steps = 1 << 12, 100
    res = gw.kinetic_vqmc(Hg, ans, p, walkers=total, steps=steps, seed=77, record=False)
    ctx = gw.Context(int(os.environ.get("LOCAL_RANK", "0")))  # fresh context without communicator
    w = gw.GenWalkers(ans.handle(ctx), total, Hg.address.words, seed=77)
    ref = w.run(p, steps)
    r = res.last
    assert r.walkers == total and r.steps == total * steps and r.hops == ref.hops, (r.walkers, r.steps, r.hops, ref.hops)
    assert abs(r.val - ref.val) <= 1e-12 * abs(ref.val) and abs(r.pooled_val - ref.pooled_val) <= 1e-12 * abs(ref.pooled_val)
    assert np.all(np.abs(r.grad - ref.grad) <= 1e-9 * np.abs(ref.grad).max())
    H8 = gw.HubbardReal1D(gw.near_uniform(8, 8), u=U)
    a8 = gw.RelativeJastrowAnsatz(H8)
    p8 = np.array([0.3, 0.05, 0.02, 0.01])
    le8 = gw.LocalEnergyEvaluator(H8, a8)
    v, g = le8.val_and_grad(p8)           # rank-range shards + all-reduce
    for _ in range(3):
        v2, g2 = le8.val_and_grad(p8)     # shard-wise representative basis
        ev, eg = abs(v2 - v) / abs(v), np.abs(g2 - g).max() / np.abs(g).max()
        assert ev <= 1e-12 and eg <= 1e-9, (rank, ev, eg, v, v2, g, g2)
    v1, g1 = gw.LocalEnergyEvaluator(H8, a8, ctx=ctx).val_and_grad(p8)  # one device, whole range
    assert abs(v - v1) <= 1e-12 * abs(v1) and np.all(np.abs(g - g1) <= 1e-11 * np.abs(g1).max())
    print(f"[rank {rank}/{world}] generic OK val={r.val:.12f} grad0={r.grad[0]:.9f} sweep={v:.12f}", flush=True)


def check_sparse(rank, world):
    """The stored-operator path under the same sharding: walkers by global id, sweep by row range."""
    import sparse_models as sm
    op = sm.hubbard_mom1d((0, 0, 0, 0, 3, 0, 0, 0, 0, 0), u=4.0)[:4]
    Hs = gw.SparseHamiltonian(*op)
    ans = gw.SparseGutzwillerAnsatz(Hs)
    total, steps = 1 << 12, 100
    res = gw.kinetic_vqmc(Hs, ans, [0.3], walkers=total, steps=steps, seed=5, record=False)
    ctx = gw.Context(int(os.environ.get("LOCAL_RANK", "0")))  # fresh context without communicator
    w = gw.SparseWalkers(Hs.handle(ctx), ans, total, 0, seed=5)
    ref = w.run([0.3], steps)
    r = res.last
    assert r.walkers == total and r.steps == total * steps and r.hops == ref.hops, (r.walkers, r.steps, r.hops, ref.hops)
    assert abs(r.val - ref.val) <= 1e-12 * abs(ref.val) and abs(r.pooled_val - ref.pooled_val) <= 1e-12 * abs(ref.pooled_val)
    assert np.all(np.abs(r.grad - ref.grad) <= 1e-9 * np.abs(ref.grad).max())
    begin, end = gw.shard_range(total, world, rank)
    assert np.array_equal(res.walkers_handle.indices(), w.indices()[begin:end])  # same trajectories as unsharded
    big = sm.random_operator(5000, 12, seed=3)
    Hb = gw.SparseHamiltonian(*big)
    ab = gw.SparseGutzwillerAnsatz(Hb)
    v, g = gw.LocalEnergyEvaluator(Hb, ab).val_and_grad(0.2)             # row shards + all-reduce
    v1, g1 = gw.LocalEnergyEvaluator(Hb, ab, ctx=ctx).val_and_grad(0.2)  # one device, all rows
    assert abs(v - v1) <= 1e-12 * abs(v1) and np.all(np.abs(g - g1) <= 1e-11 * max(1.0, np.abs(g1).max()))
    hist, _ = res.walkers_handle.sample_vector([0.3], 50)  # all-reduced residence-time histogram
    assert abs(hist.sum() - _.total_time) <= 1e-9 * _.total_time and _.walkers == total
    print(f"[rank {rank}/{world}] stored operator OK val={r.val:.12f} grad0={r.grad[0]:.9f} sweep={v:.12f}", flush=True)


def check_blocking(rank, world):
    """val_err_and_grad across ranks: every walker's series is blocked on the GPU that owns it, the 38 blocking sums
    travel in the same all-reduce as the 13 sampling sums; the result must be the single-GPU one."""
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=U)
    qmc = gw.KineticVQMC(H, gw.GutzwillerAnsatz(H), walkers=TOTAL, steps=128, seed=77)
    val, err, grad = qmc.val_err_and_grad([G])
    b = qmc.result.blocking_summary
    ctx = gw.Context(int(os.environ.get("LOCAL_RANK", "0")))
    w = gw.Walkers(H.model(ctx), TOTAL, H.address.words, seed=77)
    r = w.run([G], 128, blocking=True)
    ref = r.blocking
    assert b.walkers == ref.walkers == TOTAL and b.failed == ref.failed and (b.k_min, b.k_max) == (ref.k_min, ref.k_max)
    assert abs(b.mean - ref.mean) <= 1e-13 * abs(ref.mean) and abs(b.err_ok - ref.err_ok) <= 1e-12 * ref.err_ok
    assert abs(b.grad[0] - ref.grad[0]) <= 1e-10 * abs(ref.grad[0])
    assert val == b.mean and (err == b.err or (err != err and b.err != b.err))
    print(f"[rank {rank}/{world}] blocking OK mean={b.mean:.12f} err={b.err_ok:.3e} k={b.k_min}..{b.k_max}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--single-process", type=int, default=0)
    a = ap.parse_args()
    if a.single_process:
        n = a.single_process
        ctxs = [gw.Context(d) for d in range(n)]
        gw.comm_init_all(ctxs)
        H = gw.HubbardReal1D(gw.near_uniform(N, M), u=U)
        sets = []
        for d in range(n):
            b, e = gw.shard_range(TOTAL, n, d)
            sets.append(gw.Walkers(H.model(ctxs[d]), e - b, H.address.words, seed=123, global_offset=b))
        r = gw.run_group(sets, [G], STEPS)
        ref, ref_acc = reference_single_gpu(0)
        check(r, ref, ref_acc, ref_acc, f"single-process x{n}")
        return
    world, rank = gw.init_distributed()
    assert world > 1, "launch with torch.distributed.run"
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=U)
    res = gw.kinetic_vqmc(H, gw.GutzwillerAnsatz(H), [G], walkers=TOTAL, steps=STEPS, seed=123, record=False)
    H12 = gw.HubbardReal1D(gw.near_uniform(12, 12), u=U)
    le12 = gw.LocalEnergyEvaluator(H12, gw.GutzwillerAnsatz(H12))
    acc = le12.accumulators(0.4)  # rank-range shards + all-reduce
    for _ in range(3):  # later sweeps run over each shard's stored representative basis (GWZ_SWEEP_CACHE_AT=2 below)
        again = le12.accumulators(0.4)
        assert np.all(np.abs(again - acc) <= 1e-12 * np.abs(acc)), (again, acc)
    ref, ref_acc = reference_single_gpu(int(os.environ.get("LOCAL_RANK", "0")))
    check(res.last, ref, acc, ref_acc, f"rank {rank}/{world}")
    check_blocking(rank, world)
    check_generic(rank, world)
    check_sparse(rank, world)
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
