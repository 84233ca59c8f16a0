#!/usr/bin/env python
"""Runs the five BASELINE.json configs end to end through the public API and prints one JSON report.

    python tools/run_configs.py                      # configs 1-4 on one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/run_configs.py --cfg5   # config 5, sharded
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg5", action="store_true")
args = ap.parse_args()
report = {}

if args.cfg5:
    world, rank = gw.init_distributed()
    H = gw.HubbardReal1D(gw.near_uniform(100, 100), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    for g in (1.0, 0.33):
        qmc = gw.KineticVQMC(H, ans, walkers=1 << 24, steps=1000, record=False)
        gw.val_and_grad(qmc, g)  # warm-up block (also thermalises)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            val, grad = gw.val_and_grad(qmc, g)
        dt = (time.perf_counter() - t0) / reps
        r = qmc.result.last
        report[f"cfg5_bose100_g{g}"] = dict(n_gpus=world, walkers=1 << 24, steps_per_block=1000, s_per_block=dt,
                                            walker_steps_per_s=(1 << 24) * 1000 / dt, val=r.val, err=r.err,
                                            pooled_val=r.pooled_val, grad=r.grad[0], pooled_grad=r.pooled_grad[0],
                                            hops_per_step=r.hops / r.steps)
    if rank == 0:
        print(json.dumps(report), flush=True)
    sys.exit(0)

# ---- config 1: README example, 1 walker, 1e5 steps ----
H = gw.HubbardReal1D(gw.near_uniform(10, 10), u=2.0)
ans = gw.GutzwillerAnsatz(H)
t0 = time.perf_counter()
res = gw.kinetic_vqmc(H, ans, [1.0], steps=1e5, walkers=1)
dt = time.perf_counter() - t0
est = gw.local_energy_estimator(res)
le = gw.LocalEnergyEvaluator(H, ans)
exact, egrad = le.val_and_grad([1.0])
report["cfg1_readme_bose10_1walker_1e5steps"] = dict(
    mean=est.mean, err=est.err, exact_sweep=exact, sigmas=(est.mean - exact) / est.err, seconds=dt,
    readme="-7.8369 +- 0.016555 (README.md:218)")

# ---- config 2: LocalEnergyEvaluator full-basis sweep + val_and_grad, BoseFS{12,12} ----
H = gw.HubbardReal1D(gw.near_uniform(12, 12), u=2.0)
le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
out = {}
for g in (0.0, 0.3331889106855026, 1.0):
    for _ in range(3):
        v, gr = le.val_and_grad(g)
    t0 = time.perf_counter()
    for _ in range(100):
        v, gr = le.val_and_grad(g)
    dt = (time.perf_counter() - t0) / 100
    out[f"g={g}"] = dict(val=v, grad=float(gr[0]), ms_per_call=dt * 1e3, kernel_ms=le.last_kernel_ms(),
                         states_per_s=le.dim / dt)
report["cfg2_sweep_bose12"] = dict(states=le.dim, runs=out)

# ---- config 3: KineticVQMC samples=1e8, 2^20 walkers, BoseFS{20,20} u=4 ----
H = gw.HubbardReal1D(gw.near_uniform(20, 20), u=4.0)
ans = gw.GutzwillerAnsatz(H)
out = {}
for g in (0.25, 1.0):
    qmc = gw.KineticVQMC(H, ans, samples=1e8, walkers=1 << 20, record=False)
    for _ in range(5):
        gw.val_and_grad(qmc, g)
    t0 = time.perf_counter()
    for _ in range(10):
        val, grad = gw.val_and_grad(qmc, g)
    dt = (time.perf_counter() - t0) / 10
    r = qmc.result.last
    out[f"g={g}"] = dict(steps_per_walker=qmc.steps, ms_per_call=dt * 1e3, kernel_ms=r.kernel_ms,
                         walker_steps_per_s=qmc.steps * qmc.walkers / dt, val=r.val, err=r.err, pooled_val=r.pooled_val,
                         grad=r.grad[0], pooled_grad=r.pooled_grad[0], hops_per_step=r.hops / r.steps)
    # raw throughput with the launch amortised over longer blocks (1e3 and 1e4 steps per walker)
    for long_steps in (1000, 10000):
        t0 = time.perf_counter()
        val, grad = gw.val_and_grad(qmc, g, steps=long_steps)
        dt = time.perf_counter() - t0
        out[f"g={g}"][f"walker_steps_per_s_{long_steps}_steps"] = long_steps * qmc.walkers / dt
report["cfg3_bose20_2^20walkers_1e8samples"] = out

# ---- config 4: AMSGrad on KineticVQMC, samples=1e7 per evaluation, BoseFS{50,50} ----
H = gw.HubbardReal1D(gw.near_uniform(50, 50), u=2.0)
ans = gw.GutzwillerAnsatz(H)
out = {}
for estimator in ("walker_mean", "pooled"):
    qmc = gw.KineticVQMC(H, ans, samples=1e7, walkers=1 << 16, record=False, burn_in=500, estimator=estimator)
    t0 = time.perf_counter()
    tr = gw.amsgrad(qmc, [1.0])
    dt = time.perf_counter() - t0
    out[estimator] = dict(iterations=tr.iterations, reason=tr.reason, seconds=dt, ms_per_iteration=1e3 * dt / (tr.iterations + 1),
                          steps_per_walker=qmc.steps, first_param_delta=float(tr.param_delta[0, 0]),
                          last_param=float(tr.param[-1, 0]), last_value=float(tr.value[-1]), last_error=float(tr.error[-1]),
                          last_gradient=float(tr.gradient[-1, 0]), params_every_10=[float(x) for x in tr.param[::10, 0]])
report["cfg4_amsgrad_bose50_1e7samples"] = out
print(json.dumps(report), flush=True)
