#!/usr/bin/env python
"""Throughput of the ansatz-generic engine (walker-steps/s, sweep states/s) next to the CPU oracle port.

    python tools/gen_bench.py [--cpu]     # --cpu also times oracle/gutz_oracle_ansatz.c on the host cores
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

CASES = [  # kind class, N, M, u, v, log2 walkers, steps
    ("RelativeJastrowAnsatz", 10, 10, 2.0, 0.0, 16, 256),
    ("RelativeJastrowAnsatz", 20, 20, 4.0, 0.0, 16, 128),
    ("RelativeJastrowAnsatz", 50, 50, 2.0, 0.0, 16, 64),
    ("DensityProfileAnsatz", 50, 50, 2.0, 0.0, 16, 64),
    ("GenericGutzwillerAnsatz", 50, 50, 2.0, 0.5, 16, 64),
    ("MultinomialAnsatz", 50, 50, 2.0, 0.0, 16, 64),
    ("RelativeJastrowAnsatz", 100, 100, 2.0, 0.0, 15, 64),
    ("JastrowAnsatz", 20, 20, 2.0, 0.0, 14, 64),
]
ORACLE_KIND = {"RelativeJastrowAnsatz": "relative_jastrow", "DensityProfileAnsatz": "density_profile",
               "GenericGutzwillerAnsatz": "gutzwiller", "MultinomialAnsatz": "multinomial", "JastrowAnsatz": "jastrow"}
cpu = "--cpu" in sys.argv
report = {}
for cls, N, M, u, v, lw, steps in CASES:
    addr = gw.near_uniform(N, M)
    H = gw.ExtendedHubbardReal1D(addr, u=u, v=v) if v else gw.HubbardReal1D(addr, u=u)
    ans = getattr(gw, cls)(H)
    P = ans.N_PARAMS
    rng = np.random.default_rng(1)
    p = np.full(P, 0.0)
    if cls == "RelativeJastrowAnsatz":
        p[0] = 0.33 * u / 2  # the Gutzwiller point g = 0.33 plus small longer-range terms
        p[1:] = 0.01 * rng.random(P - 1)
    elif cls == "JastrowAnsatz":
        p[:] = 0.005 * rng.random(P)
    elif cls == "DensityProfileAnsatz":
        p[:] = 0.1 * rng.random(P)
    else:
        p[:] = 0.33
    w = gw.GenWalkers(ans.handle(), 1 << lw, H.address.words, seed=1)
    for _ in range(2):
        r = w.run(p, steps)
    ms = []
    for _ in range(3):
        r = w.run(p, steps)
        ms.append(r.kernel_ms)
    k = min(ms)
    rate = (1 << lw) * steps / k * 1e3
    line = dict(P=P, walkers=1 << lw, steps=steps, kernel_ms=k, walker_steps_per_s=rate, hops_per_step=r.hops / r.steps,
                energy=r.pooled_val)
    if cpu:
        from oracle import oracle as om
        o = om.Oracle(N, M, u, 1.0)
        ao = om.AnsatzOracle(o, ORACLE_KIND[cls], v=v)
        nthreads = os.cpu_count()
        addrs = np.tile(np.pad(H.address.words, (0, 4 - o.W)), (2 * nthreads, 1))
        s_cpu = 8
        t0 = time.perf_counter()
        ao.kinetic_vqmc(p, addrs, s_cpu, seed=1, n_threads=nthreads)
        dt = time.perf_counter() - t0
        while dt < 1.0 and s_cpu < 1 << 20:
            s_cpu *= 4
            t0 = time.perf_counter()
            ao.kinetic_vqmc(p, addrs, s_cpu, seed=1, n_threads=nthreads)
            dt = time.perf_counter() - t0
        line.update(cpu_walker_steps_per_s=2 * nthreads * s_cpu / dt, cpu_cores=nthreads,
                    cpu_sample=f"{2 * nthreads} walkers x {s_cpu} steps ({dt:.1f} s)")
    report[f"{cls}_N{N}M{M}"] = line
    print(f"{cls:26s} N={N:3d} M={M:3d} P={P:4d} 2^{lw} x {steps}: {k:8.3f} ms  {rate / 1e9:7.3f} G steps/s"
          + (f"   cpu {line['cpu_walker_steps_per_s']:.3g} steps/s on {line['cpu_cores']} cores" if cpu else ""), flush=True)

# LocalEnergyEvaluator, BASELINE config 2 shape with the 6-parameter RelativeJastrow ansatz
H = gw.HubbardReal1D(gw.near_uniform(12, 12), u=2.0)
ans = gw.RelativeJastrowAnsatz(H)
le = gw.LocalEnergyEvaluator(H, ans)
p = np.array([0.3, 0.05, 0.02, 0.01, 0.0, -0.01])
for _ in range(2):
    le.val_and_grad(p)
ms = []
for _ in range(3):
    le.val_and_grad(p)
    ms.append(le.last_kernel_ms())
report["sweep_RelativeJastrow_N12M12"] = dict(states=le.dim, kernel_ms=min(ms), states_per_s=le.dim / min(ms) * 1e3)
print(f"sweep RelativeJastrow {{12,12}}: {le.dim} states {min(ms):8.3f} ms  {le.dim / min(ms) / 1e6:8.3f} G states/s", flush=True)
print(json.dumps(report))
