#!/usr/bin/env python
"""Latency of small KineticVQMC calls (BASELINE config 4 shape: 1e7 samples per AMSGrad iteration)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

H = gw.HubbardReal1D(gw.near_uniform(50, 50), u=2.0)
ans = gw.GutzwillerAnsatz(H)
for lw in (13, 14, 15, 16, 17, 18, 20):
    walkers = 1 << lw
    qmc = gw.KineticVQMC(H, ans, samples=1e7, walkers=walkers, record=False)
    for _ in range(5):
        gw.val_and_grad(qmc, 0.5)
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        gw.val_and_grad(qmc, 0.5)
    dt = (time.perf_counter() - t0) / n
    k = qmc.result.last.kernel_ms
    steps = qmc.steps
    print(f"walkers 2^{lw} x {steps:5d} steps: call {dt * 1e3:7.3f} ms  kernel {k:7.3f} ms  "
          f"{walkers * steps / dt / 1e9:6.2f} G steps/s end to end, {walkers * steps / k / 1e6:6.2f} kernel", flush=True)
t0 = time.perf_counter()
qmc = gw.KineticVQMC(H, ans, samples=1e7, walkers=1 << 16, record=False)
tr = gw.amsgrad(qmc, [1.0], maxiter=100)
dt = time.perf_counter() - t0
print(f"amsgrad 101 iterations x 1e7 samples: {dt * 1e3:.1f} ms total, {dt / 101 * 1e3:.3f} ms / iteration, p = {tr.param[-1]}")
