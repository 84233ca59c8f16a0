#!/usr/bin/env python
"""Kernel-level micro-benchmark for iterating on the walker / sweep kernels (not the headline bench)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

CASES = [  # N, M, u, g, log2 walkers, steps
    (10, 10, 2.0, 1.0, 20, 128),
    (20, 20, 4.0, 0.25, 20, 95),
    (50, 50, 2.0, 1.0, 22, 64),
    (50, 50, 2.0, 0.33, 22, 64),
    (100, 100, 2.0, 1.0, 21, 64),
    (100, 100, 2.0, 0.33, 21, 64),
]
only = sys.argv[1:] if len(sys.argv) > 1 else None
for N, M, u, g, lw, steps in CASES:
    if only and f"{N}" not in only:
        continue
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=u)
    w = gw.Walkers(H.model(), 1 << lw, H.address.words, seed=1)
    for _ in range(3):
        r = w.run([g], steps)
    ms = []
    for _ in range(5):
        r = w.run([g], steps)
        ms.append(r.kernel_ms)
    k = min(ms)
    print(f"walker N={N:3d} M={M:3d} u={u} g={g:4.2f} 2^{lw} x {steps}: {k:8.3f} ms  {(1 << lw) * steps / k / 1e6:9.2f} G steps/s"
          f"  hops/step={r.hops / r.steps:6.1f}  E={r.pooled_val:.4f}", flush=True)
if not only:
    for n in (12, 16):
        H = gw.HubbardReal1D(gw.near_uniform(n, n), u=2.0)
        le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
        for _ in range(3):
            le.val_and_grad(1.0)
        ms = []
        for _ in range(5):
            le.val_and_grad(1.0)
            ms.append(le.last_kernel_ms())
        print(f"sweep  N=M={n}: {le.dim} states {min(ms):8.4f} ms  {le.dim / min(ms) / 1e6:8.2f} G states/s", flush=True)
