#!/usr/bin/env python
"""Sweep-kernel micro-benchmark (LocalEnergyEvaluator, ranked enumeration)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [12, 14, 16]
for n in sizes:
    H = gw.HubbardReal1D(gw.near_uniform(n, n), u=2.0)
    le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    for _ in range(3):
        le.val_and_grad(1.0)
    ms = []
    for _ in range(5):
        v, g = le.val_and_grad(1.0)
        ms.append(le.last_kernel_ms())
    print(f"sweep N=M={n}: {le.dim} states {min(ms):9.4f} ms {le.dim / min(ms) / 1e6:8.2f} G states/s val={v:.10f}", flush=True)
