"""Kernel and call time of the ranked LocalEnergyEvaluator sweep for a few system sizes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

for n, m in ((12, 12), (14, 14), (16, 16), (24, 8), (8, 24), (18, 18)):
    H = gw.HubbardReal1D(gw.near_uniform(n, m), u=2.0)
    le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    for _ in range(3):
        le.val_and_grad(1.0)
    ms = min(le.val_and_grad(1.0) and le.last_kernel_ms() for _ in range(5))
    t0 = time.perf_counter()
    for _ in range(10):
        le.val_and_grad(1.0)
    dt = (time.perf_counter() - t0) / 10
    print(f"BoseFS{{{n},{m}}} {le.dim} states: kernel {ms:.4f} ms = {le.dim / ms * 1e3:.4g} states/s, call {dt * 1e3:.4f} ms", flush=True)
