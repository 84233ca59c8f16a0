"""Kernel and call time of the ranked LocalEnergyEvaluator sweep for a few system sizes: the first sweep of a handle
(enumerates its rank range and tests every state for being an orbit representative) and the repeated sweeps (evaluate
the representative basis stored at the second call)."""
import os
import sys
import time

os.environ.setdefault("GWZ_SWEEP_CACHE_AT", "2")  # default: the call at which the sweeps so far would have paid for it

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

for n, m in ((12, 12), (14, 14), (16, 16), (24, 8), (8, 24), (18, 18)):
    H = gw.HubbardReal1D(gw.near_uniform(n, m), u=2.0)
    gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H)).val_and_grad(1.0)  # warm the context (tables, module load)
    le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
    le.val_and_grad(1.0)
    first_ms = le.last_kernel_ms()
    t0 = time.perf_counter()
    le.val_and_grad(1.0)  # builds the representative basis
    build_ms = (time.perf_counter() - t0) * 1e3
    for _ in range(3):
        le.val_and_grad(1.0)
    ms = min(le.val_and_grad(1.0) and le.last_kernel_ms() for _ in range(5))
    t0 = time.perf_counter()
    for _ in range(10):
        le.val_and_grad(1.0)
    dt = (time.perf_counter() - t0) / 10
    print(f"BoseFS{{{n},{m}}} {le.dim} states: first sweep kernel {first_ms:.4f} ms = {le.dim / first_ms * 1e3:.4g} states/s; "
          f"second call (stores the representatives) {build_ms:.3f} ms; repeated sweeps kernel {ms:.4f} ms = "
          f"{le.dim / ms * 1e3:.4g} states/s effective, {dt * 1e3:.4f} ms per call", flush=True)
