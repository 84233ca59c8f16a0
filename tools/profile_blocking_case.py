#!/usr/bin/env python
"""One val_err_and_grad configuration for ncu (-k regex:series_blocking picks the blocking kernel):
    python tools/profile_blocking_case.py N M u g log2_walkers steps"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

N, M, u, g, lw, steps = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
H = gw.HubbardReal1D(gw.near_uniform(N, M), u=u)
w = gw.Walkers(H.model(), 1 << lw, H.address.words, seed=1)
w.run([g], 100)
for _ in range(3):
    r = w.run([g], steps, blocking=True)
b = r.blocking
gb = (1 << lw) * (16.0 * steps + 80.0) / 1e9
print(f"walk {r.kernel_ms:.3f} ms, blocking {b.kernel_ms:.3f} ms = {gb / (b.kernel_ms * 1e-3):.0f} GB/s algorithmic; "
      f"err {b.err_ok:.3e} failed {b.failed} k {b.k_min}..{b.k_max}")
