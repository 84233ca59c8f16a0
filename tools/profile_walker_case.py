#!/usr/bin/env python
"""One walker-kernel configuration for ncu: python tools/profile_walker_case.py N M u g log2_walkers steps"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

N, M, u, g, lw, steps = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
H = gw.HubbardReal1D(gw.near_uniform(N, M), u=u)
w = gw.Walkers(H.model(), 1 << lw, H.address.words, seed=1)
for _ in range(4):
    r = w.run([g], steps)
print(r.kernel_ms, (1 << lw) * steps / r.kernel_ms / 1e6, "G steps/s")
