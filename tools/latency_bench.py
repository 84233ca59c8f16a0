#!/usr/bin/env python
"""Per-call latency of the two objectives: val_and_grad(qmc, p) at the BASELINE config 4 shape and
val_and_grad(le, p) on BoseFS{12,12} (config 2), through the Python mirror and straight through the C ABI."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402
from gutzwiller_b200 import capi  # noqa: E402

L = gw.lib()
H = gw.HubbardReal1D(gw.near_uniform(50, 50), u=2.0)
ans = gw.GutzwillerAnsatz(H)
for lw in (16, 18):
    qmc = gw.KineticVQMC(H, ans, samples=1e7, walkers=1 << lw, record=False)
    for _ in range(5):
        gw.val_and_grad(qmc, 0.5)
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        gw.val_and_grad(qmc, 0.5)
    dt = (time.perf_counter() - t0) / n
    h = qmc.result.walkers_handle.handle
    p = (C.c_double * 1)(0.5)
    r = capi.VqmcResult()
    t0 = time.perf_counter()
    for _ in range(n):
        L.gwz_vqmc_run(h, p, qmc.steps, 0, 1, 0, C.byref(r))
    dc = (time.perf_counter() - t0) / n
    print(f"val_and_grad(qmc): 2^{lw} walkers x {qmc.steps} steps: python {dt * 1e6:7.1f} us, C ABI {dc * 1e6:7.1f} us, "
          f"kernel {r.kernel_ms * 1e3:7.1f} us -> overhead {dc * 1e6 - r.kernel_ms * 1e3:5.1f} us", flush=True)
    for _ in range(5):  # the first call allocates the series buffers and loads the recording kernels
        L.gwz_vqmc_run(h, p, qmc.steps, 0, 1, capi.RUN_BLOCKING, C.byref(r))
    t0 = time.perf_counter()
    for _ in range(n):
        L.gwz_vqmc_run(h, p, qmc.steps, 0, 1, capi.RUN_BLOCKING, C.byref(r))
    db = (time.perf_counter() - t0) / n
    print(f"val_err_and_grad(qmc): C ABI {db * 1e6:7.1f} us (walk {r.kernel_ms * 1e3:.1f} + blocking "
          f"{r.blocking.kernel_ms * 1e3:.1f} us)", flush=True)
for nm in (12, 16):
    Hs = gw.HubbardReal1D(gw.near_uniform(nm, nm), u=2.0)
    le = gw.LocalEnergyEvaluator(Hs, gw.GutzwillerAnsatz(Hs))
    for _ in range(80):
        le.val_and_grad(1.0)
    n = 300 if nm == 12 else 30
    t0 = time.perf_counter()
    for _ in range(n):
        le.val_and_grad(1.0)
    dt = (time.perf_counter() - t0) / n
    p = (C.c_double * 1)(1.0)
    v, g = C.c_double(), (C.c_double * 1)()
    t0 = time.perf_counter()
    for _ in range(n):
        L.gwz_sweep_val_and_grad(le.handle, p, C.byref(v), g, None)
    dc = (time.perf_counter() - t0) / n
    print(f"val_and_grad(le) BoseFS{{{nm},{nm}}}: python {dt * 1e6:7.1f} us, C ABI {dc * 1e6:7.1f} us, kernel "
          f"{le.last_kernel_ms() * 1e3:6.1f} us", flush=True)
qmc = gw.KineticVQMC(H, ans, samples=1e7, walkers=1 << 16, record=False)
gw.amsgrad(qmc, [1.0], maxiter=2)  # buffers allocated, kernels loaded
t0 = time.perf_counter()
tr = gw.amsgrad(qmc, [1.0], maxiter=100)
dt = time.perf_counter() - t0
print(f"amsgrad(qmc) 101 iterations x 1e7 samples with blocked error bars: {dt * 1e3:.1f} ms, p = {tr.param[-1]}, "
      f"E = {tr.value[-1]:.4f} +- {tr.error[-1]:.4f}")
