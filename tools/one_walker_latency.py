"""Latency of BASELINE config 1 (1 walker, 1e5 steps, BoseFS{10,10}): where the time goes, per engine."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gutzwiller_b200 as gw  # noqa: E402

H = gw.HubbardReal1D(gw.near_uniform(10, 10), u=2.0)
for name, ans in (("fast path", gw.GutzwillerAnsatz(H)), ("generic engine", gw.GenericGutzwillerAnsatz(H))):
    for record in (True, False):
        for rep in range(2):
            t0 = time.perf_counter()
            res = gw.kinetic_vqmc(H, ans, [1.0], samples=1e5, walkers=1, record=record)
            t1 = time.perf_counter()
            if record:
                est = gw.local_energy_estimator(res)
            t2 = time.perf_counter()
        print(f"{name:15s} record={record!s:5s} kinetic_vqmc {t1 - t0:.4f} s (kernel {res.last.kernel_ms:.2f} ms)"
              f"  local_energy_estimator {t2 - t1:.4f} s  val {res.last.val:.4f}", flush=True)
for walkers in (1, 32, 1024):
    for name, ans in (("fast path", gw.GutzwillerAnsatz(H)), ("generic engine", gw.GenericGutzwillerAnsatz(H))):
        res = gw.kinetic_vqmc(H, ans, [1.0], steps=20000, walkers=walkers, record=False)
        res = gw.kinetic_vqmc(H, ans, [1.0], steps=20000, walkers=walkers, record=False)
        print(f"{name:15s} walkers={walkers:5d} 20000 steps: kernel {res.last.kernel_ms:.2f} ms = "
              f"{res.last.kernel_ms * 1e3 / 20000:.3f} us/step", flush=True)
