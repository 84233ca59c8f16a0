#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family, meant for compute-sanitizer (closed on this pool) and as a coverage smoke run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

for N, M in ((10, 10), (50, 50), (7, 3)):
    H = gw.HubbardReal1D(gw.near_uniform(N, M), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    w = gw.Walkers(H.model(), 300, H.address.words, seed=3)
    w.run([0.4], 40, burn_in=3)
    w.run([0.4], 20, keep_acc=True)
    w.run_record([0.4], 10)
    w.run([0.4], 10, kernel=gw.KERNEL_ENUMERATE)
    w.estimates()
    H.model().probe(w.addresses()[:50], [0.4], np.linspace(0, 0.99, 50), kernel=gw.KERNEL_ENUMERATE, want_rates=True)
    H.model().probe(w.addresses()[:50], [0.4], np.linspace(0, 0.99, 50))
H = gw.HubbardReal1D(gw.near_uniform(8, 8), u=2.0)
le = gw.LocalEnergyEvaluator(H, gw.GutzwillerAnsatz(H))
le.val_and_grad(0.4)
le.terms(0.4, 0, 100)
gw.Walkers(H.model(), 64, H.address.words, seed=1).sample_vector([0.4], 50)
for cls, M in (("RelativeJastrowAnsatz", 8), ("DensityProfileAnsatz", 8), ("JastrowAnsatz", 6), ("MultinomialAnsatz", 8)):
    Hg = gw.ExtendedHubbardReal1D(gw.near_uniform(M, M), u=2.0, v=0.3)
    a = getattr(gw, cls)(Hg)
    p = np.linspace(0.2, 0.01, a.N_PARAMS)
    g = gw.GenWalkers(a.handle(), 70, Hg.address.words, seed=2)
    g.run(p, 300, burn_in=5)
    g.run(p, 10, keep_acc=True)
    g.run_record(p, 5)
    a.handle().probe(g.addresses()[:20], p, np.linspace(0, 0.99, 20), want_rates=True)
    gw.LocalEnergyEvaluator(Hg, a).val_and_grad(p)
Hc = gw.HubbardReal1D(gw.near_uniform(6, 6), u=2.0)
combo = gw.GutzwillerAnsatz(Hc) + gw.MultinomialAnsatz(Hc)
gw.GenWalkers(combo.handle(), 40, Hc.address.words, seed=2).run([0.3, 0.4, 0.6], 50)
gw.LocalEnergyEvaluator(Hc, combo).val_and_grad([0.3, 0.4, 0.6])
addrs, vals = gw.collect_to_vec(gw.GutzwillerAnsatz(Hc), [0.3])
va = gw.VectorAnsatz((addrs, vals))
gw.kinetic_vqmc(Hc, va, [], walkers=40, steps=30, record=True)
gw.LocalEnergyEvaluator(Hc, va)([])
big = gw.HubbardReal1D(gw.near_uniform(120, 131), u=2.0)
ab = gw.DensityProfileAnsatz(big)
gw.GenWalkers(ab.handle(), 9, big.address.words, seed=2).run(np.full(131, 0.01), 20)
print("all kernel families exercised")
