#!/usr/bin/env python
"""High-statistics check of the sampler against exact values: BoseFS{12,12} (BASELINE config 2 shape), where the
LocalEnergyEvaluator gives the exact energy and gradient of the ansatz.  Prints deviations in units of the standard error."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

out = {}
H = gw.HubbardReal1D(gw.near_uniform(12, 12), u=2.0)
ans = gw.GutzwillerAnsatz(H)
le = gw.LocalEnergyEvaluator(H, ans)
for g in (0.3331889106855026, 1.0):
    exact, egrad = le.val_and_grad(g)
    vals, grads = [], []
    for seed in range(8):  # independent repetitions: the scatter of the pooled estimates is the error bar
        qmc = gw.KineticVQMC(H, ans, walkers=1 << 17, steps=2000, burn_in=500, record=False, estimator="pooled", seed=1000 + seed)
        v, gr = qmc.val_and_grad(g)
        vals.append(v)
        grads.append(gr[0])
    vals, grads = np.array(vals), np.array(grads)
    ev, eg = vals.std(ddof=1) / np.sqrt(len(vals)), grads.std(ddof=1) / np.sqrt(len(grads))
    out[f"gutzwiller_g={g}"] = dict(samples=8 * (1 << 17) * 2000, exact=exact, mean=vals.mean(), stderr=ev,
                                    sigmas=(vals.mean() - exact) / ev, exact_grad=float(egrad[0]), mean_grad=grads.mean(),
                                    stderr_grad=eg, sigmas_grad=(grads.mean() - egrad[0]) / eg)
Hj = gw.HubbardReal1D(gw.near_uniform(8, 8), u=2.0)
aj = gw.RelativeJastrowAnsatz(Hj)
p = np.array([0.3, 0.05, 0.02, 0.01])
exact, egrad = gw.LocalEnergyEvaluator(Hj, aj).val_and_grad(p)
vals, grads = [], []
for seed in range(8):
    qmc = gw.KineticVQMC(Hj, aj, walkers=1 << 14, steps=2000, burn_in=500, record=False, estimator="pooled", seed=2000 + seed)
    v, gr = qmc.val_and_grad(p)
    vals.append(v)
    grads.append(gr)
vals, grads = np.array(vals), np.array(grads)
ev = vals.std(ddof=1) / np.sqrt(len(vals))
eg = grads.std(axis=0, ddof=1) / np.sqrt(len(vals))
out["relative_jastrow_bose8"] = dict(samples=8 * (1 << 14) * 2000, exact=exact, mean=vals.mean(), stderr=ev,
                                     sigmas=(vals.mean() - exact) / ev, exact_grad=[float(x) for x in egrad],
                                     mean_grad=[float(x) for x in grads.mean(axis=0)],
                                     sigmas_grad=[float(x) for x in (grads.mean(axis=0) - egrad) / eg])
# stored-operator path: the chain {10,10} as stored rows with a 3-parameter user-defined ansatz (density-density terms)
Hc = gw.HubbardReal1D(gw.near_uniform(10, 10), u=2.0)
Hs = gw.SparseHamiltonian.from_hubbard(Hc)
n = Hs.addresses.astype(np.float64)
feats = np.stack([0.5 * (n * (n - 1)).sum(axis=1), (n * np.roll(n, -1, axis=1)).sum(axis=1),
                  (n * np.roll(n, -2, axis=1)).sum(axis=1)], axis=1)
custom = gw.TableAnsatz(Hs, lambda q: (np.exp(-feats @ q), -feats), 3)
p = np.array([0.6, 0.1, 0.03])
exact, egrad = gw.LocalEnergyEvaluator(Hs, custom).val_and_grad(p)
vals, grads = [], []
REPS = 24  # the error bar is estimated from the scatter of the repetitions: 8 of them leave heavy t-distribution tails
for seed in range(REPS):
    qmc = gw.KineticVQMC(Hs, custom, walkers=1 << 16, steps=2000, burn_in=500, record=False, estimator="pooled", seed=3000 + seed)
    v, gr = qmc.val_and_grad(p)
    vals.append(v)
    grads.append(gr)
vals, grads = np.array(vals), np.array(grads)
ev = vals.std(ddof=1) / np.sqrt(len(vals))
eg = grads.std(axis=0, ddof=1) / np.sqrt(len(vals))
out["stored_operator_bose10_custom_ansatz"] = dict(
    samples=REPS * (1 << 16) * 2000, repetitions=REPS, exact=exact, mean=vals.mean(), stderr=ev, sigmas=(vals.mean() - exact) / ev,
    exact_grad=[float(x) for x in egrad], mean_grad=[float(x) for x in grads.mean(axis=0)],
    sigmas_grad=[float(x) for x in (grads.mean(axis=0) - egrad) / eg])
print(json.dumps(out, indent=1))
