#!/usr/bin/env python
"""Are the device-side error bars (per-walker resample + blocking + M test, combined over walkers: state.jl:150-164)
calibrated?  R independent val_err_and_grad runs of the same (H, ansatz, p): the scatter of their means against the
error bar each run reports, and the means against the exact value of the LocalEnergyEvaluator."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gutzwiller_b200 as gw  # noqa: E402

out = {}
for (n, g, walkers, steps, reps) in ((12, 0.4, 1 << 10, 8192, 48), (12, 1.0, 1 << 12, 2048, 48), (10, 0.2, 1 << 8, 32768, 32)):
    H = gw.HubbardReal1D(gw.near_uniform(n, n), u=2.0)
    ans = gw.GutzwillerAnsatz(H)
    exact, _ = gw.LocalEnergyEvaluator(H, ans).val_and_grad(g)
    means, errs, fails = [], [], 0
    for seed in range(reps):
        qmc = gw.KineticVQMC(H, ans, walkers=walkers, steps=steps, burn_in=500, seed=5000 + seed)
        v, e, _ = qmc.val_err_and_grad([g])
        means.append(v)
        errs.append(e)
        fails += int(qmc.result.blocking_summary.failed) if qmc.result.blocking_summary is not None else 0
    means, errs = np.array(means), np.array(errs)
    out[f"bose{n}_g{g}_w{walkers}_s{steps}"] = dict(
        repetitions=reps, exact=exact, mean_of_means=float(means.mean()), scatter_of_means=float(means.std(ddof=1)),
        mean_reported_err=float(errs.mean()), ratio_scatter_over_reported=float(means.std(ddof=1) / errs.mean()),
        deviation_of_grand_mean_in_sigma=float((means.mean() - exact) / (means.std(ddof=1) / np.sqrt(reps))),
        chi2_per_rep=float(np.mean(((means - exact) / errs) ** 2)), walkers_failing_m_test=fails)
print(json.dumps(out, indent=1))
