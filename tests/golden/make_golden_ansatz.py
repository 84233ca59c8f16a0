"""Regenerates tests/golden/ansatz_vectors.json from part 2 of the CPU oracle (oracle/gutz_oracle_ansatz.c): per-address
amplitudes, gradients, local energies, residence times, hop rates and LocalEnergyEvaluator values / gradients for the
reference's other ansaetze (src/ansatz/*.jl) with fixed parameters.  The reference itself (Julia + Rimu.jl) cannot run in
this image; the oracle is pinned by tests/test_oracle_ansatz.py.

    python tests/golden/make_golden_ansatz.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import AnsatzOracle, Oracle  # noqa: E402

CASES = [  # kind, N, M, u, v, t, normalize
    ("gutzwiller", 5, 5, 1.0, 1.0, 0.1, False),            # GutzwillerAnsatz on ExtendedHubbardReal1D (test/ansatz.jl:46)
    ("ext_gutzwiller", 5, 5, 1.0, 1.0, 0.1, False),
    ("multinomial", 5, 5, 1.0, 0.0, 0.1, True),
    ("density_profile", 4, 6, 2.0, 0.5, 1.0, False),
    ("relative_jastrow", 6, 6, 2.0, 0.0, 1.0, False),
    ("jastrow", 5, 5, 1.0, 0.0, 0.1, False),
    (("gutzwiller", "multinomial"), 5, 5, 1.0, 0.0, 0.1, False),   # test/ansatz.jl:58
]
out = {"_source": "oracle/gutz_oracle_ansatz.c (see make_golden_ansatz.py)", "cases": []}
for kind, N, M, u, v, t, normalize in CASES:
    o = Oracle(N, M, u, t)
    ao = AnsatzOracle(o, kind, v=v, normalize=normalize)
    rng = np.random.default_rng(2024)
    p = np.round(rng.random(ao.P) * 0.4, 6)
    basis = o.build_basis()
    pick = sorted(int(i) for i in rng.choice(len(basis), size=8, replace=False))
    rows = []
    for i in pick:
        a = basis[i]
        val, grad = ao.val_and_grad(a, p)
        ks = ao.kinetic_sample(a, p, 0.37)
        rows.append(dict(addr=[int(x) for x in a[: o.W]], val=val, grad=[float(x) for x in grad], e_loc=ks["eloc"],
                         tau=ks["tau"], grad_ratio=[float(x) for x in ks["grad"]], n_hops=ks["n_hops"],
                         chosen=ks["chosen"], rates=[float(x) for x in ao.hop_rates(a, p)]))
    lv, lg = ao.lee_val_and_grad(basis, p)
    out["cases"].append(dict(kind=list(kind) if isinstance(kind, tuple) else kind, N=N, M=M, u=u, v=v, t=t,
                             normalize=normalize, params=[float(x) for x in p], dim=len(basis), le_val=lv,
                             le_grad=[float(x) for x in lg], addresses=rows))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ansatz_vectors.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)
