"""Regenerates tests/golden/sparse_vectors.json from part 3 of the CPU oracle (oracle/gutz_oracle_sparse.c): per-index
kinetic_sample! outputs, LocalEnergyEvaluator sums and short trajectories over the stored test operators of
tests/sparse_models.py (real-space chain, momentum-space chain, two-component periodic lattice).  The reference itself
(Julia + Rimu.jl) cannot run in this image; the oracle is pinned by tests/test_oracle_sparse.py.

    python tests/golden/make_golden_sparse.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
import sparse_models as sm  # noqa: E402
from oracle.oracle import Oracle, SparseOracle  # noqa: E402


def operators():
    rp, col, mel, diag, _ = sm.hubbard_real1d(Oracle(5, 5, 1.0, 0.1))
    return {"real": (rp, col, mel, diag),
            "mom": sm.hubbard_mom1d((0, 0, 0, 0, 3, 0, 0, 0, 0, 0), u=4.0)[:4],
            "fermi": sm.hubbard_2c_lattice((1, 0, 0, 1), (1, 1, 0, 0))[:4]}


def table(dim, P, seed):
    """a smooth positive table and features (the fixture stores nothing but the seed)"""
    rng = np.random.default_rng(seed)
    F = np.round(rng.normal(size=(dim, P)), 6)
    p = np.round(0.3 * rng.random(P), 6)
    return np.exp(-F @ p), -F


if __name__ == "__main__":
    out = {"_source": "oracle/gutz_oracle_sparse.c (see make_golden_sparse.py)", "cases": []}
    for name, op in operators().items():
        so = SparseOracle(*op)
        for kind in ("gutzwiller", "table3"):
            if kind == "gutzwiller":
                so.set_gutzwiller(0.4)
            else:
                so.set_table(*table(so.dim, 3, 77))
            rng = np.random.default_rng(5)
            idx = sorted(int(i) for i in rng.choice(so.dim, size=min(12, so.dim), replace=False))
            rows = []
            for k in idx:
                ks = so.kinetic_sample(k, 0.37)
                rows.append(dict(index=k, e_loc=ks["eloc"], tau=ks["tau"], n_hops=ks["n_hops"], chosen=ks["chosen"],
                                 next=ks["new_index"], grad=[float(x) for x in ks["grad"]]))
            v, g = so.lee_val_and_grad()
            tr = so.kinetic_vqmc(np.zeros(8, dtype=np.uint32), 120, seed=9)
            out["cases"].append(dict(operator=name, ansatz=kind, dim=so.dim, probes=rows, lee_val=v,
                                     lee_grad=[float(x) for x in g], final_indices=[int(x) for x in tr["idx"]],
                                     val_w=[float(x) for x in tr["val_w"]], hops=int(tr["hops"])))
    with open(os.path.join(HERE, "sparse_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")
