"""Regenerates tests/golden/sweep_vectors.json from the CPU oracle (oracle/gutz_oracle.c, mode 1 =
compensated long-double sums).  The reference itself (Julia + Rimu.jl) cannot run in this image; the
oracle is pinned bit-for-bit to the README values first (tests/test_oracle_golden.py).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle  # noqa: E402

CASES = [  # (N, M, u, t, g)
    (12, 12, 2.0, 1.0, 0.0), (12, 12, 2.0, 1.0, 0.3331889106855026), (12, 12, 2.0, 1.0, 1.0),  # BASELINE config 2
    (10, 10, 2.0, 1.0, 1.0), (5, 5, 1.0, 0.1, 0.5), (4, 4, 1.0, 1.0, 0.5), (8, 11, 4.0, 1.0, 0.15),
    (3, 2, 1.0, 1.0, 0.3),
]

out = {"_source": "oracle/gutz_oracle.c go_lee_val_and_grad mode=1 (see make_golden.py)", "cases": []}
bases = {}
for N, M, u, t, g in CASES:
    o = Oracle(N, M, u, t)
    if (N, M) not in bases:
        bases[(N, M)] = o.build_basis()
    basis = bases[(N, M)]
    val, grad = o.lee_val_and_grad(basis, g, mode=1)
    # the four sums (num, den, num_grad, den_grad), Kahan-compensated in long double
    acc = np.zeros(4, dtype=np.longdouble)
    comp = np.zeros(4, dtype=np.longdouble)
    for b in basis:
        y = o.lee_terms(b, g).astype(np.longdouble) - comp
        s = acc + y
        comp = (s - acc) - y
        acc = s
    out["cases"].append(dict(N=N, M=M, u=u, t=t, g=g, dim=len(basis), val=val, grad=grad,
                             acc=[float(x) for x in acc]))
    print(N, M, g, len(basis), val, grad, flush=True)
json.dump(out, open(os.path.join(os.path.dirname(__file__), "sweep_vectors.json"), "w"), indent=1)
