"""Throughput of the stored-operator walker and sweep kernels (csrc/gwz_sparse.cu) on synthetic operators.

    python tools/sparse_bench.py [--out gpurun_out/sparse_bench.json]

Rows of fixed length L with uniformly random targets: the gather pattern of a large unstructured basis.  Reports
walker-steps/s, hops/s and the algorithmic bytes moved (20 B of row data per hop: column, matrix element, expanded amplitude) against
the measured HBM peak when the operator does not fit in L2."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic(dim, L, seed=0):
    rng = np.random.default_rng(seed)
    row_ptr = np.arange(dim + 1, dtype=np.uint64) * np.uint64(L)
    col = rng.integers(0, dim, size=dim * L, dtype=np.uint32)
    melem = -rng.random(dim * L)
    diag = rng.normal(size=dim)
    return row_ptr, col, melem, diag


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--walkers", type=int, default=1 << 18)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--quick", action="store_true", help="fewer cases (kernel variant comparisons)")
    args = ap.parse_args()
    import gutzwiller_b200 as gw
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    rows = []
    cases = ((100_000, 8, (4, 8)), (100_000, 24, (8, 16, 32)), (100_000, 96, (16, 32)), (10_000_000, 24, (8, 16, 32)))
    if args.quick:
        cases = ((100_000, 8, (4,)), (100_000, 24, (8, 16)), (10_000_000, 24, (4, 8, 16)))
    for dim, L, groups in cases:
        op = synthetic(dim, L)
        for G in groups:
            os.environ["GWZ_SPARSE_GROUP"] = str(G)
            H = gw.SparseHamiltonian(*op)
            ans = gw.SparseGutzwillerAnsatz(H)
            h = H.handle()
            os.environ.pop("GWZ_SPARSE_GROUP", None)
            ws = gw.SparseWalkers(h, ans, args.walkers, 0, seed=1)
            ws.run([0.1], 20)  # warm-up, spreads the walkers over the basis
            best = None
            for _ in range(3):
                r = ws.run([0.1], args.steps)
                best = r if best is None or r.kernel_ms < best.kernel_ms else best
            sps = best.steps / (best.kernel_ms * 1e-3)
            hps = best.hops / (best.kernel_ms * 1e-3)
            le = gw.LocalEnergyEvaluator(H, ans)
            le.val_and_grad([0.1])
            t = []
            for _ in range(3):
                le.val_and_grad([0.1])
                t.append(le.last_kernel_ms())
            sweep_ms = min(t)
            row = dict(dim=dim, row_length=L, lanes_per_walker=h.info()["lanes_per_walker"], walkers=args.walkers,
                       steps=args.steps, walker_steps_per_s=sps, hops_per_s=hps, algorithmic_GBps=hps * 20 / 1e9,
                       sweep_ms=sweep_ms, sweep_rows_per_s=dim / (sweep_ms * 1e-3),
                       sweep_algorithmic_GBps=dim * L * 28 / (sweep_ms * 1e-3) / 1e9)
            print(json.dumps({k: (float(f"{v:.4g}") if isinstance(v, float) else v) for k, v in row.items()}), flush=True)
            rows.append(row)
            del ws, le, H, ans, h
    out = dict(rows=rows, peaks=peaks, note="algorithmic bytes: walker 20 B row data per hop (col, melem, expanded amplitude); sweep "
               "val_and_grad P=1: 12 B row data + two 8 B gathers per entry")
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
