"""One stored-operator case for ncu: walkers on a synthetic operator too large for L2 (dim 4e6, 24 entries per row),
then one gradient sweep.  ncu -k regex:sparse_ picks the kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import gutzwiller_b200 as gw  # noqa: E402
from sparse_bench import synthetic  # noqa: E402

H = gw.SparseHamiltonian(*synthetic(4_000_000, 24))
ans = gw.SparseGutzwillerAnsatz(H)
ws = gw.SparseWalkers(H.handle(), ans, 1 << 18, 0, seed=1)
ws.run([0.1], 20)
r = ws.run([0.1], 100)
print(f"walker-steps/s {r.steps / (r.kernel_ms * 1e-3):.4g}  hops/s {r.hops / (r.kernel_ms * 1e-3):.4g}")
le = gw.LocalEnergyEvaluator(H, ans)
le.val_and_grad([0.1])
print("sweep ms", le.last_kernel_ms())
