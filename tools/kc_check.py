import sys, os
sys.path.insert(0, os.getcwd())
import gutzwiller_b200 as gw
H = gw.HubbardReal1D(gw.near_uniform(12, 12), u=2.0)
ans = gw.GutzwillerAnsatz(H)
g = 0.3331889106855026
exact, egrad = gw.LocalEnergyEvaluator(H, ans).val_and_grad(g)
for seed in (1, 2):
    res = gw.kinetic_vqmc(H, ans, g, walkers=65536, steps=2000, burn_in=240, record=False, seed=seed)
    r = res.last
    print(os.environ.get("GWZ_KCAP"), seed, "val %.5f +- %.5f (exact %.5f)  grad %.5f +- %.5f pooled_grad %.5f (exact %.5f)" % (r.val, r.err, exact, r.grad[0], r.grad_err[0], r.pooled_grad[0], egrad[0]), flush=True)
