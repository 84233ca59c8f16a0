import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
import gutzwiller_b200 as gw
H = gw.HubbardReal1D(gw.near_uniform(50, 50), u=2.0)
ans = gw.RelativeJastrowAnsatz(H)
P = ans.N_PARAMS
rng = np.random.default_rng(1)
p = np.zeros(P); p[0] = 0.33; p[1:] = 0.01 * rng.random(P - 1)
w = gw.GenWalkers(ans.handle(), 1 << 16, H.address.words, seed=1)
for _ in range(3):
    r = w.run(p, 64)
print(r.kernel_ms)
