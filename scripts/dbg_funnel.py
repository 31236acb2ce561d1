import sys, numpy as np
sys.path.insert(0, ".")
from walnuts_b200 import ChainBatch
n = 65536
rng = np.random.default_rng(31)
q0 = np.empty((n, 11)); q0[:, 0] = 3.0 * rng.standard_normal(n); q0[:, 1:] = np.exp(0.5 * q0[:, :1]) * rng.standard_normal((n, 10))
for integ in ("R2P", "D", "fixed"):
    with ChainBatch("funnel", 11, n, integrator=integ, H0=0.3, delta=0.3, M=12, seed=23, dg=0) as cb:
        cb.set_state(q0)
        for k in range(5):
            out = cb.run(20, draws=False, diag=True, nevals=False)
            q = cb.get_state(); w = q[:, 0]
            z = q[:, 1] * np.exp(-0.5 * w)
            se = 3 / np.sqrt(n)
            print(integ, "iter", 20 * (k + 1), "mean w %.4f (%.1f SE) var w %.3f (%.1f SE) | z var %.4f (%.1f SE) stop999 %d" % (
                w.mean(), w.mean() / se, w.var(), (w.var() - 9) / (9 * np.sqrt(2 / n)), z.var(), (z.var() - 1) / np.sqrt(2 / n), (out["diag"][..., 19] == 999).sum()))
