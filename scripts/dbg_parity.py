import sys, numpy as np
sys.path.insert(0, ".")
import walnuts_b200 as wb
from oracle import targets as ot
from oracle import walnutspy_oracle as wo
q0 = 0.5 * np.random.default_rng(2).standard_normal((3, 6))
s, d = wb.WALNUTS(wb.targets.stdGauss, q0, integrator=wb.adaptYoshidaD, numIter=30, warmupIter=30, M=8, seed=3)
np.set_printoptions(linewidth=220, precision=10)
for c in range(1):
    so, do = wo.WALNUTS(ot.std_normal, q0[c], integrator=wo.ADAPT_YOSHIDA, numIter=30, warmupIter=30, M=8, seed=3, chain=c, adaptH=True, adaptDelta=True)
    err = np.max(np.abs(s[c] - so), axis=0)
    print("draw err per iter", err[:20])
    print("H cuda", d[c][:16, 15]); print("H orcl", do[:16, 15])
    print("delta cuda", d[c][:16, 18]); print("delta orcl", do[:16, 18])
    print("nF", d[c][:12, 6], do[:12, 6])
