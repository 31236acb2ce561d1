"""Where does the funnel call's time go?  Per-chain cost distribution of one C3 call and the time of the costliest chain
run ALONE (same seed, chain id, state): the lower bound of the call's duration."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from walnuts_b200 import ChainBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
lone = int(sys.argv[2]) if len(sys.argv) > 2 else -1      # >= 0: only run this chain alone (profiling)
rng = np.random.default_rng(0)
q0 = rng.standard_normal((n, 11)); q0[:, 0] *= 3.0; q0[:, 1:] *= np.exp(0.5 * q0[:, :1])
kw = dict(integrator="R2P", H0=0.3, delta=0.3, M=12, seed=1)
if lone >= 0:
    with ChainBatch("funnel", 11, 1, chain_offset=lone, **kw) as cb:
        cb.set_state(q0[lone:lone + 1])
        r1 = cb.run(10, draws=False)
        print(json.dumps({"chain": lone, "alone_ms": cb.last_kernel_ms(), "evals": int(r1["nevalF"][0] + r1["nevalB"][0])}))
    sys.exit(0)
with ChainBatch("funnel", 11, n, **kw) as cb:
    cb.set_state(q0)
    r = cb.run(10, draws=False)
    ms = cb.last_kernel_ms()
cost = (r["nevalF"] + r["nevalB"]).astype(np.int64)
top = np.argsort(-cost)[:5]
print(json.dumps({"chains": n, "kernel_ms": ms, "total_evals": int(cost.sum()), "mean": float(cost.mean()),
                  "quantiles_50_99_999_max": [int(np.quantile(cost, x)) for x in (0.5, 0.99, 0.999, 1.0)],
                  "top5": cost[top].tolist(), "q0_of_top5": q0[top, 0].round(2).tolist()}))
for i in top[:3]:
    with ChainBatch("funnel", 11, 1, chain_offset=int(i), **kw) as cb:
        cb.set_state(q0[i:i + 1])
        r1 = cb.run(10, draws=False)
        print(json.dumps({"chain": int(i), "alone_ms": cb.last_kernel_ms(), "evals": int(r1["nevalF"][0] + r1["nevalB"][0]),
                          "ns_per_eval": cb.last_kernel_ms() * 1e6 / float(r1["nevalF"][0] + r1["nevalB"][0])}))
