"""User-target plug-ins: one thread per chain against one warp per chain (WN_USER_LAYOUT=warp) at small d."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import walnuts_b200 as wb
from walnuts_b200 import ChainBatch

SRC = """
WN_TARGET_LP_GRAD(q, g, data, n_data) {
  double lp = 0.0;
  for (int i = 0; i < WN_D; ++i) { const double s = 1.0 + 0.1 * i; g[i] = -q[i] * s; lp -= 0.5 * q[i] * q[i] * s; }
  for (int i = 0; i + 1 < WN_D; ++i) { const double r = q[i + 1] - 0.3 * q[i]; lp -= 0.5 * r * r; g[i + 1] -= r; g[i] += 0.3 * r; }
  return lp;
}
"""
rng = np.random.default_rng(0)
n = 65536
for d in (2, 8, 20, 50):
    tg = wb.targets.cuda_target(SRC, d, name=f"sweep{d}_{os.environ.get('WN_USER_LAYOUT', 'thread')}")
    tid, data = wb.targets.resolve(tg, d)
    q0 = rng.standard_normal((n, d))
    for mode, integ in (("walnutspy", "R2P"), ("walnutspy", "fixed"), ("package", "fixed")):
        kw = dict(integrator=integ, H0=0.5, delta=0.3, M=7, seed=1, dg=0)
        if mode == "package":
            kw = dict(mode="package", H0=0.8, delta=0.2, M=7, seed=1, dg=0, data={"inv_mass": np.ones(d)})
        with ChainBatch(tid, d, n, **kw) as cb:
            cb.set_state(q0)
            cb.run_device(1)
            cb.run_device(3)
            f, b = cb.last_grad_evals()
            ms = cb.last_kernel_ms()
        print(json.dumps({"layout": os.environ.get("WN_USER_LAYOUT", "thread"), "d": d, "mode": mode, "integrator": integ,
                          "kernel_ms": round(ms, 3), "evals": int(f + b)}), flush=True)
