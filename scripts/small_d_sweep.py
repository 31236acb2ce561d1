"""Throughput of the generic small-d kernel shapes (std_normal / funnel_pkg; WALNUTSpy R2P, fixedLeapFrog and package mode):
one JSON line per case.  Used to compare threads-per-chain choices of wn_dispatch.cuh::pick_generic."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from walnuts_b200 import ChainBatch

rng = np.random.default_rng(0)
n = 131072
for d in (2, 7, 10, 20, 32, 50, 100):
    q0 = rng.standard_normal((n, d))
    for mode, integ in (("walnutspy", "R2P"), ("walnutspy", "fixed"), ("package", "fixed")):
        kw = dict(integrator=integ, H0=0.6, delta=0.3, M=8, seed=1, dg=0)
        if mode == "package":
            kw = dict(mode="package", H0=1.0, delta=0.2, M=8, seed=1, dg=0, data={"inv_mass": np.ones(d)})
        with ChainBatch("std_normal", d, n, **kw) as cb:
            cb.set_state(q0)
            cb.run_device(1)
            cb.run_device(4)
            f, b = cb.last_grad_evals()
            ms = cb.last_kernel_ms()
        print(json.dumps({"d": d, "mode": mode, "integrator": integ, "kernel_ms": round(ms, 3),
                          "grad_evals_per_s": round((f + b) / (ms * 1e-3) / 1e9, 4)}), flush=True)
