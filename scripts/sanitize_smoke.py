"""Tiny runs of every kernel path touched in round 2, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from walnuts_b200 import ChainBatch, datasets

rng = np.random.default_rng(0)


def run(name, d, n, integ, H0, M, data=None, q0=None, iters=2, **kw):
    q0 = rng.standard_normal((n, d)) if q0 is None else q0
    with ChainBatch(name, d, n, integrator=integ, H0=H0, delta=0.3, M=M, seed=3, data=data, **kw) as cb:
        cb.set_state(q0)
        out = cb.run(iters, draws=True, diag=True)
        ess, rhat = cb.ess_rhat(np.ascontiguousarray(np.repeat(out["draws"], 4, axis=0)[:, :, :min(d, 3)]))
    assert np.isfinite(out["draws"]).all()
    print("ok", name, d, integ, flush=True)


sig = np.logspace(-2, 2, 1000)
run("diag_gauss", 1000, 5, "fixed", 0.008, 6, {"inv_var": 1 / sig ** 2}, rng.standard_normal((5, 1000)) * sig)      # DiagSmT, 1 warp/chain
run("diag_gauss", 1000, 5, "R2P", 0.5, 4, {"inv_var": 1 / sig ** 2}, rng.standard_normal((5, 1000)) * sig)           # certificate path
run("std_normal", 300, 6, "fixed", 0.2, 6)                                                                        # NUTS_FAST, G = 32, 4 chains/block
run("std_normal", 2000, 3, "fixed", 0.1, 5)                                                                       # NUTS_FAST, G = 256
run("std_normal", 100, 6, "fixed", 0.3, 6)                                                                        # flat loop, compile-time fixed
run("std_normal", 10, 40, "R2P", 0.4, 6)                                                                        # 8 threads x 2 coordinates
run("std_normal", 24, 40, "D", 0.4, 6)                                                                          # 16 threads x 2 coordinates
run("std_normal", 100, 12, "R2P", 0.3, 6)                                                                       # one chain per warp, 4 coordinates per lane
run("funnel", 11, 40, "R2P", 0.3, 8)                                                                              # control block in smem
run("funnel", 11, 40, "fixed", 0.3, 8)
P = np.eye(100) * 2 + 0.01
run("dense_gauss", 100, 11, "R2P", 0.4, 5, {"precision": P})
P7 = np.eye(7) * 2 + 0.1
run("dense_gauss", 7, 11, "fixed", 0.4, 5, {"precision": P7})
y = datasets.stock_watson_series()
q0 = 0.05 * rng.standard_normal((3, 756)); q0[:, 0] = 2.4
run("stock_watson", 756, 3, "R2P", 0.1, 4, {"y": y}, q0, iters=1, minC=3)
X, yy, beta = datasets.synth_logreg(800, 100, 0)
run("logreg", 100, 9, "R2P", 0.1, 4, {"X": X, "y": yy, "tau": np.array([1.0])}, beta + 0.05 * rng.standard_normal((9, 100)))
# chain scheduler (wn_sched.cu): more chains than resident slots, two calls on one handle (the second one is served from
# the longest-first queue with exclusive warps), and a user plug-in (one warp per chain, d = 200)
qf = rng.standard_normal((19200, 11)); qf[:, 0] *= 3.0; qf[:, 1:] *= np.exp(0.5 * qf[:, :1])
with ChainBatch("funnel", 11, 19200, integrator="R2P", H0=0.3, delta=0.3, M=5, seed=3) as cb:
    cb.set_state(qf)
    a1 = cb.run(1, draws=True)
    a2 = cb.run(1, draws=True)
    assert np.isfinite(a2["draws"]).all()
print("ok sched", flush=True)
print("all ok")
