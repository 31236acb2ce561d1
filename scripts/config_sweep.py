"""Kernel-only throughput of every BASELINE.json config (device-resident), one JSON line each."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from walnuts_b200 import ChainBatch
from walnuts_b200 import datasets

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--scale", type=float, default=1.0, help="scale chain counts")
ap.add_argument("--integ", default="", help="only this integrator (fixed | D | R2P)")
a = ap.parse_args()


def run(tag, cb, q0, iters, flop_per_eval=None, pkg=False):
    cb.set_state(q0)
    cb.run_device(1)                      # warm-up
    cb.run_device(iters)
    f, b = cb.last_grad_evals()
    ms = cb.last_kernel_ms()
    line = {"config": tag, "chains": q0.shape[0], "d": q0.shape[1], "iters": iters, "kernel_ms": ms,
            "grad_evals": f + b, "grad_evals_per_s": (f + b) / (ms * 1e-3),
            "evals_per_transition": (f + b) / (iters * q0.shape[0])}
    if flop_per_eval:
        line["tflops_algorithmic"] = line["grad_evals_per_s"] * flop_per_eval / 1e12
    print(json.dumps(line), flush=True)
    cb.close()


rng = np.random.default_rng(0)
if a.only in ("", "c1"):      # package mode, 100-d std normal (test.py settings), many chains
    n = int(16384 * a.scale)
    cb = ChainBatch("std_normal", 100, n, mode="package", H0=2.0, delta=0.1, M=10, seed=1, dg=0,
                    data={"inv_mass": np.ones(100)})
    run("C1 package std_normal d=100 macro_step=2.0 depth=10", cb, rng.standard_normal((n, 100)), 1, 12 * 100)
if a.only in ("", "c2"):
    n = int(65536 * a.scale)
    sigma = np.logspace(-2, 2, 1000)
    for integ, H0 in (("R2P", 0.5), ("D", 0.5), ("fixed", 0.008)):
        if a.integ and integ != a.integ:
            continue
        cb = ChainBatch("diag_gauss", 1000, n, integrator=integ, H0=H0, delta=0.3, M=10, seed=1, dg=0,
                        data={"inv_var": 1 / sigma ** 2})
        run(f"C2 diag_gauss d=1000 {integ} H0={H0}", cb, rng.standard_normal((n, 1000)) * sigma, 1, 12 * 1000)
if a.only in ("", "c3"):
    n = int(262144 * a.scale)
    q0 = np.empty((n, 11)); q0[:, 0] = 3 * rng.standard_normal(n); q0[:, 1:] = np.exp(.5 * q0[:, :1]) * rng.standard_normal((n, 10))
    for integ in ("R2P", "fixed"):
        if a.integ and integ != a.integ:
            continue
        cb = ChainBatch("funnel", 11, n, integrator=integ, H0=0.3, delta=0.3, M=12, seed=1, dg=0)
        run(f"C3 funnel10 {integ} M=12 H0=0.3", cb, q0, 10, 200)
if a.only in ("", "c4"):
    n = int(16384 * a.scale)
    X, y, beta = datasets.synth_logreg(100_000, 100, 0)
    q0 = beta + 0.05 * rng.standard_normal((n, 100))
    for integ, H0 in (("fixed", 0.02), ("R2P", 0.05)):
        if a.integ and integ != a.integ:
            continue
        cb = ChainBatch("logreg", 100, n, integrator=integ, H0=H0, delta=0.3, M=6, seed=1, dg=0,
                        data={"X": X, "y": y, "tau": np.array([1.0])})
        run(f"C4 logreg N=100000 P=100 {integ} H0={H0}", cb, q0, 1, 4.5e7)
if a.only in ("", "c5"):
    n = int(131072 * a.scale)
    y = datasets.stock_watson_series()
    q0 = 0.05 * rng.standard_normal((n, 756)); q0[:, 0] = 2.4
    for integ, H0, minC in (("R2P", 0.1, 3), ("fixed", 0.002, 0)):
        if a.integ and integ != a.integ:
            continue
        cb = ChainBatch("stock_watson", 756, n, integrator=integ, H0=H0, delta=0.3, M=14 if integ == "fixed" else 8,
                        minC=minC, seed=1, dg=0, data={"y": y})
        run(f"C5 stock_watson T=252 {integ} H0={H0}", cb, q0, 1, 4e4)
