import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from walnuts_b200 import ChainBatch
from oracle import targets as ot
ap = argparse.ArgumentParser(); ap.add_argument("--chains", type=int, default=1184); ap.add_argument("--integrator", default="fixed")
ap.add_argument("--reps", type=int, default=2); a = ap.parse_args()
X, y, beta = ot.synth_logreg_data(100_000, 100, 0)
q0 = beta + 0.05 * np.random.default_rng(0).standard_normal((a.chains, 100))
cb = ChainBatch("logreg", 100, a.chains, integrator=a.integrator, H0=0.02 if a.integrator == "fixed" else 0.05, delta=0.3, M=6, seed=1, dg=0,
                data={"X": X, "y": y, "tau": np.array([1.0])})
cb.set_state(q0)
for r in range(a.reps):
    cb.run_device(1); f, b = cb.last_grad_evals(); ms = cb.last_kernel_ms()
    print("logreg %s chains=%d: %.4g evals/s (%.1f ms, %d evals)" % (a.integrator, a.chains, (f + b) / (ms * 1e-3), ms, f + b))
