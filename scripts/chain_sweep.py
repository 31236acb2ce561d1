"""gradient evals/s over chain counts 2^10 .. 2^18 (C2 workload, R2P), kernel-only, plus the
C1 (package, d = 100) GPU number next to the package oracle on one host core."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from walnuts_b200 import ChainBatch

sigma = bench.sigma_vec()
for lg in range(10, 19, 2):
    n = 1 << lg
    q0 = bench.init_positions(n, 0, sigma)
    cb = ChainBatch("diag_gauss", bench.D, n, integrator="R2P", H0=0.5, delta=0.3, M=10, seed=1, dg=0, data={"inv_var": 1.0 / sigma ** 2})
    cb.set_state(q0)
    cb.run_device(1)
    iters = 4 if lg <= 14 else 1
    cb.run_device(iters)
    f, b = cb.last_grad_evals(); ms = cb.last_kernel_ms()
    print(json.dumps({"config": "C2 R2P", "chains": n, "iters": iters, "kernel_ms": ms, "grad_evals_per_s": (f + b) / (ms * 1e-3)}), flush=True)
    cb.close()

# C1: package semantics, 100-d standard normal, test.py settings
from oracle import package_oracle as po, targets as ot
cnt = [0]
t0 = time.perf_counter()
po.walnuts(123, 0, np.zeros(100), ot.standard_normal_lpdf, ot.standard_normal_grad, np.ones(100), 2.0, 10, 0.1, 0, 20, counter=cnt)
dt = time.perf_counter() - t0
print(json.dumps({"config": "C1 package oracle (numpy port of walnuts.py), 1 chain, 1 core", "transitions": 20, "grad_evals": cnt[0],
                  "grad_evals_per_s": cnt[0] / dt, "transitions_per_s": 20 / dt}), flush=True)
for n in (1, 16384, 262144):
    cb = ChainBatch("std_normal", 100, n, mode="package", H0=2.0, delta=0.1, M=10, seed=123, dg=0, data={"inv_mass": np.ones(100)})
    cb.set_state(np.zeros((n, 100)))
    cb.run_device(2)
    cb.run_device(20)
    f, b = cb.last_grad_evals(); ms = cb.last_kernel_ms()
    print(json.dumps({"config": "C1 package GPU", "chains": n, "transitions": 20 * n, "kernel_ms": ms, "grad_evals_per_s": f / (ms * 1e-3),
                      "transitions_per_s": 20 * n / (ms * 1e-3)}), flush=True)
    cb.close()
