import sys, numpy as np
sys.path.insert(0, ".")
from tests.helpers import oracle_walnutspy
from tests.test_gpu_walnutspy_parity import run_cuda, sw_q0, FLOAT_COLS
from oracle import targets as ot
y = ot.load_sw_data()[:37]
q0 = sw_q0(4, 37)
out, st = run_cuda("stock_watson", q0, "R2P", 0.1, 0.3, 6, 6, 1234, 1, 10, {"y": y})
dr, dg = oracle_walnutspy("stock_watson", q0, "R2P", 0.1, 0.3, 6, 6, 1234, [0,1,2,3], 1, 10, {"y": y})
g = out["diag"]
np.set_printoptions(linewidth=250, precision=12)
err = np.abs(g - dg) / np.maximum(1, np.abs(dg))
print("max err per col", err.max(axis=(0, 1)))
i = np.unravel_index(np.argmax(err), err.shape); print(i, g[i], dg[i])
print("draw err", np.max(np.abs(out["draws"] - dr) / np.maximum(1, np.abs(dr)), axis=(1, 2)))
lp = ot.make_stock_watson(y)
print("H scale", lp(q0[0])[0])
