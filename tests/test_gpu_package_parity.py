"""GPU parity: CUDA package-mode kernel (walnuts/walnuts.py semantics) vs the numpy oracle and the
golden fixtures produced by the REAL reference, on the same keyed Philox draws.  Calls go through the
drop-in `walnuts(...)` surface -> ChainBatch -> ctypes -> C-ABI."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import package_oracle as po
from oracle import targets as ot
from tests.helpers import close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def cuda_walnuts(target, theta0, inv_mass, macro_step, depth, max_error, n_iter, seed, chain_offset=0, compat=True):
    import walnuts_b200 as wb
    tg = {"std_normal": wb.targets.standard_normal_lpdf, "funnel_pkg": wb.targets.funnel_lpdf}[target]
    return wb.walnuts(None, theta0, tg, tg, inv_mass, macro_step, depth, max_error, 0, n_iter, seed=seed,
                      chain_offset=chain_offset, compat=compat)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pkg_*.npz"))), ids=os.path.basename)
def test_package_golden(cuda_lib, path):
    """Reference's own draws (walnuts.walnuts run on /root/reference with the keyed rng shim)."""
    z = np.load(path)
    m = json.loads(str(z["meta"]))
    draws = cuda_walnuts(m["target"], z["theta0"], z["inv_mass"], m["macro_step"], m["max_depth"], m["max_error"],
                         m["n_iter"], m["seed"], chain_offset=m["chain"])
    ref = z["draws"]
    assert np.array_equal(np.isnan(draws), np.isnan(ref))
    ok, err = close(draws, ref)
    assert ok, f"{os.path.basename(path)}: max rel err {err:.3e}"


@pytest.mark.parametrize("compat", [True, False])
def test_package_vs_oracle_many_chains(cuda_lib, compat):
    """Several chains at once, d = 7 with a non-trivial mass matrix; compat and corrected semantics."""
    d, n_chains, n_iter, seed = 7, 6, 6, 31
    inv_mass = np.linspace(0.5, 2.0, d)
    theta0 = 0.4 * np.random.default_rng(1).standard_normal((n_chains, d))
    draws = cuda_walnuts("std_normal", theta0, inv_mass, 1.1, 6, 0.15, n_iter, seed, compat=compat)
    for c in range(n_chains):
        ref = po.walnuts(seed, c, theta0[c], ot.standard_normal_lpdf, ot.standard_normal_grad, inv_mass, 1.1, 6,
                         0.15, 0, n_iter, compat=compat)
        ok, err = close(draws[c], ref)
        assert ok, f"chain {c}: max rel err {err:.3e}"


def test_package_c1_d100(cuda_lib):
    """BASELINE config 1 shape: 100-d standard normal, macro_step 2.0, max_error 0.1 (depth capped for time)."""
    d, seed = 100, 123
    theta0 = np.zeros(d)
    draws = cuda_walnuts("std_normal", theta0, np.ones(d), 2.0, 5, 0.1, 4, seed)
    ref = po.walnuts(seed, 0, theta0, ot.standard_normal_lpdf, ot.standard_normal_grad, np.ones(d), 2.0, 5, 0.1, 0, 4)
    ok, err = close(draws, ref)
    assert ok, f"max rel err {err:.3e}"


def test_package_value_errors(cuda_lib):
    """Argument validation mirrors walnuts.py:309-320."""
    import walnuts_b200 as wb
    tg = wb.targets.standard_normal_lpdf
    with pytest.raises(ValueError):
        wb.walnuts(None, np.zeros(3), tg, tg, np.ones(2), 1.0, 5, 0.1, 0, 1, seed=1)
    with pytest.raises(ValueError):
        wb.walnuts(None, np.zeros(3), tg, tg, np.ones(3), 0.0, 5, 0.1, 0, 1, seed=1)
    with pytest.raises(ValueError):
        wb.walnuts(None, np.zeros(3), tg, tg, np.ones(3), 1.0, 0, 0.1, 0, 1, seed=1)
    with pytest.raises(ValueError):
        wb.walnuts(None, np.zeros(3), tg, tg, np.ones(3), 1.0, 5, -0.1, 0, 1, seed=1)
    with pytest.raises(TypeError):
        wb.walnuts(None, np.zeros(3), lambda q: 0.0, lambda q: q, np.ones(3), 1.0, 5, 0.1, 0, 1, seed=1)


def test_package_mode_rejects_diagnostics_and_reports_evals(cuda_lib):
    from walnuts_b200 import ChainBatch
    with ChainBatch("std_normal", 3, 4, mode="package", H0=1.0, delta=0.2, M=5, seed=1, data={"inv_mass": np.ones(3)}) as cb:
        cb.set_state(np.zeros((4, 3)))
        with pytest.raises(ValueError, match="WALNUTSPY mode only"):
            cb.run(1, diag=True)
        out = cb.run(2)
        assert (out["nevalF"] > 0).all() and out["draws"].shape == (2, 4, 3)
        mean, var = cb.moments()
        q = cb.get_state()
        assert np.allclose(mean, q.mean(0)) and np.allclose(var, q.var(0, ddof=1))


def test_fp64_peak_microbenchmark(cuda_lib):
    from walnuts_b200 import fp64_peak
    peak = fp64_peak(0)
    assert 5e12 < peak < 6e13        # B200: ~34 TFLOP/s measured, 37 nominal


@pytest.mark.parametrize("d", [7, 24, 40, 200])
def test_package_small_macro_step_ell_zero_defect_all_kernel_shapes(cuda_lib, d):
    """A macro step so small that the stable micro-step count is 1: choose_micro_steps then draws from {0, 1, 2}
    (walnuts.py:194), and ell = 0 makes the step size infinite and the orbit NaN (defect B3, reproduced with
    compat=True).  Covers that path on every package-kernel shape: one thread per chain (d = 7), 16 threads (24), one
    warp per chain with 4 and 16 coordinates per lane (40, 200)."""
    n_chains, n_iter, seed = 6, 8, 31
    theta0 = 0.4 * np.random.default_rng(1).standard_normal((n_chains, d))
    draws = cuda_walnuts("std_normal", theta0, np.ones(d), 0.25, 5, 0.3, n_iter, seed, compat=True)
    for c in range(n_chains):
        ref = po.walnuts(seed, c, theta0[c], ot.standard_normal_lpdf, ot.standard_normal_grad, np.ones(d), 0.25, 5, 0.3,
                         0, n_iter, compat=True)
        ok, err = close(draws[c], ref)
        assert ok, f"d = {d}, chain {c}: max rel err {err:.3e}"
