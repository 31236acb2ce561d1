"""Parity at the FULL shapes of the BASELINE configs (north_star: "per-iteration draws match in fp64 to 1e-10
relative error over the first 100 transitions"), beyond the C2 test of test_gpu_walnutspy_parity.py:

  C1  100-d standard normal, package semantics, depth 10, 100 transitions: golden from the REAL walnuts.py
      (tests/golden/pkg_std100_c1.npz, picked up by test_gpu_package_parity.py::test_package_golden) + here the
      oracle on several chains
  C4  logistic regression at N = 100 000 rows, P = 100 (the 12 500-tile DMMA accumulation)
  C5  Stock-Watson, M = 14, minC = 3 (mainSW.py:41-49), 100 transitions compared per transition on identical inputs
  C2  long run: 100 transitions from stationarity, every coordinate's mean and variance against the exact moments
      with a stated family-wise (Bonferroni) bound
All through ChainBatch -> ctypes -> C-ABI."""
import numpy as np
import pytest
from scipy.stats import norm

from tests.helpers import close
from tests.test_gpu_walnutspy_parity import EXACT_COLS, check, check_forced, sw_q0

pytestmark = pytest.mark.gpu


def test_c1_package_d100_depth10_100_transitions_vs_oracle(cuda_lib):
    """BASELINE config 1 shape on several chains (the single-chain golden of the real reference is covered by
    test_package_golden): 100-d standard normal, macro_step 2.0, max_nuts_depth 10, max_error 0.1, 100 transitions."""
    import walnuts_b200 as wb
    from oracle import package_oracle as po
    from oracle import targets as ot
    d, n_chains, n_iter, seed = 100, 4, 100, 123
    theta0 = 0.5 * np.random.default_rng(6).standard_normal((n_chains, d))
    tg = wb.targets.standard_normal_lpdf
    draws = wb.walnuts(None, theta0, tg, tg, np.ones(d), 2.0, 10, 0.1, 0, n_iter, seed=seed)
    worst = 0.0
    for c in range(n_chains):
        ref = po.walnuts(seed, c, theta0[c], ot.standard_normal_lpdf, ot.standard_normal_grad, np.ones(d), 2.0, 10,
                         0.1, 0, n_iter)
        ok, err = close(draws[c], ref)
        worst = max(worst, err)
        assert ok, f"chain {c}: max rel err {err:.3e}"
    print(f"C1 shape, 100 free-running transitions x {n_chains} chains: worst relative error {worst:.2e}")


@pytest.mark.parametrize("integrator", ["fixed", "R2P"])
def test_c4_logreg_full_size(cuda_lib, integrator):
    """BASELINE config 4 at its full size: N = 100 000 rows, P = 100 features (12 500 row tiles per gradient on the
    DMMA path), 9 chains = one full CTA of 8 plus one, a few transitions; the numpy formula on the same streams."""
    from oracle import targets as ot
    X, y, beta = ot.synth_logreg_data(N=100_000, P=100, seed=0)
    data = {"X": X, "y": y, "tau": np.array([1.0])}
    q0 = beta + 0.05 * np.random.default_rng(4).standard_normal((9, 100))
    check("logreg", q0, integrator, H0=0.05 if integrator == "R2P" else 0.02, delta=0.3, M=6, n_iter=3, data=data,
          chains=[0, 8], float_rtol=1e-7)


def test_c5_stock_watson_M14_100_transitions(cuda_lib):
    """BASELINE config 5 settings (mainSW.py:41-49: R2P, M = 14, H0 = 0.1, delta0 = 0.3, minC = 3), T = 252: all 100
    transitions compared on identical inputs (teacher-forced: the volatility model amplifies rounding differences
    between free-running chains), plus the length of the free-running prefix that agrees to 1e-10."""
    from oracle import targets as ot
    y = ot.load_sw_data()
    q0 = sw_q0(2, y.size)
    dg, worst = check_forced("stock_watson", q0, "R2P", H0=0.1, delta=0.3, M=14, n_iter=100, minC=3, data={"y": y})
    print(f"C5 (M = 14, minC = 3): worst per-transition relative error over 100 transitions {worst:.2e}; "
          f"stop codes {np.unique(dg[..., 19])}, doublings up to {int(dg[..., 1].max())}")


def _bonferroni_z(n_tests, alpha=0.0027):
    """Two-sided normal quantile such that the FAMILY of n_tests comparisons exceeds it with probability alpha
    (alpha = 0.0027 is what '3 standard errors' means for a single comparison)."""
    return float(norm.isf(alpha / (2.0 * n_tests)))


@pytest.mark.parametrize("compat", [False, True])
def test_c2_long_run_100_transitions_moments(cuda_lib, compat):
    """north_star: "over long runs, posterior means and variances match within 3 MCSE".  16 384 chains start from
    exact draws of the 1000-d ill-conditioned Gaussian and take 100 R2P transitions (H0 = 0.5, delta = 0.3, M = 10);
    afterwards the cross-chain mean and variance of EVERY coordinate are compared with the exact moments.  The chains
    are independent, so the Monte-Carlo standard errors are exact: sigma_i / sqrt(n) for the mean and
    sigma_i^2 sqrt(2 / (n - 1)) for the variance.  2000 quantities are compared, so the '3 MCSE' statement is applied
    family-wise: the bound is the Bonferroni-corrected quantile (4.8 standard errors for 2000 comparisons at the
    single-comparison level of 3 standard errors)."""
    from walnuts_b200 import ChainBatch
    d, n, n_iter = 1000, 16384, 100
    sigma = np.logspace(-2, 2, d)
    q0 = np.random.default_rng(11).standard_normal((n, d)) * sigma
    with ChainBatch("diag_gauss", d, n, integrator="R2P", H0=0.5, delta=0.3, M=10, seed=3, dg=0, compat=compat,
                    data={"inv_var": 1.0 / sigma ** 2}) as cb:
        cb.set_state(q0)
        cb.run(n_iter, draws=False, nevals=False)
        q1 = cb.get_state()
    z = q1 / sigma
    assert np.isfinite(z).all()
    zb = _bonferroni_z(2 * d)
    zm = np.abs(z.mean(0)) * np.sqrt(n)
    zv = np.abs(z.var(0, ddof=1) - 1.0) / np.sqrt(2.0 / (n - 1))
    print(f"C2 long run (compat={compat}): worst mean deviation {zm.max():.2f} SE, worst variance deviation "
          f"{zv.max():.2f} SE over {d} coordinates; family-wise bound {zb:.2f} SE")
    assert zm.max() < zb and zv.max() < zb
    # the chains moved: the fast coordinates are decorrelated from their start after 100 transitions
    fast = sigma < 1.0
    corr = np.mean(z[:, fast] * (q0 / sigma)[:, fast], axis=0)
    assert np.abs(corr).max() < 0.1


def _random_precision(d, seed, cond=50.0):
    rng = np.random.default_rng(seed)
    Qm, _ = np.linalg.qr(rng.standard_normal((d, d)))
    ev = np.exp(rng.uniform(-0.5 * np.log(cond), 0.5 * np.log(cond), d))
    P = (Qm * ev) @ Qm.T
    return 0.5 * (P + P.T)


@pytest.mark.parametrize("integrator,d", [("R2P", 100), ("fixed", 100), ("D", 24), ("R2P", 7), ("R2P", 104), ("fixed", 33),
                                          ("R2P", 120)])
def test_dense_gauss_tensor_core_target(cuda_lib, integrator, d):
    """Dense-precision Gaussian (north_star): gradient -P Q of the 8 lock-stepped chains of a CTA on the FP64 tensor
    cores (DMMA, TMA-staged rows of P) for d <= 104, per-warp FMA version beyond (d = 120); odd d exercises the
    non-bulk-copy tiles.  Oracle = numpy formula on the same streams (target not in the reference: parity unpinned
    for the FUNCTION, the transition kernel is the pinned one)."""
    P = _random_precision(d, seed=d)
    cov = np.linalg.inv(P)
    L = np.linalg.cholesky(cov)
    q0 = np.random.default_rng(2).standard_normal((11, d)) @ L.T           # more chains than one CTA holds
    H0 = (0.9 if integrator != "fixed" else 0.35) * np.sqrt(1.0 / np.linalg.eigvalsh(P).max())
    dg = check("dense_gauss", q0, integrator, H0=H0, delta=0.3, M=6, n_iter=12, data={"precision": P},
               chains=[0, 7, 8, 10], float_rtol=1e-7)
    assert len(np.unique(dg[..., 19])) >= 1


def test_dense_gauss_package_mode_and_moments(cuda_lib):
    """The per-warp version behind walnuts(...) (package semantics) and a long-run moment check of the tensor-core
    version: 4096 chains from exact draws keep the covariance P^-1."""
    import walnuts_b200 as wb
    from oracle import package_oracle as po
    from oracle import targets as ot
    d = 12
    P = _random_precision(d, seed=3, cond=10.0)
    tg = wb.targets.dense_gauss(P)
    th0 = 0.3 * np.random.default_rng(1).standard_normal((3, d))
    draws = wb.walnuts(None, th0, tg, tg, np.ones(d), 0.6, 6, 0.2, 0, 10, seed=5)
    lp = ot.make_dense_gauss(P)
    for c in range(3):
        ref = po.walnuts(5, c, th0[c], lambda q: lp(q)[0], lambda q: lp(q)[1], np.ones(d), 0.6, 6, 0.2, 0, 10)
        ok, err = close(draws[c], ref)
        assert ok, err
    n = 4096
    cov = np.linalg.inv(P)
    q0 = np.random.default_rng(4).standard_normal((n, d)) @ np.linalg.cholesky(cov).T
    s, dgn = wb.WALNUTS(tg, q0, integrator=wb.adaptLeapFrogR2P, H0=0.5, delta0=0.3, numIter=30, warmupIter=0, M=7,
                        adaptH=False, adaptDelta=False, seed=9, compat=False)
    x = s[:, :, -1]
    emp = np.cov(x.T)
    se = np.sqrt((np.outer(np.diag(cov), np.diag(cov)) + cov ** 2) / n)      # SE of a Gaussian sample covariance
    assert np.abs((emp - cov) / se).max() < 5.0
