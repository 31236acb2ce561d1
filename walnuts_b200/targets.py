"""CUDA target registry: handles that replace the reference's Python target functions
(WALNUTSpy/targetDistr.py, test/targets.py).  A handle names a hand-written CUDA log-density +
gradient (walnuts_b200/csrc/wn_targets.cuh) and carries its data; it is passed where the reference
takes `lpFun` (WALNUTS.py:111) or `logp`/`grad` (walnuts.py:362).
"""
import numpy as np


class Target:
    def __init__(self, name, data=None, d=None, ref=""):
        self.name, self.data, self.d, self.ref = name, dict(data or {}), d, ref

    def __repr__(self):
        return f"<walnuts_b200 target {self.name}{'' if self.d is None else ' d=%d' % self.d} ({self.ref})>"

    def __call__(self, *a, **k):
        raise TypeError(f"target {self.name!r} is a handle for a CUDA log density; it is evaluated "
                        "on the GPU inside walnuts_b200.WALNUTS/walnuts and cannot be called on host arrays")


# names as in the reference -------------------------------------------------------------------------
stdGauss = Target("std_normal", ref="targetDistr.py:18-21")
corrGauss = Target("corr_gauss", d=2, ref="targetDistr.py:25-31")
funnel10 = Target("funnel", d=11, ref="targetDistr.py:74-78")
standard_normal_lpdf = standard_normal_grad = Target("std_normal", ref="test/targets.py:4-7")
funnel_lpdf = funnel_grad = Target("funnel_pkg", ref="test/targets.py:23-29")


def funnel(n=10):
    """Neal's funnel with one log-scale and n conditionally normal coordinates (funnel10 for n=10)."""
    return Target("funnel", d=n + 1, ref="targetDistr.py:74-78")


def diag_gauss(sigma):
    """Zero-mean Gaussian with standard deviations `sigma` (not in the reference)."""
    sigma = np.asarray(sigma, dtype=np.float64)
    return Target("diag_gauss", data={"inv_var": 1.0 / (sigma ** 2)}, d=sigma.size, ref="SURVEY.md T2")


def ill_conditioned_gauss(d=1000, lo=-2.0, hi=2.0):
    """BASELINE config 2: sigma = logspace(lo, hi, d)."""
    return diag_gauss(np.logspace(lo, hi, d))


def stock_watson(y):
    """Stock-Watson stochastic-volatility model of the reference's example
    (WALNUTSpy_examples/StockWatson/sw_innov.stan:2-52, bridgestan default propto=True) on the
    unconstrained scale; `y` is the observed series (T values), d = 3T."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    return Target("stock_watson", data={"y": y}, d=3 * y.size, ref="sw_innov.stan:2-52")


def logreg(X, y, tau=1.0):
    """Bayesian logistic regression: Bernoulli-logit likelihood for rows of X [N, P]
    with responses y in {0, 1} and a N(0, tau^2 I) prior on the P coefficients."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if X.ndim != 2 or y.shape != (X.shape[0],):
        raise ValueError("X must be [N, P] and y [N]")
    return Target("logreg", data={"X": X, "y": y, "tau": np.array([float(tau)])}, d=X.shape[1], ref="SURVEY.md T4")


def resolve(target, d):
    """Target handle (or registry name) -> (name, data) after checking the dimension."""
    if isinstance(target, str):
        target = Target(target)
    if not isinstance(target, Target):
        raise TypeError(
            "walnuts_b200 runs hand-written CUDA targets only: pass a handle from walnuts_b200.targets "
            "(stdGauss, funnel10, diag_gauss(sigma), ...) where the reference takes a Python "
            "log-density callable; arbitrary Python callables have no GPU implementation and there is "
            "deliberately no CPU fallback")
    if target.d is not None and target.d != d:
        raise ValueError(f"target {target.name} has dimension {target.d}, state has {d}")
    return target.name, target.data
