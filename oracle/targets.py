"""CPU (numpy) target log densities for the oracles.  TEST INFRASTRUCTURE ONLY.

WALNUTSpy protocol: ``lpFun(q) -> [lp, grad]`` (reference WALNUTSpy/targetDistr.py:18).
Package protocol:   ``logp(theta) -> float``, ``grad(theta) -> (D,)`` (reference
walnuts/walnuts.py:296-297, test/targets.py:4-29).

Targets that exist in the reference are restated with the reference's own operation
order; the four the reference lacks (diag Gaussian, dense-precision Gaussian, logistic regression, Stock-Watson)
are defined here and are the specification for the CUDA targets (SURVEY.md rows T2,T4,T5; north_star).

PARITY UNPINNED for those four target functions: the reference holds no implementation, test or stored output of
the diagonal / dense-precision Gaussians and the logistic regression, and its Stock-Watson density lives in a Stan model evaluated
through bridgestan (absent here: no stanc, no bridgestan; WALNUTSpy_examples/StockWatson/mainSW.py:15-26), so
make_stock_watson() follows sw_innov.stan:2-52 line by line and is checked by finite differences only
(tests/test_oracle_golden.py).  Everything else in oracle/ -- the transition kernels, the integrators, the adaptation,
the orbit statistics and the targets the reference does contain -- is pinned against the reference itself.
"""
import os

import numpy as np

LOG_SQRT_2PI = float(np.log(np.sqrt(2.0 * np.pi)))


# ----------------------------------------------------------------------------------------
# WALNUTSpy-protocol targets
# ----------------------------------------------------------------------------------------
def std_normal(q, hessian=False):
    """targetDistr.stdGauss (targetDistr.py:18-21)."""
    lp = -0.5 * np.sum(q * q)
    return [lp, -q]


def make_diag_gauss(sigma):
    """Row T2: lp = -1/2 sum (q/sigma)^2 with s = sigma**-2 precomputed in fp64."""
    s = 1.0 / (np.asarray(sigma, dtype=np.float64) ** 2)

    def diag_gauss(q, hessian=False):
        g = -(q * s)
        lp = 0.5 * np.sum(q * g)
        return [lp, g]

    diag_gauss.inv_var = s
    return diag_gauss


def make_dense_gauss(precision):
    """Dense-precision Gaussian (north_star; not in the reference): lp = -1/2 q^T P q, grad = -P q."""
    P = np.ascontiguousarray(precision, dtype=np.float64)

    def dense_gauss(q, hessian=False):
        Pq = P @ q
        return [-0.5 * float(q @ Pq), -Pq]

    return dense_gauss


def corr_gauss(q, hessian=False):
    """targetDistr.corrGauss (targetDistr.py:25-31), rho = 0.5."""
    rho = 0.5
    tmp = 1.0 - rho ** 2
    lp = -0.5 * q[0] ** 2 - (0.5 / tmp) * (q[1] - rho * q[0]) ** 2
    grad = np.array([-(q[0] - rho * q[1]) / tmp, -(q[1] - rho * q[0]) / tmp])
    return [lp, grad]


def funnel10(q, hessian=False):
    """targetDistr.funnel10 (targetDistr.py:74-78) with scipy's norm.logpdf written out:
    logpdf(x; 0, s) = -(x/s)^2/2 - log(sqrt(2 pi)) - log(s)."""
    n = q.size - 1
    e = np.exp(-q[0])
    ss = np.sum(q[1:] * q[1:])
    lp = (-(q[0] / 3.0) ** 2 / 2.0 - LOG_SQRT_2PI - np.log(3.0)) \
        + (-0.5 * e * ss - n * LOG_SQRT_2PI - n * 0.5 * q[0])
    grad = np.empty_like(q)
    grad[0] = -0.5 * n - q[0] / 9.0 + 0.5 * e * ss
    grad[1:] = -q[1:] * e
    return [lp, grad]


def make_logreg(X, y, tau=1.0):
    """Row T4: Bernoulli-logit likelihood with N(0, tau^2 I) prior.
    lp = sum_n [y_n eta_n - log(1 + exp(eta_n))] - |beta|^2 / (2 tau^2),  eta = X beta."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    itau2 = 1.0 / (tau * tau)

    def logreg(q, hessian=False):
        eta = X @ q
        # log(1+exp(eta)) = max(eta,0) + log1p(exp(-|eta|))
        l1pe = np.maximum(eta, 0.0) + np.log1p(np.exp(-np.abs(eta)))
        lp = np.sum(y * eta - l1pe) - 0.5 * itau2 * np.sum(q * q)
        sig = 0.5 * (1.0 + np.tanh(0.5 * eta))
        grad = X.T @ (y - sig) - itau2 * q
        return [lp, grad]

    return logreg


def synth_logreg_data(N=100_000, P=100, seed=0):
    """X ~ N(0,1)/sqrt(P), beta* ~ N(0,1), y ~ Bern(sigmoid(X beta*)),
    numpy Generator(PCG64(seed))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.standard_normal((N, P)) / np.sqrt(P)
    beta = rng.standard_normal(P)
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-(X @ beta)))).astype(np.float64)
    return X, y, beta


def load_sw_data(path=None):
    """The T=252 observations `y` of the reference's Stock-Watson example
    (WALNUTSpy_examples/StockWatson/swdata.json), stored as a plain array in tests/golden/sw_y.npy so
    that the GPU box (which has no /root/reference) can run the Stock-Watson target.  Data, not code."""
    if path is None:
        path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "sw_y.npy")
    return np.load(path)


def make_stock_watson(y):
    """Row T5 / SURVEY.md appendix B: unconstrained parameterisation of
    WALNUTSpy_examples/StockWatson/sw_innov.stan:7-52 (bridgestan default propto=True drops the
    normal constants).  theta = [tS, z1, zinn[T-2], x1, xinn[T-1], tau1, tauinn[T-1]], d = 3T."""
    y = np.asarray(y, dtype=np.float64)
    T = y.size

    def stock_watson(q, hessian=False):
        tS = q[0]
        z1 = q[1]
        zinn = q[2:T]                  # T-2
        x1 = q[T]
        xinn = q[T + 1:2 * T]          # T-1
        tau1 = q[2 * T]
        tauinn = q[2 * T + 1:3 * T]    # T-1
        sigma = np.exp(-0.5 * tS)
        z = z1 + sigma * np.concatenate([[0.0], np.cumsum(zinn)])            # T-1
        x = x1 + sigma * np.concatenate([[0.0], np.cumsum(xinn)])            # T
        ez = np.exp(0.5 * z)                                                 # T-1
        tau = tau1 + np.concatenate([[0.0], np.cumsum(ez * tauinn)])         # T
        e = y - tau
        w = np.exp(-x)
        lp = (5.0 * tS - 0.5 * np.exp(tS)
              - 0.5 * (np.sum(zinn * zinn) + np.sum(xinn * xinn) + np.sum(tauinn * tauinn))
              + np.sum(-0.5 * x - 0.5 * e * e * w))
        r = e * w
        R = np.cumsum(r[::-1])[::-1]                  # R_t = sum_{s>=t} r_s
        a = -0.5 + 0.5 * e * e * w
        A = np.cumsum(a[::-1])[::-1]
        b = 0.5 * ez * tauinn * R[1:]                 # k = 1..T-1 uses R_{k+1}
        Bz = np.cumsum(b[::-1])[::-1]                 # T-1
        g = np.empty_like(q)
        g[2 * T] = R[0]
        g[2 * T + 1:3 * T] = -tauinn + ez * R[1:]
        g[T] = A[0]
        g[T + 1:2 * T] = -xinn + sigma * A[1:]
        g[1] = Bz[0]
        g[2:T] = -zinn + sigma * Bz[1:]
        g[0] = 5.0 - 0.5 * np.exp(tS) - 0.5 * sigma * (np.sum(zinn * Bz[1:]) + np.sum(xinn * A[1:]))
        return [lp, g]

    return stock_watson


# ----------------------------------------------------------------------------------------
# package-protocol targets (reference test/targets.py)
# ----------------------------------------------------------------------------------------
def standard_normal_lpdf(q):
    return -0.5 * np.dot(q, q)


def standard_normal_grad(q):
    return -q


def make_diag_gauss_pkg(sigma):
    f = make_diag_gauss(sigma)
    return (lambda q: f(q)[0]), (lambda q: f(q)[1])


def funnel_lpdf(q):
    """test/targets.py:23-24 (NB: not the funnel10 density)."""
    return -0.5 * q[0] ** 2 / 9 - 0.5 * np.dot(q[1:], q[1:]) / np.exp(0.5 * q[0])


def funnel_grad(q):
    """test/targets.py:25-29."""
    grad = np.empty(q.size)
    grad[0] = -(q[0] / 9 - 0.25 * np.dot(q[1:], q[1:]) / np.exp(0.5 * q[0]))
    grad[1:] = -1 / np.exp(0.5 * q[0]) * q[1:]
    return grad


# ---- the remaining reference targets (CUDA side: user targets of walnuts_b200/targets.py) --------------------------
def smile(q, hessian=False):
    """targetDistr.smileDistr :34-38 (= test/targets.py rosenbrock :17-21)."""
    lp = -0.5 * q[0] ** 2 - 0.5 * (q[1] - q[0] ** 2) ** 2
    return [lp, np.array([-q[0] + 2.0 * q[0] * q[1] - 2.0 * q[0] ** 3, q[0] ** 2 - q[1]])]


def mod_funnel(q, hessian=False):
    """targetDistr.modFunnel :41-51."""
    x, y = q[0], q[1]
    t1 = np.exp(-3.0 * x)
    t2 = 1.0 + t1
    t3 = 1.0 / t2
    t4 = y ** 2
    lp = -0.5 * (t2 * t4 + np.log(t3) + x ** 2)
    return [lp, np.array([1.5 * t1 * (t4 - t3) - x, -y * t2])]


def funnel1(q, hessian=False):
    """targetDistr.funnel1 :88-92 with the closed-form normal log-pdf (the reference calls scipy)."""
    ex = np.exp(-q[0])
    lp = (-((q[0] / 3.0) ** 2) / 2.0 - LOG_SQRT_2PI - np.log(3.0)) + (-0.5 * ex * q[1] ** 2 - LOG_SQRT_2PI - 0.5 * q[0])
    return [lp, np.array([-0.5 - q[0] / 9.0 + 0.5 * q[1] ** 2 * ex, -q[1] * ex])]


def funnel10_rescaled(q, hessian=False):
    """targetDistr.funnel10rescaled :81-86."""
    S = np.ones(11)
    S[0] = 3.0
    lp, g = funnel10(S * q)
    return [lp, S * g]


def correlated_normal_lpdf(q):
    """test/targets.py:9-11."""
    rho = 0.5
    return -0.5 * q[0] ** 2 - 0.5 / (1 - rho ** 2) * (q[1] - rho * q[0]) ** 2


def correlated_normal_grad(q):
    """test/targets.py:12-15 (as written there: not the derivative of the density in component 0)."""
    rho = 0.5
    return np.array([-q[0] + rho * q[1], (-q[1] + rho * q[0]) / (1 - rho ** 2)])


def rosenbrock_lpdf(q):
    return smile(q)[0]


def rosenbrock_grad(q):
    return smile(q)[1]
