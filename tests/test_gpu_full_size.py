"""Full-size (BASELINE.json chain counts) checks through size-independent properties: the transition
kernel leaves the target invariant, so chains started from exact target draws must keep the target's
moments after any number of transitions -- checked across chains within 5 Monte-Carlo standard errors
(independent chains => plain MCSE); and results must not depend on how chains are scheduled."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c2_diag_gauss_65536_chains_invariance(cuda_lib):
    """BASELINE config 2: d = 1000, sigma = logspace(-2, 2), 65 536 chains, R2P."""
    from walnuts_b200 import ChainBatch
    d, n = 1000, 65536
    sigma = np.logspace(-2, 2, d)
    rng = np.random.default_rng(11)
    q0 = rng.standard_normal((n, d)) * sigma
    with ChainBatch("diag_gauss", d, n, integrator="R2P", H0=0.5, delta=0.3, M=10, seed=3, dg=0,
                    data={"inv_var": 1.0 / sigma ** 2}) as cb:
        cb.set_state(q0)
        out = cb.run(2, draws=False, diag=True)
        mean, var = cb.moments()
        q1 = cb.get_state()
    z = q1 / sigma
    assert np.isfinite(z).all()
    # per-coordinate mean ~ N(0, 1/n), variance ~ 1 +- sqrt(2/n)
    assert np.abs(z.mean(0)).max() < 5.5 / np.sqrt(n)
    assert np.abs(z.var(0) - 1).max() < 5.5 * np.sqrt(2.0 / n)
    assert np.allclose(mean, q1.mean(0), rtol=1e-9, atol=1e-12) and np.allclose(var, q1.var(0, ddof=1), rtol=1e-9)
    # chains really moved, and the diagnostics are sane
    assert (np.abs(q1 - q0).max(1) > 0).mean() > 0.99
    dg = out["diag"]
    assert set(np.unique(dg[..., 19])) <= {-4.0, 0.0, 4.0, 5.0}
    assert (dg[..., 6] > 0).all()


def test_c3_funnel_262144_chains_invariance(cuda_lib):
    """BASELINE config 3: funnel10 (d = 11), 262 144 chains, R2P, M = 12, H0 = 0.3, delta = 0.3."""
    from walnuts_b200 import ChainBatch
    n = 262144
    rng = np.random.default_rng(12)
    q0 = np.empty((n, 11))
    q0[:, 0] = 3.0 * rng.standard_normal(n)
    q0[:, 1:] = np.exp(0.5 * q0[:, :1]) * rng.standard_normal((n, 10))
    with ChainBatch("funnel", 11, n, integrator="R2P", H0=0.3, delta=0.3, M=12, seed=4, dg=0) as cb:
        cb.set_state(q0)
        out = cb.run(3, draws=False, diag=True)
        q1 = cb.get_state()
    assert np.isfinite(q1).all()
    w = q1[:, 0]
    assert abs(w.mean()) < 5 * 3.0 / np.sqrt(n)
    assert abs(w.var() - 9.0) < 5 * 9.0 * np.sqrt(2.0 / n)
    zz = q1[:, 1:] * np.exp(-0.5 * w[:, None])                 # standardised: N(0,1) under the target
    assert np.abs(zz.mean(0)).max() < 5 / np.sqrt(n)
    assert np.abs(zz.var(0) - 1).max() < 5 * np.sqrt(2.0 / n)
    dg = out["diag"]
    assert (dg[..., 19] != 999).all()                          # no numerical rejects from stationarity
    assert dg[..., 22].max() >= 3                              # micro-step halving really exercised
