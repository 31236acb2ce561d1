"""GPU parity of user-defined CUDA targets (the role of the reference's arbitrary
Python lpFun / logp, grad callables).  The reference's remaining targets (targetDistr.py smileDistr, modFunnel,
funnel1, funnel10rescaled; test/targets.py correlated_normal, rosenbrock) are written as user targets in
walnuts_b200/targets.py; the oracle runs their numpy restatements (pinned to the reference's own functions in the
CPU suite) through the reference-pinned transition on the same Philox streams."""
import numpy as np
import pytest

from oracle import package_oracle as po
from oracle import targets as ot
from oracle import walnutspy_oracle as wo
from tests.helpers import KIND, close

pytestmark = pytest.mark.gpu

EXACT = [0, 1, 4, 5, 6, 7, 8, 9, 12, 13, 19, 20, 21, 22]


def _targets():
    import walnuts_b200 as wb
    t = wb.targets
    return {"smile": (t.smileDistr, ot.smile, 2), "modFunnel": (t.modFunnel, ot.mod_funnel, 2),
            "funnel1": (t.funnel1, ot.funnel1, 2), "funnel10rescaled": (t.funnel10rescaled, ot.funnel10_rescaled, 11)}


def _integrators():
    import walnuts_b200 as wb
    return {"fixed": wb.fixedLeapFrog, "D": wb.adaptLeapFrogD, "R2P": wb.adaptLeapFrogR2P, "Yoshida": wb.adaptYoshidaD,
            "Flow": wb.adaptLeapFrogFlowD, "Midpoint": wb.adaptImplicitMidpointD, "Rescaled": wb.adaptRescaledLeapFrogD}


@pytest.mark.parametrize("integrator", ["fixed", "D", "R2P", "Yoshida", "Flow", "Midpoint", "Rescaled"])
@pytest.mark.parametrize("name", ["smile", "modFunnel", "funnel1", "funnel10rescaled"])
def test_user_target_per_transition(cuda_lib, name, integrator):
    """Every transition from the oracle's previous state (the funnel-type targets are chaotic, see test_funnel10):
    draws to 1e-10, discrete diagnostics exactly -- for all seven integrators through ONE plug-in kernel."""
    import walnuts_b200 as wb
    from walnuts_b200 import ChainBatch
    tg, lp, d = _targets()[name]
    n, n_iter, M = 3, 25, 7
    H0 = {"fixed": 0.15, "Flow": 0.3, "Rescaled": 0.25}.get(integrator, 0.4)
    delta = {"Flow": 0.02, "Midpoint": 0.05, "Yoshida": 0.05}.get(integrator, 0.3)
    q0 = 0.6 * np.random.default_rng(8).standard_normal((n, d))
    tid, data = wb.targets.resolve(tg, d)
    ref = np.empty((n_iter, n, d))
    refd = np.empty((n_iter, n, 24))
    for c in range(n):
        with np.errstate(all="ignore"):
            s, dg = wo.WALNUTS(lp, q0[c], integrator=KIND[integrator], H0=H0, delta0=delta, numIter=n_iter, M=M,
                               seed=77, chain=c)
        ref[:, c], refd[:, c] = s[:, 1:].T, dg
    with ChainBatch(tid, d, n, integrator=_integrators()[integrator].kind, H0=H0, delta=delta, M=M, seed=77,
                    data=data) as cb:
        prev = q0
        for it in range(n_iter):
            cb.set_state(prev)
            out = cb.run(1, draws=True, diag=True)
            ok, err = close(out["draws"][0], ref[it])
            assert ok, f"transition {it}: {err:.3e}"
            assert np.array_equal(out["diag"][0][:, EXACT], refd[it][:, EXACT]), f"transition {it}"
            prev = ref[it]


def test_user_target_free_running_with_default_adaptation(cuda_lib):
    """The reference's default call -- WALNUTS(lpFun, q0) with warm-up adaptation -- on a user target."""
    import walnuts_b200 as wb
    q0 = np.array([[0.3, -0.2], [1.0, 0.5]])
    s, d = wb.WALNUTS(wb.targets.smileDistr, q0, integrator=wb.adaptLeapFrogR2P, numIter=120, warmupIter=80, M=8, seed=5)
    for c in range(2):
        with np.errstate(all="ignore"):
            so, do = wo.WALNUTS(ot.smile, q0[c], integrator=wo.ADAPT_R2P, numIter=120, warmupIter=80, M=8, seed=5,
                                chain=c, adaptH=True, adaptDelta=True)
        ok, err = close(s[c], so, rtol=1e-8, axis=-2)
        assert ok, err
        assert np.array_equal(d[c][:, [1, 6, 7, 19]], do[:, [1, 6, 7, 19]])
        ok, err = close(d[c][:, [15, 18]], do[:, [15, 18]], rtol=1e-8)
        assert ok, err


def test_user_target_with_data_matches_builtin(cuda_lib):
    """A user target reading its `data` array: the diagonal Gaussian written by hand gives the built-in target's
    draws (same arithmetic per coordinate; only the summation order of lp differs)."""
    import walnuts_b200 as wb
    d = 6
    sigma = np.logspace(-1, 1, d)
    src = """
    WN_TARGET_LP_GRAD(q, g, data, n_data) {
      double lp = 0.0;
      for (int i = 0; i < WN_D; ++i) { g[i] = -q[i] * data[i]; lp += q[i] * g[i]; }
      return 0.5 * lp;
    }"""
    tg = wb.targets.cuda_target(src, d, data=1.0 / sigma ** 2, name="my_diag")
    q0 = np.random.default_rng(1).standard_normal((5, d)) * sigma
    kw = dict(integrator=wb.adaptLeapFrogD, H0=0.3, delta0=0.3, numIter=40, warmupIter=0, M=7, adaptH=False,
              adaptDelta=False, seed=11)
    s1, d1 = wb.WALNUTS(tg, q0, **kw)
    s2, d2 = wb.WALNUTS(wb.targets.diag_gauss(sigma), q0, **kw)
    ok, err = close(s1, s2, axis=-2)
    assert ok, err
    assert np.array_equal(d1[..., EXACT], d2[..., EXACT])


def test_user_target_d40_spilled_state(cuda_lib):
    """d = 40: beyond what one thread keeps in registers (state in thread-local memory); same draws as the built-in."""
    import walnuts_b200 as wb
    d = 40
    sigma = np.logspace(-1, 1, d)
    src = """
    WN_TARGET_LP_GRAD(q, g, data, n_data) {
      double lp = 0.0;
      for (int i = 0; i < WN_D; ++i) { g[i] = -q[i] * data[i]; lp += q[i] * g[i]; }
      return 0.5 * lp;
    }"""
    tg = wb.targets.cuda_target(src, d, data=1.0 / sigma ** 2, name="my_diag40")
    q0 = np.random.default_rng(2).standard_normal((3, d)) * sigma
    kw = dict(integrator=wb.adaptLeapFrogR2P, H0=0.25, delta0=0.3, numIter=15, warmupIter=0, M=6, adaptH=False,
              adaptDelta=False, seed=12)
    s1, d1 = wb.WALNUTS(tg, q0, **kw)
    s2, d2 = wb.WALNUTS(wb.targets.diag_gauss(sigma), q0, **kw)
    ok, err = close(s1, s2, axis=-2)
    assert ok, err
    assert np.array_equal(d1[..., EXACT], d2[..., EXACT])


@pytest.mark.parametrize("name", ["correlated_normal", "rosenbrock"])
def test_user_target_package_mode(cuda_lib, name):
    """walnuts(rng, theta_init, logp, grad, ...) of the package (walnuts.py:362) on test/targets.py's remaining
    densities -- including correlated_normal's gradient exactly as the reference writes it."""
    import walnuts_b200 as wb
    tg = {"correlated_normal": wb.targets.correlated_normal_lpdf, "rosenbrock": wb.targets.rosenbrock_lpdf}[name]
    f, g = {"correlated_normal": (ot.correlated_normal_lpdf, ot.correlated_normal_grad),
            "rosenbrock": (ot.rosenbrock_lpdf, ot.rosenbrock_grad)}[name]
    theta0 = np.array([[0.2, -0.1], [0.5, 0.4], [-0.7, 0.3]])
    inv_mass = np.array([1.0, 0.8])
    n_iter = 12
    for c in range(3):
        ref = po.walnuts(31, c, theta0[c], f, g, inv_mass, 0.9, 6, 0.2, 0, n_iter)
        # per transition on identical inputs: the kernel continues the keyed streams of (seed, chain, iteration)
        draws = wb.walnuts(None, theta0[c], tg, tg, inv_mass, 0.9, 6, 0.2, 0, n_iter, seed=31, chain_offset=c)
        k = 0
        while k < n_iter and close(draws[k], ref[k])[0]:
            k += 1
        # the Gaussian agrees over the whole run; the quartic density until its dynamics amplify rounding
        assert k == n_iter if name == "correlated_normal" else k >= 4, (k, draws[:k + 1], ref[:k + 1])


def test_user_target_d200_warp_per_chain(cuda_lib):
    """64 < d <= 512: one warp per chain (the user's sequential function evaluated by every lane on the whole vector in
    shared memory); same draws as the built-in diagonal Gaussian, WALNUTSpy and package semantics."""
    import walnuts_b200 as wb
    d = 200
    sigma = np.logspace(-1, 1, d)
    src = """
    WN_TARGET_LP_GRAD(q, g, data, n_data) {
      double lp = 0.0;
      for (int i = 0; i < WN_D; ++i) { g[i] = -q[i] * data[i]; lp += q[i] * g[i]; }
      return 0.5 * lp;
    }"""
    tg = wb.targets.cuda_target(src, d, data=1.0 / sigma ** 2, name="my_diag200")
    q0 = np.random.default_rng(2).standard_normal((9, d)) * sigma
    for ig, H0 in ((wb.adaptLeapFrogR2P, 0.2), (wb.fixedLeapFrog, 0.05), (wb.adaptYoshidaD, 0.3)):
        kw = dict(integrator=ig, H0=H0, delta0=0.3, numIter=10, warmupIter=0, M=6, adaptH=False, adaptDelta=False, seed=12)
        s1, d1 = wb.WALNUTS(tg, q0, **kw)
        s2, d2 = wb.WALNUTS(wb.targets.diag_gauss(sigma), q0, **kw)
        ok, err = close(s1, s2, axis=-2)
        assert ok, (ig, err)
        assert np.array_equal(d1[..., EXACT], d2[..., EXACT]), ig
    # package semantics on the same plug-in (package_kernel<UserWarpT, ...>), incl. a macro step small enough for the
    # ell = 0 defect path (walnuts.py:194)
    bi = wb.targets.diag_gauss(sigma)
    for macro in (0.35, 0.02):
        p1 = wb.walnuts(None, q0, tg, tg, np.ones(d), macro, 5, 0.3, 0, 6, seed=9)
        p2 = wb.walnuts(None, q0, bi, bi, np.ones(d), macro, 5, 0.3, 0, 6, seed=9)
        ok, err = close(p1, p2, axis=-1)
        assert ok, (macro, err)
    with pytest.raises(ValueError):
        wb.targets.cuda_target(src, 513)
