"""GPU parity of the warm-up adaptation (reference WALNUTS.py:136-147, 313, 701-712 + P2quantile.py):
the drop-in WALNUTS(...) call with the reference's DEFAULT arguments (adaptH=True, adaptDelta=True)
against the numpy oracle, which is pinned bit-exact to the real reference (tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from oracle import targets as ot
from oracle import walnutspy_oracle as wo
from tests.helpers import KIND, close

pytestmark = pytest.mark.gpu


def run_pair(target_name, lp, q0, integrator, numIter, warmupIter, M, seed=77, **kw):
    import walnuts_b200 as wb
    tg = {"std_normal": wb.targets.stdGauss, "corr_gauss": wb.targets.corrGauss, "funnel": wb.targets.funnel10}[target_name]
    ig = {"fixed": wb.fixedLeapFrog, "D": wb.adaptLeapFrogD, "R2P": wb.adaptLeapFrogR2P}[integrator]
    s, d = wb.WALNUTS(tg, q0, integrator=ig, numIter=numIter, warmupIter=warmupIter, M=M, seed=seed, **kw)
    n_chains = q0.shape[0]
    so = np.empty_like(s)
    do = np.empty_like(d)
    for c in range(n_chains):
        so[c], do[c] = wo.WALNUTS(lp, q0[c], integrator=KIND[integrator], numIter=numIter, warmupIter=warmupIter, M=M,
                                  seed=seed, chain=c, adaptH=kw.get("adaptH", True), adaptDelta=kw.get("adaptDelta", True),
                                  H0=kw.get("H0", 0.2), delta0=kw.get("delta0", 0.05))
    return s, d, so, do


@pytest.mark.parametrize("integrator", ["fixed", "D", "R2P"])
@pytest.mark.parametrize("target", ["std_normal", "corr_gauss"])
def test_default_adaptation_matches_oracle(cuda_lib, target, integrator):
    d = 6 if target == "std_normal" else 2
    lp = ot.std_normal if target == "std_normal" else ot.corr_gauss
    q0 = 0.5 * np.random.default_rng(2).standard_normal((4, d))
    s, dg, so, do = run_pair(target, lp, q0, integrator, numIter=120, warmupIter=80, M=8)
    ok, err = close(s, so, axis=-2)
    assert ok, f"draws: {err:.3e}"
    ok, err = close(dg[..., [15, 18]], do[..., [15, 18]], rtol=1e-9)       # adapted H and delta, every iteration
    assert ok, f"adapted H / delta: {err:.3e}"
    assert np.array_equal(dg[..., [0, 1, 6, 7, 19]], do[..., [0, 1, 6, 7, 19]])
    # adaptation really happened and then froze
    assert not np.allclose(dg[:, 79, 15], 0.2) and np.array_equal(dg[:, 80, 15], dg[:, -1, 15])


def test_adapt_delta_only_and_h_only(cuda_lib):
    q0 = 0.5 * np.random.default_rng(3).standard_normal((3, 5))
    for kw in (dict(adaptH=False, adaptDelta=True), dict(adaptH=True, adaptDelta=False)):
        s, dg, so, do = run_pair("std_normal", ot.std_normal, q0, "R2P", numIter=70, warmupIter=50, M=7, **kw)
        ok, err = close(s, so, axis=-2)
        assert ok, (kw, err)
        ok, err = close(dg[..., [15, 18]], do[..., [15, 18]], rtol=1e-9)
        assert ok, (kw, err)


def test_funnel_adaptation_prefix(cuda_lib):
    """Chaotic target: free-running agreement is only expected over a prefix (see test_funnel10)."""
    rng = np.random.default_rng(5)
    q0 = np.empty((3, 11))
    q0[:, 0] = rng.standard_normal(3)
    q0[:, 1:] = np.exp(0.5 * q0[:, :1]) * rng.standard_normal((3, 10))
    s, dg, so, do = run_pair("funnel", ot.funnel10, q0, "R2P", numIter=40, warmupIter=40, M=9)
    err = np.max(np.abs(s - so) / np.maximum(1, np.abs(so)), axis=(0, 1))
    horizon = int(np.argmax(err > 1e-9)) if (err > 1e-9).any() else len(err)
    assert horizon >= 20, err[:25]
    ok, e2 = close(dg[:, :15, [15, 18]], do[:, :15, [15, 18]], rtol=1e-7)
    assert ok, e2


@pytest.mark.parametrize("integrator", ["fixed", "R2P"])
def test_record_orbit_stats(cuda_lib, integrator):
    """WALNUTS(recordOrbitStats=True) (WALNUTS.py:182-184,274-276,331-333,...): per-iteration element-wise
    min / max over every state of the orbit, returned as (dg, numIter) arrays after samples, diagnostics."""
    import walnuts_b200 as wb
    ig = {"fixed": wb.fixedLeapFrog, "R2P": wb.adaptLeapFrogR2P}[integrator]
    q0 = 0.5 * np.random.default_rng(8).standard_normal((3, 5))
    s, d, lo, hi = wb.WALNUTS(wb.targets.stdGauss, q0, integrator=ig, numIter=60, warmupIter=20, M=7, seed=5,
                              recordOrbitStats=True)
    assert lo.shape == (3, 5, 60) and hi.shape == (3, 5, 60)
    for c in range(3):
        so, do, lo_o, hi_o = wo.WALNUTS(ot.std_normal, q0[c], integrator=KIND[integrator], numIter=60, warmupIter=20,
                                        M=7, seed=5, chain=c, adaptH=True, adaptDelta=True, recordOrbitStats=True)
        for a, b in ((s[c], so), (lo[c], lo_o), (hi[c], hi_o)):
            ok, err = close(a, b, axis=-2)
            assert ok, err
    assert (lo <= s[:, :, 1:]).all() and (s[:, :, 1:] <= hi).all()


def test_record_orbit_stats_with_index_generator(cuda_lib):
    """The reference's flagship call: recordOrbitStats=True together with generated=gen, gen(q) = [q[0], q[1]]
    (WALNUTSpy_examples/funnel/mainFunnel.py:19-20,35-42): orbit minima / maxima of the selected coordinates."""
    import walnuts_b200 as wb

    def gen(q):
        return np.array([q[0], q[1]])
    rng = np.random.default_rng(3)
    q0 = np.empty((3, 11))
    q0[:, 0] = rng.standard_normal(3)
    q0[:, 1:] = np.exp(0.5 * q0[:, :1]) * rng.standard_normal((3, 10))
    kw = dict(integrator=wb.adaptLeapFrogR2P, M=8, H0=0.3, delta0=0.3, numIter=12, warmupIter=0, seed=6)
    s, d, lo, hi = wb.WALNUTS(wb.targets.funnel10, q0, generated=gen, recordOrbitStats=True, **kw)
    s_all, d_all, lo_all, hi_all = wb.WALNUTS(wb.targets.funnel10, q0, recordOrbitStats=True, **kw)
    assert s.shape == (3, 2, 13) and lo.shape == (3, 2, 12) and hi.shape == (3, 2, 12)
    assert np.array_equal(s, s_all[:, :2]) and np.array_equal(d, d_all)
    assert np.array_equal(lo, lo_all[:, :2]) and np.array_equal(hi, hi_all[:, :2])
    for c in range(3):          # and the oracle (identity statistics, sliced) on the first transitions
        so, do, lo_o, hi_o = wo.WALNUTS(ot.funnel10, q0[c], integrator=KIND["R2P"], numIter=3, M=8, H0=0.3, delta0=0.3,
                                        seed=6, chain=c, recordOrbitStats=True)
        ok, err = close(lo[c][:, :3], lo_o[:2], axis=-2)
        assert ok, err
        ok, err = close(hi[c][:, :3], hi_o[:2], axis=-2)
        assert ok, err
    with pytest.raises(NotImplementedError):
        wb.WALNUTS(wb.targets.funnel10, q0, generated=lambda q: np.array([q[0] + q[1]]), recordOrbitStats=True, **kw)
