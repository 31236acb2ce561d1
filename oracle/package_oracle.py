"""CPU restatement of the `walnuts` package transition.  TEST INFRASTRUCTURE ONLY.

Restates reference walnuts/walnuts.py:16-33 (uturn), :62-70 (sub_uturn), :74-95 (leapfrog),
:127-141 (H), :144-182 (stable_steps), :185-208 (micro-step law), :211-276 (extend_orbit),
:279-359 (walnuts_step) and :362-408 (walnuts) in the streaming form the CUDA kernel uses:

  * the dyadic sub-U-turn checks (:62-70, done after the whole extension in the reference) run
    online, in the post-order of the leaf counter, and the extension stops at the first hit;
    the set of spans is identical so the boolean is identical;
  * the orbit is never stored: ends, a running log-sum-exp and an online (reservoir) multinomial
    pick replace the lists (:326,345-358).  The pick has the same law as
    `rng.choice(orbit_ext, p=softmax(w_ext))` (:349-350) -- SURVEY.md appendix D item 7;
  * every scalar draw is keyed by (transition, depth, step) instead of a running count, so that
    early exit does not shift later draws (oracle/philox.py STREAM_PKG_*).

`compat=True` (default) reproduces the reference's two latent defects bit-for-bit
(SURVEY.md rows B3, B5): backward extensions store the negated momentum (:242-245,272), and
`choose_micro_steps` may return ell = 0 (:194) giving an infinite step.  `compat=False`
stores forward-time momenta and draws ell from {max(1, ell_s//2), ell_s, 2 ell_s}.

The rng shim `KeyedPackageRNG` feeds the REAL walnuts.py the same keyed Philox draws; it is
what tests/golden/make_golden.py and tests/test_oracle_golden.py use.
"""
import math

import numpy as np

from . import philox

MAX_N = 10              # stable_steps tries ell = 2**n for n in range(11) (:161)
NEG_LOG3 = -math.log(3.0)


def uturn(t1, r1, t2, r2, inv_mass):
    """walnuts.py:16-33."""
    diff = inv_mass * (t2 - t1)
    return bool(np.dot(r1, diff) < 0 or np.dot(r2, diff) < 0)


def hamiltonian(theta, rho, logp, inv_mass):
    """walnuts.py:97-141."""
    return -logp(theta) + 0.5 * np.dot(inv_mass, rho ** 2)


def leapfrog(grad, theta, rho, step_size, inv_mass, num_steps, counter):
    """walnuts.py:74-95 (num_steps == 0 degenerates to one step of size inf: defect B3)."""
    half = 0.5 * step_size
    sim = step_size * inv_mass
    rho = rho + half * grad(theta)
    counter[0] += 1
    for _ in range(num_steps - 1):
        theta = theta + sim * rho
        rho = rho + step_size * grad(theta)
        counter[0] += 1
    theta = theta + sim * rho
    rho = rho + half * grad(theta)
    counter[0] += 1
    return theta, rho


def stable_steps(theta0, rho0, logp, grad, inv_mass, macro_step, max_error, counter):
    """walnuts.py:144-182; returns ell (the success flag is ignored by the caller, :253,261)."""
    ell = 1
    for n in range(MAX_N + 1):
        theta, rho = theta0, rho0
        ell = 2 ** n
        step = macro_step / ell
        H_min = H_max = hamiltonian(theta, rho, logp, inv_mass)
        half = 0.5 * step
        sim = step * inv_mass
        rho = rho + half * grad(theta)
        counter[0] += 1
        for _ in range(ell - 1):
            theta = theta + sim * rho
            g = grad(theta)
            counter[0] += 1
            rho = rho + half * g
            Hc = hamiltonian(theta, rho, logp, inv_mass)
            H_min, H_max = min(H_min, Hc), max(H_max, Hc)
            rho = rho + half * g
        theta = theta + sim * rho
        rho = rho + half * grad(theta)
        counter[0] += 1
        Hc = hamiltonian(theta, rho, logp, inv_mass)
        H_min, H_max = min(H_min, Hc), max(H_max, Hc)
        if H_max - H_min <= max_error:
            return ell
    return ell


def micro_steps_logp(ell, ell_stable):
    """walnuts.py:197-208."""
    if ell == ell_stable or ell == ell_stable // 2 or ell == ell_stable * 2:
        return NEG_LOG3
    return -math.inf


def _logaddexp(a, b):
    with np.errstate(all="ignore"):
        return float(np.logaddexp(a, b))


def walnuts_step(streams, theta, logp, grad, inv_mass, macro_step, max_nuts_depth, max_error,
                 compat=True, counter=None):
    """One package transition (walnuts.py:279-359) with keyed draws.  `streams` is an
    oracle.philox.ChainStreams positioned on this transition."""
    counter = counter if counter is not None else [0]
    theta = np.array(theta, dtype=np.float64)
    inv_mass = np.array(inv_mass, dtype=np.float64)
    if theta.ndim != 1:
        raise ValueError("theta not a vector")
    if inv_mass.ndim != 1:
        raise ValueError("inv_mass not a vector")
    if theta.size != inv_mass.size:
        raise ValueError("size mismatch between theta and inv_mass")
    if not macro_step > 0:
        raise ValueError("non-positive macro_step")
    if not max_nuts_depth > 0:
        raise ValueError("non-positive max_nuts_depth")
    if not max_error > 0:
        raise ValueError("non-positive max_error")
    D = theta.size
    rho = inv_mass ** -0.5 * streams.momentum(D)                              # :322-325
    w0 = -hamiltonian(theta, rho, logp, inv_mass)                             # :326
    # ends: [0] = left (orbit[0]), [1] = right (orbit[-1]); stored momentum as the reference stores it
    end_t = [theta, theta]
    end_r = [rho, rho]
    end_w = [w0, w0]
    lse_old = w0
    selected = theta
    with np.errstate(all="ignore"):
        for depth in range(max_nuts_depth):                                   # :328
            n_new = 2 ** depth
            back = math.floor(2.0 * streams.keyed(philox.STREAM_PKG_DIR, depth)) == 1   # :330
            side = 0 if back else 1
            th, rh, weight = end_t[side], end_r[side], end_w[side]
            if back:
                rh = -rh                                                       # :245
            lse_ext = -math.inf
            cand = None
            left = {}
            sub = False
            base = n_new - 1
            for k in range(1, n_new + 1):                                      # :251
                p0 = -hamiltonian(th, rh, logp, inv_mass)                      # :252
                ell_s = stable_steps(th, rh, logp, grad, inv_mass, macro_step, max_error, counter)
                u = streams.keyed(philox.STREAM_PKG_ELL, base + k - 1)
                choices = [ell_s // 2, ell_s, ell_s * 2]                       # :194
                if not compat:
                    choices[0] = max(1, choices[0])
                ell = choices[min(2, int(math.floor(3.0 * u)))]
                step = macro_step / ell if ell > 0 else math.inf               # numpy int division -> inf
                th, rh = leapfrog(grad, th, rh, step, inv_mass, ell, counter)  # :258
                ell_n = stable_steps(th, -rh, logp, grad, inv_mass, macro_step, max_error, counter)
                p1 = -hamiltonian(th, rh, logp, inv_mass)                      # :264
                weight = (p1 - p0 + micro_steps_logp(ell, ell_n) - micro_steps_logp(ell, ell_s)
                          + weight)                                            # :265-271
                # online multinomial pick (law of :349-350)
                lse_ext = _logaddexp(lse_ext, weight)
                us = streams.keyed(philox.STREAM_PKG_SELECT, base + k - 1)
                if us < math.exp(weight - lse_ext) if lse_ext > -math.inf else False:
                    cand = th
                # stored state: the reference appends (theta, rho) with rho in integration
                # convention (:272); compat=False stores forward-time momentum instead
                rs = rh if (compat or not back) else -rh
                if k % 2 == 1:
                    lvl = depth if k == 1 else ((k - 1) & -(k - 1)).bit_length() - 1
                    left[lvl] = (th, rs)
                else:
                    s = 1
                    while s <= depth and k % (2 ** s) == 0:
                        m = k - 2 ** s + 1
                        lvl = depth if m == 1 else ((m - 1) & -(m - 1)).bit_length() - 1
                        tl, rl = left[lvl]
                        # reference order after the [::-1] of :275: earlier list position first
                        if back:
                            ut = uturn(th, rs, tl, rl, inv_mass)
                        else:
                            ut = uturn(tl, rl, th, rs, inv_mass)
                        if ut:
                            sub = True
                            break
                        s += 1
                    if sub:
                        break
            if sub:                                                            # :343-344
                break
            ua = streams.keyed(philox.STREAM_PKG_ACCEPT, depth)
            if math.log(ua) < lse_ext - lse_old:                               # :345-347
                selected = cand if cand is not None else selected              # :349-350
            end_t[side], end_r[side], end_w[side] = th, rs, weight             # :351 (new end)
            if uturn(end_t[0], end_r[0], end_t[1], end_r[1], inv_mass):        # :352
                break
            lse_old = _logaddexp(lse_old, lse_ext)                             # :354-358
    return selected


def walnuts(seed, chain, theta_init, logp, grad, inv_mass, macro_step, max_nuts_depth, max_error,
            iter_warmup, iter_sample, compat=True, first_iteration=1, counter=None):
    """walnuts.py:362-408 with Philox streams keyed by (seed, chain)."""
    streams = philox.ChainStreams(seed, chain)
    theta = np.array(theta_init, dtype=np.float64)
    draws = np.empty((iter_sample, theta.size))
    for i in range(iter_warmup + iter_sample):
        streams.begin_iteration(first_iteration + i)
        theta = walnuts_step(streams, theta, logp, grad, inv_mass, macro_step, max_nuts_depth,
                             max_error, compat=compat, counter=counter)
        if i >= iter_warmup:
            draws[i - iter_warmup] = theta
    return draws


class KeyedPackageRNG:
    """Duck-typed `rng` for the REAL walnuts.py: the same keyed Philox draws as walnuts_step()
    above, recognised by call order (normal -> [binomial -> choice(3) x 2**depth -> uniform ->
    choice(p)] per depth)."""

    def __init__(self, seed, chain, first_iteration=1):
        self.streams = philox.ChainStreams(seed, chain)
        self.it = first_iteration - 1
        self.depth = -1
        self.step = 0
        self.back = False

    def normal(self, size):
        self.it += 1
        self.streams.begin_iteration(self.it)
        self.depth = -1
        return self.streams.momentum(int(size))

    def binomial(self, n, p):
        self.depth += 1
        self.step = 0
        self.back = math.floor(2.0 * self.streams.keyed(philox.STREAM_PKG_DIR, self.depth)) == 1
        return int(self.back)

    def uniform(self, lo, hi):
        return self.streams.keyed(philox.STREAM_PKG_ACCEPT, self.depth)

    def choice(self, seq, p=None):
        base = 2 ** self.depth - 1
        if p is None:
            u = self.streams.keyed(philox.STREAM_PKG_ELL, base + self.step)
            self.step += 1
            return np.int64(seq[min(2, int(math.floor(3.0 * u)))])
        n = len(seq)
        p = np.asarray(p)
        order = range(n - 1, -1, -1) if self.back else range(n)   # generation order
        run = 0.0
        pick = None
        with np.errstate(all="ignore"):
            for k, pos in enumerate(order):
                run += p[pos]
                us = self.streams.keyed(philox.STREAM_PKG_SELECT, base + k)
                if us < p[pos] / run:
                    pick = pos
        return seq[pick if pick is not None else 0]
