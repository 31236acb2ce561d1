"""Drop-in call surfaces of the reference, running on CUDA.

WALNUTS(...)  : reference WALNUTSpy/WALNUTS.py:111-727
walnuts(...)  : reference walnuts/walnuts.py:362-408;  walnuts_step(...) : walnuts.py:279-359

Differences a caller sees: `lpFun` / `logp` / `integrator` are registry handles
(walnuts_b200.targets / walnuts_b200.integrators) instead of Python callables; `q0` / `theta_init`
may carry a leading chains axis (n_chains, d), in which case every output gains a leading chains
axis; randomness comes from per-chain Philox streams keyed by `seed` (drawn from numpy's global RNG,
respectively from `rng`, when not given -- so `np.random.seed(k)` still makes runs reproducible).
"""
import numpy as np

from . import integrators as _ig
from . import targets as _tg
from .sampler import ChainBatch


def _seed_from_global():
    return int(np.random.randint(0, 2 ** 62))


def _index_generator(generated, d):
    """If `generated` only SELECTS coordinates (the reference's example scripts: `gen(q) = np.array([q[0], q[1]])`,
    WALNUTSpy_examples/funnel/mainFunnel.py:19-20), return the selected indices, else None.  Probed with two vectors
    of distinct values: every output must reproduce one input coordinate exactly, the same one both times."""
    rng = np.random.default_rng(12345)
    idx = None
    for _ in range(2):
        x = rng.permutation(d).astype(np.float64) + rng.uniform(0.1, 0.9)
        try:
            y = np.asarray(generated(x), dtype=np.float64).ravel()
        except Exception:
            return None
        pos = {v: i for i, v in enumerate(x)}
        cur = [pos.get(v) for v in y]
        if any(c is None for c in cur) or (idx is not None and cur != idx):
            return None
        idx = cur
    return np.asarray(idx, dtype=int)


def _over_devices(fn, x, devices, chain_offset):
    """Shard the chains axis of `x` (n_chains, d) over `devices` in contiguous blocks and run fn(x_shard, device,
    chain_offset) for every shard on its own host thread (ctypes releases the GIL inside the library, so the GPUs
    sample concurrently; no exchange between them: chains are independent, and the Philox streams are keyed by the
    GLOBAL chain id, so the draws do not depend on the number of devices).  Returns the per-shard results."""
    from concurrent.futures import ThreadPoolExecutor
    n = x.shape[0]
    devices = list(devices)
    if n < len(devices):
        devices = devices[:n]
    bounds = np.linspace(0, n, len(devices) + 1).astype(int)
    jobs = [(x[bounds[k]:bounds[k + 1]], dev, chain_offset + int(bounds[k])) for k, dev in enumerate(devices)]
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        return list(ex.map(lambda j: fn(*j), jobs))


def WALNUTS(lpFun, q0, generated=None, integrator=_ig.fixedLeapFrog, H0=0.2, stepSizeRandScale=0.2,
            delta0=0.05, numIter=2000, warmupIter=1000, M=10, igrAux=None, adaptH=True,
            adaptHtarget=0.8, adaptDelta=True, adaptDeltaTarget=0.6, adaptDeltaQuantile=0.9,
            recordOrbitStats=False, *, seed=None, device=0, devices=None, chain_offset=0, compat=True):
    """Many-chain WALNUTS/NUTS with the WALNUTSpy driver semantics (WALNUTS.py:111-129 arguments).

    Returns (samples, diagnostics): for a single chain (`q0.ndim == 1`) shapes are (dg, numIter+1) and
    (numIter, 24) exactly as WALNUTS.py:163,180,724-727; with q0 of shape (n_chains, d) a leading
    chains axis is added.

    `compat=True` (default) reproduces the reference's semantics including its defects: every discrete decision
    (the integer diagnostics columns) is identical and draws agree to 1e-10 relative with the reference fed the same
    Philox streams (the kernels use FMA contraction, merged kicks and CUDA's libm, so draws are not bit-identical).
    Defect A14(i): the second leaf
    of a BACKWARD pair never adds its log-weight to the running sum (WALNUTS.py:420 has no counterpart after
    :443-459), which measurably biases adaptive runs on the funnel.  `compat=False` adds it.
    """
    if adaptH and (adaptHtarget < 0.0 or adaptHtarget > 1.0):
        raise ValueError("bad adaptHtarget")          # sys.exit in the reference, WALNUTS.py:140
    if adaptDelta and adaptDeltaTarget < 0.0:
        raise ValueError("bad adaptDeltaTarget")      # WALNUTS.py:146
    if not isinstance(integrator, _ig._Integrator):
        raise TypeError("integrator must be one of walnuts_b200.fixedLeapFrog / adaptLeapFrogD / adaptLeapFrogR2P / "
                        "adaptYoshidaD / adaptLeapFrogFlowD / adaptImplicitMidpointD / adaptRescaledLeapFrogD")
    aux = igrAux or _ig.integratorAuxPar()
    if integrator is _ig.adaptImplicitMidpointD and getattr(aux, "FPNewton", False):
        raise NotImplementedError("adaptImplicitMidpointD with FPNewton=True needs the target's Hessian "
                                  "(adaptiveIntegrators.py:503-506); only the fixed-point variant runs on the GPU")
    q0 = np.asarray(q0, dtype=np.float64)
    single = q0.ndim == 1
    q = q0.reshape(1, -1) if single else q0
    n_chains, d = q.shape
    if seed is None:
        seed = _seed_from_global()
    if devices is not None and len(devices) > 1 and n_chains > 1:
        # several GPUs of this node: contiguous blocks of chains, one host thread per GPU
        kw = dict(generated=generated, integrator=integrator, H0=H0, stepSizeRandScale=stepSizeRandScale, delta0=delta0,
                  numIter=numIter, warmupIter=warmupIter, M=M, igrAux=igrAux, adaptH=adaptH, adaptHtarget=adaptHtarget,
                  adaptDelta=adaptDelta, adaptDeltaTarget=adaptDeltaTarget, adaptDeltaQuantile=adaptDeltaQuantile,
                  recordOrbitStats=recordOrbitStats, seed=seed, compat=compat)
        parts = _over_devices(lambda xs, dev, off: WALNUTS(lpFun, xs, device=dev, chain_offset=off, **kw), q, devices,
                              chain_offset)
        return tuple(np.concatenate([p_[i] for p_ in parts]) for i in range(len(parts[0])))
    if devices is not None and len(devices) == 1:
        device = devices[0]
    gen_idx = None
    if recordOrbitStats and generated is not None:
        # orbit minima / maxima of generated(q) over the orbit's states (WALNUTS.py:274-276,332-333,...) are kept on
        # the GPU per coordinate; that equals min / max of generated(q) exactly when `generated` selects coordinates
        gen_idx = _index_generator(generated, d)
        if gen_idx is None:
            raise NotImplementedError(
                "recordOrbitStats with a `generated` that is not a selection of coordinates: the GPU keeps the orbit's "
                "per-coordinate minima / maxima, which determine min / max of generated(q) only for index selections "
                "such as gen(q) = np.array([q[0], q[1]]) (mainFunnel.py:19-20)")
    name, data = _tg.resolve(lpFun, d)
    with ChainBatch(name, d, n_chains, mode="walnutspy", integrator=integrator.kind, H0=H0,
                    jitter=stepSizeRandScale, delta=delta0, M=M, minC=aux.minC, maxC=aux.maxC,
                    r2p_prob0=aux.R2Pprob0, seed=seed, chain_offset=chain_offset, device=device,
                    compat=compat, data=data) as cb:
        cb.set_aux(getattr(aux, "maxFPiter", None), getattr(aux, "FPtol", None),
                   getattr(aux, "rescaledGradThresh", None))
        if warmupIter > 0 and (adaptH or adaptDelta):
            cb.set_adapt(min(warmupIter, numIter), adaptH, adaptHtarget, adaptDelta, adaptDeltaTarget,
                         adaptDeltaQuantile)
        cb.set_state(q)
        nw = min(warmupIter, numIter) if (adaptH or adaptDelta) else 0
        parts = []
        if nw > 0:
            parts.append(cb.run(nw, draws=True, diag=True, orbit_stats=recordOrbitStats))    # adapting iterations
        if numIter - nw > 0:
            parts.append(cb.run(numIter - nw, draws=True, diag=True, orbit_stats=recordOrbitStats))
        keys = ("draws", "diag") + (("orbit_min", "orbit_max") if recordOrbitStats else ())
        out = {k: np.concatenate([p_[k] for p_ in parts]) for k in keys}
    draws = out["draws"]                                    # (numIter, n_chains, d)
    gen = generated if generated is not None else (lambda x: x)
    g0 = np.asarray(gen(q[0]))
    samples = np.empty((n_chains, g0.size, numIter + 1))
    if generated is None:
        samples[:, :, 0] = q
        samples[:, :, 1:] = np.transpose(draws, (1, 2, 0))
    else:
        for c in range(n_chains):
            samples[c, :, 0] = gen(q[c])
            for i in range(numIter):
                samples[c, :, i + 1] = gen(draws[i, c])
    diagnostics = np.ascontiguousarray(np.transpose(out["diag"], (1, 0, 2)))   # (n_chains, numIter, 24)
    if recordOrbitStats:                                    # (dg, numIter) per chain, WALNUTS.py:183-184,724-725
        omin = np.ascontiguousarray(np.transpose(out["orbit_min"], (1, 2, 0)))
        omax = np.ascontiguousarray(np.transpose(out["orbit_max"], (1, 2, 0)))
        if gen_idx is not None:
            omin, omax = np.ascontiguousarray(omin[:, gen_idx]), np.ascontiguousarray(omax[:, gen_idx])
        if single:
            return samples[0], diagnostics[0], omin[0], omax[0]
        return samples, diagnostics, omin, omax
    if single:
        return samples[0], diagnostics[0]
    return samples, diagnostics


def _pkg_batch(rng, theta, logp, grad, inv_mass, macro_step, max_nuts_depth, max_error, seed, device,
               compat, chain_offset, first_iteration=1):
    theta = np.array(theta, dtype=np.float64)
    inv_mass = np.array(inv_mass, dtype=np.float64)
    single = theta.ndim == 1
    if theta.ndim not in (1, 2):
        raise ValueError("theta not a vector")                      # walnuts.py:309-310
    if inv_mass.ndim != 1:
        raise ValueError("inv_mass not a vector")                   # :311-312
    th = theta.reshape(1, -1) if single else theta
    if th.shape[1] != inv_mass.size:
        raise ValueError("size mismatch between theta and inv_mass")  # :313-314
    if not macro_step > 0:
        raise ValueError("non-positive macro_step")                 # :315-316
    if not max_nuts_depth > 0:
        raise ValueError("non-positive max_nuts_depth")             # :317-318
    if not max_error > 0:
        raise ValueError("non-positive max_error")                  # :319-320
    if logp is not grad and not (isinstance(logp, _tg.Target) and isinstance(grad, _tg.Target)
                                 and logp.name == grad.name):
        raise TypeError("logp and grad must be the same walnuts_b200.targets handle")
    name, data = _tg.resolve(logp, th.shape[1])
    if seed is None:
        seed = int(rng.integers(0, 2 ** 62)) if hasattr(rng, "integers") else int(rng)
    data = dict(data)
    data["inv_mass"] = inv_mass
    cb = ChainBatch(name, th.shape[1], th.shape[0], mode="package", H0=macro_step, delta=max_error,
                    M=max_nuts_depth, seed=seed, chain_offset=chain_offset, device=device, compat=compat,
                    data=data, first_iteration=first_iteration)
    cb.set_state(th)
    return cb, single


def walnuts(rng, theta_init, logp, grad, inv_mass, macro_step, max_nuts_depth, max_error, iter_warmup,
            iter_sample, *, seed=None, device=0, devices=None, compat=True, chain_offset=0):
    """walnuts.py:362-408.  Returns draws (iter_sample, D), or (n_chains, iter_sample, D) when
    `theta_init` has a leading chains axis.  `devices=[0, 1, ...]` shards the chains over several GPUs."""
    th = np.asarray(theta_init, dtype=np.float64)
    if devices is not None and len(devices) > 1 and th.ndim == 2 and th.shape[0] > 1:
        if seed is None:
            seed = int(rng.integers(0, 2 ** 62)) if hasattr(rng, "integers") else int(rng)
        parts = _over_devices(lambda xs, dev, off: walnuts(rng, xs, logp, grad, inv_mass, macro_step, max_nuts_depth,
                                                           max_error, iter_warmup, iter_sample, seed=seed, device=dev,
                                                           compat=compat, chain_offset=off), th, devices, chain_offset)
        return np.concatenate(parts)
    if devices is not None and len(devices) == 1:
        device = devices[0]
    cb, single = _pkg_batch(rng, theta_init, logp, grad, inv_mass, macro_step, max_nuts_depth, max_error,
                            seed, device, compat, chain_offset)
    with cb:
        if iter_warmup > 0:
            cb.run(iter_warmup, draws=False, nevals=False)
        draws = cb.run(iter_sample, draws=True, nevals=False)["draws"]   # (iter, chains, D)
    draws = np.ascontiguousarray(np.transpose(draws, (1, 0, 2)))
    return draws[0] if single else draws


def walnuts_step(rng, theta, logp, grad, inv_mass, macro_step, max_nuts_depth, max_error, *, seed=None,
                 iteration=None, device=0, compat=True, chain_offset=0):
    """walnuts.py:279-359: one transition; returns the next state vector(s).

    Randomness: the Philox streams are keyed by (seed, chain, iteration).  With `seed=None` a fresh seed is drawn
    from `rng` on every call, as the reference consumes `rng`.  With a FIXED `seed` the caller must tell the
    transitions apart: pass `iteration` (1, 2, 3, ... as walnuts() counts them; same seed + same iteration replays
    the same momentum, directions and uniforms); if it is omitted, one value is drawn from `rng` per call and used
    as the iteration number, so a loop over walnuts_step(rng, ..., seed=k) never repeats its streams."""
    if iteration is None:
        iteration = 1 if seed is None else (int(rng.integers(1, 2 ** 31)) if hasattr(rng, "integers") else 1)
    if not 1 <= int(iteration) < 2 ** 31:
        raise ValueError("iteration must be in [1, 2^31)")
    cb, single = _pkg_batch(rng, theta, logp, grad, inv_mass, macro_step, max_nuts_depth, max_error,
                            seed, device, compat, chain_offset, int(iteration))
    with cb:
        cb.run(1, draws=False, nevals=False)
        out = cb.get_state()
    return out[0] if single else out
