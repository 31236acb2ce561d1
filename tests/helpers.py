"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import numpy as np

from oracle import targets as otargets
from oracle import walnutspy_oracle as wo

RTOL = 1e-10   # north_star: per-iteration draws match in fp64 to 1e-10 relative error


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(1.0, np.abs(b))
    with np.errstate(invalid="ignore"):
        bad = ~((np.abs(a - b) <= rtol * scale) | (np.isnan(a) & np.isnan(b)) | (a == b))
    return not bad.any(), (np.nanmax(np.abs(a - b) / scale) if a.size else 0.0)


def oracle_target(name, d, data=None):
    if name == "std_normal":
        return otargets.std_normal
    if name == "diag_gauss":
        return otargets.make_diag_gauss(1.0 / np.sqrt(data["inv_var"]))
    if name == "funnel":
        return otargets.funnel10
    if name == "corr_gauss":
        return otargets.corr_gauss
    if name == "logreg":
        return otargets.make_logreg(data["X"], data["y"], float(np.asarray(data.get("tau", 1.0)).ravel()[0]))
    if name == "stock_watson":
        return otargets.make_stock_watson(data["y"])
    raise KeyError(name)


KIND = {"fixed": wo.FIXED, "D": wo.ADAPT_D, "R2P": wo.ADAPT_R2P, "Yoshida": wo.ADAPT_YOSHIDA,
        "Flow": wo.ADAPT_FLOW, "Midpoint": wo.ADAPT_MIDPOINT, "Rescaled": wo.ADAPT_RESCALED}


def oracle_walnutspy(name, q0, integrator, H0, delta, M, n_iter, seed, chains, minC=0, maxC=10, data=None,
                     jitter=0.2, first_iteration=1, lp=None):
    """Run the numpy oracle for the listed chain ids; returns draws (n_iter, len(chains), d), diag."""
    q0 = np.asarray(q0, dtype=np.float64)
    d = q0.shape[-1]
    lp = lp or oracle_target(name, d, data)
    draws = np.empty((n_iter, len(chains), d))
    diag = np.empty((n_iter, len(chains), 24))
    for k, c in enumerate(chains):
        with np.errstate(all="ignore"):
            s, dg = wo.WALNUTS(lp, q0[c] if q0.ndim == 2 else q0, integrator=KIND[integrator], H0=H0,
                               stepSizeRandScale=jitter, delta0=delta, numIter=n_iter, M=M,
                               igrAux=wo.AuxPar(minC, maxC), seed=seed, chain=c,
                               first_iteration=first_iteration)
        draws[:, k, :] = s[:, 1:].T
        diag[:, k, :] = dg
    return draws, diag
