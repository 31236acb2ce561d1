"""walnuts_b200: B200-native many-chain WALNUTS/NUTS sampler (CUDA behind a ctypes C-ABI).

Drop-in call surfaces of the reference (bob-carpenter/walnuts):
    walnuts(rng, theta_init, logp, grad, inv_mass, macro_step, max_nuts_depth, max_error,
            iter_warmup, iter_sample)                         reference walnuts/walnuts.py:362
    WALNUTS(lpFun, q0, generated, integrator, H0, ...)        reference WALNUTSpy/WALNUTS.py:111
with `lpFun` / `logp` taken from the CUDA target registry in `walnuts_b200.targets`.
"""
from ._ffi import WalnutsError  # noqa: F401
from .sampler import ChainBatch, fp64_peak, pinned_empty, comm_unique_id, comm_init_all  # noqa: F401
from . import targets, integrators  # noqa: F401
from .integrators import (fixedLeapFrog, adaptLeapFrogD, adaptLeapFrogR2P, adaptYoshidaD,  # noqa: F401
                          adaptLeapFrogFlowD, adaptImplicitMidpointD, adaptRescaledLeapFrogD, integratorAuxPar)
from .api import WALNUTS, walnuts, walnuts_step  # noqa: F401

__all__ = ["ChainBatch", "WALNUTS", "walnuts", "walnuts_step", "targets", "integrators",
           "fixedLeapFrog", "adaptLeapFrogD", "adaptLeapFrogR2P", "adaptYoshidaD", "adaptLeapFrogFlowD",
           "adaptImplicitMidpointD", "adaptRescaledLeapFrogD", "integratorAuxPar",
           "WalnutsError", "fp64_peak"]
