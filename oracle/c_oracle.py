"""ctypes wrapper of the C restatement oracle/c/walnuts_oracle.c.  TEST INFRASTRUCTURE / CPU BASELINE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "libwalnuts_oracle.so")
TARGET = {"std_normal": 0, "diag_gauss": 1, "funnel": 2}
KIND = {"fixed": 0, "D": 1, "R2P": 2}
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            subprocess.run(["make", "-s", "-C", os.path.join(HERE, "c")], check=True)
        _lib = C.CDLL(LIB)
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def run_chain(target, integrator, q0, H, delta, M, n_iter, seed, chain, minC=0, maxC=10, inv_var=None, jitter=0.2,
              p0=2.0 / 3.0, first_iteration=1, compat=True):
    """One chain; returns (draws (n_iter, d), diag (n_iter, 24), nevals)."""
    lib = load()
    q0 = np.ascontiguousarray(q0, dtype=np.float64)
    d = q0.size
    iv = None if inv_var is None else np.ascontiguousarray(inv_var, dtype=np.float64)
    draws = np.empty((n_iter, d))
    diag = np.empty((n_iter, 24))
    ne = C.c_uint64()
    rc = lib.wno_run_chain(TARGET[target], KIND[integrator], d, _dp(iv), _dp(q0), C.c_double(H), C.c_double(delta),
                           C.c_double(jitter), M, minC, maxC, C.c_double(p0), C.c_uint64(seed), C.c_uint32(chain),
                           C.c_uint32(first_iteration), n_iter, _dp(draws), _dp(diag), None, C.byref(ne), int(compat))
    assert rc == 0
    return draws, diag, int(ne.value)


def run_many(target, integrator, q, H, delta, M, n_iter, seed, threads, minC=0, maxC=10, inv_var=None, jitter=0.2,
             p0=2.0 / 3.0, chain0=0, first_iteration=1, compat=True):
    """Many chains on `threads` host threads; q (n_chains, d) is advanced in place.  Returns total grad evals."""
    lib = load()
    assert q.flags.c_contiguous and q.dtype == np.float64
    iv = None if inv_var is None else np.ascontiguousarray(inv_var, dtype=np.float64)
    ne = C.c_uint64()
    rc = lib.wno_run_many(TARGET[target], KIND[integrator], q.shape[1], _dp(iv), _dp(q), q.shape[0], C.c_double(H),
                          C.c_double(delta), C.c_double(jitter), M, minC, maxC, C.c_double(p0), C.c_uint64(seed),
                          C.c_uint32(chain0), C.c_uint32(first_iteration), n_iter, threads, C.byref(ne), int(compat))
    assert rc == 0
    return int(ne.value)
