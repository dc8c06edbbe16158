"""ctypes wrapper of the C++/OpenMP oracle port (oracle/ref_port.cpp).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from . import build_oracle

_lib = None
KINDS = {"basic": 0, "bulk": 1, "tail": 2, "rank": 3}
METHODS = {"direct": 0, "bda": 2}
ESTIMATORS = {"mean": 0, "median": 1, "std": 2, "mad": 3, "quantile": 4}


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.LIB if os.path.exists(build_oracle.LIB) else build_oracle.build()
        l = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        l.oracle_ess_rhat.restype = C.c_int
        l.oracle_ess_rhat.argtypes = [dp, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_double, dp, dp, C.c_int]
        l.oracle_ess_estimator.restype = C.c_int
        l.oracle_ess_estimator.argtypes = [dp, C.c_long, C.c_long, C.c_long, C.c_int, C.c_double, C.c_int, C.c_int,
                                           C.c_int, C.c_int, dp, C.c_int]
        l.oracle_tiedrank.argtypes = [dp, C.c_long, dp]
        l.oracle_norminvcdf.restype = C.c_double
        l.oracle_norminvcdf.argtypes = [C.c_double]
        l.oracle_max_threads.restype = C.c_int
        _lib = l
    return _lib


def _prep(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 3:
        raise ValueError("expects (draws, chains, params)")
    return np.asfortranarray(x)


def ess_rhat(x, kind="rank", method="direct", split_chains=2, maxlag=250, relative=False, tail_prob=0.1,
             nthreads=0, want_ess=True, want_rhat=True):
    xf = _prep(x)
    d, c, p = xf.shape
    e = np.empty(p) if want_ess else None
    r = np.empty(p) if want_rhat else None
    dp = C.POINTER(C.c_double)
    rc = lib().oracle_ess_rhat(xf.ctypes.data_as(dp), d, c, p, KINDS[kind], METHODS[method], split_chains, maxlag,
                               int(relative), tail_prob, e.ctypes.data_as(dp) if want_ess else None,
                               r.ctypes.data_as(dp) if want_rhat else None, nthreads)
    if rc not in (0,):
        raise ValueError(f"oracle_ess_rhat failed: {rc}")
    return e, r


def ess_estimator(x, estimator="mean", p=0.5, method="direct", split_chains=2, maxlag=250, relative=False, nthreads=0):
    xf = _prep(x)
    d, c, np_ = xf.shape
    e = np.empty(np_)
    dp = C.POINTER(C.c_double)
    rc = lib().oracle_ess_estimator(xf.ctypes.data_as(dp), d, c, np_, ESTIMATORS[estimator], p, METHODS[method],
                                    split_chains, maxlag, int(relative), e.ctypes.data_as(dp), nthreads)
    if rc != 0:
        raise ValueError(f"oracle_ess_estimator failed: {rc}")
    return e


def tiedrank(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty_like(v)
    dp = C.POINTER(C.c_double)
    lib().oracle_tiedrank(v.ctypes.data_as(dp), v.shape[0], out.ctypes.data_as(dp))
    return out


def norminvcdf(p):
    return lib().oracle_norminvcdf(float(p))


def max_threads():
    return lib().oracle_max_threads()
