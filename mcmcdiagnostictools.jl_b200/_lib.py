"""ctypes binding of libmcmcdiag_b200.so (the C ABI in include/mcmcdiag_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable,
the first call raises `MCDLibraryError`.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCMCDIAG_B200_LIB") or os.path.join(HERE, "libmcmcdiag_b200.so")   # override: A/B builds

MCD_OK, MCD_EINVAL, MCD_ECUDA, MCD_ENOMEM, MCD_EUNSUPPORTED, MCD_ENAN = 0, -1, -2, -3, -4, -5
MCD_F32, MCD_F64 = 0, 1
MCD_HOST, MCD_DEVICE = 0, 1
KINDS = {"basic": 0, "bulk": 1, "tail": 2, "rank": 3}
METHODS = {"direct": 0, "fft": 1, "bda": 2}
ESTIMATORS = {"mean": 0, "median": 1, "std": 2, "mad": 3, "quantile": 4}


class MCDLibraryError(RuntimeError):
    pass


_i64, _int, _dbl, _vp = C.c_int64, C.c_int, C.c_double, C.c_void_p
_SHAPE = [_vp, _vp, _int, _int, _i64, _i64, _i64]  # ctx, x, mem, dtype, draws, chains, params

# every symbol include/mcmcdiag_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mcd_abi_version": (_int, []),
    "mcd_create_error": (C.c_char_p, []),
    "mcd_create": (_int, [C.POINTER(_vp), _int]),
    "mcd_create_multi": (_int, [C.POINTER(_vp), C.POINTER(_int), _int]),
    "mcd_destroy": (None, [_vp]),
    "mcd_last_error": (C.c_char_p, [_vp]),
    "mcd_set_stream": (_int, [_vp, _vp, _int]),
    "mcd_synchronize": (_int, [_vp]),
    "mcd_set_param_mask": (_int, [_vp, C.POINTER(C.c_ubyte), _i64]),
    "mcd_set_option": (_int, [_vp, C.c_char_p, _i64]),
    "mcd_get_stat": (_i64, [_vp, C.c_char_p]),
    "mcd_ess_rhat": (_int, _SHAPE + [_int, _int, _int, _int, _int, _dbl, _int, _vp, _vp]),
    "mcd_ess_estimator": (_int, _SHAPE + [_int, _dbl, _int, _int, _int, _int, _int, _vp]),
    "mcd_mcse": (_int, _SHAPE + [_int, _dbl, _int, _int, _int, _int, _vp]),
    "mcd_summary": (_int, _SHAPE + [C.c_uint, _int, _int, _int, _dbl, _int, _vp]),
    "mcd_chain_moments": (_int, _SHAPE + [_int, _vp, _vp]),
    "mcd_bfmi": (_int, [_vp, _vp, _int, _int, _i64, _i64, _vp]),
    "mcd_rhat_nested": (_int, _SHAPE + [_vp, _i64, _i64, _int, _int, _vp]),
    "mcd_tiedrank": (_int, _SHAPE + [_vp]),
    "mcd_rank_normalize": (_int, _SHAPE + [_vp]),
    "mcd_fold_around_median": (_int, _SHAPE + [_vp]),
    "mcd_generate_ar1": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _dbl, _dbl, C.c_uint64, _vp]),
    "mcd_device_alloc": (_int, [_vp, _i64, C.POINTER(_vp)]),
    "mcd_device_free": (_int, [_vp, _vp]),
    "mcd_memcpy_h2d": (_int, [_vp, _vp, _vp, _i64]),
    "mcd_memcpy_d2h": (_int, [_vp, _vp, _vp, _i64]),
    "mcd_host_alloc": (_int, [_vp, _i64, C.POINTER(_vp)]),
    "mcd_host_free": (_int, [_vp, _vp]),
}

_lib = None


def load():
    """dlopen the library and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MCDLibraryError(
            f"{LIB_PATH} is not built (run `python -m __graft_entry__` or "
            "`python mcmcdiagnostictools.jl_b200/build.py`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
