"""Host-side mirror of the reference's public interface for the ESS / R-hat hot path.

Same names, keyword arguments, defaults and error behaviour as MCMCDiagnosticTools.jl
(`ess`, `rhat`, `ess_rhat`, `mcse`, `rhat_nested`, `AutocovMethod`, `FFTAutocovMethod`,
`BDAAutocovMethod`; /root/reference/src/MCMCDiagnosticTools.jl:19,23), written in Python
because no Julia toolchain exists in this image; the Julia shim with the identical logic is
`julia/MCMCDiagB200.jl`.  All numeric work happens in libmcmcdiag_b200.so (hand-written
sm_100a CUDA) through the C ABI in include/mcmcdiag_b200.h.  There is no CPU fallback.

What stays on the host: `kind` dispatch and the reference's exceptions
raised before the C call, the `niter <= 4` warning, missing-value masking, Int -> Float64
promotion, N-d parameter axes and scalar-vs-array outputs (`_maybescalar`), and the
superchain label -> index matrix of `_validate_superchain_ids`.

Inputs may be NumPy arrays (host memory, staged by the library in overlapped chunks),
`numpy.ma.MaskedArray` (mask = Julia `missing`), or CUDA `torch.Tensor`s (device memory,
zero copy when column-major).  Layout is Julia's: shape `(draws, [chains, [params...]])`.
"""
from __future__ import annotations

import ctypes as C
import threading
import warnings
from collections import namedtuple
from fractions import Fraction

import math

import numpy as np

from . import _lib as L

__all__ = [
    "ess", "rhat", "ess_rhat", "mcse", "rhat_nested",
    "AutocovMethod", "FFTAutocovMethod", "BDAAutocovMethod",
    "ESSMethod", "FFTESSMethod", "BDAESSMethod",
    "Quantile", "ArgumentError", "DomainError", "DimensionMismatch",
    "Context", "get_context", "tiedrank", "rank_normalize", "fold_around_median",
    "generate_ar1", "ESSRhat", "summary", "SUMMARY_FIELDS",
    "gewekediag", "heideldiag", "GewekeResult", "HeidelResult",
    "bfmi", "gelmandiag", "GelmanResult", "chain_moments",
]

ESSRhat = namedtuple("ESSRhat", ["ess", "rhat"])
GewekeResult = namedtuple("GewekeResult", ["zscore", "pvalue"])
GelmanResult = namedtuple("GelmanResult", ["psrf", "psrfci"])
HeidelResult = namedtuple("HeidelResult", ["burnin", "stationarity", "pvalue", "mean", "halfwidth", "test"])
# columns of `summary`, in the bit order of MCD_SUM_* (include/mcmcdiag_b200.h)
SUMMARY_FIELDS = ("mean", "std", "mcse_mean", "mcse_std", "ess_bulk", "ess_tail", "rhat")


class ArgumentError(ValueError):
    """Julia `ArgumentError`."""


class DomainError(ValueError):
    """Julia `DomainError`."""


class DimensionMismatch(ValueError):
    """Julia `DimensionMismatch`."""


class AbstractAutocovMethod:
    """src/ess_rhat.jl:2"""
    _code = None


class AutocovMethod(AbstractAutocovMethod):
    """Direct biased autocovariance (src/ess_rhat.jl:38, 161-179)."""
    _code = L.METHODS["direct"]


class FFTAutocovMethod(AbstractAutocovMethod):
    """FFT autocovariance, length nextprod([2,3], 2 niter - 1) (src/ess_rhat.jl:55, 130-152, 181-195)."""
    _code = L.METHODS["fft"]


class BDAAutocovMethod(AbstractAutocovMethod):
    """BDA variogram estimator (src/ess_rhat.jl:73, 197-213)."""
    _code = L.METHODS["bda"]


# north-star spellings (BASELINE.json) of the same three methods
ESSMethod, FFTESSMethod, BDAESSMethod = AutocovMethod, FFTAutocovMethod, BDAAutocovMethod


class Quantile:
    """`Base.Fix2(Statistics.quantile, p)` (src/ess_rhat.jl:647)."""

    def __init__(self, p):
        self.p = p

    def __repr__(self):
        return f"Quantile({self.p!r})"


# ---------------------------------------------------------------------------------------
# contexts
# ---------------------------------------------------------------------------------------
class Context:
    """Owns one `mcd_ctx`: one GPU (`Context(device)`), or a multi-GPU group (`Context(devices=[0, 1, ...])`,
    `mcd_create_multi`) that shards the parameter axis of HOST arrays over its devices inside the library."""

    def __init__(self, device: int = 0, devices=None):
        self._lib = L.load()
        h = C.c_void_p()
        if devices is not None:
            devs = [int(d) for d in devices]
            arr = (C.c_int * len(devs))(*devs)
            rc = self._lib.mcd_create_multi(C.byref(h), arr, len(devs))
            what = f"mcd_create_multi(devices={devs})"
            device = devs[0] if devs else 0
        else:
            rc = self._lib.mcd_create(C.byref(h), int(device))
            what = f"mcd_create(device={device})"
        if rc != L.MCD_OK:
            raise L.MCDLibraryError(f"{what} failed ({rc}): {self._lib.mcd_create_error().decode()}")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mcd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- error mapping -------------------------------------------------------------------
    def check(self, rc: int, domain: bool = False):
        if rc == L.MCD_OK:
            return
        msg = self._lib.mcd_last_error(self._h).decode()
        if rc == L.MCD_EINVAL:
            raise (DomainError if domain and "maxlag" in msg else ArgumentError)(msg)
        if rc == L.MCD_ENAN:
            raise ArgumentError(msg)
        if rc == L.MCD_ENOMEM:
            raise MemoryError(msg)
        if rc == L.MCD_EUNSUPPORTED:
            raise NotImplementedError(msg)
        raise L.MCDLibraryError(f"libmcmcdiag_b200 error {rc}: {msg}")

    def set_option(self, key: str, value: int):
        self.check(self._lib.mcd_set_option(self._h, key.encode(), int(value)))

    def stat(self, key: str) -> int:
        return int(self._lib.mcd_get_stat(self._h, key.encode()))

    def set_stream(self, cuda_stream: int | None):
        """cuda_stream: a cudaStream_t handle (0 = CUDA's default stream); None = the context's own stream."""
        own = cuda_stream is None
        self.check(self._lib.mcd_set_stream(self._h, C.c_void_p(0 if own else cuda_stream), int(own)))

    def synchronize(self):
        self.check(self._lib.mcd_synchronize(self._h))


def _default_host_device() -> int:
    """Device for host (NumPy) input when no context is passed: the process's current CUDA device if torch is loaded
    (torchrun workers set it from LOCAL_RANK), else LOCAL_RANK, else 0."""
    import os
    import sys
    t = sys.modules.get("torch")
    if t is not None:
        try:
            if t.cuda.is_available():
                return int(t.cuda.current_device())
        except Exception:
            pass
    try:
        return int(os.environ.get("LOCAL_RANK", "0"))
    except ValueError:
        return 0


_contexts: dict[int, Context] = {}
_ctx_lock = threading.Lock()


def get_context(device: int = 0) -> Context:
    with _ctx_lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = _contexts[device] = Context(device)
        return ctx


# ---------------------------------------------------------------------------------------
# array plumbing
# ---------------------------------------------------------------------------------------
class _Arr:
    """A (draws, chains, P) column-major view of the caller's samples + how to build outputs."""

    def __init__(self, samples, min_ndim=1):
        self.is_torch = type(samples).__module__.startswith("torch")
        self.missing = None   # boolean per parameter (True = contains missing)
        if self.is_torch:
            self._init_torch(samples)
        else:
            self._init_numpy(samples)

    # numpy / masked arrays: host memory
    def _init_numpy(self, samples):
        mask = None
        if isinstance(samples, np.ma.MaskedArray):
            mask = np.ma.getmaskarray(samples)
            samples = samples.filled(0)
        x = np.asarray(samples)
        if x.ndim == 0:
            raise ArgumentError("samples must have at least one dimension")
        if x.dtype == np.float32:
            dt = np.float32
        elif x.dtype.kind in "fiub":
            dt = np.float64           # promote_type(eltype, typeof(zero(eltype)/1))
        else:
            raise ArgumentError(f"unsupported eltype {x.dtype}")
        self.dtype = np.dtype(dt)
        self.pshape = tuple(x.shape[2:])
        shape3 = (x.shape[0], x.shape[1] if x.ndim > 1 else 1, int(np.prod(self.pshape, dtype=np.int64)))
        self.draws, self.chains, self.P = shape3
        self.scalar = x.ndim < 3
        x3 = np.asfortranarray(x.astype(dt, copy=False)).reshape(shape3, order="F")
        if mask is not None and mask.any():
            # a parameter that contains `missing` yields `missing` (src/ess_rhat.jl:382-385,519-523): the library gets the
            # array as it is plus a per-parameter skip mask (mcd_set_param_mask); no compacted copy is made
            m3 = np.asfortranarray(mask).reshape(shape3, order="F")
            self.missing = m3.any(axis=(0, 1))
        self.x3 = x3
        self.mem = L.MCD_HOST
        self.ptr = x3.ctypes.data if x3.size else 0
        self.nparams = x3.shape[2]
        self.device = _default_host_device()

    # torch CUDA tensors: device memory
    def _init_torch(self, samples):
        import torch
        x = samples
        if not x.is_cuda:
            raise ArgumentError("torch inputs must be CUDA tensors; pass NumPy arrays for host data")
        if x.ndim == 0:
            raise ArgumentError("samples must have at least one dimension")
        if x.dtype == torch.float32:
            dt = np.float32
        elif x.dtype == torch.float64:
            dt = np.float64
        else:
            x = x.to(torch.float64)
            dt = np.float64
        self.dtype = np.dtype(dt)
        self.pshape = tuple(x.shape[2:])
        self.scalar = x.ndim < 3
        self.draws = x.shape[0]
        self.chains = x.shape[1] if x.ndim > 1 else 1
        self.P = int(np.prod(self.pshape, dtype=np.int64))
        # column-major <=> the dims-reversed view is C-contiguous
        xt = x.permute(*reversed(range(x.ndim)))
        if not xt.is_contiguous():
            xt = xt.contiguous()
        self._keep = xt
        self.torch_dtype = xt.dtype
        self.torch_device = xt.device
        self.mem = L.MCD_DEVICE
        self.ptr = xt.data_ptr()
        self.nparams = self.P
        self.device = xt.device.index or 0

    @property
    def code(self):
        return L.MCD_F64 if self.dtype == np.float64 else L.MCD_F32

    def new_out(self):
        if self.is_torch:
            import torch
            return torch.empty(self.nparams, dtype=self.torch_dtype, device=self.torch_device)
        return np.empty(self.nparams, dtype=self.dtype)

    @staticmethod
    def out_ptr(o):
        if o is None:
            return None
        if isinstance(o, np.ndarray):
            return C.c_void_p(o.ctypes.data)
        return C.c_void_p(o.data_ptr())

    def finish(self, o):
        """Scatter missing parameters back, restore parameter axes, `_maybescalar`."""
        if o is None:
            return None
        if self.is_torch:
            if self.scalar:
                return o.reshape(())
            return o.reshape(tuple(reversed(self.pshape))).permute(*reversed(range(len(self.pshape))))
        if self.missing is not None:
            o = np.ma.array(o, mask=self.missing)
        if self.scalar:
            v = o.reshape(())[()]
            return v if v is np.ma.masked else self.dtype.type(v)
        return o.reshape(self.pshape, order="F")

    def context(self, ctx):
        ctx = ctx or get_context(self.device)
        if self.is_torch:
            import torch
            ctx.set_stream(torch.cuda.current_stream(self.torch_device).cuda_stream)
        else:
            ctx.set_stream(None)
        if self.missing is not None:
            skip = np.ascontiguousarray(self.missing, dtype=np.uint8)
            ctx.check(ctx._lib.mcd_set_param_mask(ctx._h, skip.ctypes.data_as(C.POINTER(C.c_ubyte)), skip.size))
        return ctx


_SYMBOLS = ("rank", "bulk", "tail", "basic")
# north-star symbol spellings -> reference estimators (SURVEY.md vocabulary map)
_ESTIMATOR_ALIASES = {"mean": "mean", "median": "median", "std": "std", "squared": "std",
                      "mad": "mad", "abs": "mad", "folded": "mad"}


def _estimator(kind):
    """Map an estimator `kind` to (code, p, p_is_f64) or None (src/ess_rhat.jl:628-659)."""
    if isinstance(kind, Quantile):
        p = kind.p
        if isinstance(p, Fraction):
            return "quantile", float(p), False
        return "quantile", float(p), not isinstance(p, np.float32)
    if isinstance(kind, str):
        name = _ESTIMATOR_ALIASES.get(kind)
        return (name, 0.0, False) if name else None
    if kind is np.mean:
        return "mean", 0.0, False
    if kind is np.median:
        return "median", 0.0, False
    if kind is np.std:
        return "std", 0.0, False
    name = getattr(kind, "__name__", "")
    if name == "median_abs_deviation":      # scipy.stats.median_abs_deviation ~ StatsBase.mad
        return "mad", 0.0, False
    tag = getattr(kind, "_mcd_estimator", None)
    return (tag, 0.0, False) if tag in L.ESTIMATORS else None


def _method_code(m):
    if isinstance(m, type) and issubclass(m, AbstractAutocovMethod):
        m = m()
    if not isinstance(m, AbstractAutocovMethod) or m._code is None:
        raise ArgumentError(f"autocov_method must be an AbstractAutocovMethod, got {m!r}")
    return m._code


def _tail_prob(tp):
    if isinstance(tp, (Fraction, int)):
        return float(Fraction(tp)), 0
    return float(tp), (0 if isinstance(tp, np.float32) else 1)


def _warn_niter(a: _Arr, split_chains: int):
    niter = a.draws // split_chains
    if not niter > 4:
        warnings.warn(f"number of draws after splitting must be >4 but is {niter}. ESS cannot be computed.")


def _check_split(split_chains):
    if not isinstance(split_chains, (int, np.integer)) or split_chains < 1:
        raise ArgumentError("split_chains must be a positive integer")


def _call_ess_rhat(a, kind, want_ess, want_rhat, relative=False, autocov_method=None, split_chains=2,
                   maxlag=250, tail_prob=Fraction(1, 10), ctx=None):
    _check_split(split_chains)
    ctx = a.context(ctx)
    method = _method_code(autocov_method if autocov_method is not None else AutocovMethod())
    if want_ess:
        _warn_niter(a, split_chains)
        if a.draws // split_chains > 4 and not maxlag > 0:
            raise DomainError(f"maxlag must be >0. (got {maxlag})")
    tp, tp64 = _tail_prob(tail_prob)
    e = a.new_out() if want_ess else None
    r = a.new_out() if want_rhat else None
    rc = ctx._lib.mcd_ess_rhat(ctx._h, C.c_void_p(a.ptr), a.mem, a.code, a.draws, a.chains, a.nparams,
                               L.KINDS[kind], method, int(split_chains), int(max(min(maxlag, 2**31 - 1), -1)),
                               int(bool(relative)), tp, tp64, a.out_ptr(e), a.out_ptr(r))
    ctx.check(rc, domain=True)
    return a.finish(e), a.finish(r)


def _clamp_maxlag(maxlag) -> int:
    """maxlag as a C int: ctypes would otherwise truncate modulo 2^32 (the library clamps to niter - 4 anyway)."""
    return int(max(min(int(maxlag), 2**31 - 1), -1))


def _call_estimator(a, est, relative=False, autocov_method=None, split_chains=2, maxlag=250, ctx=None,
                    mcse_mode=False):
    _check_split(split_chains)
    name, p, p64 = est
    ctx = a.context(ctx)
    method = _method_code(autocov_method if autocov_method is not None else AutocovMethod())
    _warn_niter(a, split_chains)
    if a.draws // split_chains > 4 and not maxlag > 0:
        raise DomainError(f"maxlag must be >0. (got {maxlag})")
    out = a.new_out()
    if mcse_mode:
        rc = ctx._lib.mcd_mcse(ctx._h, C.c_void_p(a.ptr), a.mem, a.code, a.draws, a.chains, a.nparams,
                               L.ESTIMATORS[name], p, int(p64), method, int(split_chains), _clamp_maxlag(maxlag),
                               a.out_ptr(out))
    else:
        rc = ctx._lib.mcd_ess_estimator(ctx._h, C.c_void_p(a.ptr), a.mem, a.code, a.draws, a.chains, a.nparams,
                                        L.ESTIMATORS[name], p, int(p64), method, int(split_chains), _clamp_maxlag(maxlag),
                                        int(bool(relative)), a.out_ptr(out))
    ctx.check(rc, domain=True)
    return a.finish(out)


# ---------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------
def ess(samples, *, kind="bulk", ctx=None, **kwargs):
    """`ess(samples; kind=:bulk, relative=false, autocov_method=AutocovMethod(), split_chains=2,
    maxlag=250, [tail_prob=1//10])`  (src/ess_rhat.jl:215-311)."""
    a = _Arr(samples)
    if isinstance(kind, str) and kind in ("bulk", "tail", "basic"):
        if kind != "tail" and "tail_prob" in kwargs:
            raise TypeError("ess() got an unexpected keyword argument 'tail_prob' for this kind")
        return _call_ess_rhat(a, kind, True, False, ctx=ctx, **kwargs)[0]
    if isinstance(kind, str) and kind == "rank":
        raise ArgumentError(f"the `kind` `{kind}` is not supported by `ess`")
    est = _estimator(kind)
    if est is None:
        if isinstance(kind, str):
            raise ArgumentError(f"the `kind` `{kind}` is not supported by `ess`")
        raise ArgumentError(f"the estimator {kind} is not yet supported by `ess`")
    return _call_estimator(a, est, ctx=ctx, **kwargs)


def rhat(samples, *, kind="rank", split_chains=2, ctx=None):
    """`rhat(samples; kind=:rank, split_chains=2)`  (src/ess_rhat.jl:313-420)."""
    if kind not in _SYMBOLS:
        raise ArgumentError(f"the `kind` `{kind}` is not supported by `rhat`")
    a = _Arr(samples)
    return _call_ess_rhat(a, kind, False, True, split_chains=split_chains, ctx=ctx)[1]


def ess_rhat(samples, *, kind="rank", ctx=None, **kwargs):
    """`ess_rhat(samples; kind=:rank, kwargs...) -> (; ess, rhat)`  (src/ess_rhat.jl:422-455)."""
    if kind not in _SYMBOLS:
        raise ArgumentError(f"the `kind` `{kind}` is not supported by `ess_rhat`")
    a = _Arr(samples)
    if kind != "tail" and "tail_prob" in kwargs:
        raise TypeError("ess_rhat() got an unexpected keyword argument 'tail_prob' for this kind")
    e, r = _call_ess_rhat(a, kind, True, True, ctx=ctx, **kwargs)
    return ESSRhat(e, r)


def _mad(v, axis=None):
    """StatsBase.mad(x) (normalize = true): 1.4826... * median(|x - median(x)|)."""
    med = np.median(v, axis=axis, keepdims=axis is not None)
    return 1.4826022185056018 * np.median(np.abs(v - med), axis=axis)


def _mcse_sbm(f, samples, batch_size=None):
    """`_mcse_sbm(f, x; batch_size)` (src/mcse.jl:120-148): the subsampling-bootstrap fallback for estimators without
    an ESS rule.  As in the reference it is host logic: `f` is an arbitrary host callable evaluated on every
    overlapping batch of the flattened draws, so nothing of it can cross the C ABI."""
    masked = isinstance(samples, np.ma.MaskedArray)
    mask = np.ma.getmaskarray(samples) if masked else None
    x = np.asarray(samples.filled(0) if masked else samples)
    if type(samples).__module__.startswith("torch"):
        x = samples.detach().cpu().numpy()
    T = np.float32 if x.dtype == np.float32 else np.float64
    draws = x.shape[0]
    chains = x.shape[1] if x.ndim > 1 else 1
    pshape = tuple(x.shape[2:])
    x3 = np.asfortranarray(x.astype(T, copy=False)).reshape((draws, chains, -1), order="F")
    m3 = None if mask is None else np.asfortranarray(mask).reshape((draws, chains, -1), order="F")
    n = draws * chains
    b = int(math.floor(math.sqrt(n))) if batch_size is None else int(batch_size)
    out = np.empty(x3.shape[2], dtype=T)
    miss = np.zeros(x3.shape[2], dtype=bool)
    for p in range(x3.shape[2]):
        v = x3[:, :, p].reshape(-1, order="F")
        if m3 is not None and m3[:, :, p].any():
            miss[p] = True
            out[p] = np.nan
            continue
        if np.all(v == v[0]):
            out[p] = np.nan
            continue
        win = np.lib.stride_tricks.sliding_window_view(v, b)
        try:
            vals = np.asarray(f(win, axis=1), dtype=np.float64)
            if vals.shape != (win.shape[0],):
                raise TypeError
        except TypeError:
            vals = np.array([f(w) for w in win], dtype=np.float64)
        out[p] = T(math.sqrt(vals.var() * (b / n)))
    if x.ndim < 3:
        return np.ma.masked if miss[0] else T(out[0])
    res = np.ma.array(out, mask=miss) if miss.any() else out
    return res.reshape(pshape, order="F")


def mcse(samples, *, kind=np.mean, ctx=None, **kwargs):
    """`mcse(samples; kind=Statistics.mean, kwargs...)`  (src/mcse.jl:5-42).

    mean / std / median / quantile use the ESS-based rules on the GPU (src/mcse.jl:45-118).  Any other estimator
    (`mad`, an arbitrary callable) uses the reference's subsampling-bootstrap fallback `_mcse_sbm`
    (src/mcse.jl:120-148; keyword `batch_size`), which evaluates a host callable per batch and therefore runs on the host
    here exactly as it does in the reference."""
    est = _estimator(kind)
    if est is None or est[0] == "mad":
        f = _mad if (est is not None or kind in ("mad", "abs", "folded")) else kind
        if not callable(f):
            raise ArgumentError(f"the estimator {kind!r} is not supported by `mcse`")
        extra = set(kwargs) - {"batch_size"}
        if extra:
            raise TypeError(f"mcse() got unexpected keyword arguments {sorted(extra)} for the subsampling-bootstrap fallback")
        return _mcse_sbm(f, samples, kwargs.get("batch_size"))
    if kwargs.pop("relative", False):
        # the reference forwards `relative` to `_ess`, i.e. it would plug the RELATIVE effective sample size into the
        # standard-error rules; the C ABI's mcd_mcse has no such mode, so it is rejected rather than ignored
        raise NotImplementedError("mcse(...; relative=true) is not supported by the accelerated path")
    a = _Arr(samples)
    return _call_estimator(a, est, ctx=ctx, mcse_mode=True, **kwargs)


def summary(samples, *, fields=None, autocov_method=None, split_chains=2, maxlag=250,
            tail_prob=Fraction(1, 10), ctx=None):
    """Fused per-parameter summary: a dict of the columns `SUMMARY_FIELDS`,

        mean      = Statistics.mean(x; dims=(1,2))        std       = Statistics.std(x; dims=(1,2))
        mcse_mean = mcse(x; kind=mean, kw...)             mcse_std  = mcse(x; kind=std, kw...)
        ess_bulk  = ess(x; kind=:bulk, kw...)             ess_tail  = ess(x; kind=:tail, tail_prob, kw...)
        rhat      = rhat(x; kind=:rank, split_chains)

    each identical to the separate reference call (src/mcse.jl:45-69, src/ess_rhat.jl:298-311,
    410-420, 604-624), computed by one library call that stages a host array once."""
    names = SUMMARY_FIELDS if fields is None else tuple(fields)
    for f in names:
        if f not in SUMMARY_FIELDS:
            raise ArgumentError(f"unknown summary field `{f}`; expected a subset of {SUMMARY_FIELDS}")
    if not names:
        raise ArgumentError("no summary field requested")
    names = tuple(f for f in SUMMARY_FIELDS if f in names)   # library column order
    mask = sum(1 << SUMMARY_FIELDS.index(f) for f in names)
    _check_split(split_chains)
    a = _Arr(samples)
    ctx = a.context(ctx)
    method = _method_code(autocov_method if autocov_method is not None else AutocovMethod())
    if mask & 0b0111100:
        _warn_niter(a, split_chains)
        if a.draws // split_chains > 4 and not maxlag > 0:
            raise DomainError(f"maxlag must be >0. (got {maxlag})")
    tp, tp64 = _tail_prob(tail_prob)
    if a.is_torch:
        import torch
        out = torch.empty((len(names), a.nparams), dtype=a.torch_dtype, device=a.torch_device)
    else:
        out = np.empty((len(names), a.nparams), dtype=a.dtype)
    rc = ctx._lib.mcd_summary(ctx._h, C.c_void_p(a.ptr), a.mem, a.code, a.draws, a.chains, a.nparams,
                              mask, method, int(split_chains), int(max(min(maxlag, 2**31 - 1), -1)),
                              tp, tp64, a.out_ptr(out))
    ctx.check(rc, domain=True)
    return {f: a.finish(out[i]) for i, f in enumerate(names)}


def _validate_superchain_ids(superchain_ids, nchains):
    """`_validate_superchain_ids` + `unique_indices` (src/rhat_nested.jl:68-81, src/utils.jl:50-64):
    superchains ordered by sorted label, chains within a superchain by first appearance.
    Returns a 0-based int32 matrix (chains_per_super x nsuper), column-major."""
    ids = list(superchain_ids)
    if len(ids) != nchains:
        raise DimensionMismatch(
            f"`superchain_ids` has length {len(ids)} but `samples` has {nchains} chains")
    groups: dict = {}
    for i, s in enumerate(ids):
        groups.setdefault(s, []).append(i)
    keys = sorted(groups)
    if len(keys) < 2:
        raise ArgumentError(f"at least 2 superchains are required, got {len(keys)}")
    if len({len(groups[k]) for k in keys}) != 1:
        raise ArgumentError("all superchains must contain the same number of chains")
    return np.asfortranarray(np.stack([np.asarray(groups[k], dtype=np.int32) for k in keys], axis=1))


def rhat_nested(samples, superchain_ids, *, kind="rank", split_chains=2, ctx=None):
    """`rhat_nested(samples, superchain_ids; kind=:rank, split_chains=2)`  (src/rhat_nested.jl:1-66)."""
    ndim = samples.ndim if hasattr(samples, "ndim") else np.asarray(samples).ndim
    if ndim < 2:
        raise ArgumentError("`samples` must have at least 2 dimensions `(draws, chains[, parameters…])`")
    a = _Arr(samples)
    inds = _validate_superchain_ids(superchain_ids, a.chains)
    if kind not in _SYMBOLS:
        raise ArgumentError(f"the `kind` `{kind}` is not supported by `rhat_nested`")
    _check_split(split_chains)
    ctx = a.context(ctx)
    out = a.new_out()
    rc = ctx._lib.mcd_rhat_nested(ctx._h, C.c_void_p(a.ptr), a.mem, a.code, a.draws, a.chains, a.nparams,
                                  C.c_void_p(inds.ctypes.data), inds.shape[0], inds.shape[1], L.KINDS[kind],
                                  int(split_chains), a.out_ptr(out))
    ctx.check(rc)
    return a.finish(out)


# ---------------------------------------------------------------------------------------
# transforms (exposed for parity checks) and the synthetic generator
# ---------------------------------------------------------------------------------------
def _transform(samples, fn_name, out_dtype, ctx=None):
    a = _Arr(samples)
    if a.missing is not None:
        raise ArgumentError("masked input is not supported by the transform helpers")
    ctx = a.context(ctx)
    shape3 = (a.draws, a.chains, a.nparams)
    if a.is_torch:
        import torch
        tdt = torch.float64 if out_dtype == np.float64 else torch.float32
        out = torch.empty((a.nparams, a.chains, a.draws), dtype=tdt, device=a.torch_device)
        ptr = C.c_void_p(out.data_ptr())
    else:
        out = np.empty(shape3, dtype=out_dtype, order="F")
        ptr = C.c_void_p(out.ctypes.data)
    rc = getattr(ctx._lib, fn_name)(ctx._h, C.c_void_p(a.ptr), a.mem, a.code, a.draws, a.chains, a.nparams, ptr)
    ctx.check(rc)
    if a.is_torch:
        # the library wrote (params, chains, draws) C-contiguous with the parameter axis flattened in column-major order
        # (as `_init_torch` / `finish` flatten it): view it as (*reversed(pshape), chains, draws) and reverse all axes
        nd = np.ndim(samples) if not hasattr(samples, "ndim") else samples.ndim
        full = out.reshape(tuple(reversed(a.pshape)) + (a.chains, a.draws))
        full = full.permute(*reversed(range(full.ndim)))          # (draws, chains, *pshape)
        return full.reshape(a.draws) if nd == 1 else full
    full_shape = (a.draws,) + ((a.chains,) if np.ndim(samples) > 1 else ()) + a.pshape
    return out.reshape(full_shape, order="F")


def tiedrank(samples, ctx=None):
    """StatsBase.tiedrank of every parameter's flattened draws x chains slab, as Float64."""
    return _transform(samples, "mcd_tiedrank", np.float64, ctx)


def rank_normalize(samples, ctx=None):
    """`_rank_normalize` (src/utils.jl:169-193)."""
    a_dt = np.float32 if getattr(samples, "dtype", None) in (np.float32,) or str(getattr(samples, "dtype", "")) == "torch.float32" else np.float64
    return _transform(samples, "mcd_rank_normalize", a_dt, ctx)


def fold_around_median(samples, ctx=None):
    """`_fold_around_median` (src/utils.jl:148-158)."""
    a_dt = np.float32 if getattr(samples, "dtype", None) in (np.float32,) or str(getattr(samples, "dtype", "")) == "torch.float32" else np.float64
    return _transform(samples, "mcd_fold_around_median", a_dt, ctx)


def generate_ar1(phi, sigma, draws, chains, params, *, dtype="float64", seed=1, param_offset=0, device=0, ctx=None):
    """AR(1) chains as test/helpers.jl:4-12, generated on the GPU; returns a CUDA torch tensor of
    logical shape (draws, chains, params) in column-major layout."""
    import torch
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    dev = torch.device("cuda", device)
    buf = torch.empty((params, chains, draws), dtype=tdt, device=dev)
    ctx = ctx or get_context(device)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    rc = ctx._lib.mcd_generate_ar1(ctx._h, L.MCD_F64 if tdt == torch.float64 else L.MCD_F32, draws, chains,
                                   params, int(param_offset), float(phi), float(sigma), int(seed),
                                   C.c_void_p(buf.data_ptr()))
    ctx.check(rc)
    return buf.permute(2, 1, 0)


# ---------------------------------------------------------------------------------------
# in-package callers of the path: thin host wrappers over the device mcse
# ---------------------------------------------------------------------------------------
def _series(x):
    """(draws,) or (draws, params) -> host float matrix (draws, P), the input itself viewed as
    (draws, 1, P) for the device calls, and whether the input was a single vector."""
    is_torch = type(x).__module__.split(".")[0] == "torch"
    nd = x.ndim if hasattr(x, "ndim") else np.asarray(x).ndim
    if nd not in (1, 2):
        raise ArgumentError("expected a vector of draws or a (draws, parameters) matrix")
    if is_torch:
        x3 = x.reshape(x.shape[0], 1, -1)
        host = x3[:, 0, :].detach().cpu().numpy()
    else:
        x = np.asarray(x)
        x3 = x.reshape(x.shape[0], 1, -1)
        host = x3[:, 0, :]
    T = np.float32 if host.dtype == np.float32 else np.float64
    return host.astype(T, copy=False), x3, nd == 1, T


def _window_mean_mcse(x3, lo, hi, ctx, kwargs):
    """mean and mcse(...; split_chains=1, kwargs...) of draws [lo, hi) of every series, on the device."""
    w = x3[lo:hi]
    if "kind" in kwargs:
        se = mcse(w, split_chains=1, ctx=ctx, **kwargs)
        mu = summary(w, fields=("mean",), split_chains=1, ctx=ctx)["mean"]
    else:
        r = summary(w, fields=("mean", "mcse_mean"), split_chains=1, ctx=ctx, **kwargs)
        mu, se = r["mean"], r["mcse_mean"]
    to_np = lambda v: v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    return to_np(mu).reshape(-1), to_np(se).reshape(-1)


def gewekediag(x, *, first=0.1, last=0.5, ctx=None, **kwargs):
    """`gewekediag(x::AbstractVector; first=0.1, last=0.5, kwargs...)` (src/gewekediag.jl:19-35):
    z-score and p-value comparing the means of the first and last windows, their standard errors
    from `mcse(...; split_chains=1, kwargs...)` on the device.  Extension: a `(draws, params)`
    matrix runs every column in the same two device calls and returns arrays."""
    from scipy import special
    if not 0 < first < 1:
        raise ArgumentError("`first` is not in (0, 1)")
    if not 0 < last < 1:
        raise ArgumentError("`last` is not in (0, 1)")
    if not first + last <= 1:
        raise ArgumentError("`first` and `last` proportions overlap")
    host, x3, vector, T = _series(x)
    n = host.shape[0]
    hi1 = int(np.round(first * n))                      # x[1:round(Int, first * n)]
    lo2 = int(np.round(n - last * n + 1)) - 1           # x[round(Int, n - last * n + 1):n]
    m1, s1 = _window_mean_mcse(x3, 0, hi1, ctx, kwargs)
    m2, s2 = _window_mean_mcse(x3, lo2, n, ctx, kwargs)
    with np.errstate(all="ignore"):
        z = ((m1.astype(T) - m2.astype(T)) / np.hypot(s1.astype(T), s2.astype(T))).astype(T)
        p = special.erfc(np.abs(z.astype(np.float64)) / np.sqrt(2.0)).astype(T)
    if vector:
        return GewekeResult(T(z[0]), T(p[0]))
    return GewekeResult(z, p)


def _pcramer(q):
    """Csorgo & Faraway (1996) series for the Cramer-von Mises distribution (src/heideldiag.jl:60-71)."""
    from scipy import special
    q = np.asarray(q, dtype=np.float64)
    p = np.zeros_like(q)
    with np.errstate(all="ignore"):
        for k in range(4):
            c1 = 4.0 * k + 1.0
            c2 = c1 * c1 / (16.0 * q)
            p += special.gamma(k + 0.5) / math.factorial(k) * math.sqrt(c1) * np.exp(-c2) * special.kv(0.25, c2)
        return p / (math.pi ** 1.5 * np.sqrt(q))


def heideldiag(x, *, alpha=Fraction(1, 20), eps=0.1, start=1, ctx=None, **kwargs):
    """`heideldiag(x::AbstractVector; alpha=1//20, eps=0.1, start=1, kwargs...)` (src/heideldiag.jl:16-54):
    Heidelberger-Welch stationarity (Cramer-von Mises on the Brownian-bridge statistic, discarding
    10 % of the draws at a time) and half-width tests; both spectral-density-at-zero estimates are
    `mcse(...; split_chains=1, kwargs...)` on the device.  Extension: a `(draws, params)` matrix is
    processed column-wise with the device calls batched over the columns that share a burn-in."""
    from scipy import special
    host, x3, vector, T = _series(x)
    n, P = host.shape
    delta = int(0.10 * n)
    lo = int(n / 2) - 1                                   # y = x[trunc(Int, n / 2):end]
    _, s = _window_mean_mcse(x3, lo, n, ctx, kwargs)
    S0 = T(n - lo) * s.astype(T) * s.astype(T)
    burn = np.ones(P, dtype=np.int64)                     # i of the reference loop, per series
    y_start = np.full(P, max(int(n / 2), 1), dtype=np.int64)   # 1-based start of the last y a series tested
    pvalue = np.ones(P, dtype=T)
    converged = np.zeros(P, dtype=bool)
    ybar = np.full(P, np.nan, dtype=T)
    active = np.ones(P, dtype=bool)
    i = 1
    while i < n / 2 and active.any():
        idx = np.flatnonzero(active)
        y = host[i - 1:, idx]
        m = y.shape[0]
        yb = y.mean(axis=0, dtype=T).astype(T)
        with np.errstate(all="ignore"):
            B = np.cumsum(y, axis=0, dtype=T) - yb[None, :] * np.arange(1, m + 1, dtype=T)[:, None]
            I = ((B * B) / (T(m) * S0[idx])[None, :]).sum(axis=0, dtype=T) / T(m)
            pv = (T(1) - _pcramer(I).astype(T)).astype(T)
        ok = pv > float(alpha)
        burn[idx], y_start[idx], pvalue[idx], ybar[idx], converged[idx] = i, i, pv, yb, ok
        active[idx[ok]] = False
        if delta == 0:                                    # n < 10: the reference loop would never advance
            break
        i += delta
        burn[active] = i                                  # the reference adds delta before its loop test fails
    halfwidth = np.empty(P, dtype=T)
    for ys in np.unique(y_start):                         # s = mcse(y) on the last y each series tested
        cols = np.flatnonzero(y_start == ys)
        _, se = _window_mean_mcse(x3[:, :, cols], int(ys) - 1, n, ctx, kwargs)
        halfwidth[cols] = (T(math.sqrt(2.0)) * T(special.erfcinv(float(T(float(alpha))))) * se.astype(T)).astype(T)
    with np.errstate(all="ignore"):
        passed = halfwidth / np.abs(ybar) <= eps
    out = HeidelResult(burn + start - 2, converged, pvalue, ybar, halfwidth, passed)
    if vector:
        return HeidelResult(int(out.burnin[0]), bool(out.stationarity[0]), T(out.pvalue[0]), T(out.mean[0]),
                            T(out.halfwidth[0]), bool(out.test[0]))
    return out


# ---------------------------------------------------------------------------------------
# the same moment kernels behind a different combine
# ---------------------------------------------------------------------------------------
def chain_moments(samples, *, split_chains=1, ctx=None):
    """Mean and corrected variance of every (split) chain of every parameter, as two arrays of shape
    `(chains * split_chains, params)`: what `_rhat_basic!` (src/ess_rhat.jl:387-399) and `_gelmandiag`
    (src/gelmandiag.jl:9-17) start from."""
    _check_split(split_chains)
    a = _Arr(samples, min_ndim=2)
    if a.missing is not None and a.missing.any():
        raise ArgumentError("chain_moments does not take missing values")
    ctx = a.context(ctx)
    nch = a.chains * int(split_chains)
    if a.is_torch:
        import torch
        mk = lambda: torch.empty((a.nparams, nch), dtype=a.torch_dtype, device=a.torch_device)
    else:
        mk = lambda: np.empty((a.nparams, nch), dtype=a.dtype)
    mean, var = mk(), mk()
    rc = ctx._lib.mcd_chain_moments(ctx._h, C.c_void_p(a.ptr), a.mem, a.code, a.draws, a.chains, a.nparams,
                                    int(split_chains), a.out_ptr(mean), a.out_ptr(var))
    ctx.check(rc)
    return mean.T, var.T          # (nch, params): the library writes column-major


def bfmi(energy, *, dims=1, ctx=None):
    """`bfmi(energy::AbstractVector)` / `bfmi(energy::AbstractMatrix; dims=1)` (src/bfmi.jl:36-43):
    `mean(abs2, diff(energy)) / var(energy)` per chain.  `dims` is the (1-based, as in the reference)
    dimension holding the draws."""
    is_torch = type(energy).__module__.split(".")[0] == "torch"
    nd = energy.ndim if hasattr(energy, "ndim") else np.asarray(energy).ndim
    if nd == 1:
        e2 = energy.reshape(-1, 1) if is_torch else np.asarray(energy).reshape(-1, 1)
    elif nd == 2:
        if dims not in (1, 2):
            raise ArgumentError("dims must be 1 or 2")
        e2 = energy if dims == 1 else (energy.T if not is_torch else energy.t())
        if not is_torch:
            e2 = np.asarray(e2)
    else:
        raise ArgumentError("energy must be a vector or a matrix")
    a = _Arr(e2.reshape(e2.shape[0], e2.shape[1], 1), min_ndim=3)   # (draws, chains, 1): chains are contiguous columns
    ctx = a.context(ctx)
    if a.is_torch:
        import torch
        out = torch.empty(a.chains, dtype=a.torch_dtype, device=a.torch_device)
    else:
        out = np.empty(a.chains, dtype=a.dtype)
    rc = ctx._lib.mcd_bfmi(ctx._h, C.c_void_p(a.ptr), a.mem, a.code, a.draws, a.chains, a.out_ptr(out))
    ctx.check(rc)
    if nd == 1:
        return out[0] if not a.is_torch else out.reshape(())
    return out


def gelmandiag(samples, *, alpha=0.05, ctx=None):
    """`gelmandiag(samples::AbstractArray{<:Real,3}; alpha=0.05) -> (psrf, psrfci)` (src/gelmandiag.jl:1-76):
    Gelman-Rubin-Brooks potential scale reduction factors and their upper confidence limits.  Only the
    diagonals of the within / between covariance matrices are needed, i.e. the per-chain means and
    variances, which come from the device (`mcd_chain_moments`); the combine is O(chains * params)."""
    from scipy import stats
    nd = samples.ndim if hasattr(samples, "ndim") else np.asarray(samples).ndim
    if nd != 3:
        raise ArgumentError("`samples` must have shape (draws, chains, parameters)")
    niters, nchains = int(samples.shape[0]), int(samples.shape[1])
    if not nchains > 1:
        raise RuntimeError("Gelman diagnostic requires at least 2 chains")
    mean, var = chain_moments(samples, split_chains=1, ctx=ctx)
    to_np = lambda v: v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    psibar = to_np(mean).astype(np.float64)       # (chains, params)
    s2 = to_np(var).astype(np.float64)
    rfixed = (niters - 1) / niters
    rrandomscale = (nchains + 1) / (nchains * niters)
    with np.errstate(all="ignore"):
        w = s2.mean(axis=0)
        b = niters * psibar.var(axis=0, ddof=1)
        psibar2 = psibar.mean(axis=0)

        def cov(u, v):
            return ((u - u.mean(axis=0)) * (v - v.mean(axis=0))).sum(axis=0) / (nchains - 1)

        var_w = s2.var(axis=0, ddof=1) / nchains
        var_b = (2 / (nchains - 1)) * b ** 2
        var_wb = (niters / nchains) * (cov(s2, psibar ** 2) - 2 * psibar2 * cov(s2, psibar))
        V = rfixed * w + rrandomscale * b
        var_V = rfixed ** 2 * var_w + rrandomscale ** 2 * var_b + 2 * rfixed * rrandomscale * var_wb
        df = 2 * V ** 2 / var_V
        W_df = 2 * w ** 2 / var_w
        correction = (df + 3) / (df + 1)
        rrandom = rrandomscale * b / w
        psrf = np.sqrt(correction * (rfixed + rrandom))
        q = stats.f.ppf(1 - alpha / 2, nchains - 1, W_df)
        upper = np.where(np.isnan(rrandom), rrandom, rrandom * q)
        psrfci = np.sqrt(correction * (rfixed + upper))
    return GelmanResult(psrf, psrfci)
