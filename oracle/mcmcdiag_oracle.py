"""CPU oracle for the ess / rhat / ess_rhat / mcse / rhat_nested hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (`mcmcdiagnostictools.jl_b200/`,
the C-ABI library, `bench.py`'s GPU arm) may import this file; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` leg do.

What it is: a line-by-line NumPy/SciPy restatement of the reference algorithm
(MCMCDiagnosticTools.jl v0.3.19, `/root/reference/src`), dtype-faithful (Float32 input
keeps Float32 arithmetic, Float64 keeps Float64).  Every function cites the reference
file:line it follows.

PARITY STATUS: **pinned on every known-answer test and identity the reference's own test-suite
holds for this path; no reference-output pin exists.**  The reference is 100 % Julia and no
Julia binary exists in this image (nor on the GPU box), so the reference itself cannot be
run here, and its test-suite holds no literal golden vectors for ESS / R-hat / MCSE (SURVEY
§8(c)): parity against the reference's *outputs* is therefore unpinned.  What
*is* pinned (tests/test_oracle_anchors.py): the `copyto_split!` index goldens
(test/utils.jl:26-56), the antithetic cap identity `max(ess) == ntotal*log10(ntotal)`
(test/ess_rhat.jl:314-327), constants => NaN (:242-257), monotone-transform invariance of
bulk ESS (:329-335), slice consistency (:167-204), direct ~ FFT and identical R-hat across
methods (:228-230), direct/FFT ~ StatsBase.autocov(demean=true) (:259-266), rank-normalised
mean/std (test/utils.jl:98-107), fold identity (:109-123), the nested identity
sqrt(rhat^2 + 1/n) (test/rhat_nested.jl:132-146), rank == max(bulk, tail) (:148-155), and an
independent rank check against scipy.stats.rankdata(method="average").  For the callers added from
SURVEY §8(f): the literal golden values of test/bfmi.jl (0.6 by hand, ArviZ's 0.2406937229) and the
classical Cramer-von Mises critical values for `pcramer` (src/heideldiag.jl:60-71).

Third-party arithmetic that lives outside /root/reference (versions bounded by
Project.toml:23-36 only) is restated from its published algorithm:
  StatsBase 0.34  tiedrank           -> `tiedrank`
  StatsFuns       norminvcdf         -> scipy.special.ndtri   (= -erfcinv(2p)*sqrt(2))
  StatsFuns       betainvcdf         -> scipy.special.betaincinv
  Statistics      median / quantile  -> `jl_median`, `jl_quantile` (type 7, fma aleph)
  Base            nextprod([2,3], n) -> `nextprod23`
  AbstractFFTs/FFTW                  -> numpy.fft
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np
from scipy import special as _sp

__all__ = [
    "AutocovMethod", "FFTAutocovMethod", "BDAAutocovMethod",
    "ess", "rhat", "ess_rhat", "mcse", "rhat_nested",
    "copyto_split", "tiedrank", "rank_normalize", "fold_around_median",
    "jl_median", "jl_quantile", "nextprod23", "ar1", "Quantile",
]

NORMCDF1 = 0.8413447460685429    # src/mcse.jl:1
NORMCDFN1 = 0.15865525393145705  # src/mcse.jl:2


# ----------------------------------------------------------------------------------------
# autocovariance method tags (src/ess_rhat.jl:38,55,73)
# ----------------------------------------------------------------------------------------
class AutocovMethod:
    name = "direct"


class FFTAutocovMethod:
    name = "fft"


class BDAAutocovMethod:
    name = "bda"


class Quantile:
    """Stand-in for `Base.Fix2(Statistics.quantile, p)` (src/ess_rhat.jl:647)."""

    def __init__(self, p):
        self.p = p


class DomainError(ValueError):
    pass


class DimensionMismatch(ValueError):
    pass


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
def _float_dtype(x):
    """promote_type(eltype, typeof(zero(eltype)/1)): Float32 stays, ints -> Float64."""
    dt = np.asarray(x).dtype
    if dt == np.float32:
        return np.float32
    return np.float64


def _as3d(x):
    """(draws,[chains,[params...]]) -> (draws, chains, P), param_shape  (src/utils.jl:197-211)."""
    x = np.asarray(x)
    if x.ndim == 0:
        raise ValueError("samples must have at least one dimension")
    if x.ndim == 1:
        return x.reshape(x.shape[0], 1, 1), ()
    if x.ndim == 2:
        return x.reshape(x.shape[0], x.shape[1], 1), ()
    pshape = x.shape[2:]
    # Julia reshape is column-major over the trailing dims
    return x.reshape(x.shape[0], x.shape[1], -1, order="F"), pshape


def _restore(v, pshape):
    """_maybescalar + parameter axes restored (src/utils.jl:214-215)."""
    v = np.asarray(v)
    if pshape == ():
        return v.reshape(()).item() if v.dtype != object else v.reshape(()).item()
    return v.reshape(pshape, order="F")


def nextprod23(n: int) -> int:
    """Base.nextprod([2,3], n): smallest 2^a 3^b >= n  (src/ess_rhat.jl:110)."""
    best = None
    p3 = 1
    while True:
        v = p3
        while v < n:
            v *= 2
        if best is None or v < best:
            best = v
        if p3 >= n:
            break
        p3 *= 3
    return best


def copyto_split(x2d: np.ndarray, split: int) -> np.ndarray:
    """copyto_split!  (src/utils.jl:13-41).

    x2d is (m, c); returns (m // split, c * split).  When d = m % split > 0 one row is
    skipped after each of the first d splits of every chain.
    """
    m, c = x2d.shape
    niter, d = divmod(m, split)
    out = np.empty((niter, c * split), dtype=x2d.dtype)
    for j in range(c):
        off = 0
        for k in range(split):
            out[:, j * split + k] = x2d[off:off + niter, j]
            off += niter + (1 if (k + 1) <= d else 0)
    return out


def tiedrank(v: np.ndarray) -> np.ndarray:
    """StatsBase.tiedrank (call site src/utils.jl:180): sortperm (isless: NaN last, stable),
    walk equal runs with `!=`, assign (first+last)/2 as Float64."""
    v = np.asarray(v)
    n = v.shape[0]
    p = np.argsort(v, kind="stable")  # NaN last, stable; -0.0/0.0 tie under != anyway
    rks = np.empty(n, dtype=np.float64)
    if n == 0:
        return rks
    sv = v[p]
    # run boundaries: sv[i] != sv[i-1]  (NaN != NaN is True -> each NaN its own run)
    with np.errstate(invalid="ignore"):
        newrun = np.ones(n, dtype=bool)
        newrun[1:] = sv[1:] != sv[:-1]
    starts = np.flatnonzero(newrun)            # 0-based starts
    ends = np.append(starts[1:], n)            # exclusive ends
    runid = np.cumsum(newrun) - 1
    ar = (starts + 1 + ends) / 2.0             # (s + e - 1)/2 with 1-based s, pass-by-end e
    rks[p] = ar[runid]
    return rks


def _norminvcdf(q: np.ndarray, T) -> np.ndarray:
    """StatsFuns.norminvcdf in eltype T (src/utils.jl:182)."""
    return _sp.ndtri(q.astype(np.float64)).astype(T)


def rank_normalize(x) -> np.ndarray:
    """_rank_normalize / _rank_normalize! / _normal_quantiles_from_ranks!  (src/utils.jl:169-193)."""
    x3, _ = _as3d(x)
    T = _float_dtype(x3)
    y = np.empty(x3.shape, dtype=T)
    n = x3.shape[0] * x3.shape[1]
    for i in range(x3.shape[2]):
        v = x3[:, :, i].reshape(-1, order="F")
        r = tiedrank(v)
        # q .= (r .- 3//8) ./ (n - 2*(3//8) + 1): Float64 arithmetic, stored as T
        q = ((r - 0.375) / (n + 0.25)).astype(T)
        y[:, :, i] = _norminvcdf(q, T).reshape(x3.shape[0], x3.shape[1], order="F")
    return y.reshape(np.asarray(x).shape, order="F")


def jl_median(v: np.ndarray):
    """Statistics.median(vec): NaN if any NaN; middle(a, b) = a/2 + b/2 for even n."""
    v = np.asarray(v)
    T = _float_dtype(v)
    if v.dtype.kind == "f" and np.isnan(v).any():
        return T(np.nan)
    n = v.shape[0]
    s = np.sort(v)
    if n % 2 == 1:
        return T(s[n // 2])
    a, b = T(s[n // 2 - 1]), T(s[n // 2])
    return T(a / T(2) + b / T(2))


def jl_quantile(v: np.ndarray, p):
    """Statistics.quantile(v, p) (type 7: alpha = beta = 1), call site src/ess_rhat.jl:655.

    aleph = fma(n, p, 1 - p) in p's type; j = clamp(trunc(aleph), 1, n-1); g = aleph - j;
    a + g*(b - a).  NaNs raise (Julia throws ArgumentError).  Returns a numpy scalar of
    type promote(eltype(v), typeof(p))."""
    v = np.asarray(v)
    if v.dtype.kind == "f" and np.isnan(v).any():
        raise ValueError("quantiles are undefined in presence of NaNs or missing values")
    Tp = np.float32 if isinstance(p, np.float32) else np.float64
    n = v.shape[0]
    s = np.sort(v)
    if v.dtype.kind != "f":
        s = s.astype(np.float64)
    m = Tp(1.0 - float(p))  # oftype(p, alpha + p*(1 - alpha - beta)) with alpha = beta = 1.0
    if Tp is np.float32:
        # fma in Float32: exact product+sum in Float64, rounded once to Float32
        aleph = np.float32(float(np.float32(n)) * float(p) + float(m))
    else:
        aleph = np.float64(math.fma(float(n), float(p), float(m))) if hasattr(math, "fma") else np.float64(n * float(p) + float(m))
    j = int(min(max(math.trunc(float(aleph)), 1), n - 1)) if n > 1 else 1
    g = Tp(min(max(float(Tp(aleph - Tp(j))), 0.0), 1.0))
    if n == 1:
        a = b = s[0]
    else:
        a, b = s[j - 1], s[j]
    R = np.result_type(s.dtype, Tp).type
    a, b, g = R(a), R(b), R(g)
    if np.isfinite(a) and np.isfinite(b):
        return R(a + g * (b - a))
    return R((R(1) - g) * a + g * b)


def fold_around_median(x) -> np.ndarray:
    """_fold_around_median (src/utils.jl:148-158)."""
    x3, _ = _as3d(x)
    T = _float_dtype(x3)
    y = np.empty(x3.shape, dtype=T)
    for i in range(x3.shape[2]):
        xi = x3[:, :, i]
        med = jl_median(xi.reshape(-1, order="F"))
        y[:, :, i] = np.abs(xi.astype(T) - med)
    return y.reshape(np.asarray(x).shape, order="F")


# ----------------------------------------------------------------------------------------
# moments as the reference computes them
# ----------------------------------------------------------------------------------------
def _chain_stats(samples: np.ndarray):
    """mean!(chain_mean, samples); var(view(samples,:,j); mean, corrected=true)
    (src/ess_rhat.jl:391-399, 529-537)."""
    T = samples.dtype.type
    niter = samples.shape[0]
    chain_mean = (samples.sum(axis=0, dtype=T) / T(niter)).astype(T)
    dev = samples - chain_mean[None, :]
    chain_var = ((dev * dev).sum(axis=0, dtype=T) / T(niter - 1)).astype(T)
    return chain_mean, chain_var


def _var_vec(v: np.ndarray, corrected: bool):
    """Statistics.var(v; corrected)."""
    T = v.dtype.type
    n = v.shape[0]
    m = T(v.sum(dtype=T) / T(n))
    d = v - m
    return T((d * d).sum(dtype=T) / T(n - (1 if corrected else 0)))


def _rhat_basic(x3: np.ndarray, split_chains: int) -> np.ndarray:
    """_rhat_basic!  (src/ess_rhat.jl:362-409)."""
    T = _float_dtype(x3)
    draws, chains, P = x3.shape
    niter = draws // split_chains
    nchains = split_chains * chains
    out = np.empty(P, dtype=T)
    cf = T(niter - 1) / T(niter) if niter > 0 else T(np.nan)
    with np.errstate(all="ignore"):
        for i in range(P):
            samples = copyto_split(x3[:, :, i].astype(T), split_chains)
            chain_mean, chain_var = _chain_stats(samples)
            W = T(chain_var.sum(dtype=T) / T(nchains))
            var_plus = T(cf * W + _var_vec(chain_mean, nchains > 1))
            out[i] = np.sqrt(T(var_plus / W))
    return out


def _mean_autocov_factory(method, samples: np.ndarray, chain_var: np.ndarray):
    """build_cache / update! / mean_autocov for the three methods
    (src/ess_rhat.jl:95-213).  `samples` is the centred (niter x nchains) matrix."""
    T = samples.dtype.type
    niter, nchains = samples.shape
    name = method.name
    if name == "direct":
        def f(k):
            # mean_i dot(x[1:n-k,i], x[k+1:n,i]) / niter   (:161-179)
            s = T(0)
            for i in range(nchains):
                s = T(s + T(np.dot(samples[: niter - k, i], samples[k:, i])))
            return T(T(s / T(nchains)) / T(niter))
        return f
    if name == "fft":
        n = nextprod23(2 * niter - 1)
        CT = np.complex64 if T is np.float32 else np.complex128
        buf = np.zeros((n, nchains), dtype=CT)
        buf[:niter, :] = samples
        f1 = np.fft.fft(buf, axis=0).astype(CT)
        f1 = (f1.real * f1.real + f1.imag * f1.imag).astype(CT)
        # plan_ifft! is the normalised inverse
        c = np.fft.ifft(f1, axis=0).astype(CT)
        unc = T(niter - 1) / T(niter)

        def f(k):
            # mean_i(real(c[k+1,i]) / real(c[1,i]) * var_i) * (niter-1)//niter   (:181-195)
            s = T(0)
            for i in range(nchains):
                s = T(s + T(T(c[k, i].real) / T(c[0, i].real)) * chain_var[i])
            return T(T(s / T(nchains)) * unc)
        return f
    if name == "bda":
        mean_chain_var = T(chain_var.sum(dtype=T) / T(nchains))

        def f(k):
            # mean(var) - mean_i sum_{t<=n-k}(x[t,i]-x[t+k,i])^2 / (2(n-k))   (:197-213)
            n = niter - k
            s = T(0)
            for j in range(nchains):
                d = samples[:n, j] - samples[k:k + n, j]
                s = T(s + T((d * d).sum(dtype=T)))
            s = T(s / T(nchains))
            return T(mean_chain_var - T(s / T(2 * n)))
        return f
    raise ValueError(f"unknown autocov method {method!r}")


def _jl_min(a, b, T):
    """Julia min: NaN-propagating."""
    if np.isnan(a) or np.isnan(b):
        return T(np.nan)
    return a if a < b else b


def _jl_max(a, b, T):
    if np.isnan(a) or np.isnan(b):
        return T(np.nan)
    return a if a > b else b


def _ess_rhat_basic(x3, relative, autocov_method, split_chains, maxlag):
    """_ess_rhat_basic!  (src/ess_rhat.jl:488-603).  `maxlag` already clamped by the caller."""
    T = _float_dtype(x3)
    draws, chains, P = x3.shape
    niter = draws // split_chains
    nchains = split_chains * chains
    ntotal = niter * nchains
    ess = np.empty(P, dtype=T)
    rhat = np.empty(P, dtype=T)
    cf = T(niter - 1) / T(niter)
    rel_ess_max = T(np.log10(T(ntotal)))
    one, zero = T(1), T(0)
    with np.errstate(all="ignore"):
        for i in range(P):
            samples = copyto_split(x3[:, :, i].astype(T), split_chains)
            chain_mean, chain_var = _chain_stats(samples)
            W = T(chain_var.sum(dtype=T) / T(nchains))
            var_plus = T(cf * W + _var_vec(chain_mean, nchains > 1))
            inv_var_plus = T(one / var_plus)
            rhat[i] = np.sqrt(T(var_plus / W))
            samples = (samples - chain_mean[None, :]).astype(T)
            mac = _mean_autocov_factory(autocov_method, samples, chain_var)

            def rho(k):
                return T(one - inv_var_plus * T(W - mac(k)))

            rho_odd = rho(1)
            rho_even = one
            p_t = T(rho_even + rho_odd)
            sum_p = p_t
            k = 2
            while k < maxlag - 1:
                rho_even = rho(k)
                rho_odd = rho(k + 1)
                delta = T(rho_even + rho_odd)
                if not (delta > zero):
                    break
                p_t = _jl_min(delta, p_t, T)
                sum_p = T(sum_p + p_t)
                k += 2
            rho_even = rho(k) if maxlag > 1 else zero
            tau = _jl_max(zero, T(T(2) * sum_p + _jl_max(zero, rho_even, T) - one), T)
            ess[i] = _jl_min(T(one / tau), rel_ess_max, T)
    if not relative:
        ess = (ess * T(ntotal)).astype(T)
    return ess, rhat


def _ess_rhat_kind_basic(x3, relative=False, autocov_method=None, split_chains=2, maxlag=250):
    """_ess_rhat(Val(:basic))  (src/ess_rhat.jl:456-487)."""
    if autocov_method is None:
        autocov_method = AutocovMethod()
    T = _float_dtype(x3)
    niter = x3.shape[0] // split_chains
    P = x3.shape[2]
    if not (niter > 4):
        ess = np.full(P, np.nan, dtype=T)
        rh = _rhat_basic(x3, split_chains)
        return ess, rh
    if not maxlag > 0:
        raise DomainError(f"maxlag must be >0 (got {maxlag})")
    maxlag = min(maxlag, niter - 4)
    return _ess_rhat_basic(x3, relative, autocov_method, split_chains, maxlag)


# ----------------------------------------------------------------------------------------
# expectand proxies (src/ess_rhat.jl:628-659)
# ----------------------------------------------------------------------------------------
def _expectand_proxy(kind, x3):
    T = _float_dtype(x3)
    name = _estimator_name(kind)
    if name == "mean":
        return x3
    if name == "median":
        # `y = similar(x)`: for integer input the indicator array keeps the integer eltype
        y = np.empty(x3.shape, dtype=x3.dtype)
        for i in range(x3.shape[2]):
            xi = x3[:, :, i]
            with np.errstate(invalid="ignore"):
                y[:, :, i] = xi <= jl_median(xi.reshape(-1, order="F"))
        return y
    if name == "std":
        xf = x3.astype(T)
        n = x3.shape[0] * x3.shape[1]
        m = (xf.sum(axis=(0, 1), dtype=T) / T(n)).astype(T)
        d = xf - m[None, None, :]
        return (d * d).astype(T)
    if name == "mad":
        xf = fold_around_median(x3)
        return _expectand_proxy("median", xf)
    if name == "quantile":
        p = kind.p
        y = np.empty(x3.shape, dtype=x3.dtype)
        for i in range(x3.shape[2]):
            xi = x3[:, :, i]
            y[:, :, i] = xi <= jl_quantile(xi.reshape(-1, order="F"), p)
        return y
    return None


def _estimator_name(kind):
    if isinstance(kind, Quantile):
        return "quantile"
    if isinstance(kind, str):
        return kind if kind in ("mean", "median", "std", "mad") else None
    if kind is np.mean:
        return "mean"
    if kind is np.median:
        return "median"
    if kind is np.std:
        return "std"
    return getattr(kind, "_mcd_estimator", None)


# ----------------------------------------------------------------------------------------
# kind dispatch
# ----------------------------------------------------------------------------------------
_SYMBOL_KINDS = ("rank", "bulk", "tail", "basic")


def _tail_probs(x3, tail_prob):
    """pl, pu as `_ess(Val(:tail))` makes them (src/ess_rhat.jl:301-311)."""
    T = _float_dtype(x3)
    if isinstance(tail_prob, (int, Fraction)):
        Tp = T                      # Rational promotes to the array's float type
        tp = Fraction(tail_prob)
        pl = Tp(float(tp / 2))
        pu = Tp(float(1 - tp / 2))
    else:
        Tp = np.float32 if (isinstance(tail_prob, np.float32) and T is np.float32) else np.float64
        pl = Tp(Tp(tail_prob) / Tp(2))
        pu = Tp(Tp(1) - Tp(tail_prob) / Tp(2))
    return pl, pu


def _ess_rhat_val(kind, x3, split_chains=2, tail_prob=Fraction(1, 10), **kw):
    """_ess_rhat(::Val{kind}) compositions  (src/ess_rhat.jl:604-624)."""
    if kind == "basic":
        return _ess_rhat_kind_basic(x3, split_chains=split_chains, **kw)
    if kind == "bulk":
        return _ess_rhat_kind_basic(rank_normalize(x3), split_chains=split_chains, **kw)
    if kind == "tail":
        S = _ess_val("tail", x3, split_chains=split_chains, tail_prob=tail_prob, **kw)
        R = _rhat_val("tail", x3, split_chains=split_chains)
        return S, R
    if kind == "rank":
        Sb, Rb = _ess_rhat_val("bulk", x3, split_chains=split_chains, **kw)
        Rt = _rhat_val("tail", x3, split_chains=split_chains)
        return Sb, _map_max(Rt, Rb)
    raise ValueError(f"the `kind` `{kind}` is not supported by `ess_rhat`")


def _map_max(a, b):
    """map(max, a, b) with Julia's NaN-propagating max."""
    out = np.where(a > b, a, b)
    out = np.where(np.isnan(a) | np.isnan(b), np.nan, out)
    return out.astype(a.dtype)


def _map_min(a, b):
    out = np.where(a < b, a, b)
    out = np.where(np.isnan(a) | np.isnan(b), np.nan, out)
    return out.astype(a.dtype)


def _rhat_val(kind, x3, split_chains=2):
    """_rhat(::Val{kind})  (src/ess_rhat.jl:350-361, 410-420)."""
    if kind == "basic":
        return _rhat_basic(x3, split_chains)
    if kind == "bulk":
        return _rhat_basic(rank_normalize(x3), split_chains)
    if kind == "tail":
        return _rhat_val("bulk", fold_around_median(x3), split_chains)
    if kind == "rank":
        return _map_max(_rhat_val("tail", x3, split_chains), _rhat_val("bulk", x3, split_chains))
    raise ValueError(f"the `kind` `{kind}` is not supported by `rhat`")


def _ess_val(kind, x3, tail_prob=Fraction(1, 10), **kw):
    """_ess(...)  (src/ess_rhat.jl:291-311)."""
    if kind == "tail":
        pl, pu = _tail_probs(x3, tail_prob)
        Sl = _ess_estimator(Quantile(pl), x3, **kw)
        Su = _ess_estimator(Quantile(pu), x3, **kw)
        return _map_min(Sl, Su)
    if kind in ("bulk", "basic"):
        return _ess_rhat_val(kind, x3, **kw)[0]
    raise ValueError(f"the `kind` `{kind}` is not supported by `ess`")


def _ess_estimator(kind, x3, **kw):
    y = _expectand_proxy(kind, x3)
    if y is None:
        raise ValueError(f"the estimator {kind} is not yet supported by `ess`")
    return _ess_rhat_kind_basic(y, **kw)[0]


# ----------------------------------------------------------------------------------------
# public API (src/ess_rhat.jl:276-290, 335-349, 438-455; src/mcse.jl:40-42;
#             src/rhat_nested.jl:43-66)
# ----------------------------------------------------------------------------------------
def ess(samples, kind="bulk", **kw):
    x3, pshape = _as3d(samples)
    if isinstance(kind, str) and kind in ("bulk", "tail", "basic"):
        return _restore(_ess_val(kind, x3, **kw), pshape)
    if isinstance(kind, str) and kind not in ("mean", "median", "std", "mad"):
        raise ValueError(f"the `kind` `{kind}` is not supported by `ess`")
    kw.pop("tail_prob", None)
    return _restore(_ess_estimator(kind, x3, **kw), pshape)


def rhat(samples, kind="rank", split_chains=2):
    x3, pshape = _as3d(samples)
    if kind not in _SYMBOL_KINDS:
        raise ValueError(f"the `kind` `{kind}` is not supported by `rhat`")
    return _restore(_rhat_val(kind, x3, split_chains), pshape)


def ess_rhat(samples, kind="rank", **kw):
    x3, pshape = _as3d(samples)
    if kind not in _SYMBOL_KINDS:
        raise ValueError(f"the `kind` `{kind}` is not supported by `ess_rhat`")
    if kind != "tail":
        kw.pop("tail_prob", None)
    S, R = _ess_rhat_val(kind, x3, **kw)
    return _restore(S, pshape), _restore(R, pshape)


# ----------------------------------------------------------------------------------------
# mcse (src/mcse.jl:40-118)
# ----------------------------------------------------------------------------------------
def _mcse_quantile(v, p, Seff):
    """_mcse_quantile (src/mcse.jl:96-118)."""
    T = _float_dtype(v)
    if np.isnan(Seff):
        return T(np.nan)
    S = v.shape[0]
    Seff = float(Seff)
    p = float(p)
    a = Seff * p + 1
    b = Seff * (1 - p) + 1
    pu = _sp.betaincinv(a, b, NORMCDF1)
    pl = _sp.betaincinv(a, b, NORMCDFN1)
    l = max(math.floor(pl * S), 1)
    u = min(math.ceil(pu * S), S)
    s = np.sort(v)
    xl, xu = T(s[l - 1]), T(s[u - 1])
    return T(T(xu - xl) / T(2))


def mcse(samples, kind="mean", **kw):
    x3, pshape = _as3d(samples)
    T = _float_dtype(x3)
    name = _estimator_name(kind)
    draws, chains, P = x3.shape
    n = draws * chains
    with np.errstate(all="ignore"):
        if name == "mean":
            # std(samples; dims=(1,2)) ./ sqrt.(S)   (src/mcse.jl:45-51)
            S = _ess_estimator("mean", x3, **kw)
            xf = x3.astype(T)
            m = (xf.sum(axis=(0, 1), dtype=T) / T(n)).astype(T)
            d = xf - m[None, None, :]
            sd = np.sqrt(((d * d).sum(axis=(0, 1), dtype=T) / T(n - 1)).astype(T))
            return _restore((sd / np.sqrt(S)).astype(T), pshape)
        if name == "std":
            # (src/mcse.jl:52-65)
            xf = x3.astype(T)
            m = (xf.sum(axis=(0, 1), dtype=T) / T(n)).astype(T)
            d = xf - m[None, None, :]
            proxy = (d * d).astype(T)
            S = _ess_estimator("mean", proxy, **kw)
            mean_var = (proxy.sum(axis=(0, 1), dtype=T) / T(n)).astype(T)
            mean_m4 = ((proxy * proxy).sum(axis=(0, 1), dtype=T) / T(n)).astype(T)
            out = np.sqrt((mean_m4 / mean_var - mean_var) / S) / T(2)
            return _restore(out.astype(T), pshape)
        if name in ("median", "quantile"):
            # (src/mcse.jl:66-94)
            p = Fraction(1, 2) if name == "median" else kind.p
            S = _ess_estimator(kind, x3, **kw)
            out = np.empty(P, dtype=T)
            for i in range(P):
                out[i] = _mcse_quantile(x3[:, :, i].reshape(-1, order="F").astype(T), p, S[i])
            return _restore(out, pshape)
    raise NotImplementedError("SBM fallback (src/mcse.jl:120-148) is out of scope (SURVEY §2)")


SUMMARY_FIELDS = ("mean", "std", "mcse_mean", "mcse_std", "ess_bulk", "ess_tail", "rhat")


def summary(samples, fields=None, autocov_method=None, split_chains=2, maxlag=250, tail_prob=Fraction(1, 10)):
    """The per-parameter columns downstream summaries assemble from separate calls of the
    reference: every column is literally the reference call named beside it."""
    x3, pshape = _as3d(samples)
    T = _float_dtype(x3)
    n = x3.shape[0] * x3.shape[1]
    kw = dict(split_chains=split_chains, maxlag=maxlag)
    if autocov_method is not None:
        kw["autocov_method"] = autocov_method
    names = SUMMARY_FIELDS if fields is None else tuple(fields)

    def moments():
        with np.errstate(all="ignore"):
            xf = x3.astype(T)
            m = (xf.sum(axis=(0, 1), dtype=T) / T(n)).astype(T)                   # Statistics.mean(x; dims=(1,2))
            d = xf - m[None, None, :]
            return m, np.sqrt(((d * d).sum(axis=(0, 1), dtype=T) / T(n - 1)).astype(T))  # Statistics.std

    column = {
        "mean": lambda: _restore(moments()[0], pshape),
        "std": lambda: _restore(moments()[1], pshape),
        "mcse_mean": lambda: mcse(samples, kind="mean", **kw),                    # src/mcse.jl:45-51
        "mcse_std": lambda: mcse(samples, kind="std", **kw),                      # src/mcse.jl:52-65
        "ess_bulk": lambda: ess(samples, kind="bulk", **kw),                      # src/ess_rhat.jl:604-624
        "ess_tail": lambda: ess(samples, kind="tail", tail_prob=tail_prob, **kw), # src/ess_rhat.jl:298-311
        "rhat": lambda: rhat(samples, kind="rank", split_chains=split_chains),    # src/ess_rhat.jl:410-420
    }
    return {k: column[k]() for k in SUMMARY_FIELDS if k in names}


# ----------------------------------------------------------------------------------------
# nested R-hat (src/rhat_nested.jl:43-188)
# ----------------------------------------------------------------------------------------
def _validate_superchain_ids(ids, nchains):
    """_validate_superchain_ids + unique_indices (src/rhat_nested.jl:68-81, src/utils.jl:50-64).
    Superchains ordered by sorted label; chains within by first appearance.
    Returns an int matrix (chains_per_super x nsuper), 0-based."""
    ids = list(ids)
    if len(ids) != nchains:
        raise DimensionMismatch(
            f"`superchain_ids` has length {len(ids)} but `samples` has {nchains} chains")
    groups = {}
    for i, s in enumerate(ids):
        groups.setdefault(s, []).append(i)
    keys = sorted(groups.keys())
    if len(keys) < 2:
        raise ValueError(f"at least 2 superchains are required, got {len(keys)}")
    sizes = {len(groups[k]) for k in keys}
    if len(sizes) != 1:
        raise ValueError("all superchains must contain the same number of chains")
    return np.stack([np.asarray(groups[k], dtype=np.int64) for k in keys], axis=1)


def _rhat_nested_basic(x3, chain_inds, split_chains):
    """_rhat_nested_basic!  (src/rhat_nested.jl:127-188)."""
    T = _float_dtype(x3)
    draws, chains, P = x3.shape
    m = chain_inds.shape[0] * split_chains
    K = chain_inds.shape[1]
    out = np.empty(P, dtype=T)
    with np.errstate(all="ignore"):
        for i in range(P):
            sl = x3[:, :, i].astype(T)
            vw = T(0)
            sc_mean = np.empty(K, dtype=T)
            for k in range(K):
                samples = copyto_split(sl[:, chain_inds[:, k]], split_chains)
                chain_mean, chain_var = _chain_stats(samples)
                sc_mean[k] = T(chain_mean.sum(dtype=T) / T(m))
                Wk = T(chain_var.sum(dtype=T) / T(m))
                Bk = _var_vec(chain_mean, m > 1)
                vw = T(vw + T(Wk + Bk))
            vw = T(vw / T(K))
            vb = _var_vec(sc_mean, True)
            out[i] = np.sqrt(T(T(1) + T(vb / vw)))
    return out


def rhat_nested(samples, superchain_ids, kind="rank", split_chains=2):
    x = np.asarray(samples)
    if x.ndim < 2:
        raise ValueError("`samples` must have at least 2 dimensions `(draws, chains[, parameters…])`")
    x3, pshape = _as3d(x)
    inds = _validate_superchain_ids(superchain_ids, x3.shape[1])

    def val(kind, x3):
        if kind == "basic":
            return _rhat_nested_basic(x3, inds, split_chains)
        if kind == "bulk":
            return val("basic", rank_normalize(x3))
        if kind == "tail":
            return val("bulk", fold_around_median(x3))
        if kind == "rank":
            return _map_max(val("bulk", x3), val("tail", x3))
        raise ValueError(f"the `kind` `{kind}` is not supported by `rhat_nested`")

    return _restore(val(kind, x3), pshape)


# ----------------------------------------------------------------------------------------
# synthetic input (test/helpers.jl:4-12)
# ----------------------------------------------------------------------------------------
def ar1(phi, sigma, *shape, rng=None, dtype=np.float64):
    """x = sigma*randn; accumulate along dim 1 with muladd(phi, x_prev, eps)."""
    rng = np.random.default_rng(1) if rng is None else rng
    x = (rng.standard_normal(shape) * sigma).astype(dtype)
    phi = dtype(phi)
    for t in range(1, shape[0]):
        x[t] = phi * x[t - 1] + x[t]
    return x


# ----------------------------------------------------------------------------------------
# callers of the path: gewekediag (src/gewekediag.jl:19-35), heideldiag (src/heideldiag.jl:16-71)
# ----------------------------------------------------------------------------------------
def _jl_round_int(v):
    return int(np.round(v))          # Julia round(Int, x): ties to even, as numpy


def pcramer(q):
    """Csorgo & Faraway (1996) series for the Cramer-von Mises distribution (src/heideldiag.jl:60-71)."""
    from scipy import special
    q = float(q)
    p = 0.0
    for k in range(4):
        c1 = 4.0 * k + 1.0
        c2 = c1 * c1 / (16.0 * q)
        p += special.gamma(k + 0.5) / math.factorial(k) * math.sqrt(c1) * math.exp(-c2) * special.kv(0.25, c2)
    return p / (math.pi ** 1.5 * math.sqrt(q))


def gewekediag(x, first=0.1, last=0.5, **kw):
    """src/gewekediag.jl:19-35 on one vector."""
    from scipy import special
    if not 0 < first < 1:
        raise ValueError("`first` is not in (0, 1)")
    if not 0 < last < 1:
        raise ValueError("`last` is not in (0, 1)")
    if not first + last <= 1:
        raise ValueError("`first` and `last` proportions overlap")
    x = np.asarray(x)
    T = _float_dtype(x)
    n = x.shape[0]
    x1 = x[: _jl_round_int(first * n)]
    x2 = x[_jl_round_int(n - last * n + 1) - 1:]
    s1 = mcse(x1.reshape(-1, 1, 1), kind="mean", split_chains=1, **kw)[0]
    s2 = mcse(x2.reshape(-1, 1, 1), kind="mean", split_chains=1, **kw)[0]
    s = T(np.hypot(s1, s2))
    z = T((T(x1.astype(T).mean(dtype=T)) - T(x2.astype(T).mean(dtype=T))) / s)
    pv = T(special.erfc(abs(float(z)) / math.sqrt(2.0)))
    return {"zscore": z, "pvalue": pv}


def heideldiag(x, alpha=Fraction(1, 20), eps=0.1, start=1, **kw):
    """src/heideldiag.jl:16-54 on one vector."""
    from scipy import special
    x = np.asarray(x)
    T = _float_dtype(x)
    xf = x.astype(T)
    n = x.shape[0]
    delta = int(0.10 * n)
    y = xf[int(n / 2) - 1:]
    s = mcse(y.reshape(-1, 1, 1), kind="mean", split_chains=1, **kw)[0]
    S0 = T(len(y)) * s * s
    i, pvalue, converged, ybar = 1, T(1), False, T(np.nan)
    while i < n / 2:
        y = xf[i - 1:]
        m = len(y)
        ybar = T(y.mean(dtype=T))
        B = np.cumsum(y, dtype=T) - ybar * np.arange(1, m + 1, dtype=T)
        Bsq = (B * B) / (T(m) * S0)
        I = T(Bsq.sum(dtype=T) / T(m))
        with np.errstate(all="ignore"):
            pvalue = T(1) - T(pcramer(I))
        converged = bool(pvalue > float(alpha))
        if converged or delta == 0:     # delta == 0 (n < 10) never advances in the reference
            break
        i += delta
    s = mcse(y.reshape(-1, 1, 1), kind="mean", split_chains=1, **kw)[0]
    halfwidth = T(math.sqrt(2.0)) * T(special.erfcinv(float(T(float(alpha))))) * s
    passed = bool(halfwidth / abs(ybar) <= eps)
    return {"burnin": i + start - 2, "stationarity": converged, "pvalue": pvalue, "mean": ybar,
            "halfwidth": T(halfwidth), "test": passed}


# ----------------------------------------------------------------------------------------
# bfmi (src/bfmi.jl:36-43), gelmandiag (src/gelmandiag.jl:1-76)
# ----------------------------------------------------------------------------------------
def bfmi(energy, dims=1):
    e = np.asarray(energy)
    T = _float_dtype(e)
    e = e.astype(T)
    if e.ndim == 1:
        return T(np.mean(np.diff(e) ** 2, dtype=T) / np.var(e, ddof=1, dtype=T))
    ax = dims - 1
    return (np.mean(np.diff(e, axis=ax) ** 2, axis=ax, dtype=T) / np.var(e, axis=ax, ddof=1, dtype=T)).astype(T)


def gelmandiag(psi, alpha=0.05):
    """Line-by-line restatement of `_gelmandiag` (full covariance matrices, then their diagonals)."""
    from scipy import stats
    psi = np.asarray(psi, dtype=np.float64)
    niters, nchains, nparams = psi.shape
    if not nchains > 1:
        raise RuntimeError("Gelman diagnostic requires at least 2 chains")
    rfixed = (niters - 1) / niters
    rrandomscale = (nchains + 1) / (nchains * niters)
    S2 = [np.atleast_2d(np.cov(psi[:, i, :], rowvar=False)) for i in range(nchains)]
    W = sum(S2) / nchains
    psibar = psi.mean(axis=0)                                   # (chains, params)
    B = niters * np.atleast_2d(np.cov(psibar, rowvar=False))
    w, b = np.diag(W), np.diag(B)
    s2 = np.stack([np.diag(S) for S in S2], axis=0)             # (chains, params)
    psibar2 = psibar.mean(axis=0)

    def cov_diag(u, v):
        return np.array([np.cov(u[:, j], v[:, j])[0, 1] for j in range(nparams)])

    with np.errstate(all="ignore"):
        var_w = s2.var(axis=0, ddof=1) / nchains
        var_b = (2 / (nchains - 1)) * b ** 2
        var_wb = (niters / nchains) * (cov_diag(s2, psibar ** 2) - 2 * psibar2 * cov_diag(s2, psibar))
        V = rfixed * w + rrandomscale * b
        var_V = rfixed ** 2 * var_w + rrandomscale ** 2 * var_b + 2 * rfixed * rrandomscale * var_wb
        df = 2 * V ** 2 / var_V
        W_df = 2 * w ** 2 / var_w
        est = np.empty(nparams)
        up = np.empty(nparams)
        for i in range(nparams):
            correction = (df[i] + 3) / (df[i] + 1)
            rrandom = rrandomscale * b[i] / w[i]
            est[i] = np.sqrt(correction * (rfixed + rrandom))
            if not np.isnan(rrandom):
                rrandom *= stats.f.ppf(1 - alpha / 2, nchains - 1, W_df[i])
            up[i] = np.sqrt(correction * (rfixed + rrandom))
    return {"psrf": est, "psrfci": up}
