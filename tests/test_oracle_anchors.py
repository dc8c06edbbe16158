"""Pins the CPU oracle (oracle/mcmcdiag_oracle.py) against every exact / identity anchor the
reference's own test-suite holds for the hot path.  CPU only.

The reference (Julia) cannot run in this image and publishes no literal golden vectors for
ESS / R-hat / MCSE, so these anchors are what "pinning" means here; see the oracle header.
"""
from fractions import Fraction

import numpy as np
import pytest
from scipy import stats

from oracle import mcmcdiag_oracle as o


def rng(seed=1):
    return np.random.default_rng(seed)


# --- test/utils.jl:26-56 : copyto_split! literal index goldens ---------------------------
def test_copyto_split_even():
    x = rng().random((50, 20))
    y = o.copyto_split(x, 2)
    assert y.shape == (25, 40)
    assert np.array_equal(y.reshape(50, 20, order="F"), x)


def test_copyto_split_odd_two_splits():
    x = rng().random((51, 20))
    y = o.copyto_split(x, 2)
    rows = np.r_[0:25, 26:51]                       # Julia vcat(1:25, 27:51)
    assert np.array_equal(y.reshape(50, 20, order="F"), x[rows, :])


def test_copyto_split_three_splits():
    x = rng().random((50, 20))
    y = o.copyto_split(x, 3)
    rows = np.r_[0:16, 17:33, 34:50]                # vcat(1:16, 18:33, 35:50)
    assert np.array_equal(y.reshape(48, -1, order="F"), x[rows, :])
    x = rng().random((49, 20))
    y = o.copyto_split(x, 3)
    rows = np.r_[0:16, 17:33, 33:49]                # vcat(1:16, 18:33, 34:49)
    assert np.array_equal(y.reshape(48, -1, order="F"), x[rows, :])


# --- StatsBase.tiedrank vs an independent implementation ----------------------------------
@pytest.mark.parametrize("kind", ["cont", "ties", "signedzero"])
def test_tiedrank_matches_scipy(kind):
    r = rng(3)
    if kind == "cont":
        v = r.standard_normal(4000)
    elif kind == "ties":
        v = r.integers(1, 11, 4000).astype(np.float64)
    else:
        v = np.where(r.random(200) < 0.5, 0.0, -0.0)
    assert np.array_equal(o.tiedrank(v), stats.rankdata(v, method="average"))


def test_tiedrank_nan_last_distinct_in_index_order():
    v = np.array([np.nan, 2.0, np.nan, 1.0, 2.0])
    assert np.array_equal(o.tiedrank(v), np.array([4.0, 2.5, 5.0, 1.0, 2.5]))


# --- test/utils.jl:98-107 : rank-normalised mean ~ 0 (1e-13), std ~ 1 (1e-2) --------------
@pytest.mark.parametrize("sz", [(1000,), (1000, 4), (1000, 4, 8), (1000, 4, 8, 2)])
def test_rank_normalize_moments(sz):
    x = rng(5).exponential(size=sz)
    z = o.rank_normalize(x)
    assert z.shape == x.shape
    dims = tuple(range(min(2, x.ndim)))
    assert np.all(np.abs(z.mean(axis=dims)) < 1e-13)
    assert np.allclose(z.std(axis=dims, ddof=1), 1, rtol=1e-2)


# --- test/utils.jl:109-123 : fold ~ abs(x - median) ----------------------------------------
@pytest.mark.parametrize("sz", [(1000,), (1000, 4), (1000, 4, 8)])
def test_fold_identity(sz):
    x = rng(6).random(sz)
    dims = tuple(range(min(2, x.ndim)))
    med = np.median(x, axis=dims, keepdims=True)
    assert np.allclose(o.fold_around_median(x), np.abs(x - med))


# --- Statistics.quantile type 7 / median -----------------------------------------------------
def test_quantile_type7_matches_numpy_linear():
    v = rng(7).standard_normal(4000)
    for p in (0.05, 0.25, 0.5, 0.95):
        assert np.isclose(o.jl_quantile(v, p), np.quantile(v, p), rtol=1e-13)
    assert o.jl_median(v) == np.sort(v)[1999] / 2 + np.sort(v)[2000] / 2
    assert np.isnan(o.jl_median(np.array([1.0, np.nan])))
    with pytest.raises(ValueError):
        o.jl_quantile(np.array([1.0, np.nan]), 0.5)


def test_nextprod23():
    assert o.nextprod23(999) == 1024
    assert o.nextprod23(99) == 108
    assert o.nextprod23(999_999) == 2 ** 20
    assert o.nextprod23(1) == 1
    assert o.nextprod23(19999) == 20736


# --- test/ess_rhat.jl:314-327 : antithetic cap identity (exact ==) ---------------------------
@pytest.mark.parametrize("ndraws", [10, 100])
@pytest.mark.parametrize("phi", [-0.3, -0.9])
def test_antithetic_cap(ndraws, phi):
    x = o.ar1(phi, np.sqrt(1 - phi ** 2), ndraws, 4, 1000, rng=rng(11))
    S = o.ess(x, kind="mean")
    ntotal = ndraws * 4
    assert S.max() == ntotal * np.log10(ntotal)
    assert S.min() > 0


# --- test/ess_rhat.jl:242-257 : constants => NaN for all three methods -----------------------
@pytest.mark.parametrize("method", [o.AutocovMethod, o.FFTAutocovMethod, o.BDAAutocovMethod])
def test_identical_samples_nan(method):
    x = np.ones((1000, 10, 4))
    S, R = o.ess_rhat(x, autocov_method=method())
    assert np.all(np.isnan(S)) and np.all(np.isnan(R))


# --- test/ess_rhat.jl:329-335 : monotone-transform invariance of bulk ESS (==) ---------------
def test_bulk_ess_monotone_invariance():
    xn = rng(12).standard_normal((1000, 4, 10))
    xc = stats.cauchy.ppf(stats.norm.cdf(xn))
    assert np.array_equal(o.ess(xn, kind="bulk"), o.ess(xc, kind="bulk"))
    assert np.array_equal(o.ess(xn, kind="bulk"), o.ess(o.rank_normalize(xn), kind="basic"))


# --- test/ess_rhat.jl:167-204 : ess / rhat / ess_rhat / slice consistency (==) ---------------
@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
@pytest.mark.parametrize("split_chains", [1, 2])
def test_consistency(kind, split_chains):
    x = rng(13).standard_normal((200, 4, 3, 2))
    R1 = o.rhat(x, kind=kind, split_chains=split_chains)
    kind_ess = "bulk" if kind == "rank" else kind
    for method in (o.AutocovMethod(), o.BDAAutocovMethod()):
        for maxlag in (100, 10):
            kw = dict(split_chains=split_chains, autocov_method=method, maxlag=maxlag)
            S1 = o.ess(x, kind=kind_ess, **kw)
            S2, R2 = o.ess_rhat(x, kind=kind, **kw)
            assert np.array_equal(S1, S2) and np.array_equal(R1, R2)
            for i in range(2):
                Si, Ri = o.ess_rhat(x[:, :, :, i], kind=kind, **kw)
                assert np.array_equal(Si, S1[:, i]) and np.array_equal(Ri, R1[:, i])
                for j in range(3):
                    Sji, Rji = o.ess_rhat(x[:, :, j, i], kind=kind, **kw)
                    assert Sji == S1[j, i] and Rji == R1[j, i]


# --- test/ess_rhat.jl:210-240 : IID; methods agree; R-hat identical across methods -----------
@pytest.mark.parametrize("nchains", [1, 10])
@pytest.mark.parametrize("split_chains", [1, 2])
def test_iid(nchains, split_chains):
    x = 50 * rng(14).standard_normal((10_000, nchains, 8))
    ntotal = 10_000 * nchains
    S, R = o.ess_rhat(x, split_chains=split_chains)
    Sf, Rf = o.ess_rhat(x, split_chains=split_chains, autocov_method=o.FFTAutocovMethod())
    Sb, Rb = o.ess_rhat(x, split_chains=split_chains, autocov_method=o.BDAAutocovMethod())
    assert np.allclose(S, Sf, rtol=1e-8)
    assert np.array_equal(R, Rf) and np.array_equal(R, Rb)
    assert np.allclose(S, ntotal, rtol=0.15) and np.allclose(Sb, ntotal, rtol=0.15)  # statistical band (seed differs from Julia)
    assert np.allclose(R, 1, rtol=0.1)


# --- test/ess_rhat.jl:259-266 : direct / FFT ~ StatsBase.autocov(demean=true) -----------------
def test_autocov_definition():
    x = rng(15).standard_normal((1000, 10, 6))
    niter, nch = 500, 20
    S = o.ess(x, kind="basic")
    Sf = o.ess(x, kind="basic", autocov_method=o.FFTAutocovMethod())

    class Explicit:
        name = "explicit"

    # StatsBase.autocov(x, k:k; demean=true)[1] = sum((x-m)[1:n-k] .* (x-m)[k+1:n]) / n
    def explicit_ess(x3):
        out = np.empty(x3.shape[2])
        for i in range(x3.shape[2]):
            s = o.copyto_split(x3[:, :, i], 2)
            m = s.mean(axis=0)
            v = s.var(axis=0, ddof=1)
            W = v.mean()
            vp = (niter - 1) / niter * W + m.var(ddof=1)
            c = s - m

            def mac(k):
                return np.mean([(c[: niter - k, j] * c[k:, j]).sum() / niter for j in range(nch)])

            rho = lambda k: 1 - (W - mac(k)) / vp
            pt = 1 + rho(1)
            sp = pt
            k = 2
            while k < 249:
                d = rho(k) + rho(k + 1)
                if not d > 0:
                    break
                pt = min(d, pt)
                sp += pt
                k += 2
            tau = max(0, 2 * sp + max(0, rho(k)) - 1)
            out[i] = min(1 / tau, np.log10(niter * nch)) * niter * nch
        return out

    Se = explicit_ess(x)
    assert np.allclose(S, Se, rtol=1e-10) and np.allclose(Sf, Se, rtol=1e-8)


# --- test/ess_rhat.jl:268-276 : two unmixed epochs ---------------------------------------------
def test_two_epochs():
    x = rng(16).standard_normal((1000, 4, 10)) + np.repeat([0.0, 10.0], 500)[:, None, None]
    S1, R1 = o.ess_rhat(x, kind="basic", split_chains=1)
    S2, R2 = o.ess_rhat(x, kind="basic", split_chains=2)
    assert np.allclose(R1, 1, rtol=0.1)
    assert np.all(S2 < S1) and np.all(R2 > 2)


# --- test/ess_rhat.jl:66-98 : errors / NaN rules -------------------------------------------------
def test_errors_and_short_chains():
    r = rng(17)
    x, x2, x3, x4 = r.random((4, 3, 5)), r.random((5, 3, 5)), r.random((100, 3, 5)), r.random((1, 3, 5))
    for kind in ("rank", "bulk", "tail", "basic"):
        S, R = o.ess_rhat(x, split_chains=1, kind=kind)
        assert np.all(np.isnan(S))
        assert np.array_equal(R, o.rhat(x, split_chains=1, kind=kind), equal_nan=True)
        assert np.all(np.isnan(o.ess_rhat(x4, split_chains=2, kind=kind)[0]))
        o.ess_rhat(x2, split_chains=1, kind=kind)
        S, R = o.ess_rhat(x2, split_chains=2, kind=kind)
        assert np.all(np.isnan(S))
        o.ess_rhat(x3, maxlag=1, kind=kind)
        with pytest.raises(o.DomainError):
            o.ess_rhat(x3, maxlag=0, kind=kind)
    with pytest.raises(ValueError):
        o.ess_rhat(x2, kind="foo")
    with pytest.raises(ValueError):
        o.rhat(x2, kind="foo")
    with pytest.raises(ValueError):
        o.ess(x2, kind=lambda v: v.mean())
    with pytest.raises(ValueError):
        o.ess(x2, kind="rank")


def test_relative():
    x = rng(18).random((100, 4, 2))
    for kind in ("rank", "bulk", "tail", "basic"):
        S, R = o.ess_rhat(x, kind=kind)
        S2, R2 = o.ess_rhat(x, kind=kind, relative=True)
        assert np.allclose(S2, S / 400) and np.array_equal(R, R2)


# --- test/ess_rhat.jl:337-375 : tail detection, integer input ------------------------------------
def test_tail_detects_scale_mismatch():
    phi = 0.1
    sig = np.sqrt(1 - phi ** 2) * np.array([0.1, 1, 10, 100])
    r = rng(19)
    x = 10 + np.concatenate([o.ar1(phi, s, 1000, 1, 20, rng=r) for s in sig], axis=1)
    S, R = o.ess_rhat(x, kind="basic")
    assert np.all(S >= 400) and np.all(R <= 1.01)
    S, R = o.ess_rhat(x, kind="bulk")
    assert np.all(S >= 400) and np.all(R <= 1.01)
    S, R = o.ess_rhat(x, kind="tail")
    assert np.all(S < 400) and np.all(R > 1.01)


def test_integer_input():
    x = rng(20).integers(1, 10_001, (1000, 4, 5))
    S, _ = o.ess_rhat(x, kind="tail")
    assert S.dtype == np.float64 and np.all(np.isfinite(S))


def test_float32_stays_float32():
    x = rng(21).standard_normal((100, 4, 2)).astype(np.float32)
    for kind in ("rank", "bulk", "tail", "basic"):
        S, R = o.ess_rhat(x, kind=kind)
        assert S.dtype == np.float32 and R.dtype == np.float32
    x64 = x.astype(np.float64)
    assert np.allclose(o.ess_rhat(x)[0], o.ess_rhat(x64)[0], rtol=1e-3)


# --- mcse (test/mcse.jl) ---------------------------------------------------------------------------
def test_mcse_constant_nan_and_slices():
    x = np.ones((100, 4, 3))
    for kind in ("mean", "std", "median", o.Quantile(0.25)):
        assert np.all(np.isnan(o.mcse(x, kind=kind)))
    x = rng(22).standard_normal((200, 4, 3))
    for kind in ("mean", "std", "median", o.Quantile(0.25)):
        full = o.mcse(x, kind=kind)
        for j in range(3):
            assert np.isclose(o.mcse(x[:, :, j], kind=kind), full[j])


def test_mcse_mean_iid_scale():
    x = rng(23).standard_normal((1000, 4, 20))
    se = o.mcse(x, kind="mean")
    assert np.allclose(se, 1 / np.sqrt(4000), rtol=0.2)


# --- nested R-hat (test/rhat_nested.jl) --------------------------------------------------------------
def test_nested_errors():
    x = rng(24).standard_normal((100, 4, 2))
    with pytest.raises(ValueError):
        o.rhat_nested(x[:, 0, 0], [1])
    with pytest.raises(o.DimensionMismatch):
        o.rhat_nested(x, [1, 1, 2])
    with pytest.raises(ValueError):
        o.rhat_nested(x, [1, 1, 1, 1])
    with pytest.raises(ValueError):
        o.rhat_nested(x, [1, 1, 1, 2])
    with pytest.raises(ValueError):
        o.rhat_nested(x, [1, 1, 2, 2], kind="foo")


def test_nested_identity_one_chain_per_superchain():
    # test/rhat_nested.jl:132-146: rhat_nested ~ sqrt(rhat^2 + 1/ndraws) with split_chains=1
    x = rng(25).standard_normal((100, 8, 5))
    for kind in ("basic", "bulk", "tail", "rank"):
        Rn = o.rhat_nested(x, list(range(8)), kind=kind, split_chains=1)
        R = o.rhat(x, kind=kind, split_chains=1)
        assert np.allclose(Rn, np.sqrt(R ** 2 + 1 / 100))


def test_nested_rank_is_max_and_invariances():
    x = rng(26).standard_normal((100, 8, 5))
    ids = [1, 1, 2, 2, 3, 3, 4, 4]
    Rb = o.rhat_nested(x, ids, kind="bulk")
    Rt = o.rhat_nested(x, ids, kind="tail")
    assert np.array_equal(o.rhat_nested(x, ids, kind="rank"), np.maximum(Rb, Rt))
    # label-type invariance (test/rhat_nested.jl:102-111)
    assert np.array_equal(o.rhat_nested(x, ["a", "a", "b", "b", "c", "c", "d", "d"]), o.rhat_nested(x, ids))
    # joint permutation invariance (:113-130) -- exact up to summation order
    perm = rng(27).permutation(8)
    assert np.allclose(o.rhat_nested(x[:, perm, :], [ids[p] for p in perm]), o.rhat_nested(x, ids), rtol=1e-12)


def test_nested_iid_many_chains():
    x = rng(28).standard_normal((20, 512, 3))
    ids = np.repeat(np.arange(8), 64)
    R = o.rhat_nested(x, ids)
    assert np.all(R > 1) and np.all(R < 1.01)


# --- callers of the path -------------------------------------------------------
def test_pcramer_known_critical_values():
    """Cramer-von Mises limiting distribution: the classical critical values 0.34730 / 0.46136 / 0.74346
    have probabilities 0.90 / 0.95 / 0.99 (Anderson & Darling 1952), src/heideldiag.jl:60-71."""
    from oracle import mcmcdiag_oracle as o
    for q, p in ((0.34730, 0.90), (0.46136, 0.95), (0.74346, 0.99)):
        assert abs(o.pcramer(q) - p) < 2e-5


def test_gewekediag_heideldiag_oracle_behaviour():
    """test/gewekediag.jl:10-18 (exceptions), result fields and types (test/gewekediag.jl:2-7,
    test/heideldiag.jl:2-7), and the qualitative contract: a stationary series passes, a series with
    a strong transient fails both diagnostics."""
    from oracle import mcmcdiag_oracle as o
    rng = np.random.default_rng(5)
    x = rng.standard_normal(100)
    for v in (-0.3, 0, 1, 1.2):
        with pytest.raises(ValueError):
            o.gewekediag(x, first=v)
        with pytest.raises(ValueError):
            o.gewekediag(x, last=v)
    with pytest.raises(ValueError):
        o.gewekediag(x, first=0.6, last=0.5)
    for T in (np.float32, np.float64):
        g = o.gewekediag(x.astype(T))
        assert set(g) == {"zscore", "pvalue"} and g["zscore"].dtype == T and g["pvalue"].dtype == T
        h = o.heideldiag(x.astype(T))
        assert set(h) == {"burnin", "stationarity", "pvalue", "mean", "halfwidth", "test"}
    y = o.ar1(0.5, 0.8, 2000, 1, 1, rng=rng)[:, 0, 0]
    assert o.gewekediag(y)["pvalue"] > 0.01 and o.heideldiag(y)["stationarity"]
    bad = y + np.linspace(3, 0, 2000) ** 2
    assert o.gewekediag(bad)["pvalue"] < 1e-6 and not o.heideldiag(bad)["stationarity"]


BFMI_ENERGY = [42, 44, 45, 46, 42, 43, 36, 36, 31, 36, 36, 32, 36, 31, 31, 29, 29, 30, 25, 26, 29, 29, 27, 30, 31, 29]


def test_bfmi_golden_values():
    """test/bfmi.jl:2-43: the hand value 0.6 and ArviZ's 0.2406937229 for the 26-draw energy series."""
    from oracle import mcmcdiag_oracle as o
    assert np.isclose(o.bfmi([1, 2, 3, 4]), 0.6)
    assert np.isclose(o.bfmi(BFMI_ENERGY), 0.2406937229, rtol=1e-9)
    multi = np.repeat(np.asarray(BFMI_ENERGY, dtype=float)[:, None], 4, axis=1)
    assert np.allclose(o.bfmi(multi), 0.2406937229, rtol=1e-9)
    assert np.allclose(o.bfmi(multi), o.bfmi(multi.T, dims=2))


def test_gelmandiag_oracle_behaviour():
    """test/gelmandiag.jl: shapes / types / exceptions; PSRF ~ 1 for IID chains, > 1.2 when chains disagree."""
    from oracle import mcmcdiag_oracle as o
    rng = np.random.default_rng(8)
    x = rng.standard_normal((100, 2, 4))
    r = o.gelmandiag(x)
    assert r["psrf"].shape == (4,) and r["psrf"].dtype == np.float64 and r["psrfci"].shape == (4,)
    assert np.all(np.abs(r["psrf"] - 1) < 0.1) and np.all(r["psrfci"] >= r["psrf"])
    with pytest.raises(RuntimeError):
        o.gelmandiag(x[:, :1, :])
    y = rng.standard_normal((500, 4, 2)); y[:, 0, 0] += 3.0
    assert o.gelmandiag(y)["psrf"][0] > 1.2 and abs(o.gelmandiag(y)["psrf"][1] - 1) < 0.05


def test_host_pcramer_matches_oracle():
    """The host mirror's vectorised Cramer-von Mises series (api._pcramer) equals the oracle's scalar one."""
    from oracle import mcmcdiag_oracle as o
    import mcmcdiag_b200 as m
    q = np.array([0.01, 0.05, 0.2, 0.34730, 0.46136, 0.74346, 1.5, 3.0])
    got = m.api._pcramer(q)
    want = np.array([o.pcramer(v) for v in q])
    assert np.allclose(got, want, rtol=1e-14, atol=0)


def test_callers_argument_errors_need_no_gpu():
    """Argument validation of the host wrappers happens before any device call (test/gewekediag.jl:10-18)."""
    import mcmcdiag_b200 as m
    x = np.random.default_rng(0).standard_normal(100)
    for v in (-0.3, 0, 1, 1.2):
        with pytest.raises(m.ArgumentError):
            m.gewekediag(x, first=v)
    with pytest.raises(m.ArgumentError):
        m.gewekediag(x, first=0.6, last=0.5)
    with pytest.raises(RuntimeError):
        m.gelmandiag(np.zeros((10, 1, 2)))
    with pytest.raises(m.ArgumentError):
        m.summary(np.zeros((100, 4, 2)), fields=("bogus",))
    with pytest.raises(m.ArgumentError):
        m.bfmi(np.zeros((4, 4, 4)))
