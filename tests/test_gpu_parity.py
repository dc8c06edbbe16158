"""Parity of the CUDA path (through the Python host -> C ABI -> sm_100a kernels) against the
CPU oracle on identical seeded inputs.  Tolerances are the north star's: ranks and ties
bit-exact; ESS / R-hat / MCSE within 1e-8 relative for Float64 and 1e-4 for Float32.
"""
from fractions import Fraction

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu

RTOL64 = 1e-8
RTOL32 = 1e-4


@pytest.fixture(scope="module")
def mcd():
    import mcmcdiag_b200 as m
    m.get_context(0)
    return m


@pytest.fixture(scope="module")
def o():
    from oracle import mcmcdiag_oracle
    return mcmcdiag_oracle


def rng(seed):
    return np.random.default_rng(seed)


def close(a, b, rtol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.allclose(a, b, rtol=rtol, atol=0, equal_nan=True)


METHODS = ["AutocovMethod", "FFTAutocovMethod", "BDAAutocovMethod"]


# --- C1: the reference's own test-suite scale ---------------------------------------------------
def test_c1_ess_rhat_rank(mcd, o):
    x = o.ar1(0.5, np.sqrt(0.75), 1000, 4, 10, rng=rng(1))
    S, R = mcd.ess_rhat(x)
    So, Ro = o.ess_rhat(x)
    assert S.shape == (10,) and S.dtype == np.float64
    assert close(S, So, RTOL64) and close(R, Ro, RTOL64)


@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("split_chains", [1, 2, 3])
def test_kinds_methods_splits(mcd, o, kind, method, split_chains):
    x = o.ar1(0.7, np.sqrt(1 - 0.49), 301, 4, 6, rng=rng(2))
    for maxlag in (250, 10, 1):
        S, R = mcd.ess_rhat(x, kind=kind, autocov_method=getattr(mcd, method)(), split_chains=split_chains, maxlag=maxlag)
        So, Ro = o.ess_rhat(x, kind=kind, autocov_method=getattr(o, method)(), split_chains=split_chains, maxlag=maxlag)
        assert close(S, So, RTOL64), (S, So)
        assert close(R, Ro, RTOL64), (R, Ro)
    Rr = mcd.rhat(x, kind=kind, split_chains=split_chains)
    assert close(Rr, o.rhat(x, kind=kind, split_chains=split_chains), RTOL64)
    if kind != "rank":
        Se = mcd.ess(x, kind=kind, autocov_method=getattr(mcd, method)(), split_chains=split_chains)
        assert close(Se, o.ess(x, kind=kind, autocov_method=getattr(o, method)(), split_chains=split_chains), RTOL64)


@pytest.mark.parametrize("phi", [-0.9, -0.3, 0.0, 0.5, 0.9, 0.99])
def test_ar1_sweep(mcd, o, phi):
    x = o.ar1(phi, np.sqrt(1 - phi ** 2), 1000, 4, 8, rng=rng(3))
    for kind in ("rank", "tail", "basic"):
        S, R = mcd.ess_rhat(x, kind=kind)
        So, Ro = o.ess_rhat(x, kind=kind)
        assert close(S, So, RTOL64) and close(R, Ro, RTOL64)


# --- ranks: bit exact ---------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["normal", "cauchy", "ties", "signedzero", "constant", "nan", "inf", "tiny"])
def test_tiedrank_bit_exact(mcd, o, case):
    r = rng(4)
    shape = (500, 4, 5)
    if case == "normal":
        x = r.standard_normal(shape)
    elif case == "cauchy":
        x = r.standard_cauchy(shape)
    elif case == "ties":
        x = r.integers(1, 11, shape).astype(np.float64)
    elif case == "signedzero":
        x = np.where(r.random(shape) < 0.5, 0.0, -0.0) * np.where(r.random(shape) < 0.3, 1.0, 0.0) + np.where(r.random(shape) < 0.2, 1.0, 0.0)
    elif case == "constant":
        x = np.full(shape, 3.25)
    elif case == "nan":
        x = r.standard_normal(shape)
        x[r.random(shape) < 0.01] = np.nan
        x[:, :, 4] = np.nan
    elif case == "inf":
        x = r.standard_normal(shape)
        x[0, 0, :] = np.inf
        x[1, 1, :] = -np.inf
    else:
        x = r.standard_normal(shape) * 1e-310
    got = mcd.tiedrank(x)
    for p in range(shape[2]):
        want = o.tiedrank(x[:, :, p].reshape(-1, order="F"))
        assert np.array_equal(got[:, :, p].reshape(-1, order="F"), want), case


def test_tiedrank_float32_and_scipy(mcd):
    x = rng(5).standard_normal((333, 3, 4)).astype(np.float32)
    x[::7] = x[1::7][: x[::7].shape[0]]  # ties
    got = mcd.tiedrank(x)
    for p in range(4):
        want = stats.rankdata(x[:, :, p].reshape(-1, order="F"), method="average")
        assert np.array_equal(got[:, :, p].reshape(-1, order="F"), want)


def test_rank_normalize_and_fold(mcd, o):
    x = rng(6).standard_exponential((1000, 4, 8))
    assert close(mcd.rank_normalize(x), o.rank_normalize(x), 1e-13)
    assert close(mcd.fold_around_median(x), o.fold_around_median(x), 1e-15)
    z = mcd.rank_normalize(x)
    assert np.all(np.abs(z.mean(axis=(0, 1))) < 1e-13)          # test/utils.jl:98-107
    xo = rng(6).standard_normal((999, 3, 2))                        # odd n: median is an element
    assert close(mcd.fold_around_median(xo), o.fold_around_median(xo), 1e-15)


# --- estimators and mcse -------------------------------------------------------------------------
@pytest.mark.parametrize("method", METHODS)
def test_estimator_ess(mcd, o, method):
    x = o.ar1(0.3, np.sqrt(1 - 0.09), 1000, 4, 6, rng=rng(7)) * 3 + 1
    pairs = [("mean", "mean"), ("median", "median"), ("std", "std"), ("mad", "mad"),
             (mcd.Quantile(0.25), o.Quantile(0.25)), (np.mean, np.mean), (np.median, np.median), (np.std, np.std)]
    for km, ko in pairs:
        S = mcd.ess(x, kind=km, autocov_method=getattr(mcd, method)())
        So = o.ess(x, kind=ko, autocov_method=getattr(o, method)())
        assert close(S, So, RTOL64), (km, S, So)


def test_mcse(mcd, o):
    x = o.ar1(0.5, np.sqrt(0.75), 1000, 4, 12, rng=rng(8)) * 2 - 5
    for km, ko in [("mean", "mean"), ("std", "std"), ("median", "median"), (mcd.Quantile(0.25), o.Quantile(0.25)),
                   (mcd.Quantile(0.9), o.Quantile(0.9))]:
        assert close(mcd.mcse(x, kind=km), o.mcse(x, kind=ko), RTOL64), km
    assert np.all(np.isnan(mcd.mcse(np.ones((100, 4, 3)), kind="median")))
    assert np.all(np.isnan(mcd.mcse(np.ones((100, 4, 3)), kind="mean")))
    assert close(mcd.mcse(x), o.mcse(x), RTOL64)                   # default kind = mean


# --- nested R-hat ---------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
@pytest.mark.parametrize("split_chains", [1, 2])
def test_rhat_nested(mcd, o, kind, split_chains):
    x = rng(9).standard_normal((101, 16, 5)) + rng(10).standard_normal((1, 16, 1)) * 0.1
    ids = [3, 1, 2, 0, 1, 3, 0, 2, 2, 2, 0, 0, 1, 1, 3, 3]
    R = mcd.rhat_nested(x, ids, kind=kind, split_chains=split_chains)
    assert close(R, o.rhat_nested(x, ids, kind=kind, split_chains=split_chains), RTOL64)
    labels = ["d", "b", "c", "a", "b", "d", "a", "c", "c", "c", "a", "a", "b", "b", "d", "d"]
    assert np.array_equal(mcd.rhat_nested(x, labels, kind=kind, split_chains=split_chains), R)


def test_rhat_nested_identity(mcd):
    x = rng(11).standard_normal((100, 8, 5))
    for kind in ("basic", "bulk", "tail", "rank"):
        Rn = mcd.rhat_nested(x, list(range(8)), kind=kind, split_chains=1)
        R = mcd.rhat(x, kind=kind, split_chains=1)
        assert np.allclose(Rn, np.sqrt(R ** 2 + 1 / 100))           # test/rhat_nested.jl:132-146


# --- Float32 ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
def test_float32(mcd, o, kind):
    x = o.ar1(0.5, np.sqrt(0.75), 1000, 4, 10, rng=rng(12)).astype(np.float32)
    for method in METHODS:
        S, R = mcd.ess_rhat(x, kind=kind, autocov_method=getattr(mcd, method)())
        So, Ro = o.ess_rhat(x, kind=kind, autocov_method=getattr(o, method)())
        assert S.dtype == np.float32 and R.dtype == np.float32
        assert close(S, So, RTOL32) and close(R, Ro, RTOL32)
    for km, ko in [("median", "median"), ("std", "std"), ("mad", "mad"), (mcd.Quantile(np.float32(0.3)), o.Quantile(np.float32(0.3)))]:
        assert close(mcd.ess(x, kind=km), o.ess(x, kind=ko), RTOL32)
        if km != "mad":
            assert close(mcd.mcse(x, kind=km), o.mcse(x, kind=ko), RTOL32)


def test_int_input_promotes(mcd, o):
    x = rng(13).integers(1, 10_001, (1000, 4, 5))
    S, R = mcd.ess_rhat(x, kind="tail")
    assert S.dtype == np.float64 and np.all(np.isfinite(S))
    So, Ro = o.ess_rhat(x, kind="tail")
    assert close(S, So, RTOL64) and close(R, Ro, RTOL64)
    xi = rng(13).integers(1, 11, (1000, 4, 5))                       # heavy ties -> sort fallback path
    S, R = mcd.ess_rhat(xi)
    So, Ro = o.ess_rhat(xi)
    assert close(S, So, RTOL64) and close(R, Ro, RTOL64)


# --- exact anchors of the reference test-suite, on the GPU path -------------------------------------
@pytest.mark.parametrize("ndraws", [10, 100])
@pytest.mark.parametrize("phi", [-0.3, -0.9])
def test_antithetic_cap_exact(mcd, o, ndraws, phi):
    x = o.ar1(phi, np.sqrt(1 - phi ** 2), ndraws, 4, 1000, rng=rng(14))
    S = mcd.ess(x, kind="mean")
    ntotal = ndraws * 4
    assert S.max() == ntotal * np.log10(ntotal)
    assert S.min() > 0
    assert close(S, o.ess(x, kind="mean"), RTOL64)


def test_constant_gives_nan(mcd):
    x = np.ones((1000, 10, 4))
    for method in METHODS:
        S, R = mcd.ess_rhat(x, autocov_method=getattr(mcd, method)())
        assert np.all(np.isnan(S)) and np.all(np.isnan(R))


def test_monotone_invariance_exact(mcd):
    xn = rng(15).standard_normal((1000, 4, 10))
    xc = stats.cauchy.ppf(stats.norm.cdf(xn))
    assert np.array_equal(mcd.ess(xn, kind="bulk"), mcd.ess(xc, kind="bulk"))
    assert np.array_equal(mcd.ess(xn, kind="bulk"), mcd.ess(mcd.rank_normalize(xn), kind="basic"))


@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
def test_slice_consistency_exact(mcd, kind):
    x = rng(16).standard_normal((400, 4, 5, 3))
    for method in ("AutocovMethod", "BDAAutocovMethod"):
        S, R = mcd.ess_rhat(x, kind=kind, autocov_method=getattr(mcd, method)(), maxlag=100)
        assert S.shape == (5, 3)
        for i in range(3):
            Si, Ri = mcd.ess_rhat(x[:, :, :, i], kind=kind, autocov_method=getattr(mcd, method)(), maxlag=100)
            assert np.array_equal(Si, S[:, i]) and np.array_equal(Ri, R[:, i])
            for j in range(5):
                Sji, Rji = mcd.ess_rhat(x[:, :, j, i], kind=kind, autocov_method=getattr(mcd, method)(), maxlag=100)
                assert Sji == S[j, i] and Rji == R[j, i]
                assert isinstance(Sji, np.float64)
    assert np.array_equal(mcd.rhat(x, kind=kind), R)


def test_methods_share_rhat_and_fft_close(mcd):
    x = 50 * rng(17).standard_normal((2000, 3, 8))
    S, R = mcd.ess_rhat(x, kind="basic")
    Sf, Rf = mcd.ess_rhat(x, kind="basic", autocov_method=mcd.FFTAutocovMethod())
    Sb, Rb = mcd.ess_rhat(x, kind="basic", autocov_method=mcd.BDAAutocovMethod())
    assert np.array_equal(R, Rf) and np.array_equal(R, Rb)
    assert np.allclose(S, Sf, rtol=1e-8)


# --- errors, warnings, edge shapes --------------------------------------------------------------------
def test_errors(mcd):
    r = rng(18)
    x2, x3 = r.random((5, 3, 5)), r.random((100, 3, 5))
    for kind in ("rank", "bulk", "tail", "basic"):
        with pytest.raises(mcd.DomainError):
            mcd.ess_rhat(x3, maxlag=0, kind=kind)
        mcd.ess_rhat(x3, maxlag=1, kind=kind)
    with pytest.raises(mcd.ArgumentError):
        mcd.ess_rhat(x2, kind="foo")
    with pytest.raises(mcd.ArgumentError):
        mcd.rhat(x2, kind="foo")
    with pytest.raises(mcd.ArgumentError):
        mcd.ess(x2, kind="rank")
    with pytest.raises(mcd.ArgumentError):
        mcd.ess(x2, kind=lambda v: v.mean())
    xn = r.random((100, 3, 2))
    xn[3, 1, 1] = np.nan
    with pytest.raises(mcd.ArgumentError):
        mcd.ess(xn, kind="tail")                 # Statistics.quantile throws on NaN
    with pytest.raises(mcd.DimensionMismatch):
        mcd.rhat_nested(x3, [1, 2])
    with pytest.raises(mcd.ArgumentError):
        mcd.rhat_nested(x3, [1, 1, 1])
    with pytest.raises(mcd.ArgumentError):
        mcd.rhat_nested(x3[:, 0, 0], [1])


def test_short_chains(mcd, o):
    r = rng(19)
    x, x4 = r.random((4, 3, 5)), r.random((1, 3, 5))
    for kind in ("rank", "bulk", "tail", "basic"):
        with pytest.warns(UserWarning):
            S, R = mcd.ess_rhat(x, split_chains=1, kind=kind)
        assert np.all(np.isnan(S))
        assert np.array_equal(R, mcd.rhat(x, split_chains=1, kind=kind), equal_nan=True)
        assert close(R, o.rhat(x, split_chains=1, kind=kind), RTOL64)
        with pytest.warns(UserWarning):
            S, R = mcd.ess_rhat(x4, split_chains=2, kind=kind)
        assert np.all(np.isnan(S))


def test_nan_data_bulk_is_finite(mcd, o):
    x = rng(20).standard_normal((200, 4, 3))
    x[5, 2, 1] = np.nan
    S, R = mcd.ess_rhat(x, kind="bulk")
    So, Ro = o.ess_rhat(x, kind="bulk")
    assert np.all(np.isfinite(S)) and close(S, So, RTOL64) and close(R, Ro, RTOL64)
    Rt = mcd.rhat(x, kind="tail")                 # fold -> all NaN -> ranks by index order
    assert close(Rt, o.rhat(x, kind="tail"), RTOL64)


def test_relative_and_shapes(mcd):
    x = rng(21).random((100, 4, 2, 3))
    S, R = mcd.ess_rhat(x, kind="bulk")
    S2, R2 = mcd.ess_rhat(x, kind="bulk", relative=True)
    assert S.shape == (2, 3) and np.allclose(S2, S / 400) and np.array_equal(R, R2)
    assert isinstance(mcd.ess(x[:, 0, 0, 0]), np.float64)
    assert isinstance(mcd.rhat(x[:, :, 0, 0]), np.float64)
    assert isinstance(mcd.ess(x[:, :, 0, 0].astype(np.float32)), np.float32)


def test_missing_values(mcd):
    x = np.ma.masked_array(rng(22).standard_normal((1000, 4, 3)))
    x[0, 0, 0] = np.ma.masked
    S, R = mcd.ess_rhat(x)
    assert S.mask.tolist() == [True, False, False] and R.mask.tolist() == [True, False, False]
    S2, R2 = mcd.ess_rhat(np.asarray(x.data)[:, :, 1:])
    assert np.array_equal(S.compressed(), S2) and np.array_equal(R.compressed(), R2)


def test_device_tensor_input_matches_host(mcd):
    import torch
    x = rng(23).standard_normal((1000, 4, 7))
    xd = torch.from_numpy(np.ascontiguousarray(x.transpose(2, 1, 0))).cuda().permute(2, 1, 0)
    for kind in ("rank", "tail", "basic"):
        S, R = mcd.ess_rhat(x, kind=kind)
        Sd, Rd = mcd.ess_rhat(xd, kind=kind)
        assert np.array_equal(Sd.cpu().numpy(), S) and np.array_equal(Rd.cpu().numpy(), R)


def test_generator_is_shard_invariant(mcd, o):
    a = mcd.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, 64, seed=1)
    b = mcd.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, 32, seed=1, param_offset=32)
    assert np.array_equal(a[:, :, 32:].cpu().numpy(), b.cpu().numpy())
    x = a.cpu().numpy()
    assert abs(x.mean()) < 0.02 and abs(x.std() - 1) < 0.02
    lag1 = np.mean(x[1:] * x[:-1])
    assert abs(lag1 - 0.5) < 0.02
    S, R = mcd.ess_rhat(a)
    So, Ro = o.ess_rhat(x)
    assert close(S.cpu().numpy(), So, RTOL64) and close(R.cpu().numpy(), Ro, RTOL64)


def test_host_chunking_matches(mcd):
    x = rng(24).standard_normal((200, 4, 50))
    S, R = mcd.ess_rhat(x)
    ctx = mcd.get_context(0)
    ctx.set_option("h2d_chunk_bytes", 200 * 4 * 8 * 7)        # 7 parameters per chunk
    try:
        S2, R2 = mcd.ess_rhat(x)
    finally:
        ctx.set_option("h2d_chunk_bytes", 256 << 20)
    assert np.array_equal(S, S2) and np.array_equal(R, R2)


def test_empty_and_degenerate_shapes(mcd, o):
    import warnings
    S, R = mcd.ess_rhat(np.zeros((100, 4, 0)))
    assert S.shape == (0,) and R.shape == (0,)
    assert mcd.rhat(np.zeros((100, 4, 0, 3))).shape == (0, 3)
    r = rng(70)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for shape, split in (((3, 2, 4), 2), ((2, 1, 3), 2), ((1, 4, 2), 2), ((5, 1, 2), 1), ((6, 3, 2), 3)):
            x = r.standard_normal(shape)
            for kind in ("rank", "basic", "tail"):
                S, R = mcd.ess_rhat(x, kind=kind, split_chains=split)
                So, Ro = o.ess_rhat(x, kind=kind, split_chains=split)
                assert close(S, So, RTOL64) and close(R, Ro, RTOL64), (shape, split, kind, R, Ro)
    # one chain: uncorrected between-chain variance (src/ess_rhat.jl:403,541)
    x1 = o.ar1(0.3, 1.0, 400, 1, 3, rng=rng(71))
    for split in (1, 2):
        S, R = mcd.ess_rhat(x1, split_chains=split)
        So, Ro = o.ess_rhat(x1, split_chains=split)
        assert close(S, So, RTOL64) and close(R, Ro, RTOL64)


def test_mcse_quantile_over_a_range_of_ess(mcd, o):
    """The block-cooperative betainvcdf (quadrature + Newton) against the oracle's scipy inverse over a
    wide range of Beta parameters (ESS from a few to the cap; src/mcse.jl:96-118)."""
    r = rng(77)
    cols = []
    for phi in (-0.6, 0.0, 0.3, 0.6, 0.9, 0.98, 0.995):
        cols.append(o.ar1(phi, np.sqrt(1 - phi * phi), 400, 4, 12, rng=r))
    x = np.concatenate(cols, axis=2)
    for kind, okind in ((np.median, "median"), (mcd.Quantile(0.1), o.Quantile(0.1)), (mcd.Quantile(0.93), o.Quantile(0.93))):
        got = mcd.mcse(x, kind=kind)
        want = o.mcse(x, kind=okind)
        assert close(got, want, RTOL64), (kind, np.max(np.abs(got / want - 1)))


@pytest.mark.parametrize("draws", [100, 500, 1000, 2000])
def test_many_nans_rank_in_place(mcd, o, draws):
    """Regression: with many NaNs the in-place rank-normalisation of the general kernel raced (a NaN slot
    was overwritten before its owner had read it).  NaNs rank last, each distinct, in index order."""
    r = rng(91)
    x = np.full((draws, 4, 2), np.inf)
    x[..., 1] = r.standard_normal((draws, 4))
    x[np.cumsum(r.standard_normal((draws, 4, 2)), axis=0) > 0] = np.nan      # clustered NaNs
    for _ in range(3):
        assert np.allclose(mcd.rank_normalize(x), o.rank_normalize(x), rtol=1e-12, equal_nan=True)
        assert np.array_equal(mcd.tiedrank(x), np.stack([o.tiedrank(x[..., j].reshape(-1, order="F")).reshape(draws, 4, order="F")
                                                         for j in range(2)], axis=2))
        assert close(mcd.rhat(x, kind="bulk"), o.rhat(x, kind="bulk"), RTOL64)


def test_all_infinite_parameter_tail_rhat(mcd, o):
    """A parameter of +-Inf only: the fold gives NaN (Inf - Inf) for one sign and Inf for the other."""
    r = rng(92)
    for draws in (50, 1000, 2000):
        x = np.where(np.cumsum(r.standard_normal((draws, 4, 2)), axis=0) > 0, np.inf, -np.inf)
        assert close(mcd.rhat(x, kind="tail"), o.rhat(x, kind="tail"), RTOL64)
        assert close(mcd.rhat(x, kind="rank"), o.rhat(x, kind="rank"), RTOL64)
