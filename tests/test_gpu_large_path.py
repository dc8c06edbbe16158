"""Parity of the global-memory large-slab pipeline (mcd_large.cuh), forced on inputs small
enough for the oracle, plus agreement with the shared-memory slab kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL64 = 1e-8
RTOL32 = 1e-4
METHODS = ["AutocovMethod", "FFTAutocovMethod", "BDAAutocovMethod"]


@pytest.fixture(scope="module")
def o():
    from oracle import mcmcdiag_oracle
    return mcmcdiag_oracle


@pytest.fixture()
def mcd():
    import mcmcdiag_b200 as m
    ctx = m.get_context(0)
    ctx.set_option("force_path", 2)
    yield m
    ctx.set_option("force_path", 0)


def rng(seed):
    return np.random.default_rng(seed)


def close(a, b, rtol):
    return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=0, equal_nan=True)


@pytest.mark.parametrize("shape", [(1000, 4, 6), (3000, 4, 3), (1237, 5, 4), (40, 64, 3)])
@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
def test_large_kinds(mcd, o, shape, kind):
    x = o.ar1(0.6, np.sqrt(1 - 0.36), *shape, rng=rng(31))
    for method in METHODS:
        S, R = mcd.ess_rhat(x, kind=kind, autocov_method=getattr(mcd, method)())
        assert mcd.get_context(0).stat("last_path") == 2
        So, Ro = o.ess_rhat(x, kind=kind, autocov_method=getattr(o, method)())
        assert close(S, So, RTOL64), (method, S, So)
        assert close(R, Ro, RTOL64)
    for split_chains in (1, 3):
        S, R = mcd.ess_rhat(x, kind=kind, split_chains=split_chains, maxlag=20)
        So, Ro = o.ess_rhat(x, kind=kind, split_chains=split_chains, maxlag=20)
        assert close(S, So, RTOL64) and close(R, Ro, RTOL64)
    assert close(mcd.rhat(x, kind=kind), o.rhat(x, kind=kind), RTOL64)


@pytest.mark.parametrize("case", ["normal", "ties", "nan", "constant", "signedzero"])
def test_large_tiedrank_bit_exact(mcd, o, case):
    r = rng(32)
    shape = (2600, 4, 3)
    if case == "normal":
        x = r.standard_cauchy(shape)
    elif case == "ties":
        x = r.integers(1, 11, shape).astype(np.float64)
    elif case == "nan":
        x = r.standard_normal(shape)
        x[r.random(shape) < 0.01] = np.nan
        x[:, :, 2] = np.nan
    elif case == "constant":
        x = np.full(shape, -1.5)
    else:
        x = np.where(r.random(shape) < 0.5, 0.0, -0.0) + np.where(r.random(shape) < 0.2, 1.0, 0.0)
    got = mcd.tiedrank(x)
    for p in range(shape[2]):
        assert np.array_equal(got[:, :, p].reshape(-1, order="F"), o.tiedrank(x[:, :, p].reshape(-1, order="F"))), case


def test_large_transforms(mcd, o):
    x = rng(33).standard_exponential((2500, 4, 3))
    assert close(mcd.rank_normalize(x), o.rank_normalize(x), 1e-13)
    assert close(mcd.fold_around_median(x), o.fold_around_median(x), 1e-15)


def test_large_estimators_and_mcse(mcd, o):
    x = o.ar1(0.3, np.sqrt(1 - 0.09), 2200, 4, 4, rng=rng(34)) * 3 + 1
    for km, ko in [("mean", "mean"), ("median", "median"), ("std", "std"), ("mad", "mad"),
                   (mcd.Quantile(0.25), o.Quantile(0.25))]:
        assert close(mcd.ess(x, kind=km), o.ess(x, kind=ko), RTOL64), km
        if km != "mad":
            assert close(mcd.mcse(x, kind=km), o.mcse(x, kind=ko), RTOL64), km
    assert close(mcd.ess(x, kind="tail"), o.ess(x, kind="tail"), RTOL64)


@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
def test_large_nested(mcd, o, kind):
    x = rng(35).standard_normal((100, 64, 3)) + rng(36).standard_normal((1, 64, 1)) * 0.05
    ids = np.repeat(np.arange(8), 8)
    assert close(mcd.rhat_nested(x, ids, kind=kind), o.rhat_nested(x, ids, kind=kind), RTOL64)


def test_large_float32(mcd, o):
    x = o.ar1(0.5, np.sqrt(0.75), 4000, 8, 3, rng=rng(37)).astype(np.float32)
    for kind in ("rank", "tail", "basic"):
        S, R = mcd.ess_rhat(x, kind=kind, autocov_method=mcd.BDAAutocovMethod())
        So, Ro = o.ess_rhat(x, kind=kind, autocov_method=o.BDAAutocovMethod())
        assert S.dtype == np.float32 and close(S, So, RTOL32) and close(R, Ro, RTOL32)
    for km, ko in [("median", "median"), ("std", "std")]:
        assert close(mcd.ess(x, kind=km, autocov_method=mcd.BDAAutocovMethod()),
                     o.ess(x, kind=ko, autocov_method=o.BDAAutocovMethod()), RTOL32)


def test_large_matches_slab_and_chunks(mcd):
    x = rng(38).standard_normal((1000, 4, 40))
    ctx = mcd.get_context(0)
    S2, R2 = mcd.ess_rhat(x)
    ctx.set_option("workspace_bytes", 1 << 20)        # forces several parameter chunks
    try:
        S3, R3 = mcd.ess_rhat(x)
    finally:
        ctx.set_option("workspace_bytes", 6 << 30)
    assert np.array_equal(S2, S3) and np.array_equal(R2, R3)
    ctx.set_option("force_path", 1)
    S1, R1 = mcd.ess_rhat(x)
    assert ctx.stat("last_path") == 1
    assert close(S1, S2, 1e-10) and close(R1, R2, 1e-12)


def test_auto_path_picks_large_for_big_slabs(o):
    import mcmcdiag_b200 as m
    ctx = m.get_context(0)
    ctx.set_option("force_path", 0)
    x = o.ar1(0.5, np.sqrt(0.75), 20000, 4, 2, rng=rng(39))
    S, R = m.ess_rhat(x)
    assert ctx.stat("last_path") == 2
    So, Ro = o.ess_rhat(x)
    assert close(S, So, RTOL64) and close(R, Ro, RTOL64)


@pytest.mark.parametrize("draws,chains,split", [(8000, 2, 2), (10000, 2, 2), (9001, 3, 1), (30000, 1, 2)])
def test_large_four_step_fft(o, draws, chains, split):
    """FFTAutocovMethod on chains whose transform (nextprod([2,3], 2 niter - 1) points) does not fit
    shared memory: four-step FFT (8192 = 2^13, 10368 = 2^7 3^4, 18432 = 2^11 3^2, 32768 points)."""
    import mcmcdiag_b200 as m
    x = o.ar1(0.8, np.sqrt(1 - 0.64), draws, chains, 3, rng=rng(40))
    for kind in ("basic", "bulk"):
        S, R = m.ess_rhat(x, kind=kind, split_chains=split, autocov_method=m.FFTAutocovMethod())
        assert m.get_context(0).stat("last_path") == 2
        So, Ro = o.ess_rhat(x, kind=kind, split_chains=split, autocov_method=o.FFTAutocovMethod())
        assert close(S, So, RTOL64), (S, So)
        assert close(R, Ro, RTOL64)
        Sd, _ = m.ess_rhat(x, kind=kind, split_chains=split)
        assert close(S, Sd, 1e-8)                     # FFT ~ direct (test/ess_rhat.jl:228-230)
    S = m.ess(x, kind="basic", split_chains=split, autocov_method=m.FFTAutocovMethod(), maxlag=5000)
    assert close(S, o.ess(x, kind="basic", split_chains=split, autocov_method=o.FFTAutocovMethod(), maxlag=5000), RTOL64)
    x32 = x.astype(np.float32)
    S32 = m.ess(x32, kind="basic", split_chains=split, autocov_method=m.FFTAutocovMethod())
    assert close(S32, o.ess(x32, kind="basic", split_chains=split, autocov_method=o.FFTAutocovMethod()), 2e-3)
