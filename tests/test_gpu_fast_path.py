"""The register-resident fast kernel (mcd_fast.cuh: 8 split chains of <= 512 draws, one warp per
split chain) against the oracle and against the general slab kernel, including the slabs it
hands back to the general kernel (NaN, infinities, heavy ties, constants)."""
import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu
RTOL64, RTOL32 = 1e-8, 1e-4


@pytest.fixture(scope="module")
def o():
    from oracle import mcmcdiag_oracle
    return mcmcdiag_oracle


@pytest.fixture()
def mcd():
    import mcmcdiag_b200 as m
    ctx = m.get_context(0)
    ctx.set_option("force_path", 0)
    yield m
    ctx.set_option("force_path", 0)


def rng(s):
    return np.random.default_rng(s)


def close(a, b, rtol):
    return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=0, equal_nan=True)


@pytest.mark.parametrize("shape,split", [((1000, 4), 2), ((500, 8), 1), ((1024, 4), 2), ((64, 4), 2), ((300, 2), 4),
                                          ((999, 8), 1), ((10, 4), 2), ((1002, 4), 2)])
@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
def test_fast_shapes(mcd, o, shape, split, kind):
    x = o.ar1(0.5, np.sqrt(0.75), shape[0], shape[1], 12, rng=rng(51))
    ctx = mcd.get_context(0)
    fast_ok = shape[0] % split == 0 and shape[0] // split <= 512
    for maxlag in (250, 9, 2, 1):
        S, R = mcd.ess_rhat(x, kind=kind, split_chains=split, maxlag=maxlag)
        if kind != "tail":
            assert ctx.stat("last_path") == (3 if fast_ok else 1)
        So, Ro = o.ess_rhat(x, kind=kind, split_chains=split, maxlag=maxlag)
        assert close(S, So, RTOL64), (maxlag, S, So)
        assert close(R, Ro, RTOL64)
    Rr = mcd.rhat(x, kind=kind, split_chains=split)
    assert ctx.stat("last_path") == (3 if fast_ok else 1)
    assert close(Rr, o.rhat(x, kind=kind, split_chains=split), RTOL64)
    if kind in ("bulk", "basic"):
        assert close(mcd.ess(x, kind=kind, split_chains=split), o.ess(x, kind=kind, split_chains=split), RTOL64)


@pytest.mark.parametrize("phi", [-0.9, -0.3, 0.0, 0.9, 0.99])
def test_fast_phi(mcd, o, phi):
    x = o.ar1(phi, np.sqrt(1 - phi ** 2), 1000, 4, 16, rng=rng(52))
    S, R = mcd.ess_rhat(x)
    So, Ro = o.ess_rhat(x)
    assert close(S, So, RTOL64) and close(R, Ro, RTOL64)
    S, R = mcd.ess_rhat(x, relative=True)
    assert close(S, So / 4000, RTOL64)


def test_fast_matches_general_kernel(mcd):
    x = rng(53).standard_normal((1000, 4, 64))
    ctx = mcd.get_context(0)
    out = {}
    for path in (3, 1):
        ctx.set_option("force_path", path)
        out[path] = [mcd.ess_rhat(x, kind=k) for k in ("rank", "bulk", "basic")] + [(None, mcd.rhat(x, kind="tail"))]
    for (S3, R3), (S1, R1) in zip(out[3], out[1]):
        if S3 is not None:
            assert close(S3, S1, 1e-11)
        assert close(R3, R1, 1e-12)


def test_fast_fallback_slabs(mcd, o):
    r = rng(54)
    x = r.standard_normal((1000, 4, 12))
    x[3, 1, 1] = np.nan                      # NaN -> general kernel (ranks by index order)
    x[:, :, 2] = 7.5                         # constant
    x[:, :, 3] = r.integers(1, 6, (1000, 4))  # heavy ties -> bucket overflow -> general kernel
    x[5, 2, 4] = np.inf
    x[6, 3, 5] = -np.inf
    x[:, :, 6] = np.round(x[:, :, 6], 2)     # light ties stay on the fast path
    x[:, :, 7] *= 1e-300
    x[:, :, 8] *= 1e300
    x[:, :, 9] = stats.cauchy.ppf(stats.norm.cdf(x[:, :, 9]))
    x[:10, :, 10] = 1e6                      # outliers squeeze the rest into few buckets
    for kind in ("rank", "bulk", "basic"):
        S, R = mcd.ess_rhat(x, kind=kind)
        So, Ro = o.ess_rhat(x, kind=kind)
        assert close(S, So, RTOL64), (kind, S, So)
        assert close(R, Ro, RTOL64), (kind, R, Ro)
    assert close(mcd.rhat(x, kind="tail"), o.rhat(x, kind="tail"), RTOL64)


def test_fast_exact_anchors(mcd, o):
    xn = rng(55).standard_normal((1000, 4, 10))
    xc = stats.cauchy.ppf(stats.norm.cdf(xn))
    assert np.array_equal(mcd.ess(xn, kind="bulk"), mcd.ess(xc, kind="bulk"))
    xa = o.ar1(-0.9, np.sqrt(1 - 0.81), 100, 4, 500, rng=rng(56))
    S = mcd.ess(xa, kind="basic")
    assert S.max() == 400 * np.log10(400) and S.min() > 0
    x = rng(57).standard_normal((1000, 4, 6, 2))
    S, R = mcd.ess_rhat(x)
    for i in range(2):
        for j in range(6):
            s, r_ = mcd.ess_rhat(x[:, :, j, i])
            assert s == S[j, i] and r_ == R[j, i]
    S1, R1 = mcd.ess_rhat(np.ones((1000, 4, 3)))
    assert np.all(np.isnan(S1)) and np.all(np.isnan(R1))


def test_fast_float32(mcd, o):
    x = o.ar1(0.5, np.sqrt(0.75), 1000, 4, 32, rng=rng(58)).astype(np.float32)
    for kind in ("rank", "bulk", "basic"):
        S, R = mcd.ess_rhat(x, kind=kind)
        assert mcd.get_context(0).stat("last_path") == 3
        So, Ro = o.ess_rhat(x, kind=kind)
        assert S.dtype == np.float32 and close(S, So, RTOL32) and close(R, Ro, RTOL32)
    r = mcd.tiedrank(x)
    for p in range(3):
        assert np.array_equal(r[:, :, p].reshape(-1, order="F"), o.tiedrank(x[:, :, p].reshape(-1, order="F")))


def test_fast_many_params_statistics(mcd, o):
    """Parity reported as a fraction within tolerance (Geyer truncation is discontinuous)."""
    x = mcd.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, 3000, seed=7)
    S, R = mcd.ess_rhat(x)
    from oracle import ref_port as rp
    So, Ro = rp.ess_rhat(x.cpu().numpy())
    relS = np.abs(S.cpu().numpy() - So) / np.abs(So)
    relR = np.abs(R.cpu().numpy() - Ro) / np.abs(Ro)
    assert (relS < RTOL64).mean() == 1.0 and (relR < RTOL64).mean() == 1.0, (relS.max(), relR.max())


def test_fast_tail_and_estimators(mcd, o):
    """Tail ESS (quantile indicators), estimator ESS kinds and tail R-hat on the fast kernel."""
    ctx = mcd.get_context(0)
    x = o.ar1(0.6, np.sqrt(1 - 0.36), 1000, 4, 24, rng=rng(59)) * 2 + 1
    for split, shape in ((2, x), (1, x.reshape(500, 8, 24, order="F"))):
        S, R = mcd.ess_rhat(shape, kind="tail", split_chains=split)
        assert ctx.stat("last_path") == 3
        So, Ro = o.ess_rhat(shape, kind="tail", split_chains=split)
        assert close(S, So, RTOL64), (S, So)
        assert close(R, Ro, RTOL64)
        assert close(mcd.ess(shape, kind="tail", split_chains=split, tail_prob=0.3),
                     o.ess(shape, kind="tail", split_chains=split, tail_prob=0.3), RTOL64)
        assert ctx.stat("last_path") == 3
        for km, ko in (("median", "median"), ("std", "std"), ("mad", "mad"), ("mean", "mean"),
                       (mcd.Quantile(0.25), o.Quantile(0.25)), (mcd.Quantile(0.999), o.Quantile(0.999)),
                       (mcd.Quantile(0.0), o.Quantile(0.0)), (mcd.Quantile(1.0), o.Quantile(1.0))):
            for maxlag in (250, 3):
                got = mcd.ess(shape, kind=km, split_chains=split, maxlag=maxlag)
                assert ctx.stat("last_path") == 3, km
                assert close(got, o.ess(shape, kind=ko, split_chains=split, maxlag=maxlag), RTOL64), (km, maxlag)
    # odd n (median is an element), ties, Float32 quantile arithmetic
    xo = o.ar1(0.3, np.sqrt(1 - 0.09), 333, 8, 10, rng=rng(60))
    xo[:, :, 3] = np.round(xo[:, :, 3], 1)
    for km, ko in (("median", "median"), ("mad", "mad"), (mcd.Quantile(0.1), o.Quantile(0.1))):
        assert close(mcd.ess(xo, kind=km, split_chains=1), o.ess(xo, kind=ko, split_chains=1), RTOL64), km
    S, R = mcd.ess_rhat(xo, kind="tail", split_chains=1)
    So, Ro = o.ess_rhat(xo, kind="tail", split_chains=1)
    assert close(S, So, RTOL64) and close(R, Ro, RTOL64)
    x32 = x.astype(np.float32)
    S, R = mcd.ess_rhat(x32, kind="tail")
    So, Ro = o.ess_rhat(x32, kind="tail")
    assert close(S, So, RTOL32) and close(R, Ro, RTOL32)
    for km, ko in (("median", "median"), ("mad", "mad"), (mcd.Quantile(np.float32(0.3)), o.Quantile(np.float32(0.3))), (mcd.Quantile(0.3), o.Quantile(0.3))):
        assert close(mcd.ess(x32, kind=km), o.ess(x32, kind=ko), RTOL32), km
    # NaN data: the quantile error still surfaces (general kernel handles the declined slab)
    xn = x.copy(); xn[3, 1, 2] = np.nan
    with pytest.raises(mcd.ArgumentError):
        mcd.ess(xn, kind="tail")
    assert close(mcd.ess(xn, kind="median"), o.ess(xn, kind="median"), RTOL64)


def test_fast_mcse_mean_std(mcd, o):
    ctx = mcd.get_context(0)
    x = o.ar1(0.4, np.sqrt(1 - 0.16), 1000, 4, 16, rng=rng(61)) * 3 - 2
    for kind in ("mean", "std"):
        got = mcd.mcse(x, kind=kind)
        assert ctx.stat("last_path") == 3
        assert close(got, o.mcse(x, kind=kind), RTOL64), kind
        got32 = mcd.mcse(x.astype(np.float32), kind=kind)
        assert close(got32, o.mcse(x.astype(np.float32), kind=kind), RTOL32), kind
    assert np.all(np.isnan(mcd.mcse(np.ones((1000, 4, 3)), kind="mean")))
    assert np.all(np.isnan(mcd.mcse(np.ones((1000, 4, 3)), kind="std")))


@pytest.mark.parametrize("shape,split", [((1000, 1), 1), ((4000, 1), 1), ((1000, 1), 2), ((2000, 2), 1), ((1000, 2), 2),
                                          ((1000, 4), 1), ((500, 8), 2), ((256, 16), 1), ((250, 8), 2), ((128, 16), 2),
                                          ((64, 32), 1), ((20, 16), 2), ((1024, 4), 1), ((333, 2), 1)])
def test_other_chain_counts(mcd, o, shape, split):
    """Other split-chain counts (1, 2, 4, 16, 32; whichever kernel the dispatcher picks)."""
    x = o.ar1(0.6, np.sqrt(1 - 0.36), shape[0], shape[1], 9, rng=rng(62)) + 0.3
    x[:, :, 4] = np.round(x[:, :, 4], 1)
    for kind in ("rank", "bulk", "tail", "basic"):
        for maxlag in (250, 5):
            S, R = mcd.ess_rhat(x, kind=kind, split_chains=split, maxlag=maxlag)
            So, Ro = o.ess_rhat(x, kind=kind, split_chains=split, maxlag=maxlag)
            assert close(S, So, RTOL64), (shape, split, kind, maxlag, S, So)
            assert close(R, Ro, RTOL64), (shape, split, kind, R, Ro)
        assert close(mcd.rhat(x, kind=kind, split_chains=split), o.rhat(x, kind=kind, split_chains=split), RTOL64)
    for km, ko in (("median", "median"), ("std", "std"), ("mad", "mad"), (mcd.Quantile(0.8), o.Quantile(0.8))):
        assert close(mcd.ess(x, kind=km, split_chains=split), o.ess(x, kind=ko, split_chains=split), RTOL64), (shape, km)
    assert close(mcd.mcse(x, kind="mean", split_chains=split), o.mcse(x, kind="mean", split_chains=split), RTOL64)
    x32 = x.astype(np.float32)
    S, R = mcd.ess_rhat(x32, split_chains=split)
    So, Ro = o.ess_rhat(x32, split_chains=split)
    assert close(S, So, 3 * RTOL32) and close(R, Ro, RTOL32)


@pytest.mark.parametrize("shape,split", [((1000, 2), 2), ((1000, 1), 2), ((1000, 3), 2), ((512, 7), 1), ((300, 1), 1), ((999, 2), 3),
                                          ((64, 3), 2), ((10, 1), 2), ((1000, 5), 1)])
@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
def test_fewer_than_eight_split_chains_on_the_staged_kernel(mcd, o, shape, split, kind):
    """1..7 split chains of <= 512 draws run on the TMA-staged kernel too (warps without a chain idle) whenever a slab
    is a whole number of 16-byte units; results against the oracle and against the general kernel (which sums the
    lagged products in another order: 1e-10)."""
    x = o.ar1(0.6, 0.8, shape[0], shape[1], 9, rng=rng(58))
    x[:3, 0, 4] = x[3:6, 0, 4]                                    # a few ties
    ctx = mcd.get_context(0)
    n = shape[0] * shape[1]
    fast_ok = shape[0] % split == 0 and shape[0] // split <= 512 and (n * 8) % 16 == 0
    for maxlag in (250, 3):
        S, R = mcd.ess_rhat(x, kind=kind, split_chains=split, maxlag=maxlag)
        if kind != "tail":
            assert ctx.stat("last_path") == (3 if fast_ok else 1)
        So, Ro = o.ess_rhat(x, kind=kind, split_chains=split, maxlag=maxlag)
        assert close(S, So, RTOL64), (maxlag, S, So)
        assert close(R, Ro, RTOL64)
    Rr = mcd.rhat(x, kind=kind, split_chains=split)
    assert close(Rr, o.rhat(x, kind=kind, split_chains=split), RTOL64)
    if kind in ("rank", "bulk", "basic"):
        S3, R3 = mcd.ess_rhat(x, kind=kind, split_chains=split)
        ctx.set_option("force_path", 1)
        try:
            S1, R1 = mcd.ess_rhat(x, kind=kind, split_chains=split)
        finally:
            ctx.set_option("force_path", 0)
        assert close(S3, S1, 1e-10) and close(R3, R1, 1e-12)
    xf = x.astype(np.float32)
    if (n * 4) % 16 == 0:
        Sf, Rf = mcd.ess_rhat(xf, kind=kind, split_chains=split)
        Sof, Rof = o.ess_rhat(xf, kind=kind, split_chains=split)
        assert close(Sf, Sof, 1e-4) and close(Rf, Rof, 1e-4)
