"""Fused per-parameter summary (`mcd_summary`): every column equals the separate
reference call it stands for (oracle), on host and device inputs, on each kernel path."""
import numpy as np
import pytest

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


def check(got, want, rtol, names=None):
    assert list(got) == list(want)
    for k in want:
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(want[k], dtype=np.float64)
        assert a.shape == b.shape, k
        assert np.allclose(a, b, rtol=rtol, atol=1e-300 if rtol < 1e-6 else 1e-6, equal_nan=True), (k, a, b)


@pytest.mark.parametrize("shape", [(1000, 4, 12), (301, 3, 5), (100, 16, 4), (64, 1, 3)])
def test_summary_matches_separate_calls_f64(mcd, o, shape):
    x = o.ar1(0.6, 0.8, *shape, rng=np.random.default_rng(5))
    x[..., 0] += np.linspace(0, 2, shape[0])[:, None]          # a drifting parameter
    got = mcd.summary(x)
    check(got, o.summary(x), RTOL64)
    assert list(got) == list(mcd.SUMMARY_FIELDS)
    # and the library's own separate calls give the same bits
    assert np.array_equal(got["ess_bulk"], mcd.ess(x, kind="bulk"), equal_nan=True)
    assert np.array_equal(got["ess_tail"], mcd.ess(x, kind="tail"), equal_nan=True)
    assert np.array_equal(got["rhat"], mcd.rhat(x), equal_nan=True)
    assert np.array_equal(got["mcse_mean"], mcd.mcse(x, kind="mean"), equal_nan=True)
    assert np.array_equal(got["mcse_std"], mcd.mcse(x, kind="std"), equal_nan=True)


def test_summary_f32_and_keywords(mcd, o):
    x = o.ar1(0.4, 0.9, 400, 4, 9, rng=np.random.default_rng(6)).astype(np.float32)
    kw = dict(split_chains=3, maxlag=20, tail_prob=0.2)
    got = mcd.summary(x, autocov_method=mcd.BDAAutocovMethod(), **kw)
    want = o.summary(x, autocov_method=o.BDAAutocovMethod(), **kw)
    assert all(v.dtype == np.float32 for v in got.values())
    check(got, want, RTOL32)


def test_summary_field_subset_and_order(mcd, o):
    x = o.ar1(0.5, 0.8, 200, 4, 7, rng=np.random.default_rng(7))
    got = mcd.summary(x, fields=("rhat", "mean", "ess_tail"))
    assert list(got) == ["mean", "ess_tail", "rhat"]
    check(got, o.summary(x, fields=("mean", "ess_tail", "rhat")), RTOL64)
    only_bulk = mcd.summary(x, fields=["ess_bulk"])
    assert np.array_equal(only_bulk["ess_bulk"], mcd.ess(x, kind="bulk"))
    with pytest.raises(mcd.ArgumentError):
        mcd.summary(x, fields=("mean", "nope"))
    with pytest.raises(mcd.DomainError):
        mcd.summary(x, maxlag=0)


def test_summary_device_tensor_param_axes_and_single_staging(mcd, o):
    import torch
    x = o.ar1(0.5, 0.8, 1000, 4, 6, rng=np.random.default_rng(8)).reshape(1000, 4, 2, 3, order="F")
    want = o.summary(x)
    got_d = mcd.summary(torch.as_tensor(x, device="cuda"))
    assert all(v.is_cuda and tuple(v.shape) == (2, 3) for v in got_d.values())
    check({k: v.cpu().numpy() for k, v in got_d.items()}, want, RTOL64)
    ctx = mcd.get_context(0)
    before = ctx.stat("h2d_bytes")
    got_h = mcd.summary(x)
    assert ctx.stat("h2d_bytes") - before == x.nbytes        # the host array crossed PCIe once for 7 columns
    check(got_h, want, RTOL64)
    ctx.set_option("h2d_chunk_bytes", 2 * 4000 * 8)          # several staged chunks
    try:
        check(mcd.summary(x), want, RTOL64)
    finally:
        ctx.set_option("h2d_chunk_bytes", 256 << 20)


def test_summary_edge_cases(mcd, o):
    rng = np.random.default_rng(9)
    x = rng.standard_normal((100, 4, 6))
    x[:, :, 1] = 3.0                                           # constant
    x[:, :, 2] = rng.integers(1, 10, size=(100, 4))            # heavy ties
    x[5, 2, 3] = np.nan                                        # NaN: tail quantile throws in the reference
    ok = np.delete(x, 3, axis=2)
    check(mcd.summary(ok), o.summary(ok), RTOL64)
    with pytest.raises(mcd.ArgumentError):
        mcd.summary(x)
    nz = mcd.summary(x, fields=("mean", "std", "ess_bulk", "rhat"))
    check(nz, o.summary(x, fields=("mean", "std", "ess_bulk", "rhat")), RTOL64)
    tiny = rng.standard_normal((8, 2, 3))                      # niter <= 4: ESS columns NaN + warning
    with pytest.warns(UserWarning):
        got = mcd.summary(tiny)
    want = o.summary(tiny)
    check(got, want, RTOL64)
    m = np.ma.masked_invalid(x)
    gm = mcd.summary(m, fields=("mean", "rhat"))
    assert gm["mean"].mask[3] and not gm["mean"].mask[0]


def test_summary_large_path(mcd, o):
    x = o.ar1(0.3, 0.95, 20000, 2, 2, rng=np.random.default_rng(10))
    check(mcd.summary(x, autocov_method=mcd.FFTAutocovMethod()), o.summary(x, autocov_method=o.FFTAutocovMethod()), RTOL64)
