"""In-package callers of the path: gewekediag and heideldiag, whose standard
errors come from the device `mcse(...; split_chains=1)`, against the oracle's restatement of
src/gewekediag.jl:19-35 and src/heideldiag.jl:16-71."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mcd():
    import mcmcdiag_b200 as m
    m.get_context(0)
    return m


@pytest.fixture(scope="module")
def o():
    from oracle import mcmcdiag_oracle
    return mcmcdiag_oracle


def series(o, seed, n=1000, P=6):
    x = o.ar1(0.5, 0.8, n, 1, P, rng=np.random.default_rng(seed))[:, 0, :]
    x[:, 1] += np.linspace(3, 0, n) ** 2          # strong initial transient
    x[:, 2] += np.linspace(1, 0, n)               # mild drift
    x[:, 3] += 50.0                               # large mean: half-width test passes
    x[: n // 12, 4] += 2.0                        # short transient: converges after discarding some draws
    return x


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-8), (np.float32, 2e-4)])
def test_gewekediag_vector_and_batched(mcd, o, dtype, rtol):
    x = series(o, 11).astype(dtype)
    batched = mcd.gewekediag(x)
    assert batched.zscore.dtype == dtype and batched.zscore.shape == (6,)
    for j in range(x.shape[1]):
        want = o.gewekediag(x[:, j])
        got = mcd.gewekediag(x[:, j])
        assert isinstance(got.zscore, dtype)
        for g in (got, mcd.GewekeResult(batched.zscore[j], batched.pvalue[j])):
            assert np.isclose(g.zscore, want["zscore"], rtol=rtol, atol=1e-12)
            assert np.isclose(g.pvalue, want["pvalue"], rtol=50 * rtol, atol=1e-12)
    got = mcd.gewekediag(x[:, 0], first=0.2, last=0.3, autocov_method=mcd.BDAAutocovMethod(), maxlag=30)
    want = o.gewekediag(x[:, 0], first=0.2, last=0.3, autocov_method=o.BDAAutocovMethod(), maxlag=30)
    assert np.isclose(got.zscore, want["zscore"], rtol=rtol)


def test_gewekediag_exceptions(mcd):
    x = np.random.default_rng(0).standard_normal(100)
    for v in (-0.3, 0, 1, 1.2):
        with pytest.raises(mcd.ArgumentError):
            mcd.gewekediag(x, first=v)
        with pytest.raises(mcd.ArgumentError):
            mcd.gewekediag(x, last=v)
    with pytest.raises(mcd.ArgumentError):
        mcd.gewekediag(x, first=0.6, last=0.5)


@pytest.mark.parametrize("dtype,rtol", [(np.float64, 1e-8), (np.float32, 5e-4)])
def test_heideldiag_vector_and_batched(mcd, o, dtype, rtol):
    x = series(o, 12).astype(dtype)
    batched = mcd.heideldiag(x)
    seen = set()
    for j in range(x.shape[1]):
        want = o.heideldiag(x[:, j])
        got = mcd.heideldiag(x[:, j])
        seen.add((want["burnin"], want["stationarity"]))
        for g in (got, mcd.HeidelResult(*(f[j] for f in batched))):
            assert g.burnin == want["burnin"] and bool(g.stationarity) == want["stationarity"]
            assert bool(g.test) == want["test"]
            assert np.isclose(g.pvalue, want["pvalue"], rtol=rtol, atol=1e-7 if dtype == np.float64 else 1e-3)
            assert np.isclose(g.mean, want["mean"], rtol=rtol, atol=1e-12)
            assert np.isclose(g.halfwidth, want["halfwidth"], rtol=rtol)
    assert len(seen) >= 3                         # converged at once, later, and never: every branch ran
    got = mcd.heideldiag(x[:, 3], alpha=0.1, eps=0.05, start=11)
    want = o.heideldiag(x[:, 3], alpha=0.1, eps=0.05, start=11)
    assert got.burnin == want["burnin"] and bool(got.test) == want["test"] and want["test"]


def test_callers_accept_device_tensors(mcd, o):
    import torch
    x = series(o, 13)
    g_h, g_d = mcd.gewekediag(x), mcd.gewekediag(torch.as_tensor(x, device="cuda"))
    assert np.array_equal(g_h.zscore, g_d.zscore)
    h_h, h_d = mcd.heideldiag(x), mcd.heideldiag(torch.as_tensor(x, device="cuda"))
    assert np.array_equal(h_h.burnin, h_d.burnin) and np.array_equal(h_h.halfwidth, h_d.halfwidth)


BFMI_ENERGY = [42, 44, 45, 46, 42, 43, 36, 36, 31, 36, 36, 32, 36, 31, 31, 29, 29, 30, 25, 26, 29, 29, 27, 30, 31, 29]


def test_bfmi_golden_values_and_parity(mcd, o):
    """test/bfmi.jl: 0.6 by hand, 0.2406937229 from ArviZ; matrix form, dims=2; Float32; device tensors."""
    import torch
    assert np.isclose(mcd.bfmi([1, 2, 3, 4]), 0.6)
    assert np.isclose(mcd.bfmi(BFMI_ENERGY), 0.2406937229, rtol=1e-9)
    multi = np.repeat(np.asarray(BFMI_ENERGY, dtype=float)[:, None], 4, axis=1)
    assert np.allclose(mcd.bfmi(multi), 0.2406937229, rtol=1e-9)
    assert np.allclose(mcd.bfmi(multi), mcd.bfmi(multi.T, dims=2))
    e = np.random.default_rng(3).standard_normal((1000, 7)).cumsum(axis=0) * 0.1 + np.random.default_rng(4).standard_normal((1000, 7))
    assert np.allclose(mcd.bfmi(e), o.bfmi(e), rtol=1e-10)
    assert np.allclose(mcd.bfmi(e.astype(np.float32)), o.bfmi(e.astype(np.float32)), rtol=1e-4)
    assert np.allclose(mcd.bfmi(torch.as_tensor(e, device="cuda")).cpu().numpy(), o.bfmi(e), rtol=1e-10)


def test_chain_moments_and_gelmandiag(mcd, o):
    rng = np.random.default_rng(21)
    x = o.ar1(0.4, 0.9, 301, 5, 9, rng=rng)
    x[:, 0, 0] += 2.0
    x[:, :, 3] = 1.5                                      # constant parameter: NaN, as in the reference
    for split in (1, 2, 3):
        m, v = mcd.chain_moments(x, split_chains=split)
        sp = np.stack([o.copyto_split(x[:, :, p], split) for p in range(x.shape[2])], axis=2)   # (niter, nch, P)
        assert m.shape == (5 * split, 9)
        assert np.allclose(m, sp.mean(axis=0), rtol=1e-12, atol=1e-15)
        assert np.allclose(v, sp.var(axis=0, ddof=1), rtol=1e-11, atol=1e-30)
    got, want = mcd.gelmandiag(x), o.gelmandiag(x)
    assert got.psrf.dtype == np.float64 and got.psrf.shape == (9,)
    assert np.allclose(got.psrf, want["psrf"], rtol=1e-9, equal_nan=True)
    assert np.allclose(got.psrfci, want["psrfci"], rtol=1e-8, equal_nan=True)
    assert got.psrf[0] > 1.2
    got2, want2 = mcd.gelmandiag(x, alpha=0.2), o.gelmandiag(x, alpha=0.2)
    assert np.allclose(got2.psrfci, want2["psrfci"], rtol=1e-8, equal_nan=True)
    with pytest.raises(RuntimeError):
        mcd.gelmandiag(x[:, :1, :])
    ctx = mcd.get_context(0)
    ctx.set_option("h2d_chunk_bytes", 3 * 301 * 5 * 8)    # several staged chunks
    try:
        m2, v2 = mcd.chain_moments(x, split_chains=2)
    finally:
        ctx.set_option("h2d_chunk_bytes", 256 << 20)
    m1, v1 = mcd.chain_moments(x, split_chains=2)
    assert np.array_equal(m1, m2) and np.array_equal(v1, v2)
