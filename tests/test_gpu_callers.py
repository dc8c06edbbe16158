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
