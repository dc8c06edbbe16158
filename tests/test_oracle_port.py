"""The C++/OpenMP oracle port (oracle/ref_port.cpp, also the timed CPU baseline) must agree with
the NumPy oracle to <= 1e-12 — two independent restatements of the reference."""
import numpy as np
import pytest
from scipy import special, stats

from oracle import mcmcdiag_oracle as o
from oracle import ref_port as rp


def rng(s):
    return np.random.default_rng(s)


def test_norminvcdf_matches_scipy():
    ps = np.concatenate([np.logspace(-300, -1, 200), np.linspace(0.01, 0.99, 500), 1 - np.logspace(-15, -1, 100)])
    got = np.array([rp.norminvcdf(p) for p in ps])
    want = special.ndtri(ps)
    assert np.allclose(got, want, rtol=4e-15, atol=1e-16)


@pytest.mark.parametrize("case", ["normal", "ties", "nan", "zeros"])
def test_tiedrank_port(case):
    r = rng(41)
    v = r.standard_normal(3000)
    if case == "ties":
        v = r.integers(0, 7, 3000).astype(float)
    elif case == "nan":
        v[r.random(3000) < 0.02] = np.nan
    elif case == "zeros":
        v = np.where(r.random(3000) < 0.5, 0.0, -0.0)
    assert np.array_equal(rp.tiedrank(v), o.tiedrank(v))


@pytest.mark.parametrize("kind", ["rank", "bulk", "tail", "basic"])
@pytest.mark.parametrize("method", ["direct", "bda"])
@pytest.mark.parametrize("split_chains", [1, 2, 3])
def test_port_matches_numpy_oracle(kind, method, split_chains):
    x = o.ar1(0.6, np.sqrt(1 - 0.36), 501, 4, 6, rng=rng(42))
    om = o.AutocovMethod() if method == "direct" else o.BDAAutocovMethod()
    for maxlag in (250, 7, 1):
        S, R = rp.ess_rhat(x, kind=kind, method=method, split_chains=split_chains, maxlag=maxlag, nthreads=2)
        So, Ro = o.ess_rhat(x, kind=kind, autocov_method=om, split_chains=split_chains, maxlag=maxlag)
        assert np.allclose(S, So, rtol=1e-12, atol=0, equal_nan=True)
        assert np.allclose(R, Ro, rtol=1e-12, atol=0, equal_nan=True)


def test_port_estimators():
    x = o.ar1(0.3, np.sqrt(1 - 0.09), 1000, 4, 5, rng=rng(43)) * 2 + 3
    for name, kind in [("mean", "mean"), ("median", "median"), ("std", "std"), ("mad", "mad")]:
        assert np.allclose(rp.ess_estimator(x, name), o.ess(x, kind=kind), rtol=1e-12)
    assert np.allclose(rp.ess_estimator(x, "quantile", p=0.25), o.ess(x, kind=o.Quantile(0.25)), rtol=1e-12)


def test_port_anchors():
    xa = o.ar1(-0.9, np.sqrt(1 - 0.81), 100, 4, 200, rng=rng(44))
    S, _ = rp.ess_rhat(xa, kind="basic")
    assert S.max() == 400 * np.log10(400) and S.min() > 0
    S, R = rp.ess_rhat(np.ones((100, 4, 3)), kind="rank")
    assert np.all(np.isnan(S)) and np.all(np.isnan(R))
    xn = rng(45).standard_normal((1000, 4, 5))
    xc = stats.cauchy.ppf(stats.norm.cdf(xn))
    assert np.array_equal(rp.ess_rhat(xn, kind="bulk")[0], rp.ess_rhat(xc, kind="bulk")[0])
