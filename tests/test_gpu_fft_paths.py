"""FFTAutocovMethod (src/ess_rhat.jl:103-152,181-195) on its three device implementations — the shared-memory slab
kernel, one CTA per parameter for chains that fit shared memory, the four-step transform through global memory — in
the paired / summed data flow (two real chains per complex transform, one inverse per parameter, transform length
nextprod(niter + maxlag)) against the reference's per-chain recipe (fft_pair = 0, fft_full = 1), the direct method
and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def env():
    import mcmcdiag_b200 as m
    from oracle import mcmcdiag_oracle as o
    ctx = m.get_context(0)
    yield m, o, ctx
    for k, v in (("fft_pair", 1), ("fft_full", 0), ("fft_tc", 0), ("force_path", 0)):
        ctx.set_option(k, v)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(a)
    return float(np.max(np.abs(a[ok] - b[ok]) / np.abs(b[ok]))) if ok.any() else 0.0


# (draws, chains, params, split): slab kernel / per-parameter kernel / four-step (even and odd chain counts)
SHAPES = [(1000, 4, 8, 2), (301, 3, 5, 1), (6000, 4, 3, 2), (7001, 3, 2, 1), (9001, 3, 2, 1), (30000, 1, 2, 2), (40000, 5, 1, 1)]


@pytest.mark.parametrize("shape", SHAPES)
def test_paired_flow_matches_per_chain_flow_and_oracle(env, shape):
    m, o, ctx = env
    d, c, p, split = shape
    x = o.ar1(0.7, np.sqrt(1 - 0.49), d, c, p, rng=np.random.default_rng(d + c))
    call = lambda: m.ess_rhat(x, kind="bulk", autocov_method=m.FFTAutocovMethod(), split_chains=split)
    S1, R1 = call()
    ctx.set_option("fft_pair", 0); ctx.set_option("fft_full", 1)
    S0, R0 = call()
    assert rel(S1, S0) < 1e-12 and np.array_equal(R1, R0)
    So, Ro = o.ess_rhat(x, kind="bulk", autocov_method=o.FFTAutocovMethod(), split_chains=split)
    assert rel(S1, So) < 1e-8 and rel(R1, Ro) < 1e-8
    Sd, _ = m.ess_rhat(x, kind="bulk", split_chains=split)       # the direct method computes the same lags
    assert rel(S1, Sd) < 1e-9


def test_constant_chain_keeps_the_reference_nan(env):
    """A chain whose centred values are all exactly zero has c[0] = 0: the reference's c[k] / c[0] is NaN."""
    m, o, ctx = env
    x = o.ar1(0.5, np.sqrt(0.75), 400, 4, 6, rng=np.random.default_rng(4))
    x[:200, 1, 2] = 1.0                      # first split chain of chain 2, parameter 3: constant
    x[:, 3, 4] = -2.0                        # a whole chain of parameter 5
    for force in (0, 2):
        ctx.set_option("force_path", force)
        S, R = m.ess_rhat(x, kind="basic", autocov_method=m.FFTAutocovMethod())
        So, Ro = o.ess_rhat(x, kind="basic", autocov_method=o.FFTAutocovMethod())
        assert np.isnan(S[[2, 4]]).all() and np.isfinite(S[[0, 1, 3, 5]]).all()
        assert rel(S, So) < 1e-8 and rel(R, Ro) < 1e-8
