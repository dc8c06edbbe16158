"""Multi-GPU behind the C ABI: one `mcd_create_multi` context shards the parameter axis of a HOST array over its devices
(one host thread + one staging pipeline per device) and writes every device's results into the caller's single
output buffer.  Results must be bit-identical to the single-device context whatever the device count (parameters are
independent; the parallel axis of src/ess_rhat.jl:517).  With one visible GPU the group has one member (the code
path is the same); with two or more the test uses them all (run it with `gpurun --gpus 2`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import mcmcdiag_b200 as m
    from oracle import mcmcdiag_oracle as o
    ndev = torch.cuda.device_count()
    grp = m.Context(devices=list(range(ndev)))
    yield m, o, grp, ndev
    grp.close()


def test_group_reports_its_devices(env):
    m, o, grp, ndev = env
    assert grp.stat("ndev") == ndev
    assert m.get_context(0).stat("ndev") == 1


@pytest.mark.parametrize("P", [1, 7, 1003])
def test_group_matches_single_device_bitwise(env, P):
    m, o, grp, ndev = env
    rng = np.random.default_rng(P)
    x = o.ar1(0.6, np.sqrt(1 - 0.36), 400, 4, P, rng=rng)
    x[:, :, 0] = 2.5                      # a constant parameter (NaN outputs)
    S0, R0 = m.ess_rhat(x)
    S1, R1 = m.ess_rhat(x, ctx=grp)
    assert np.array_equal(S0, S1, equal_nan=True) and np.array_equal(R0, R1, equal_nan=True)
    for kind in ("bulk", "tail", "basic"):
        a, b = m.ess(x, kind=kind), m.ess(x, kind=kind, ctx=grp)
        assert np.array_equal(a, b, equal_nan=True), kind
    assert np.array_equal(m.mcse(x, kind="median"), m.mcse(x, kind="median", ctx=grp), equal_nan=True)
    s0, s1 = m.summary(x), m.summary(x, ctx=grp)
    for k in s0:
        assert np.array_equal(s0[k], s1[k], equal_nan=True), k
    assert np.array_equal(m.tiedrank(x), m.tiedrank(x, ctx=grp))
    assert np.array_equal(m.rank_normalize(x), m.rank_normalize(x, ctx=grp), equal_nan=True)


def test_group_nested_rhat_and_float32(env):
    m, o, grp, ndev = env
    rng = np.random.default_rng(3)
    x = o.ar1(0.5, np.sqrt(0.75), 100, 16, 37, rng=rng).astype(np.float32)
    ids = np.repeat(np.arange(4), 4)
    assert np.array_equal(m.rhat_nested(x, ids), m.rhat_nested(x, ids, ctx=grp), equal_nan=True)
    assert np.array_equal(m.ess(x, kind="std", autocov_method=m.BDAAutocovMethod()),
                          m.ess(x, kind="std", autocov_method=m.BDAAutocovMethod(), ctx=grp), equal_nan=True)


def test_group_rejects_device_resident_input(env):
    import torch
    m, o, grp, ndev = env
    xd = torch.randn(5, 4, 100, dtype=torch.float64, device="cuda").permute(2, 1, 0)
    with pytest.raises(NotImplementedError):
        m.ess_rhat(xd, ctx=grp)


def test_group_errors_carry_the_device_message(env):
    m, o, grp, ndev = env
    x = np.random.default_rng(0).standard_normal((100, 4, 9))
    with pytest.raises(m.DomainError):
        m.ess_rhat(x, maxlag=0, ctx=grp)
