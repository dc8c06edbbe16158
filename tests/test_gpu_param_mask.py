"""`missing` handling on the device: the host passes the array as it is plus a per-parameter skip
mask (`mcd_set_param_mask`); the library computes the runs of kept parameters in place (a skipped parameter's bytes are
never staged or read) and NaN-fills the skipped outputs.  The reference's rule: a parameter that contains `missing`
yields `missing` (src/ess_rhat.jl:382-385,519-523)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import mcmcdiag_b200 as m
    from oracle import mcmcdiag_oracle as o
    m.get_context(0)
    return m, o


def masked_case(o, P=41, seed=8):
    rng = np.random.default_rng(seed)
    x = o.ar1(0.5, np.sqrt(0.75), 300, 4, P, rng=rng)
    xm = np.ma.masked_array(x.copy())
    miss = np.zeros(P, dtype=bool)
    miss[[0, 5, 6, 7, 20, P - 1]] = True
    for p in np.flatnonzero(miss):
        xm[int(rng.integers(300)), int(rng.integers(4)), p] = np.ma.masked
    return x, xm, miss


def test_masked_parameters_are_skipped_and_the_rest_is_unchanged(env):
    m, o = env
    x, xm, miss = masked_case(o)
    ctx = m.get_context(0)
    h0 = ctx.stat("h2d_bytes")
    S, R = m.ess_rhat(xm)
    staged = ctx.stat("h2d_bytes") - h0
    assert staged == int((~miss).sum()) * 300 * 4 * 8          # skipped parameters never cross PCIe
    Sk, Rk = m.ess_rhat(x[:, :, ~miss])
    assert S.mask.tolist() == miss.tolist() and R.mask.tolist() == miss.tolist()
    assert np.array_equal(S.compressed(), Sk) and np.array_equal(R.compressed(), Rk)
    for kind in ("median", "std"):
        e = m.ess(xm, kind=kind)
        assert e.mask.tolist() == miss.tolist() and np.array_equal(e.compressed(), m.ess(x[:, :, ~miss], kind=kind))
    s = m.summary(xm)
    sk = m.summary(x[:, :, ~miss])
    for k in s:
        assert s[k].mask.tolist() == miss.tolist() and np.array_equal(s[k].compressed(), sk[k], equal_nan=True), k
    ids = np.repeat(np.arange(2), 2)
    rn = m.rhat_nested(xm, ids)
    assert rn.mask.tolist() == miss.tolist() and np.array_equal(rn.compressed(), m.rhat_nested(x[:, :, ~miss], ids))


def test_all_missing_and_scalar(env):
    m, o = env
    x = masked_case(o)[0][:, :, :3]
    xa = np.ma.masked_array(x.copy())
    xa[0, 0, :] = np.ma.masked
    S, R = m.ess_rhat(xa)
    assert S.mask.all() and R.mask.all()
    one = np.ma.masked_array(x[:, :, 0].copy())
    one[2, 1] = np.ma.masked
    assert m.rhat(one) is np.ma.masked


def test_mask_on_device_resident_input_and_on_a_group(env):
    import torch
    m, o = env
    x, xm, miss = masked_case(o, P=64, seed=9)
    ctx = m.get_context(0)
    xd = torch.from_numpy(np.ascontiguousarray(x.transpose(2, 1, 0))).cuda().permute(2, 1, 0)
    skip = np.ascontiguousarray(miss, dtype=np.uint8)
    ctx.check(ctx._lib.mcd_set_param_mask(ctx._h, skip.ctypes.data_as(C.POINTER(C.c_ubyte)), skip.size))
    S, R = m.ess_rhat(xd)
    S, R = S.cpu().numpy(), R.cpu().numpy()
    Sk, Rk = m.ess_rhat(x[:, :, ~miss])
    assert np.isnan(S[miss]).all() and np.isnan(R[miss]).all()
    assert np.array_equal(S[~miss], Sk) and np.array_equal(R[~miss], Rk)
    S2, _ = m.ess_rhat(xd)                                       # the mask is consumed by one call
    assert np.isfinite(S2.cpu().numpy()).all()
    grp = m.Context(devices=list(range(torch.cuda.device_count())))
    Sg, Rg = m.ess_rhat(xm, ctx=grp)
    assert Sg.mask.tolist() == miss.tolist() and np.array_equal(Sg.compressed(), Sk) and np.array_equal(Rg.compressed(), Rk)
    grp.close()
    with pytest.raises(m.ArgumentError):                         # a stale mask of the wrong length is an error, not ignored
        ctx.check(ctx._lib.mcd_set_param_mask(ctx._h, skip.ctypes.data_as(C.POINTER(C.c_ubyte)), 5))
        m.ess_rhat(x)
