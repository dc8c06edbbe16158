"""`mcse` for estimators without an ESS rule: the reference's subsampling-bootstrap fallback `_mcse_sbm`
(src/mcse.jl:120-148) is host logic in the reference and stays host logic here (no GPU needed)."""
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def m():
    import mcmcdiag_b200 as mod
    return mod


def sbm_literal(f, v, b):
    n = len(v)
    vals = np.array([f(v[i:i + b]) for i in range(n - b + 1)])
    return math.sqrt(vals.var() * b / n)


def test_sbm_matches_the_literal_formula(m):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((200, 3, 2))
    f = lambda w, axis=None: np.max(w, axis=axis) - np.min(w, axis=axis)   # no ESS rule: an arbitrary callable
    got = m.mcse(x, kind=f)
    b = int(math.floor(math.sqrt(600)))
    for p in range(2):
        v = x[:, :, p].reshape(-1, order="F")
        assert got[p] == pytest.approx(sbm_literal(lambda w: w.max() - w.min(), v, b), rel=1e-12)
    got7 = m.mcse(x, kind=f, batch_size=7)
    assert got7[0] == pytest.approx(sbm_literal(lambda w: w.max() - w.min(), x[:, :, 0].reshape(-1, order="F"), 7), rel=1e-12)


def test_sbm_mad_constant_scalar_and_missing(m):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((400, 4, 3))
    x[:, :, 1] = 1.5
    r = m.mcse(x, kind="mad")
    assert np.isfinite(r[0]) and np.isnan(r[1]) and np.isfinite(r[2])
    assert 0.2 < r[0] / (1.4826 * 0.6745 * math.sqrt(math.pi / 2) / math.sqrt(1600) * 1.0) < 5      # right order of magnitude
    assert isinstance(m.mcse(x[:, :, 0], kind="mad"), np.floating)                                    # _maybescalar
    xm = np.ma.masked_array(x.copy())
    xm[3, 0, 2] = np.ma.masked
    rm = m.mcse(xm, kind="mad")
    assert rm.mask.tolist() == [False, False, True] and rm[0] == r[0]
    with pytest.raises(TypeError):
        m.mcse(x, kind="mad", maxlag=3)                                                               # not a keyword of _mcse_sbm


def test_mcse_relative_is_rejected_not_ignored(m):
    with pytest.raises(NotImplementedError):
        m.mcse(np.zeros((10, 2, 1)), kind="mean", relative=True)
