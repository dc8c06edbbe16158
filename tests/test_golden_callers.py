"""Committed ORACLE-derived fixtures for the SURVEY §8(f) entry points (tests/golden/callers_oracle_vectors.npz,
made by tests/golden/make_golden_callers.py).  One checker runs against two backends: the oracle on the CPU
(it must still reproduce its own frozen numbers) and the CUDA path on the GPU."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "callers_oracle_vectors.npz"))
MAIN = np.load(os.path.join(HERE, "golden", "hot_path_oracle_vectors.npz"))


def _fields(res, names):
    """dict (oracle) or namedtuple (product) -> list of values in the order of `names`"""
    return [res[k] if isinstance(res, dict) else getattr(res, k) for k in names]


def check(b, rtol64, rtol32):
    def close(a, g, rt):
        return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(g, dtype=np.float64), rtol=rt, atol=1e-300, equal_nan=True)

    for name in ("x", "y", "z"):
        rt = rtol32 if name == "z" else rtol64
        got = b.summary(MAIN[name])
        for k in ("mean", "std", "mcse_mean", "mcse_std", "ess_bulk", "ess_tail", "rhat"):
            assert close(got[k], GOLD[f"{name}.summary.{k}"], rt), (name, k)
    s = GOLD["s"]
    for j in range(s.shape[1]):
        z, p = _fields(b.gewekediag(s[:, j]), ("zscore", "pvalue"))
        assert close([z], GOLD[f"s.geweke.{j}"][:1], rtol64) and np.isclose(p, GOLD[f"s.geweke.{j}"][1], rtol=100 * rtol64, atol=1e-12)
        h = _fields(b.heideldiag(s[:, j]), ("burnin", "stationarity", "pvalue", "mean", "halfwidth", "test"))
        want = GOLD[f"s.heidel.{j}"]
        assert int(h[0]) == int(want[0]) and bool(h[1]) == bool(want[1]) and bool(h[5]) == bool(want[5])
        assert np.isclose(h[2], want[2], rtol=rtol64, atol=1e-7) and close([h[3], h[4]], want[3:5], rtol64)
    assert close(b.bfmi(GOLD["e"]), GOLD["e.bfmi"], max(rtol64, 1e-10))
    psrf, psrfci = _fields(b.gelmandiag(MAIN["x"]), ("psrf", "psrfci"))
    assert close(psrf, GOLD["x.gelman.psrf"], 1e-8) and close(psrfci, GOLD["x.gelman.psrfci"], 1e-8)


def test_oracle_reproduces_callers_golden():
    from oracle import mcmcdiag_oracle as o
    check(o, 1e-13, 1e-6)


@pytest.mark.gpu
def test_gpu_matches_callers_golden():
    import mcmcdiag_b200 as m
    check(m, 1e-8, 1e-3)
