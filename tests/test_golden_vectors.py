"""Committed fixtures (tests/golden/hot_path_oracle_vectors.npz, made by tests/golden/make_golden.py).

They are ORACLE-derived (the Julia reference cannot run in this image): the CPU tests check that
today's oracle and the independent C++ port still reproduce them, the GPU tests check the CUDA
path against them — the same numbers on every side."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hot_path_oracle_vectors.npz"))
SPLITS = {"x": (2,), "y": (1, 2, 3), "z": (2,)}


def close(a, b, rtol):
    return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=0, equal_nan=True)


def rtol_for(name):
    return 1e-4 if name == "z" else 1e-8


# ---- CPU: the oracle is frozen, the C++ port agrees ---------------------------------------------------
def test_oracle_reproduces_golden():
    from oracle import mcmcdiag_oracle as o
    for name in ("x", "y", "z"):
        arr = GOLD[name]
        for kind in ("rank", "tail"):
            for split in SPLITS[name]:
                S, R = o.ess_rhat(arr, kind=kind, split_chains=split)
                assert np.array_equal(S, GOLD[f"{name}.ess_rhat.{kind}.s{split}.direct.ess"], equal_nan=True)
                assert np.array_equal(R, GOLD[f"{name}.ess_rhat.{kind}.s{split}.direct.rhat"], equal_nan=True)
    assert np.array_equal(o.rhat_nested(GOLD["n"], list(GOLD["n.ids"])), GOLD["n.rhat_nested.rank.s2"])


def test_cpp_port_matches_golden():
    from oracle import ref_port as rp
    for name in ("x", "y"):
        arr = GOLD[name]
        for kind in ("rank", "bulk", "tail", "basic"):
            for split in SPLITS[name]:
                for mname in ("direct", "bda"):
                    S, R = rp.ess_rhat(arr, kind=kind, method=mname, split_chains=split)
                    assert close(S, GOLD[f"{name}.ess_rhat.{kind}.s{split}.{mname}.ess"], 1e-12)
                    assert close(R, GOLD[f"{name}.ess_rhat.{kind}.s{split}.{mname}.rhat"], 1e-12)
        for p in range(arr.shape[2]):
            assert np.array_equal(rp.tiedrank(arr[:, :, p].reshape(-1, order="F")), GOLD[f"{name}.tiedrank"][p])


# ---- GPU: the CUDA path against the same vectors --------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["x", "y", "z"])
def test_gpu_matches_golden(name):
    import mcmcdiag_b200 as m
    arr = GOLD[name]
    rt = rtol_for(name)
    methods = {"direct": m.AutocovMethod(), "fft": m.FFTAutocovMethod(), "bda": m.BDAAutocovMethod()}
    for kind in ("rank", "bulk", "tail", "basic"):
        for split in SPLITS[name]:
            for mname, meth in methods.items():
                S, R = m.ess_rhat(arr, kind=kind, split_chains=split, autocov_method=meth)
                assert close(S, GOLD[f"{name}.ess_rhat.{kind}.s{split}.{mname}.ess"], rt), (kind, split, mname)
                assert close(R, GOLD[f"{name}.ess_rhat.{kind}.s{split}.{mname}.rhat"], rt), (kind, split, mname)
    for est, k in (("mean", "mean"), ("median", "median"), ("std", "std"), ("mad", "mad"), ("q25", m.Quantile(0.25))):
        assert close(m.ess(arr, kind=k), GOLD[f"{name}.ess.{est}"], rt), est
        if est != "mad":
            assert close(m.mcse(arr, kind=k), GOLD[f"{name}.mcse.{est}"], rt), est
    assert close(m.ess(arr, kind="bulk", maxlag=7, relative=True), GOLD[f"{name}.ess.maxlag7.relative"], rt)
    ranks = m.tiedrank(arr)
    for p in range(arr.shape[2]):
        assert np.array_equal(ranks[:, :, p].reshape(-1, order="F"), GOLD[f"{name}.tiedrank"][p])   # bit exact


@pytest.mark.gpu
def test_gpu_nested_matches_golden():
    import mcmcdiag_b200 as m
    n, ids = GOLD["n"], list(GOLD["n.ids"])
    for kind in ("rank", "bulk", "tail", "basic"):
        for split in (1, 2):
            assert close(m.rhat_nested(n, ids, kind=kind, split_chains=split), GOLD[f"n.rhat_nested.{kind}.s{split}"], 1e-8)
