"""Reference-output pin: when `tests/golden/julia_reference_vectors.npz` exists -- the outputs of the
REAL MCMCDiagnosticTools.jl for `tests/golden/julia_inputs.npz`, dumped by `baseline/ref_run.jl` on a machine with
Julia -- the CPU oracle (not gpu) and the CUDA path (gpu) are checked against it with the north star's tolerances.
There is no Julia in this image or on the GPU box, so until someone runs the script the tests below skip and parity
stays "pinned on the reference's identities, unpinned on its outputs" (DESIGN.md §4)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = os.path.join(GOLD, "julia_reference_vectors.npz")
INP = os.path.join(GOLD, "julia_inputs.npz")
EST = {"mean": "mean", "median": "median", "std": "std", "mad": "mad"}


def tol_for(x):
    return 1e-4 if x.dtype == np.float32 else 1e-8


def check_all(impl, methods, quantile):
    ref, inp = np.load(REF), np.load(INP)
    checked = 0
    for key in ref.files:
        parts = key.split(".")
        want = ref[key]
        if parts[0] == "ess_rhat":
            _, tag, kind, var, *rest = parts
            x = inp["x_" + tag]
            if var.startswith("maxlag"):
                got = impl.ess_rhat(x, kind=kind, maxlag=int(var[6:]))[0]
            else:
                split, field = int(rest[0][1:]), rest[1]
                r = impl.ess_rhat(x, kind=kind, autocov_method=methods[var](), split_chains=split)
                got = r[0] if field == "ess" else r[1]
        elif parts[0] == "ess" and parts[-1] == "relative":
            x = inp["x_" + parts[1]]
            got = impl.ess(x, kind="bulk" if parts[2] == "rank" else parts[2], relative=True)
        elif parts[0] in ("ess", "mcse"):
            x = inp["x_" + parts[1]]
            kind = quantile(0.25) if parts[2] == "q25" else EST[parts[2]]
            got = getattr(impl, parts[0])(x, kind=kind)
        elif parts[0] == "rhat_nested":
            got = impl.rhat_nested(inp["nested_x"], inp["nested_ids"], kind=parts[1], split_chains=int(parts[2][1:]))
            x = inp["nested_x"]
        elif parts[0] == "ranknorm":
            x = inp["x_" + parts[1]]; got = impl.rank_normalize(x).reshape(-1, order="F")
        elif parts[0] == "fold":
            x = inp["x_" + parts[1]]; got = impl.fold_around_median(x).reshape(-1, order="F")
        else:
            continue
        got = np.asarray(got, dtype=np.float64).reshape(-1)
        assert got.shape == want.reshape(-1).shape, key
        assert np.allclose(got, want.reshape(-1), rtol=tol_for(x), atol=0, equal_nan=True), key
        checked += 1
    assert checked > 100


@pytest.mark.skipif(not os.path.exists(REF), reason="no Julia reference vectors (run baseline/ref_run.jl with Julia)")
def test_oracle_matches_julia_reference_vectors():
    from oracle import mcmcdiag_oracle as o
    check_all(o, {"direct": o.AutocovMethod, "fft": o.FFTAutocovMethod, "bda": o.BDAAutocovMethod}, o.Quantile)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason="no Julia reference vectors (run baseline/ref_run.jl with Julia)")
def test_cuda_path_matches_julia_reference_vectors():
    import mcmcdiag_b200 as m
    check_all(m, {"direct": m.AutocovMethod, "fft": m.FFTAutocovMethod, "bda": m.BDAAutocovMethod}, m.Quantile)


def test_julia_inputs_are_committed_and_reproducible():
    """The inputs the Julia script reads are committed and are what the generator writes."""
    import subprocess, sys, tempfile, shutil
    inp = np.load(INP)
    assert {"x_ar1_f64", "x_sticky_f64", "x_iid_f32", "x_ties_f64", "x_skew_f64", "nested_x", "nested_ids"} <= set(inp.files)
    assert inp["x_iid_f32"].dtype == np.float32 and inp["x_ar1_f64"].shape == (1000, 4, 6)
    assert os.path.exists(os.path.join(os.path.dirname(GOLD), "..", "baseline", "ref_run.jl"))
