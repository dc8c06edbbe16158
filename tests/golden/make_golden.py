"""Generates tests/golden/*.npz: seeded inputs and the outputs of the NumPy oracle
(oracle/mcmcdiag_oracle.py) for every public call on the hot path.

The reference (MCMCDiagnosticTools.jl) is Julia and cannot be run in this image, so these are
ORACLE-derived regression vectors, not reference outputs: they freeze today's oracle (itself
pinned on the reference's known-answer tests, tests/test_oracle_anchors.py) so that the C++ port,
the CUDA path and future edits of the oracle are all checked against the same numbers.

    python tests/golden/make_golden.py        # rewrites the fixtures
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mcmcdiag_oracle as o  # noqa: E402


def case_inputs():
    r = np.random.default_rng(20261017)
    x = o.ar1(0.5, np.sqrt(0.75), 1000, 4, 6, rng=r)            # C1-like
    x[:, :, 1] = np.round(x[:, :, 1], 1)                         # ties
    x[:, :, 2] = r.standard_cauchy((1000, 4))                    # heavy tails
    y = o.ar1(-0.4, np.sqrt(1 - 0.16), 301, 3, 4, rng=r) * 5 + 2  # odd draws: discard rule
    z = o.ar1(0.7, np.sqrt(1 - 0.49), 600, 8, 3, rng=r).astype(np.float32)
    n = r.standard_normal((60, 16, 3)) + r.standard_normal((1, 16, 1)) * 0.2
    return x, y, z, n


def main():
    x, y, z, n = case_inputs()
    out = {"x": x, "y": y, "z": z, "n": n}
    methods = {"direct": o.AutocovMethod(), "fft": o.FFTAutocovMethod(), "bda": o.BDAAutocovMethod()}
    for name, arr, splits in (("x", x, (2,)), ("y", y, (1, 2, 3)), ("z", z, (2,))):
        for kind in ("rank", "bulk", "tail", "basic"):
            for split in splits:
                for mname, m in methods.items():
                    S, R = o.ess_rhat(arr, kind=kind, split_chains=split, autocov_method=m)
                    out[f"{name}.ess_rhat.{kind}.s{split}.{mname}.ess"] = S
                    out[f"{name}.ess_rhat.{kind}.s{split}.{mname}.rhat"] = R
        for est, k in (("mean", "mean"), ("median", "median"), ("std", "std"), ("mad", "mad"), ("q25", o.Quantile(0.25))):
            out[f"{name}.ess.{est}"] = o.ess(arr, kind=k)
            if est != "mad":
                out[f"{name}.mcse.{est}"] = o.mcse(arr, kind=k)
        out[f"{name}.ess.maxlag7.relative"] = o.ess(arr, kind="bulk", maxlag=7, relative=True)
        out[f"{name}.tiedrank"] = np.stack([o.tiedrank(arr[:, :, p].reshape(-1, order="F")) for p in range(arr.shape[2])])
    ids = [3, 1, 2, 0, 1, 3, 0, 2, 2, 2, 0, 0, 1, 1, 3, 3]
    out["n.ids"] = np.asarray(ids)
    for kind in ("rank", "bulk", "tail", "basic"):
        for split in (1, 2):
            out[f"n.rhat_nested.{kind}.s{split}"] = o.rhat_nested(n, ids, kind=kind, split_chains=split)
    np.savez_compressed(os.path.join(HERE, "hot_path_oracle_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
