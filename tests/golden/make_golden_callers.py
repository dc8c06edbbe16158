"""Generates tests/golden/callers_oracle_vectors.npz: ORACLE-derived regression vectors (see make_golden.py for
why they are not reference outputs) for the entry points added from SURVEY.md §8(f): the fused summary,
gewekediag / heideldiag, bfmi, gelmandiag.

    python tests/golden/make_golden_callers.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mcmcdiag_oracle as o  # noqa: E402


def main():
    main_gold = np.load(os.path.join(HERE, "hot_path_oracle_vectors.npz"))
    out = {}
    for name in ("x", "y", "z"):                                   # same inputs as the main fixture
        for k, v in o.summary(main_gold[name]).items():
            out[f"{name}.summary.{k}"] = v
    r = np.random.default_rng(20261018)
    s = o.ar1(0.5, 0.8, 800, 1, 5, rng=r)[:, 0, :]
    s[:, 1] += np.linspace(3, 0, 800) ** 2
    s[:80, 2] += 2.0
    s[:, 3] += 40.0
    out["s"] = s
    for j in range(s.shape[1]):
        g = o.gewekediag(s[:, j])
        h = o.heideldiag(s[:, j])
        out[f"s.geweke.{j}"] = np.array([g["zscore"], g["pvalue"]])
        out[f"s.heidel.{j}"] = np.array([h["burnin"], float(h["stationarity"]), h["pvalue"], h["mean"], h["halfwidth"], float(h["test"])])
    e = r.standard_normal((500, 6)).cumsum(axis=0) * 0.05 + r.standard_normal((500, 6))
    out["e"] = e
    out["e.bfmi"] = o.bfmi(e)
    gd = o.gelmandiag(main_gold["x"])
    out["x.gelman.psrf"], out["x.gelman.psrfci"] = gd["psrf"], gd["psrfci"]
    np.savez_compressed(os.path.join(HERE, "callers_oracle_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
