"""Writes tests/golden/julia_inputs.npz: the small seeded arrays baseline/ref_run.jl feeds to the REAL
MCMCDiagnosticTools.jl to produce tests/golden/julia_reference_vectors.npz (see tests/test_julia_reference_vectors.py).
Deterministic: python tests/golden/make_julia_inputs.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mcmcdiag_oracle as o  # noqa: E402  (only its AR(1) generator: test/helpers.jl:4-12)

rng = np.random.default_rng(20261017)
arrs = {
    "x_ar1_f64": o.ar1(0.5, np.sqrt(0.75), 1000, 4, 6, rng=rng),                       # BASELINE configs[0] shape
    "x_sticky_f64": o.ar1(0.95, np.sqrt(1 - 0.95 ** 2), 301, 4, 4, rng=rng),           # odd draws: the discard rule of copyto_split!
    "x_iid_f32": rng.standard_normal((400, 8, 5)).astype(np.float32),
    "x_ties_f64": rng.integers(1, 11, size=(200, 4, 3)).astype(np.float64),            # heavy ties
    "x_skew_f64": np.exp(1.5 * o.ar1(0.3, np.sqrt(1 - 0.09), 500, 4, 3, rng=rng)),
}
arrs["x_ar1_f64"][:, :, 5] = 3.25                                                      # a constant parameter: NaN outputs
nested = o.ar1(0.5, np.sqrt(0.75), 100, 64, 5, rng=rng) + rng.standard_normal((1, 64, 1)) * 0.3
arrs["nested_x"] = nested
arrs["nested_ids"] = np.repeat(np.arange(1, 9), 8).astype(np.int64)
np.savez_compressed(os.path.join(HERE, "julia_inputs.npz"), **arrs)
print({k: v.shape for k, v in arrs.items()})
