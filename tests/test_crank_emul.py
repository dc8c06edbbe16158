"""CPU emulation of the counting-rank kernels (csrc/mcd_crank.cuh) against SciPy's average ranks and NumPy's median.

The per-element bodies of the kernels are __host__ __device__ functions; tests/emul/crank_emul.cu replays the
launch sequence thread by thread on the host (nvcc builds it, nothing runs on a GPU).  This pins the algorithm
(bucket map, 4-bit packed counters with arrival offsets, scan, placement, exact ties, median selection, the
flags that send a slab to the sort path) where no GPU exists; the GPU tests then check the real launches
bit-for-bit against the sort-based path (tests/test_gpu_large_path.py)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
from scipy.stats import rankdata

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "crank_emul.cu")
OUT = os.path.join(HERE, "emul", "_build", "libcrank_emul.so")
HDR = os.path.join(HERE, "..", "mcmcdiagnostictools.jl_b200", "csrc", "mcd_crank.cuh")


@pytest.fixture(scope="module")
def lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC], check=True,
                       capture_output=True)
    return ctypes.CDLL(OUT)


def run(lib, x, factor=4):
    """x: (params, n) C-contiguous -> ranks (params, n), medians (params,), flags (params,)"""
    pc, n = x.shape
    fn = lib.crank_emul_f64 if x.dtype == np.float64 else lib.crank_emul_f32
    ranks = np.zeros((pc, n)); med = np.zeros(pc); flags = np.zeros(pc, dtype=np.int32)
    rc = fn(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(n), ctypes.c_longlong(pc), ctypes.c_int(factor),
            ranks.ctypes.data_as(ctypes.c_void_p), med.ctypes.data_as(ctypes.c_void_p), flags.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return ranks, med, flags


def check(lib, x, factor=4, expect_flag=None):
    ranks, med, flags = run(lib, np.ascontiguousarray(x), factor)
    for p in range(x.shape[0]):
        if expect_flag is not None:
            assert bool(flags[p]) == expect_flag, (p, flags[p])
        if flags[p]:
            continue
        assert np.array_equal(ranks[p], rankdata(x[p].astype(np.float64), method="average")), p
        xs = np.sort(x[p]); n = x.shape[1]
        m = xs[n // 2] if n % 2 else xs[n // 2 - 1] / x.dtype.type(2) + xs[n // 2] / x.dtype.type(2)
        assert med[p] == float(m), (p, med[p], float(m))
    return flags


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 7, 100, 4097, 20000])
def test_continuous(lib, dtype, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).astype(dtype)
    if n == 1:
        check(lib, x, expect_flag=True)      # a constant slab goes to the sort path
    else:
        check(lib, x, expect_flag=False)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_ties_and_signed_zero(lib, dtype):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((4, 5000)).astype(dtype)
    x[0] = np.round(x[0] * 300) / 300          # mild ties: resolved exactly inside the buckets
    x[1, :6] = 0.0; x[1, 6:12] = -0.0          # -0.0 == 0.0: one run of twelve tied values
    x[2, :2500] = x[2, 2500:]                  # every value twice
    x[3] = np.abs(x[3])                        # folded shape: minimum at the edge of the range
    flags = check(lib, x)
    assert not flags.any()


def test_heavy_ties_are_flagged_or_exact(lib):
    rng = np.random.default_rng(6)
    x = np.round(rng.standard_normal((2, 6000)) * 3.0)     # ~ 20 distinct values: >= 15 in a bucket
    flags = check(lib, x)
    assert flags.all()


def test_nan_inf_constant_are_flagged(lib):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((4, 3000))
    x[0, 17] = np.nan; x[1, 5] = np.inf; x[2, 9] = -np.inf; x[3] = 2.5
    _, _, flags = run(lib, x)
    assert flags.tolist() == [1, 1, 1, 1]


@pytest.mark.parametrize("case", ["offset", "tiny", "huge", "skew", "negative"])
def test_ranges(lib, case):
    rng = np.random.default_rng(8)
    x = rng.standard_normal((2, 8000))
    x = {"offset": x + 1e9, "tiny": x * 1e-300, "huge": x * 1e300, "skew": np.exp(3 * x), "negative": -np.abs(x) - 5}[case]
    check(lib, x)


def test_c4_shaped_slab_and_bucket_factors(lib):
    rng = np.random.default_rng(9)
    x = rng.standard_normal((1, 204800))
    for factor in (1, 2, 4, 8):
        flags = check(lib, x, factor)
        assert not flags.any()
    check(lib, np.abs(x - np.median(x)), 4)      # the folded series of the same slab


def test_ar1_float32_c5_shape(lib):
    rng = np.random.default_rng(10)
    e = rng.standard_normal((2, 32000))
    x = np.zeros_like(e)
    for t in range(1, e.shape[1]):
        x[:, t] = 0.5 * x[:, t - 1] + e[:, t]
    check(lib, x.astype(np.float32))
