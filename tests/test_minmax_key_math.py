"""The order-preserving integer keys of the headline kernel's min / max (csrc/mcd_rk2.cuh, RkKey) stated in NumPy.

The kernel bounds a slab's range from the minimum / maximum KEY of the leading 32 bits of each value (for Float64: sign,
exponent and 20 mantissa bits) instead of FP64 compares.  What the bucket map needs, and what this test pins:
key(a) < key(b) implies a < b; lower(min key) <= every value <= upper(max key); NaN and +-Inf have keys beyond the
finite range (those slabs go to the redo list); for Float32 the key is exact."""
import numpy as np
import pytest

M = np.int32(0x7FFFFFFF)


def key64(x):
    h = (np.asarray(x, dtype=np.float64).view(np.int64) >> 32).astype(np.int32)
    return h ^ ((h >> 31) & M)


def unkey(k):
    return k ^ ((k >> 31) & M)


def lower64(k):
    h = unkey(np.int32(k)).astype(np.int64)
    lo = np.int64(0xFFFFFFFF) if h < 0 else np.int64(0)
    return np.array([(h << 32) | lo], dtype=np.int64).view(np.float64)[0]


def upper64(k):
    h = unkey(np.int32(k)).astype(np.int64)
    lo = np.int64(0) if h < 0 else np.int64(0xFFFFFFFF)
    return np.array([(h << 32) | lo], dtype=np.int64).view(np.float64)[0]


def key32(x):
    h = np.asarray(x, dtype=np.float32).view(np.int32)
    return h ^ ((h >> 31) & M)


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-300, 0.0), (1e300, 0.0), (1e-3, 1e9), (1.0, -5.0), (1e-12, 1.0)])
def test_keys_order_and_bound_float64(scale, offset):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(20000) * scale + offset
    k = key64(x)
    o = np.argsort(k, kind="stable")
    ks, xs = k[o], x[o]
    strictly = ks[1:] > ks[:-1]
    assert np.all(xs[1:][strictly] > xs[:-1][strictly])            # key(a) < key(b)  =>  a < b
    lo, hi = lower64(k.min()), upper64(k.max())
    assert lo <= x.min() and x.max() <= hi and np.isfinite(lo) and np.isfinite(hi)
    # one step of the 20 leading mantissa bits on each side at most
    assert x.min() - lo <= abs(x.min()) * 2.0 ** -20 and hi - x.max() <= abs(x.max()) * 2.0 ** -20


def test_special_values():
    pos_inf, neg_inf = np.int32(0x7FF00000), np.int32(-2146435073)      # RkKey<double>::pos_inf / neg_inf (0x800fffff)
    assert key64(np.inf) == pos_inf and key64(-np.inf) == neg_inf
    assert key64(np.nan) >= pos_inf and key64(-np.nan) <= neg_inf
    assert key64(np.array([np.nan]).view(np.int64).__or__(1).view(np.float64))[0] >= pos_inf
    finite = np.array([np.finfo(np.float64).max, -np.finfo(np.float64).max, 0.0, -0.0, 5e-324, -5e-324])
    k = key64(finite)
    assert np.all(k < pos_inf) and np.all(k > neg_inf)
    assert key64(-0.0) < key64(0.0)                                       # (only bounds: -0.0 and 0.0 tie in the ranks)


def test_float32_keys_are_exact():
    rng = np.random.default_rng(2)
    x = (rng.standard_normal(20000) * 3).astype(np.float32)
    k = key32(x)
    assert unkey(k.min()).view(np.float32) == x.min() and unkey(k.max()).view(np.float32) == x.max()
    o = np.argsort(k, kind="stable")
    assert np.all(np.diff(x[o]) >= 0)
