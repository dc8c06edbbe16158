"""The folded ranks of the headline kernel (csrc/mcd_rk2.cuh, step 4) stated in NumPy and checked against SciPy.

`_fold_around_median` (src/utils.jl:148-158) ranks |x - median(x)|.  The kernel never sorts that series: with S the
sorted slab, the folded values of the slots below the median decrease towards it and those above increase, so the
folded series in rank order is the MERGE of two sorted runs of S (run A = slots [0, L) read downwards, run B = slots
[L, n)), taken A-first among equals, on the COMPUTED values fl(|S[p] - med|); ties are runs of equal neighbours of the
merged sequence and get their average rank (StatsBase.tiedrank).  This test pins that statement — including ties that
only exist after the rounding of x - med, exact mirror pairs and duplicated draws — on the CPU; the kernel itself is
checked against the oracle on the GPU (tests/test_gpu_fast_path.py, tests/test_gpu_parity.py)."""
import numpy as np
import pytest
from scipy.stats import rankdata


def folded_ranks_by_merge(x):
    n = x.size
    order = np.argsort(x, kind="stable")
    S = x[order]
    med = S[n // 2] if n % 2 else S[n // 2 - 1] / x.dtype.type(2) + S[n // 2] / x.dtype.type(2)   # Statistics.median
    f = np.abs(S - med)                                  # computed folded value per sorted slot
    L = n // 2
    pa, pb = L - 1, L                                    # next slot of run A (downwards) and of run B (upwards)
    merged_slot = np.empty(n, dtype=np.int64)
    for i in range(n):
        take_a = pa >= 0 and (pb >= n or f[pa] <= f[pb])
        if take_a:
            merged_slot[i] = pa; pa -= 1
        else:
            merged_slot[i] = pb; pb += 1
    fm = f[merged_slot]
    assert np.all(np.diff(fm) >= 0), "the merge of the two runs must be sorted"
    ranks_sorted = np.empty(n)
    i = 0
    while i < n:                                         # runs of equal neighbours -> average rank
        j = i
        while j + 1 < n and fm[j + 1] == fm[i]:
            j += 1
        ranks_sorted[merged_slot[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    ranks = np.empty(n)
    ranks[order] = ranks_sorted
    return ranks, med


def check(x):
    r, med = folded_ranks_by_merge(x)
    assert np.array_equal(r, rankdata(np.abs(x - med), method="average"))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [2, 3, 10, 999, 4000])
def test_continuous(dtype, n):
    rng = np.random.default_rng(n)
    check(rng.standard_normal(n).astype(dtype))
    check(np.exp(2 * rng.standard_normal(n)).astype(dtype))          # skewed: the runs have very different lengths


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_ties(dtype):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(4000).astype(dtype)
    y = x.copy(); y[:50] = y[50:100]; check(y)                       # duplicated draws (ties inside a run)
    m = np.median(x)
    y = x.copy(); y[:30] = (2 * m - y[30:60]).astype(dtype); check(y)  # mirror images about the median (ties across the runs)
    check(np.round(x * 8) / 8)                                       # heavy ties
    check(np.round(x * 300) / 300)                                   # mild ties
    check(np.full(64, 1.5, dtype=dtype))                             # constant
    z = np.zeros(100, dtype=dtype); z[::2] = -0.0; z[:7] = x[:7]; check(z)


def test_ties_created_by_rounding():
    """x - med loses the low bits of values that are tiny next to the median: distinct draws, equal folded values."""
    rng = np.random.default_rng(2)
    x = np.concatenate([1.0 + rng.standard_normal(2001) * 1e-3, rng.standard_normal(2000) * 1e-20])
    r, med = folded_ranks_by_merge(x)
    f = np.abs(x - med)
    assert np.unique(x).size == x.size and np.unique(f).size < x.size   # the ties exist only after the subtraction
    assert np.array_equal(r, rankdata(f, method="average"))
    y = (1.0 + np.arange(4000) * 2.0 ** -52).astype(np.float64)         # neighbours one ulp apart around the median
    check(y)
