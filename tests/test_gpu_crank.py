"""The counting rank of the large-slab pipeline (csrc/mcd_crank.cuh) against the sort-based path it replaces:
every output bit-identical on the same device-resident input, and the statistics show which path ranked a chunk.
(The CPU emulation of the same kernels against SciPy is tests/test_crank_emul.py; parity of the large path with the
oracle is tests/test_gpu_large_path.py, which now runs on the counting rank where it applies.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SUP = np.repeat(np.arange(32), 64)


@pytest.fixture()
def env():
    import torch
    import mcmcdiag_b200 as m
    ctx = m.get_context(0)
    yield m, ctx, torch
    ctx.set_option("use_crank", 1)
    ctx.set_option("use_big", 1)
    ctx.set_option("force_path", 0)
    ctx.set_option("crank_factor", 4)
    ctx.set_option("workspace_bytes", 6 << 30)


def both(ctx, torch, fn):
    out, used = [], None
    for v in (0, 1):
        ctx.set_option("use_crank", v)
        c0, f0 = ctx.stat("crank_chunks"), ctx.stat("crank_fallbacks")
        r = fn()
        torch.cuda.synchronize()
        out.append([torch.as_tensor(t).clone() for t in (r if isinstance(r, tuple) else (r,))])
        used = (ctx.stat("crank_chunks") - c0, ctx.stat("crank_fallbacks") - f0)
    for a, b in zip(*out):
        assert bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
    return used


def ar1(m, d, c, p, dtype="float64", seed=3):
    return m.generate_ar1(0.5, np.sqrt(0.75), d, c, p, seed=seed, dtype=dtype)


@pytest.mark.parametrize("kind", ["rank", "bulk", "tail"])
def test_nested_many_short_chains(env, kind):
    m, ctx, torch = env
    x = ar1(m, 100, 2048, 12)
    used = both(ctx, torch, lambda: m.rhat_nested(x, SUP, kind=kind, split_chains=2))
    assert used[0] >= 1 and used[1] == 0


@pytest.mark.parametrize("method", ["AutocovMethod", "FFTAutocovMethod", "BDAAutocovMethod"])
def test_long_chains(env, method):
    m, ctx, torch = env
    x = ar1(m, 60000, 4, 3)
    used = both(ctx, torch, lambda: m.ess_rhat(x, kind="rank", autocov_method=getattr(m, method)()))
    assert used == (2, 0)


def test_float32_and_transforms(env):
    m, ctx, torch = env
    ctx.set_option("use_big", 0)
    ctx.set_option("force_path", 2)
    x = ar1(m, 4000, 8, 20, "float32")
    assert both(ctx, torch, lambda: m.ess_rhat(x, kind="rank")) == (2, 0)
    y = ar1(m, 30000, 2, 5)
    for fn in (m.tiedrank, m.rank_normalize, m.fold_around_median):
        assert both(ctx, torch, lambda: fn(y))[1] == 0


def test_flagged_slabs_take_the_sort_path(env):
    m, ctx, torch = env
    x = ar1(m, 30000, 2, 6).clone()
    x[5, 0, 1] = float("nan"); x[7, 1, 2] = float("inf")
    x[:, :, 3] = torch.round(x[:, :, 3] * 2); x[:, :, 4] = 1.5
    used = both(ctx, torch, lambda: m.ess_rhat(x, kind="rank"))
    assert used[0] == 0 and used[1] >= 1
    mild = torch.round(ar1(m, 30000, 2, 4) * 1000) / 1000      # about 30 values per distinct level: buckets overflow
    both(ctx, torch, lambda: m.ess_rhat(mild, kind="rank"))
    few = ar1(m, 30000, 2, 4).clone(); few[:40, 0, :] = few[40:80, 0, :]   # a few exact duplicates: resolved in the buckets
    assert both(ctx, torch, lambda: m.ess_rhat(few, kind="rank")) == (2, 0)


def test_chunks_and_bucket_factors(env):
    m, ctx, torch = env
    x = ar1(m, 100, 2048, 24)
    ctx.set_option("workspace_bytes", 64 << 20)
    used = both(ctx, torch, lambda: m.rhat_nested(x, SUP, kind="rank", split_chains=2))
    assert used[0] >= 4 and used[1] == 0
    ctx.set_option("workspace_bytes", 6 << 30)
    for factor in (1, 2, 8):
        ctx.set_option("crank_factor", factor)
        both(ctx, torch, lambda: m.rhat_nested(x, SUP, kind="rank", split_chains=2))
