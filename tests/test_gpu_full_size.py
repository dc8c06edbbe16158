"""BASELINE.json full size (1000 draws x 4 chains x 1e6 parameters, Float64, 32 GB in HBM):
size-independent properties of the CUDA path, plus a random-subset check against the C++ oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 1_000_000


@pytest.fixture(scope="module")
def big():
    import torch
    import mcmcdiag_b200 as m
    free, _ = torch.cuda.mem_get_info()
    if free < 80e9:
        pytest.skip("needs ~70 GB of free device memory")
    x = m.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, P, seed=1)
    S, R = m.ess_rhat(x)
    torch.cuda.synchronize()
    yield m, x, S, R
    del x


def test_full_size_deterministic_and_shard_invariant(big):
    import torch
    m, x, S, R = big
    S2, R2 = m.ess_rhat(x)
    assert torch.equal(S, S2) and torch.equal(R, R2)                     # run-to-run bit identical
    cuts = [0, 137, 500_000, 812_345, P]
    for lo, hi in zip(cuts, cuts[1:]):                                   # any sharding gives the same bits
        Ss, Rs = m.ess_rhat(x[:, :, lo:hi])
        assert torch.equal(Ss, S[lo:hi]) and torch.equal(Rs, R[lo:hi])
    assert bool(torch.isfinite(S).all()) and bool(torch.isfinite(R).all())
    assert float(S.min()) > 100 and float(R.max()) < 1.1
    # AR(1) phi = 0.5: ESS ~ N (1 - phi) / (1 + phi)
    assert abs(float(S.mean()) / (4000 / 3) - 1) < 0.05


def test_full_size_matches_oracle_on_random_subset(big):
    from oracle import ref_port as rp
    m, x, S, R = big
    idx = np.sort(np.random.default_rng(5).choice(P, 400, replace=False))
    xs = np.stack([x[:, :, int(i)].cpu().numpy() for i in idx], axis=2)
    So, Ro = rp.ess_rhat(xs, kind="rank")
    Sg, Rg = S.cpu().numpy()[idx], R.cpu().numpy()[idx]
    relS, relR = np.abs(Sg - So) / So, np.abs(Rg - Ro) / Ro
    assert (relS < 1e-8).all() and (relR < 1e-8).all(), (relS.max(), relR.max())


def test_full_size_rank_invariances(big):
    import torch
    m, x, S, R = big
    # monotone map leaves ranks, hence bulk ESS / R-hat, untouched (test/ess_rhat.jl:329-335), bit for bit
    # (the lognormal image overflows some fine buckets, so a few slabs take the general-kernel redo path:
    #  the two kernels sum in the same order, so even those agree to the last bit)
    sub = x[:, :, :200_000]
    Sb, Rb = m.ess_rhat(sub, kind="bulk")
    y = torch.exp(sub.permute(2, 1, 0).contiguous()).permute(2, 1, 0)
    Sb2, Rb2 = m.ess_rhat(y, kind="bulk")
    assert torch.equal(Sb, Sb2) and torch.equal(Rb, Rb2)
    # kind = :rank is (bulk ESS, max(bulk R-hat, tail R-hat))
    Rt = m.rhat(sub, kind="tail")
    assert torch.equal(S[:200_000], Sb) and torch.equal(R[:200_000], torch.maximum(Rb, Rt))
    # relative ESS
    Srel, _ = m.ess_rhat(sub, relative=True)
    assert torch.allclose(Srel * 4000, S[:200_000], rtol=1e-14)


def test_full_size_sentinels(big):
    import torch
    m, x, S, R = big
    sub = x[:, :, 300_000:300_064].clone()
    buf = sub.permute(2, 1, 0)                      # (params, chains, draws) contiguous view
    buf[3] = 2.5                                    # constant parameter -> NaN, NaN
    buf[7, 1, 17] = float("nan")                    # NaN -> general-kernel redo path
    buf[11] = torch.round(buf[11])                  # heavy ties
    Ss, Rs = m.ess_rhat(sub)
    keep = [i for i in range(64) if i not in (3, 7, 11)]
    assert torch.equal(Ss[keep], S[300_000:300_064][keep]) and torch.equal(Rs[keep], R[300_000:300_064][keep])
    assert bool(torch.isnan(Ss[3])) and bool(torch.isnan(Rs[3]))
    from oracle import mcmcdiag_oracle as o
    xs = sub[:, :, [7, 11]].cpu().numpy()
    So, Ro = o.ess_rhat(xs)
    assert np.allclose(Ss[[7, 11]].cpu().numpy(), So, rtol=1e-8) and np.allclose(Rs[[7, 11]].cpu().numpy(), Ro, rtol=1e-8)


def test_full_size_summary_equals_separate_calls(big):
    """The fused seven-column summary over all 1e6 parameters gives the bits of the separate calls
    (ESS / R-hat / MCSE columns) and the rank-kind results the fixture computed."""
    import torch
    m, x, S, R = big
    out = m.summary(x)
    assert torch.equal(out["ess_bulk"], S) and torch.equal(out["rhat"], R)
    assert torch.equal(out["ess_tail"], m.ess(x, kind="tail"))
    assert torch.equal(out["mcse_mean"], m.mcse(x, kind="mean"))
    assert torch.equal(out["mcse_std"], m.mcse(x, kind="std"))
    # mean / std against torch on a slice (different summation order: tolerance, not bits)
    sl = x[:, :, :50_000].reshape(4000, -1) if x.shape[0] == 4000 else x[:, :, :50_000].reshape(-1, 50_000)
    assert torch.allclose(out["mean"][:50_000], sl.mean(dim=0), rtol=0, atol=1e-13)
    assert torch.allclose(out["std"][:50_000], sl.std(dim=0), rtol=1e-12, atol=0)
    ctx = m.get_context(0)
    assert ctx.stat("redo_count") == 0 and ctx.stat("last_path") == 3
