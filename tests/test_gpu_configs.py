"""BASELINE.json configs 2-5 at their exact per-parameter shapes: the CUDA path against the CPU restatement of the
reference algorithm on identical device-generated inputs, reported as the FRACTION of parameters within the north
star's tolerance (1e-8 Float64, 1e-4 Float32) with the outliers listed (Geyer's truncation is
discontinuous, so a Float32 rounding difference can flip one lag pair of a rare parameter).

The workloads are bench.py's own `CONFIGS` (same call, same generator, same seed); the parameter count is what the
oracle finishes in seconds.  `scripts/parity_configs.py` runs the same comparison on as many parameters as asked.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import bench
    import mcmcdiag_b200 as m
    from oracle import build_oracle
    build_oracle.build()
    m.get_context(0)
    return bench, m


def compare(bench, m, name, nparams, oracle_only=False):
    """Returns (fraction within tolerance, worst relative difference, outliers [(param, column, gpu, cpu)])."""
    import torch
    cfg = bench.CONFIGS[name]
    x = m.generate_ar1(bench.PHI, np.sqrt(1 - bench.PHI ** 2), cfg.draws, cfg.chains, nparams, seed=1, dtype=cfg.dtype)
    res = cfg.run(m, x)
    torch.cuda.synchronize()
    xs = np.asfortranarray(x.cpu().numpy())
    if oracle_only:
        cres = run_oracle(name, xs)
    else:
        cres, _, _ = cfg.run_cpu(xs, bench.host_threads())
    bad = np.zeros(nparams, dtype=bool)
    worst, outliers = 0.0, []
    for col, (g, c) in enumerate(zip(res, cres)):
        if c is None:
            continue
        g = g.double().cpu().numpy(); c = np.asarray(c, dtype=np.float64)
        with np.errstate(all="ignore"):
            rel = np.abs(g - c) / np.abs(c)
        rel[np.isnan(g) & np.isnan(c)] = 0.0
        rel[np.isnan(rel)] = np.inf
        worst = max(worst, float(rel.max()))
        for p in np.flatnonzero(rel > cfg.tol):
            outliers.append((int(p), col, float(g[p]), float(c[p]), float(rel[p])))
        bad |= rel > cfg.tol
    return 1.0 - bad.mean(), worst, outliers


def run_oracle(name, xs):
    """dtype-faithful NumPy oracle (Float32 arithmetic for Float32 input)."""
    from oracle import mcmcdiag_oracle as o
    if name == "c5bda":
        bda = o.BDAAutocovMethod()
        return (o.ess(xs, kind="median", autocov_method=bda), o.ess(xs, kind="std", autocov_method=bda))
    if name == "c2summary":
        r = o.summary(xs)
        return tuple(r[k] for k in ("mean", "std", "mcse_mean", "mcse_std", "ess_bulk", "ess_tail", "rhat"))
    raise ValueError(name)


def test_c2_rank_fraction_within_tolerance(env):
    bench, m = env
    frac, worst, out = compare(bench, m, "c2rank", 3000)
    print(f"C2 ess_rhat(kind=:rank) 1000x4 f64: {frac:.6f} of 3000 parameters within 1e-8, worst {worst:.2e}, outliers {out}")
    assert frac == 1.0, out


def test_c2_summary_fraction_within_tolerance(env):
    bench, m = env
    frac, worst, out = compare(bench, m, "c2summary", 300, oracle_only=True)
    print(f"C2 summary (7 columns) 1000x4 f64: {frac:.6f} of 300 parameters within 1e-8, worst {worst:.2e}, outliers {out}")
    assert frac == 1.0, out


def test_c3_fft_full_length_chains(env):
    bench, m = env
    frac, worst, out = compare(bench, m, "c3fft", 3)
    print(f"C3 ess(kind=:bulk, FFT) 1e6x4 f64: {frac:.6f} of 3 parameters within 1e-8, worst {worst:.2e}, outliers {out}")
    assert frac == 1.0, out


def test_c4_nested_fraction_within_tolerance(env):
    bench, m = env
    frac, worst, out = compare(bench, m, "c4nested", 40)
    print(f"C4 rhat_nested(kind=:rank) 100x2048 f64: {frac:.6f} of 40 parameters within 1e-8, worst {worst:.2e}, outliers {out}")
    assert frac == 1.0, out


def test_c5_bda_float32_fraction_within_tolerance(env):
    bench, m = env
    n = 1000
    frac, worst, out = compare(bench, m, "c5bda", n, oracle_only=True)
    print(f"C5 ess(median)+ess(std) BDA 4000x8 f32: {frac:.6f} of {n} parameters within 1e-4, worst {worst:.2e}, outliers {out}")
    # Float32: a summation-order difference of ~1e-7 can flip one Geyer lag pair (`delta > 0`) of a rare parameter;
    # everything else must be inside the tolerance and the flips must stay rare
    assert frac >= 0.995, out
    assert all(rel < 0.2 for *_, rel in out), out
