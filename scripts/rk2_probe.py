"""Developer probe: the TMA-staged headline kernel (rk2) against the round-1 register-resident kernel on the
same device-resident input: agreement of every output, then CUDA-event timings of both.
python scripts/rk2_probe.py [P] [phi]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mcmcdiag_b200 as m

P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
phi = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
print(torch.cuda.get_device_name(0), "P =", P, "phi =", phi, flush=True)
ctx = m.get_context(0)


def both(fn):
    out = []
    for v in (0, 1):
        ctx.set_option("use_rk2", v)
        r = fn()
        torch.cuda.synchronize()
        out.append([t.clone() for t in (r if isinstance(r, tuple) else (r,))])
    return out


def cmp(name, fn):
    o, n = both(fn)
    worst = 0.0
    for a, b in zip(o, n):
        a = a.double().cpu().numpy(); b = b.double().cpu().numpy()
        same_nan = np.array_equal(np.isnan(a), np.isnan(b))
        ok = ~np.isnan(a)
        rel = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(a[ok]), 1e-300)
        worst = max(worst, rel.max() if rel.size else 0.0)
        if not same_nan:
            worst = np.inf
    print(f"{name:40s} max rel diff old vs rk2 = {worst:.3e}  redo={ctx.stat('redo_count')}", flush=True)
    return worst


def timeit(name, fn, reps=3):
    res = []
    for v in (0, 1):
        ctx.set_option("use_rk2", v)
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res.append(min(ts))
    f = [P * 32016 / t / 1e6 / 6548.2 for t in res]
    print(f"{name:28s} old {res[0]:8.3f} ms ({f[0]:.3f})   rk2 {res[1]:8.3f} ms ({f[1]:.3f})   x{res[0] / res[1]:.2f}", flush=True)


# ---- parity on assorted inputs -------------------------------------------------------------------
for dt in ("float64", "float32"):
    for (d, c, ph) in ((1000, 4, 0.5), (1000, 4, 0.95), (998, 4, 0.0), (600, 4, 0.5), (1024, 4, 0.3), (100, 4, 0.5), (22, 4, 0.2)):
        xs = m.generate_ar1(ph, np.sqrt(1 - ph * ph), d, c, 3000, seed=7, dtype=dt)
        for kind in ("rank", "bulk", "basic"):
            cmp(f"{dt} {d}x{c} phi={ph} ess_rhat {kind}", lambda: m.ess_rhat(xs, kind=kind))
        cmp(f"{dt} {d}x{c} phi={ph} rhat tail", lambda: m.rhat(xs, kind="tail"))
        cmp(f"{dt} {d}x{c} phi={ph} rhat rank", lambda: m.rhat(xs, kind="rank"))
    # ties / discrete / skewed / constants
    g = torch.Generator(device="cuda").manual_seed(3)
    tdt = torch.float64 if dt == "float64" else torch.float32
    base = m.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, 2000, seed=11, dtype=dt)
    disc = torch.round(base * 3.0)                     # heavy ties (about 20 distinct values)
    mild = torch.round(base * 300.0) / 300.0           # mild ties
    skew = torch.exp(base * 1.5)                       # lognormal: crowded buckets near zero
    const = torch.ones_like(base); const[:5] = base[:5]
    off6 = base + 1.0e6                                # large offset: few leading-word steps between min and max
    off9 = base * 1.0e-3 + 1.0e9                       # offset >> spread (Float32: constant after rounding)
    tiny = base * 1.0e-300 if dt == "float64" else base * 1.0e-30
    zeros = torch.zeros_like(base); zeros[::2] = -0.0; zeros[:, :, :5] = base[:, :, :5]
    naninf = base.clone(); naninf[3, 1, 0] = float("nan"); naninf[5, 2, 1] = float("inf"); naninf[7, 0, 2] = float("-inf")
    neg = -torch.abs(base) - 2.0                       # all negative
    cases = (("disc", disc), ("mild", mild), ("skew", skew), ("const", const), ("off6", off6), ("off9", off9),
             ("tiny", tiny), ("zeros", zeros), ("naninf", naninf), ("neg", neg))
    for nm, xs in cases:
        for kind in ("rank", "bulk"):
            cmp(f"{dt} {nm} ess_rhat {kind}", lambda: m.ess_rhat(xs, kind=kind))
        cmp(f"{dt} {nm} rhat tail", lambda: m.rhat(xs, kind="tail"))

# ---- timing ---------------------------------------------------------------------------------------
x = m.generate_ar1(phi, np.sqrt(1 - phi * phi), 1000, 4, P, seed=1)
torch.cuda.synchronize()
cmp("headline ess_rhat rank", lambda: m.ess_rhat(x))
timeit("ess_rhat rank", lambda: m.ess_rhat(x))
timeit("rhat rank", lambda: m.rhat(x))
timeit("ess_rhat bulk", lambda: m.ess_rhat(x, kind="bulk"))
timeit("ess_rhat basic", lambda: m.ess_rhat(x, kind="basic"))
timeit("rhat basic", lambda: m.rhat(x, kind="basic"))
timeit("rhat tail", lambda: m.rhat(x, kind="tail"))
for mult in (2, 3):
    ctx.set_option("fast_grid_mult", mult)
    timeit(f"ess_rhat rank grid x{mult}", lambda: m.ess_rhat(x))
ctx.set_option("fast_grid_mult", 0)
