"""Developer probe: the counting-rank path of the large-slab pipeline (mcd_crank.cuh) against the sort-based path
on the same device-resident input: bit-identity of every output, then CUDA-event timings for a few settings.
python scripts/crank_probe.py [P_c4]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mcmcdiag_b200 as m

ctx = m.get_context(0)
print(torch.cuda.get_device_name(0), flush=True)
sup = np.repeat(np.arange(32), 64)


def both(fn):
    out = []
    for v in (0, 1):
        ctx.set_option("use_crank", v)
        c0, f0 = ctx.stat("crank_chunks"), ctx.stat("crank_fallbacks")
        r = fn()
        torch.cuda.synchronize()
        out.append([torch.as_tensor(t).clone() for t in (r if isinstance(r, tuple) else (r,))])
        used = (ctx.stat("crank_chunks") - c0, ctx.stat("crank_fallbacks") - f0)
    return out[0], out[1], used


def cmp(name, fn):
    try:
        o, n, used = both(fn)
    except Exception as e:  # keep going: one failing case must not hide the others
        print(f"{name:58s} ERROR {type(e).__name__}: {e}", flush=True)
        return False
    same = all(bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all()) for a, b in zip(o, n))
    worst = 0.0
    for a, b in zip(o, n):
        a = a.double(); b = b.double(); ok = ~torch.isnan(a)
        if ok.any():
            worst = max(worst, float(((a[ok] - b[ok]).abs() / a[ok].abs().clamp_min(1e-300)).max()))
    print(f"{name:58s} identical={same} max rel diff={worst:.2e} crank chunks/fallbacks={used}", flush=True)
    return same


ok = True
g = lambda d, c, p, dt="float64", seed=3: m.generate_ar1(0.5, np.sqrt(0.75), d, c, p, seed=seed, dtype=dt)
x4 = g(100, 2048, 24)
ok &= cmp("C4 rhat_nested rank 100x2048", lambda: m.rhat_nested(x4, sup, kind="rank", split_chains=2))
ok &= cmp("C4 rhat_nested bulk", lambda: m.rhat_nested(x4, sup, kind="bulk", split_chains=2))
ok &= cmp("C4 rhat_nested tail", lambda: m.rhat_nested(x4, sup, kind="tail", split_chains=2))
ok &= cmp("C4 rhat rank (plain)", lambda: m.rhat(x4, kind="rank"))
x3 = g(100000, 4, 3)
ok &= cmp("C3-like ess bulk direct 1e5x4", lambda: m.ess(x3, kind="bulk"))
ok &= cmp("C3-like ess_rhat rank 1e5x4", lambda: m.ess_rhat(x3, kind="rank"))
ok &= cmp("C3-like ess bulk FFT 1e5x4", lambda: m.ess(x3, kind="bulk", autocov_method=m.FFTAutocovMethod()))
ctx.set_option("use_big", 0)
x5 = g(4000, 8, 50, "float32")
ok &= cmp("C5-shape f32 ess_rhat rank (large path)", lambda: m.ess_rhat(x5, kind="rank"))
ctx.set_option("use_big", 1)
xs = g(30000, 2, 6)
ok &= cmp("tiedrank 30000x2", lambda: m.tiedrank(xs))
ok &= cmp("rank_normalize 30000x2", lambda: m.rank_normalize(xs))
ok &= cmp("fold_around_median 30000x2", lambda: m.fold_around_median(xs))
# slabs the counting rank hands back: NaN, Inf, heavy ties, constants (the whole chunk takes the sort path)
xb = g(30000, 2, 6).clone(); xb[5, 0, 1] = float("nan"); xb[7, 1, 2] = float("inf"); xb[:, :, 3] = torch.round(xb[:, :, 3] * 2); xb[:, :, 4] = 1.5
ok &= cmp("NaN/Inf/ties/constant slabs ess_rhat rank", lambda: m.ess_rhat(xb, kind="rank"))
xm = torch.round(g(30000, 2, 4) * 1000) / 1000      # mild ties: resolved inside the buckets
ok &= cmp("mild ties ess_rhat rank", lambda: m.ess_rhat(xm, kind="rank"))
ctx.set_option("workspace_bytes", 64 << 20)          # several chunks per call
ok &= cmp("C4 small workspace (multi-chunk)", lambda: m.rhat_nested(x4, sup, kind="rank", split_chains=2))
ctx.set_option("workspace_bytes", 6 << 30)
print("ALL IDENTICAL" if ok else "MISMATCH", flush=True)

# ---- timing -------------------------------------------------------------------------------------------
P = int(sys.argv[1]) if len(sys.argv) > 1 else 400
x = g(100, 2048, P, seed=1)
torch.cuda.synchronize()


def timeit(label):
    fn = lambda: m.rhat_nested(x, sup, kind="rank", split_chains=2)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = min(ts)
    print(f"C4 P={P} {label:40s} {t:9.3f} ms  {P / t * 1e3:10.4g} params/s  frac {P / t * 1e3 / 3.997e6:.4f}", flush=True)


ctx.set_option("use_crank", 0); timeit("sort path")
ctx.set_option("use_crank", 1)
for factor, chunk in ((2, 0), (4, 0), (8, 0), (4, 200), (2, 200)):
    ctx.set_option("crank_factor", factor); ctx.set_option("crank_chunk", chunk)
    timeit(f"crank factor={factor} chunk={chunk}")
ctx.set_option("crank_factor", 4); ctx.set_option("crank_chunk", 0)
x = g(4000, 8, 4000, "float32", seed=1)
ctx.set_option("use_big", 0)
for v in (0, 1):
    ctx.set_option("use_crank", v)
    fn = lambda: m.ess_rhat(x, kind="rank")
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    print(f"C5-shape f32 ess_rhat rank P=4000 use_crank={v}: {a.elapsed_time(b):.3f} ms", flush=True)
