"""Developer probe: the big-slab estimator kernel (mcd_big.cuh) against the previous path (use_big = 0) and timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcmcdiag_b200 as m
ctx = m.get_context(0)

def both(fn):
    out = []
    for v in (0, 1):
        ctx.set_option("use_big", v)
        r = fn(); torch.cuda.synchronize()
        out.append((r.clone() if not isinstance(r, tuple) else tuple(t.clone() for t in r), ctx.stat("last_path")))
    return out

def cmp(name, fn, tol):
    (o, po), (n_, pn) = both(fn)
    o = o if isinstance(o, tuple) else (o,); n_ = n_ if isinstance(n_, tuple) else (n_,)
    worst = 0.0
    for a, b in zip(o, n_):
        a = a.double().cpu().numpy(); b = b.double().cpu().numpy()
        if not np.array_equal(np.isnan(a), np.isnan(b)): worst = np.inf
        ok = ~np.isnan(a)
        if ok.any(): worst = max(worst, float((np.abs(a[ok] - b[ok]) / np.abs(a[ok])).max()))
    print(f"{name:60s} paths {po}->{pn}  max rel diff {worst:.3e} {'OK' if worst <= tol else 'FAIL'}", flush=True)

bda, direct = m.BDAAutocovMethod(), m.AutocovMethod()
for dt, tol in (("float32", 2e-4), ("float64", 1e-9)):
    for (d, c) in ((4000, 8), (4001, 8), (3000, 5), (20000, 1), (1500, 16)):
        if dt == "float64" and d * c * 8 > 200_000: continue
        for phi in (0.5, 0.95):
            x = m.generate_ar1(phi, np.sqrt(1 - phi * phi), d, c, 300, seed=5, dtype=dt)
            for kind in ("mean", "std", "median"):
                for meth, mn in ((bda, "bda"), (direct, "direct")):
                    for split in (2, 1, 3):
                        cmp(f"{dt} {d}x{c} phi={phi} ess {kind} {mn} split={split}", lambda: m.ess(x, kind=kind, autocov_method=meth, split_chains=split), tol)
            cmp(f"{dt} {d}x{c} phi={phi} ess_rhat basic", lambda: m.ess_rhat(x, kind="basic"), tol)
    tdt = torch.float32 if dt == "float32" else torch.float64
    base = m.generate_ar1(0.5, np.sqrt(0.75), 4000, 8 if dt == "float32" else 4, 200, seed=9, dtype=dt)
    disc = torch.round(base * 2.0); nanv = base.clone(); nanv[5, 1, 3] = float("nan"); infv = base.clone(); infv[7, 0, 2] = float("inf"); infv[9, 1, 2] = -float("inf")
    const = torch.ones_like(base); const[:, :, :3] = base[:, :, :3]
    for nm, xs in (("discrete", disc), ("nan", nanv), ("inf", infv), ("const", const)):
        for kind in ("mean", "std", "median"):
            cmp(f"{dt} {nm} ess {kind} bda", lambda: m.ess(xs, kind=kind, autocov_method=bda), tol)

P = 20000
x = m.generate_ar1(0.5, np.sqrt(0.75), 4000, 8, P, seed=1, dtype="float32")
for kind in ("median", "std", "mean"):
    for v in (0, 1):
        ctx.set_option("use_big", v)
        fn = lambda: m.ess(x, kind=kind, autocov_method=bda)
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        t = min(ts)
        print(f"C5 ess {kind} BDA P={P} use_big={v}: {t:8.3f} ms  {P / t * 1e3:.4e} params/s  frac {P * 128004 / t / 1e6 / 6548.2:.4f}", flush=True)
