"""Larger instances of the large-slab configs: exercises workspace chunking; prints timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcmcdiag_b200 as m
ctx = m.get_context(0)
def run(name, fn, P, bpp):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    o = out[0] if isinstance(out, tuple) else out
    print(f"{name:44s} P={P:7d} {dt*1e3:10.1f} ms {P/dt:11.4e} params/s frac={P*bpp/dt/1e9/6548.2:.4f} finite={float(torch.isfinite(o).float().mean()):.3f} path={ctx.stat('last_path')}", flush=True)
x = m.generate_ar1(0.5, np.sqrt(0.75), 100, 2048, 2000, seed=1)
ids = np.repeat(np.arange(32), 64)
run("C4 rhat_nested(rank) 100x2048", lambda: m.rhat_nested(x, ids), 2000, 1_638_408)
del x
x = m.generate_ar1(0.5, np.sqrt(0.75), 4000, 8, 20000, seed=1, dtype="float32")
run("C5 ess(median, BDA) 4000x8 f32", lambda: m.ess(x, kind="median", autocov_method=m.BDAAutocovMethod()), 20000, 128_004)
run("C5 ess(std, BDA) 4000x8 f32", lambda: m.ess(x, kind="std", autocov_method=m.BDAAutocovMethod()), 20000, 128_004)
del x
x = m.generate_ar1(0.5, np.sqrt(0.75), 1_000_000, 4, 64, seed=1)
run("C3 ess(bulk, FFT) 1e6x4", lambda: m.ess(x, kind="bulk", autocov_method=m.FFTAutocovMethod()), 64, 32_000_008)
run("C3 ess(bulk, direct) 1e6x4", lambda: m.ess(x, kind="bulk"), 64, 32_000_008)
run("C3 ess_rhat(rank, FFT) 1e6x4", lambda: m.ess_rhat(x, autocov_method=m.FFTAutocovMethod()), 64, 32_000_016)
