"""Developer probe: the four-step FFT on the C3 shape (1e6 draws x 4 chains): transform length, columns per tile,
paired / summed data flow; then agreement of the paired path with the per-chain path and with the direct method on
other shapes.   python scripts/fft_probe.py [P]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mcmcdiag_b200 as m

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = m.get_context(0)
x = m.generate_ar1(0.5, np.sqrt(0.75), 1_000_000, 4, P, seed=1)
torch.cuda.synchronize()
fn = lambda: m.ess(x, kind="bulk", autocov_method=m.FFTAutocovMethod())
ref = None
for full, tc, pair in ((1, 4, 0), (0, 2, 0), (0, 2, 1)):
    ctx.set_option("fft_tc", tc); ctx.set_option("fft_full", full); ctx.set_option("fft_pair", pair)
    r = fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    if ref is None:
        ref = r.clone()
    rel = float(((r - ref).abs() / ref.abs()).max())
    print(f"C3 P={P} fft_full={full} fft_tc={tc} fft_pair={pair}: {min(ts):8.3f} ms  {P / min(ts) * 1e3:8.1f} params/s  "
          f"frac {P / min(ts) * 1e3 / 204631:.4f}  max rel diff vs first = {rel:.2e}", flush=True)
ctx.set_option("fft_tc", 0); ctx.set_option("fft_full", 0); ctx.set_option("fft_pair", 1)
# other shapes of the paired path: odd chain count, N1 with a factor 3, Float32
for (d, c, split, dt) in ((9001, 3, 1, "float64"), (30000, 1, 2, "float64"), (200000, 3, 2, "float32"), (70000, 5, 1, "float64")):
    y = m.generate_ar1(0.7, np.sqrt(1 - 0.49), d, c, 6, seed=5, dtype=dt)
    out = []
    for pair in (0, 1):
        ctx.set_option("fft_pair", pair)
        out.append(m.ess(y, kind="bulk", autocov_method=m.FFTAutocovMethod(), split_chains=split).double())
    direct = m.ess(y, kind="bulk", split_chains=split).double()
    print(f"{d}x{c} split={split} {dt}: pair vs per-chain max rel diff {float(((out[0] - out[1]).abs() / out[0].abs()).max()):.2e}; "
          f"FFT vs direct {float(((out[1] - direct).abs() / direct.abs()).max()):.2e}", flush=True)
ctx.set_option("fft_pair", 1)

# ---- shared-memory FFT paths (slab kernel; large path with chains that fit shared memory) ----------------------
def t_ms(f, reps=3):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)

xh = m.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, 200_000, seed=1)
f_fft = lambda: m.ess_rhat(xh, kind="basic", autocov_method=m.FFTAutocovMethod())
d = m.ess_rhat(xh, kind="basic")[0].double(); f = f_fft()[0].double()
print(f"headline shape, FFT method (slab kernel): {t_ms(f_fft):.2f} ms per 200k (round 1: 97.3); FFT vs direct max rel diff "
      f"{float(((f - d).abs() / d.abs()).max()):.2e}", flush=True)
ym = m.generate_ar1(0.6, 0.8, 6000, 4, 2000, seed=2)
for pair in (0, 1):
    ctx.set_option("fft_pair", pair)
    ff = lambda: m.ess(ym, kind="bulk", autocov_method=m.FFTAutocovMethod())
    r = ff().double()
    if pair == 0:
        r0 = r
    print(f"6000x4 x 2000 params, FFT in shared memory per chain / per parameter (fft_pair={pair}): {t_ms(ff):.2f} ms, path {ctx.stat('last_path')}, "
          f"max rel diff vs per-chain {float(((r - r0).abs() / r0.abs()).max()):.2e}", flush=True)
ctx.set_option("fft_pair", 1)
for mb in (32, 160):
    ctx.set_option("ztab_max_mb", mb)
    print(f"C3 P={P} z table up to {mb} MB: {t_ms(fn, 2):.3f} ms", flush=True)
ctx.set_option("ztab_max_mb", 32)
