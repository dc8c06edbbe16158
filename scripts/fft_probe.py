"""Developer probe: columns per tile of the four-step FFT (C3 shape: 1e6 draws x 4 chains, N = 2^20 per split chain).
python scripts/fft_probe.py [P]"""
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
for full, tc, crank in ((1, 4, 0), (1, 2, 1), (0, 2, 1), (0, 4, 1), (0, 1, 1)):
    ctx.set_option("use_crank", crank); ctx.set_option("fft_tc", tc); ctx.set_option("fft_full", full)
    r = fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    if ref is None:
        ref = r.clone()
    rel = float(((r - ref).abs() / ref.abs()).max())
    print(f"C3 P={P} fft_full={full} fft_tc={tc} use_crank={crank}: {min(ts):8.3f} ms  {P / min(ts) * 1e3:8.1f} params/s  "
          f"frac {P / min(ts) * 1e3 / 204631:.4f}  max rel diff vs first = {rel:.2e}", flush=True)
ctx.set_option("fft_tc", 0); ctx.set_option("use_crank", 1); ctx.set_option("fft_full", 0)
