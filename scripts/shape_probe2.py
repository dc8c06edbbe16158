"""Developer probe: shapes with fewer than eight split chains on the TMA-staged kernel against the general kernel.
python scripts/shape_probe2.py [P]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mcmcdiag_b200 as m

P = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
ctx = m.get_context(0)


def t_ms(f, reps=3):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


for (d, c) in ((1000, 4), (1000, 3), (1000, 2), (1000, 1), (500, 4), (250, 8)):
    x = m.generate_ar1(0.5, np.sqrt(0.75), d, c, P, seed=1)
    fn = lambda: m.ess_rhat(x, kind="rank")
    out = []
    for force in (1, 0):
        ctx.set_option("force_path", force)
        r = fn(); torch.cuda.synchronize()
        out.append((t_ms(fn), ctx.stat("last_path"), [t.double() for t in r]))
    ctx.set_option("force_path", 0)
    rel = max(float(((a - b).abs() / b.abs()).max()) for a, b in zip(out[1][2], out[0][2]))
    bytes_pp = d * c * 8 + 16
    print(f"{d}x{c} x {P}: general kernel {out[0][0]:7.2f} ms -> path {out[1][1]} {out[1][0]:7.2f} ms  (x{out[0][0] / out[1][0]:.2f}, "
          f"{P * bytes_pp / out[1][0] / 1e6 / 6548.2 * 100:.1f} % of the roofline)  max rel diff {rel:.1e}", flush=True)
