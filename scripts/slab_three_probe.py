"""Developer probe: the general kernel's 80-register entry (three CTAs per SM, spills) against its 128-register entry
(two CTAs per SM, no spills) on slabs small enough for three.   python scripts/slab_three_probe.py [P]"""
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


for (d, c, kind, kw) in ((250, 8, "rank", {}), (100, 8, "rank", {}), (200, 16, "rank", {}), (250, 8, "tail", {}), (250, 8, "basic", {}),
                         (500, 4, "bulk", {"autocov_method": m.BDAAutocovMethod()}), (300, 4, "rank", {"autocov_method": m.FFTAutocovMethod()})):
    x = m.generate_ar1(0.5, np.sqrt(0.75), d, c, P, seed=1)
    fn = lambda: m.ess_rhat(x, kind=kind, **kw)
    res = []
    for three in (1, 0):
        ctx.set_option("slab_three", three)
        r = fn(); torch.cuda.synchronize()
        res.append((t_ms(fn), ctx.stat("last_path"), [t.double() for t in r]))
    ctx.set_option("slab_three", 1)
    same = all(bool(((a == b) | (a.isnan() & b.isnan())).all()) for a, b in zip(res[0][2], res[1][2]))
    print(f"{d}x{c} {kind} {list(kw)}: three CTAs/SM (80 regs) {res[0][0]:7.2f} ms, two (128 regs) {res[1][0]:7.2f} ms  path {res[0][1]} identical={same}", flush=True)
