"""Developer probe: time ess_rhat(kind=:rank) on the fast kernel for several bucket counts / phi."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcmcdiag_b200 as m
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000
ctx = m.get_context(0)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
for phi in (0.5,):
    x = m.generate_ar1(phi, np.sqrt(1 - phi * phi), 1000, 4, P, seed=1)
    for ver, B in ((0, 0), (1, 0), (2, 0), (8, 0)):
        ctx.set_option('fast_grid_mult', ver)
        for name, fn in (("rank", lambda: m.ess_rhat(x)), ("rhat rank", lambda: m.rhat(x)), ("bulk", lambda: m.ess_rhat(x, kind="bulk")), ("basic", lambda: m.ess_rhat(x, kind="basic"))):
            ms = t(fn)
            print(f"gridmult{ver} phi={phi} B={B:6d} {name:10s} {ms:8.3f} ms  {P/ms*1e3:10.4e} params/s  frac={P*32016/ms/1e6/6548.2:.4f}")
