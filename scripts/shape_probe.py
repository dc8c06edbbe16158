"""Developer probe: rank-kind throughput for several slab shapes, auto path vs forced general slab kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcmcdiag_b200 as m
ctx = m.get_context(0)
def t(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
P = 100_000
for draws, chains, split in ((1000, 4, 2), (500, 8, 2), (2000, 2, 2), (1000, 2, 2), (4000, 1, 1), (125, 32, 1), (1000, 3, 2), (2000, 4, 2)):
    x = m.generate_ar1(0.5, np.sqrt(0.75), draws, chains, P, seed=1)
    bpp = draws * chains * 8 + 16
    row = f"{draws:5d}x{chains:<3d} split {split}:"
    for fp in (0, 1):
        ctx.set_option("force_path", fp)
        try:
            ms = t(lambda: m.ess_rhat(x, split_chains=split))
            row += f"  path {ctx.stat('last_path')}: {ms:8.2f} ms {P/ms*1e3:10.3e}/s ({P*bpp/ms/1e6/6548.2*100:4.1f}%)"
        except Exception as e:
            row += f"  forced {fp}: {type(e).__name__}"
    ctx.set_option("force_path", 0)
    ms = t(lambda: m.ess_rhat(x, kind="tail", split_chains=split))
    row += f"  | tail path {ctx.stat('last_path')}: {ms:8.2f} ms"
    print(row, flush=True)
    del x
