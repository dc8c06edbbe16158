import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcmcdiag_b200 as m
x = m.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, 20000, seed=1)
ctx = m.get_context(0)
for fp in (0, 3):
    ctx.set_option("force_path", fp)
    try:
        l0 = ctx.stat("kernel_launches")
        r = m.mcse(x, kind="median")
        torch.cuda.synchronize()
        print("force", fp, "redo", ctx.stat("redo_count"), "last_path", ctx.stat("last_path"), "launches", ctx.stat("kernel_launches") - l0, float(r.mean()))
    except Exception as e:
        print("force", fp, "error", e)
ctx.set_option("force_path", 0)

import time
for kind in ("median", m.Quantile(0.1), "mean"):
    m.mcse(x, kind=kind); torch.cuda.synchronize()
    t0 = time.perf_counter(); m.mcse(x, kind=kind); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(kind, "ms per 20k", (t1 - t0) * 1e3, "redo", ctx.stat("redo_count"))
m.ess(x, kind="median"); torch.cuda.synchronize()
t0 = time.perf_counter(); m.ess(x, kind="median"); torch.cuda.synchronize(); print("ess median", (time.perf_counter() - t0) * 1e3)
