"""Small workload for compute-sanitizer (memcheck / racecheck): touches every kernel family once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcmcdiag_b200 as m
r = np.random.default_rng(1)
x = r.standard_normal((1000, 4, 3)); x[:, :, 1] = np.round(x[:, :, 1], 1)
ctx = m.get_context(0)
for kind in ("rank", "tail", "basic"):
    m.ess_rhat(x, kind=kind)                                   # fast / fastgen kernels
m.ess(x, kind="median"); m.ess(x, kind="mad"); m.mcse(x, kind="std"); m.mcse(x, kind="median")
xn = x.copy(); xn[0, 0, 0] = np.nan; m.ess_rhat(xn)             # redo list -> general kernel
ctx.set_option("force_path", 1)
for meth in (m.AutocovMethod(), m.FFTAutocovMethod(), m.BDAAutocovMethod()):
    m.ess_rhat(x[:300], kind="rank", autocov_method=meth)
m.rhat_nested(r.standard_normal((50, 8, 2)), [0, 0, 1, 1, 2, 2, 3, 3])
m.tiedrank(x[:200])
ctx.set_option("force_path", 2)
m.ess_rhat(x, kind="rank"); m.ess(x, kind="tail"); m.mcse(x, kind=m.Quantile(0.3))
m.ess_rhat(r.standard_normal((9000, 2, 2)), kind="basic", autocov_method=m.FFTAutocovMethod())   # four-step FFT
m.rhat_nested(r.standard_normal((50, 8, 2)), [0, 0, 1, 1, 2, 2, 3, 3])
ctx.set_option("force_path", 0)
print("probe done")
