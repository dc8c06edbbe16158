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
# many NaNs / infinities (in-place NaN ranking), summary (fused and composed), callers
xm = r.standard_normal((600, 4, 2)); xm[np.cumsum(r.standard_normal((600, 4, 2)), axis=0) > 0] = np.nan
m.rank_normalize(xm); m.rhat(xm, kind="bulk"); m.ess_rhat(xm, kind="basic")
xi = np.where(r.random((1000, 4, 2)) < 0.5, np.inf, -np.inf); m.rhat(xi, kind="tail"); m.rhat(xi[:300], kind="rank")
m.summary(x); m.summary(x[:300]); m.summary(x, fields=("ess_tail", "rhat"))
m.gewekediag(x[:, 0, :]); m.heideldiag(x[:, 0, :])
print("probe done")
