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

# ---- round-2 kernels ------------------------------------------------------------------------------------------
ctx.set_option("force_path", 0)
for d in (1000, 998, 600, 100, 22):                              # rk2: LONG and short chains, every mode
    xs = r.standard_normal((d, 4, 3)); xs[:5, 0, 1] = xs[5:10, 0, 1]
    m.ess_rhat(xs, kind="rank"); m.ess_rhat(xs, kind="bulk"); m.ess_rhat(xs, kind="basic"); m.rhat(xs, kind="tail"); m.rhat(xs, kind="rank")
m.ess_rhat(r.standard_normal((1000, 4, 3)).astype(np.float32), kind="rank")
xo = r.standard_normal((1000, 4, 3)) * 1e-3 + 1e9; m.ess_rhat(xo, kind="rank")     # few key steps of range: exact min / max path
xb = r.standard_normal((4000, 8, 3)).astype(np.float32)                             # big-slab kernel
for kind in ("median", "std", "mean"):
    m.ess(xb, kind=kind, autocov_method=m.BDAAutocovMethod()); m.ess(xb, kind=kind)
xc = r.standard_normal((100, 256, 2))                                               # counting rank (n = 25 600)
ids = np.repeat(np.arange(8), 32)
m.rhat_nested(xc, ids, kind="rank"); m.rhat(xc, kind="rank"); m.tiedrank(xc); m.fold_around_median(xc)
xc[3, 7, 0] = np.nan; m.rhat_nested(xc, ids, kind="rank")                           # flagged chunk -> sort path
ctx.set_option("workspace_bytes", 4 << 20); m.rhat(r.standard_normal((100, 256, 3)), kind="rank"); ctx.set_option("workspace_bytes", 6 << 30)
fft = m.FFTAutocovMethod()
m.ess_rhat(r.standard_normal((6000, 4, 2)), kind="bulk", autocov_method=fft)        # one CTA per parameter, paired chains
m.ess_rhat(r.standard_normal((9001, 3, 2)), kind="bulk", autocov_method=fft, split_chains=1)   # four-step, odd chain count
m.ess_rhat(r.standard_normal((30000, 1, 2)), kind="bulk", autocov_method=fft)       # four-step, N1 with a factor 3
m.ess_rhat(r.standard_normal((1000, 4, 3)), kind="tail", autocov_method=fft)        # slab kernel, paired chains
xm2 = np.ma.masked_array(r.standard_normal((300, 4, 5))); xm2[3, 1, 2] = np.ma.masked; m.ess_rhat(xm2)   # skip mask
print("round-2 probe done")
