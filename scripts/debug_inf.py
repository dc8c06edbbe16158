import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcmcdiag_b200 as m
from oracle import mcmcdiag_oracle as o
import importlib.util
spec = importlib.util.spec_from_file_location('fz', os.path.join(os.path.dirname(__file__), '..', 'tests', 'test_gpu_fuzz.py'))
fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
warnings.simplefilter("ignore")
r, x, kw, _ = fz.make_case(o, 1516)
x = x.astype(np.float64)
d, c, P = x.shape
for _ in range(int(r.integers(1, 4))):
    x[int(r.integers(0, d)), int(r.integers(0, c)), 0] = r.choice([np.nan, np.inf, -np.inf])
print(x.shape, kw, np.isinf(x[..., 0]).sum(), (x[..., 0] > 0).sum())
print("oracle rhat tail", o.rhat(x, kind="tail"), " ess_rhat tail", o.ess_rhat(x, kind="tail"))
print("gpu    rhat tail", m.rhat(x, kind="tail"), " ess_rhat tail", m.ess_rhat(x, kind="tail"))
print("gpu ess tail", m.ess(x, kind="tail"), "rank", m.ess_rhat(x, kind="rank"), o.ess_rhat(x, kind="rank"))
x0 = x[..., :1]
print("P=1: gpu ess_rhat tail", m.ess_rhat(x0, kind="tail"), "oracle", o.ess_rhat(x0, kind="tail"))
ctx = m.get_context(0)
for fp in (1, 2):
    ctx.set_option("force_path", fp)
    try:
        print("force_path", fp, m.ess_rhat(x0, kind="tail"), m.rhat(x0, kind="tail"))
    except Exception as e:
        print("force_path", fp, "error", e)
ctx.set_option("force_path", 0)
print("---- pieces")
fo = o.fold_around_median(x0); fg = m.fold_around_median(x0)
print("fold equal", np.array_equal(fo, fg, equal_nan=True), np.isnan(fo).sum(), np.isnan(fg).sum())
for rep in range(3):
    to, tg = o.tiedrank(fo.reshape(-1, order="F")), m.tiedrank(fo).reshape(-1, order="F")
    bad = np.flatnonzero(to != tg)
    print("tiedrank(fold) equal", bad.size == 0, bad[:10], to[bad[:10]], tg[bad[:10]])
    print("rhat tail gpu", m.rhat(x0, kind="tail"), "basic on oracle z:", m.rhat(o.rank_normalize(fo), kind="basic"), o.rhat(o.rank_normalize(fo), kind="basic"))
print("---- variants")
print("rhat bulk on folded data: gpu", m.rhat(fo, kind="bulk"), "oracle", o.rhat(fo, kind="bulk"))
xf = np.where(x0 > 0, 1e300, -1e300)
print("finite +-1e300: gpu tail", m.rhat(xf, kind="tail"), "oracle", o.rhat(xf, kind="tail"))
xs = x0[:1000]
print("first 1000 draws: gpu tail", m.rhat(xs, kind="tail"), m.rhat(xs, kind="tail"), "oracle", o.rhat(xs, kind="tail"))
for fp in (1, 2):
    ctx.set_option("force_path", fp)
    print(" force", fp, m.rhat(xs, kind="tail"), m.rhat(fo, kind="bulk"))
ctx.set_option("force_path", 0)
print("---- rank_normalize on folded data (NaN + inf)")
zo = o.rank_normalize(fo).reshape(-1, order="F")
for rep in range(3):
    zg = m.rank_normalize(fo).reshape(-1, order="F")
    bad = np.flatnonzero(~np.isclose(zo, zg, rtol=1e-12, equal_nan=True))
    print(rep, "mismatches", bad.size, bad[:8], zo[bad[:4]], zg[bad[:4]], "nan in out", np.isnan(zg).sum())
fs = fo.reshape(-1, order="F")
print("nan positions head", np.flatnonzero(np.isnan(fs))[:10])
