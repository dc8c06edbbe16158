import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mcmcdiag_b200 as m
from oracle import mcmcdiag_oracle as o
warnings.simplefilter("ignore")
r = np.random.default_rng(3)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
x = np.full((n, 4, 1), np.inf)
# clustered NaNs
mask = (np.cumsum(r.standard_normal((n, 4, 1)), axis=0) > 0)
x[mask] = np.nan
zo = o.rank_normalize(x).reshape(-1, order="F")
zg = m.rank_normalize(x).reshape(-1, order="F")
bad = np.flatnonzero(~np.isclose(zo, zg, rtol=1e-12, equal_nan=True))
print("n", 4 * n, "nan", int(np.isnan(x).sum()), "mismatches", bad.size, bad[:8], zg[bad[:4]])
