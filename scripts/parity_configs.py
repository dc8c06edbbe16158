"""Parity of the BASELINE.json configs at their exact per-parameter shapes on as many device-generated parameters
as asked: fraction of parameters within the north star's tolerance, worst relative difference, outliers.
    python scripts/parity_configs.py c2rank=20000 c2summary=1000 c3fft=8 c4nested=100 c5bda=3000"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
import mcmcdiag_b200 as m
from oracle import build_oracle
from test_gpu_configs import compare

build_oracle.build()
m.get_context(0)
for spec in sys.argv[1:] or ["c2rank=20000", "c2summary=300", "c3fft=3", "c4nested=40", "c5bda=1000"]:
    name, n = spec.split("="); n = int(n)
    t0 = time.time()
    frac, worst, out = compare(bench, m, name, n, oracle_only=name in ("c5bda", "c2summary"))
    cfg = bench.CONFIGS[name]
    print(f"{name:10s} {n:6d} params  within {cfg.tol:g}: {frac:.6f}  worst rel diff {worst:.3e}  outliers {out[:8]}  ({time.time() - t0:.0f} s)", flush=True)
