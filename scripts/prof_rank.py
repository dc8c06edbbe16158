"""Profiling target: ess_rhat(kind=:rank) on device-resident AR(1) data, a few launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcmcdiag_b200 as m
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20000
kind = sys.argv[2] if len(sys.argv) > 2 else "rank"
x = m.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, P, seed=1)
for _ in range(3):
    S, R = m.ess_rhat(x, kind=kind)
torch.cuda.synchronize()
print(float(S.mean()), float(R.mean()))
