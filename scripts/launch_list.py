"""One pass of a bench.py config at a reduced parameter count (profiling target for `ncu --metrics gpu__time_duration.sum`).
python scripts/launch_list.py c4nested 200"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import mcmcdiag_b200 as m
cfg = bench.CONFIGS[sys.argv[1]]
P = int(sys.argv[2])
x = m.generate_ar1(0.5, np.sqrt(0.75), cfg.draws, cfg.chains, P, seed=1, dtype=cfg.dtype)
torch.cuda.synchronize()
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 2):
    r = cfg.run(m, x)
torch.cuda.synchronize()
print([float(t.double().mean()) for t in r])
