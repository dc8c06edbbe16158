"""End-to-end throughput of ONE process driving D GPUs through a mcd_create_multi context: host (pinned) array in,
host results out.  python scripts/multi_e2e.py [P]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcmcdiag_b200 as m

P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000
ndev = torch.cuda.device_count()
xd = m.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, P, seed=1)
xh_t = torch.empty((P, 4, 1000), dtype=torch.float64, pin_memory=True)
xh_t.copy_(xd.permute(2, 1, 0)); torch.cuda.synchronize()
xh = xh_t.numpy().transpose(2, 1, 0)
S0, R0 = m.ess_rhat(xd); S0 = S0.cpu().numpy()
for d in [k for k in (1, 2, 4, 8) if k <= ndev]:
    grp = m.Context(devices=list(range(d)))
    S, R = m.ess_rhat(xh, ctx=grp)                       # warm-up: allocates the staging buffers
    t0 = time.perf_counter()
    for _ in range(2):
        S, R = m.ess_rhat(xh, ctx=grp)
    dt = (time.perf_counter() - t0) / 2
    print(json.dumps({"probe": "multi_ctx_e2e", "devices": d, "params": P, "ms": dt * 1e3, "params_per_s": P / dt,
                      "h2d_GBs": P * 32000 / dt / 1e9, "bitwise_equal_to_device_resident": bool(np.array_equal(S, S0))}), flush=True)
    grp.close()
