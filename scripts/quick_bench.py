"""Developer timing probe (not the contract bench): device-resident AR(1) input, CUDA-event timing."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mcmcdiag_b200 as m

P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000
phi = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
reps = 3
print("host:", os.cpu_count(), "cpus;", torch.cuda.get_device_name(0))
os.system("free -g | head -2; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,memory.total --format=csv")
x = m.generate_ar1(phi, np.sqrt(1 - phi * phi), 1000, 4, P, seed=1)
torch.cuda.synchronize()
ctx = m.get_context(0)

def timeit(name, fn, bytes_per_param=32016):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = min(ts)
    print(f"{name:34s} {t:9.3f} ms  {P / t * 1e3:12.4e} params/s  {P * bytes_per_param / t / 1e6:8.1f} GB/s  frac={P * bytes_per_param / t / 1e6 / 6548.2:.3f}")

timeit("ess_rhat rank direct", lambda: m.ess_rhat(x))
timeit("rhat rank", lambda: m.rhat(x))
timeit("ess_rhat bulk", lambda: m.ess_rhat(x, kind="bulk"))
timeit("ess_rhat basic", lambda: m.ess_rhat(x, kind="basic"))
timeit("rhat basic", lambda: m.rhat(x, kind="basic"))
timeit("ess_rhat tail", lambda: m.ess_rhat(x, kind="tail"))
timeit("ess_rhat basic bda", lambda: m.ess_rhat(x, kind="basic", autocov_method=m.BDAAutocovMethod()))
timeit("ess_rhat basic fft", lambda: m.ess_rhat(x, kind="basic", autocov_method=m.FFTAutocovMethod()))
timeit("ess median", lambda: m.ess(x, kind="median"))
timeit("mcse mean", lambda: m.mcse(x, kind="mean"))
timeit("mcse median", lambda: m.mcse(x, kind="median"))
print("launches", ctx.stat("kernel_launches"))
timeit("summary (7 columns, fused)", lambda: m.summary(x))
timeit("summary bulk+tail+rhat", lambda: m.summary(x, fields=("ess_bulk", "ess_tail", "rhat")))
