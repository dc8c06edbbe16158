"""Developer probe: the other BASELINE.json configs at reduced parameter counts (device-resident,
CUDA-event timing), with a spot parity check against the CPU oracle on the first parameters."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mcmcdiag_b200 as m
from oracle import mcmcdiag_oracle as o

which = sys.argv[1:] or ["c1", "c2b", "c3d", "c4", "c5"]
ctx = m.get_context(0)

def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best, out

def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.nanmax(np.abs(a - b) / np.abs(b)))

def report(name, P, bytes_per_param, ms, path):
    print(f"{name:34s} P={P:8d} {ms:10.3f} ms {P/ms*1e3:11.4e} params/s  roofline frac={P*bytes_per_param/ms/1e6/6548.2:.4f}  path={path}", flush=True)

if "c1" in which:
    x = m.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, 10, seed=1)
    ms, (S, R) = timeit(lambda: m.ess_rhat(x), reps=5)
    xh = x.cpu().numpy(); t0 = time.perf_counter(); m.ess_rhat(xh); host_ms = (time.perf_counter() - t0) * 1e3
    So, Ro = o.ess_rhat(xh)
    print(f"C1 ess_rhat(rank) 1000x4x10: device {ms*1e3:.1f} us, host-array call {host_ms*1e3:.1f} us, rel err {rel(S.cpu(), So):.2e} {rel(R.cpu(), Ro):.2e}")

if "c2b" in which:
    P = 100_000
    x = m.generate_ar1(0.5, np.sqrt(0.75), 1000, 4, P, seed=1)
    for name, fn, nout in (("C2b ess_rhat bulk", lambda: m.ess_rhat(x, kind="bulk"), 2), ("C2b ess_rhat tail", lambda: m.ess_rhat(x, kind="tail"), 2),
                           ("C2b mcse mean", lambda: m.mcse(x), 1), ("C2b mcse median", lambda: m.mcse(x, kind="median"), 1),
                           ("C2b ess median", lambda: m.ess(x, kind="median"), 1), ("C2b ess std", lambda: m.ess(x, kind="std"), 1)):
        ms, out = timeit(fn)
        report(name, P, 32000 + 8 * nout, ms, ctx.stat("last_path"))

if "c3d" in which:
    for P, meth, name in ((4, m.AutocovMethod(), "C3 ess bulk direct 1e6x4"), (4, m.FFTAutocovMethod(), "C3 ess bulk FFT 1e6x4")):
        x = m.generate_ar1(0.5, np.sqrt(0.75), 1_000_000, 4, P, seed=1)
        try:
            ms, S = timeit(lambda: m.ess(x, kind="bulk", autocov_method=meth), reps=1)
            report(name, P, 32_000_008, ms, ctx.stat("last_path"))
        except Exception as e:
            print(name, "->", type(e).__name__, str(e)[:120])
    x = m.generate_ar1(0.5, np.sqrt(0.75), 100_000, 4, 2, seed=1)
    S = m.ess(x, kind="bulk"); So = o.ess(x.cpu().numpy(), kind="bulk")
    print("   parity 1e5x4x2 bulk direct rel err", rel(S.cpu(), So))

if "c4" in which:
    P = 200
    x = m.generate_ar1(0.5, np.sqrt(0.75), 100, 2048, P, seed=1)
    ids = np.repeat(np.arange(32), 64)
    ms, R = timeit(lambda: m.rhat_nested(x, ids, kind="rank"))
    report("C4 rhat_nested rank 100x2048", P, 1_638_408, ms, ctx.stat("last_path"))
    Ro = o.rhat_nested(x[:, :, :2].cpu().numpy(), ids, kind="rank")
    print("   parity rel err", rel(R[:2].cpu(), Ro))

if "c5" in which:
    P = 2000
    x = m.generate_ar1(0.5, np.sqrt(0.75), 4000, 8, P, seed=1, dtype="float32")
    for kind in ("median", "std"):
        ms, S = timeit(lambda: m.ess(x, kind=kind, autocov_method=m.BDAAutocovMethod()))
        report(f"C5 ess {kind} BDA 4000x8 f32", P, 128_004, ms, ctx.stat("last_path"))
        So = o.ess(x[:, :, :3].cpu().numpy(), kind=kind, autocov_method=o.BDAAutocovMethod())
        print("   parity rel err", rel(S[:3].cpu(), So))
