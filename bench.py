#!/usr/bin/env python
"""bench.py — params/sec of the ESS / R-hat hot path on N B200s, for the workloads BASELINE.json names.

    python bench.py --gpus N --steps K --warmup W                     # headline: ess_rhat(kind=:rank), 1000x4x1e6 f64
    python bench.py --config {c2rank,c2summary,c3fft,c4nested,c5bda}  # the other BASELINE.json configs at their named sizes
    python bench.py --impl reference [--config ...]                   # CPU arm: the reference algorithm on the host cores

One "step" = one complete pass of the named call(s) over the whole (sharded) array, inputs resident in HBM.
The parameter axis is sharded contiguously over ranks (parameters are independent); the only
collective is the gather of the per-parameter results to rank 0 (NCCL), inside the timed region.  Total work
is fixed (the named array), so scaling is "strong".

`e2e` is the same metric through the public host API with HOST (pinned) input: every step stages the shard
over PCIe in overlapped chunks and reads the results back.

The CPU baseline / reference arm is a restatement of the reference algorithm (oracle/ref_port.cpp, C++/OpenMP,
for the calls it covers; the NumPy oracle for FFT / nested R-hat): the real package is Julia, and no Julia
exists in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PHI = 0.5


# ---------------------------------------------------------------------------------------------
# workloads (BASELINE.json `configs`; SURVEY.md §8(d) for shapes and algorithmic bytes)
# ---------------------------------------------------------------------------------------------
class Config:
    def __init__(self, name, draws, chains, params, dtype, reads, outputs, metric, workload, kernel):
        self.name, self.draws, self.chains, self.params, self.dtype = name, draws, chains, params, dtype
        self.elem = 8 if dtype == "float64" else 4
        self.reads, self.outputs = reads, outputs      # reference calls per step (each reads x once), result columns
        # input once per call + every output once
        self.bytes_per_param = reads * draws * chains * self.elem + outputs * self.elem
        self.metric, self.workload, self.kernel = metric, workload, kernel

    # the step on an array (torch CUDA tensor or NumPy host array) -> tuple of per-parameter result vectors
    def run(self, m, x):
        n = self.name
        if n == "c2rank":
            return tuple(m.ess_rhat(x, kind="rank"))
        if n == "c2summary":
            r = m.summary(x)
            return tuple(r[k] for k in ("mean", "std", "mcse_mean", "mcse_std", "ess_bulk", "ess_tail", "rhat"))
        if n == "c3fft":
            return (m.ess(x, kind="bulk", autocov_method=m.FFTAutocovMethod()),)
        if n == "c4nested":
            import numpy as np
            return (m.rhat_nested(x, np.repeat(np.arange(32), 64), kind="rank", split_chains=2),)
        if n == "c5bda":
            bda = m.BDAAutocovMethod()
            return (m.ess(x, kind="median", autocov_method=bda), m.ess(x, kind="std", autocov_method=bda))
        raise ValueError(n)

    # the same step by the CPU restatement of the reference algorithm; returns (results, cores used, description)
    def run_cpu(self, xs, cores):
        import numpy as np
        from oracle import mcmcdiag_oracle as o
        from oracle import ref_port as rp
        n = self.name
        port = "C++/OpenMP restatement of the reference algorithm (oracle/ref_port.cpp)"
        numpy_oracle = "NumPy/SciPy restatement of the reference algorithm (oracle/mcmcdiag_oracle.py), one thread"
        if n == "c2rank":
            return tuple(rp.ess_rhat(xs, kind="rank", nthreads=cores)), cores, port
        if n == "c2summary":
            # the ESS-shaped work of the seven columns: bulk + tail ESS / R-hat and the two MCSE proxies' ESS
            eb, rb = rp.ess_rhat(xs, kind="bulk", nthreads=cores)
            et, rt = rp.ess_rhat(xs, kind="tail", nthreads=cores)
            _, rr = rp.ess_rhat(xs, kind="rank", nthreads=cores, want_ess=False)
            em = rp.ess_estimator(xs, "mean", nthreads=cores)
            es = rp.ess_estimator(xs, "std", nthreads=cores)
            mean = xs.mean(axis=(0, 1)); std = xs.std(axis=(0, 1), ddof=1)
            return (mean, std, std / np.sqrt(em), None, eb, et, rr), cores, port + "; mcse_std column not restated in C++"
        if n == "c3fft":
            return (o.ess(xs, kind="bulk", autocov_method=o.FFTAutocovMethod()),), 1, numpy_oracle
        if n == "c4nested":
            return (o.rhat_nested(xs, np.repeat(np.arange(32), 64), kind="rank", split_chains=2),), 1, numpy_oracle
        if n == "c5bda":
            x64 = xs.astype(np.float64)
            return (rp.ess_estimator(x64, "median", method="bda", nthreads=cores),
                    rp.ess_estimator(x64, "std", method="bda", nthreads=cores)), cores, port + " in Float64"
        raise ValueError(n)

    def cpu_sample(self, cores):
        return {"c2rank": max(2000, 1500 * cores), "c2summary": max(500, 300 * cores), "c3fft": 2, "c4nested": 16,
                "c5bda": max(200, 100 * cores)}[self.name]

    @property
    def tol(self):
        return 1e-8 if self.dtype == "float64" else 1e-4


CONFIGS = {c.name: c for c in (
    Config("c2rank", 1000, 4, 1_000_000, "float64", 1, 2,
           "params/sec for ess_rhat(kind=:rank), 1000x4x1e6 f64",
           "ess_rhat(kind=:rank, split_chains=2, maxlag=250, AutocovMethod) on 1000 draws x 4 chains x 1e6 params Float64, AR(1) phi=0.5",
           "mcd::rk2_kernel<double, LONG, rank>"),
    Config("c2summary", 1000, 4, 1_000_000, "float64", 1, 7,
           "params/sec for fused summary (mean, std, mcse_mean, mcse_std, ess_bulk, ess_tail, rhat), 1000x4x1e6 f64",
           "ess_rhat bulk+tail + mcse as ONE fused call (mcd_summary: 7 columns from one read) on 1000 draws x 4 chains x 1e6 params Float64, AR(1) phi=0.5",
           "mcd::fastgen_kernel<double>"),
    Config("c3fft", 1_000_000, 4, 1000, "float64", 1, 1,
           "params/sec for ess(kind=:bulk, FFTAutocovMethod), 1e6x4x1000 f64",
           "ess(kind=:bulk, FFTAutocovMethod, split_chains=2, maxlag=250) on 1e6 draws x 4 chains x 1000 params Float64, AR(1) phi=0.5",
           "large-slab pipeline (counting rank mcd_crank + four-step FFT, N = 2^19 >= niter + maxlag)"),
    Config("c4nested", 100, 2048, 10_000, "float64", 1, 1,
           "params/sec for rhat_nested(kind=:rank), 2048 chains in 32 superchains x 100 draws x 1e4 params f64",
           "rhat_nested(kind=:rank, split_chains=2), superchain_ids = repeat(1:32, inner=64), on 100 draws x 2048 chains x 1e4 params Float64, AR(1) phi=0.5",
           "large-slab pipeline (counting rank mcd_crank + split-chain moments + nested R-hat)"),
    Config("c5bda", 4000, 8, 100_000, "float32", 2, 2,
           "params/sec for ess(kind=median) + ess(kind=std) with BDAAutocovMethod, 4000x8x1e5 f32",
           "ess(kind=median) and ess(kind=std), BDAAutocovMethod, split_chains=2, maxlag=250, on 4000 draws x 8 chains x 1e5 params Float32, AR(1) phi=0.5 (two reference calls per step)",
           "mcd::big_kernel<float> (slab resident in one SM's shared memory via TMA)"),
)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores
# ---------------------------------------------------------------------------------------------
def host_threads():
    """Host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 for its workers; the CPU arm
    sets its own thread count (omp_set_num_threads), so the affinity mask is what counts."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def host_ar1(cfg, params, seed=1):
    """AR(1) chains as test/helpers.jl:4-12 on the host: (draws, chains, params), column-major."""
    import numpy as np
    from scipy.signal import lfilter
    rng = np.random.default_rng(seed)
    sigma = (1 - PHI * PHI) ** 0.5
    eps = rng.standard_normal((params, cfg.chains, cfg.draws))
    x = lfilter([sigma], [1.0, -PHI], eps, axis=2)
    return np.ascontiguousarray(x.astype(cfg.dtype)).transpose(2, 1, 0)


def time_cpu(cfg, xs, steps, warmup):
    cores = host_threads()
    for _ in range(warmup):
        cfg.run_cpu(xs, cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        res, used, desc = cfg.run_cpu(xs, cores)
    dt = time.perf_counter() - t0
    return xs.shape[2] * steps / dt, used, dt / steps, desc, res


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import build_oracle
    build_oracle.build()
    cores = host_threads()
    sample = cfg.cpu_sample(cores)
    x = host_ar1(cfg, sample)
    value, used, sec, desc, _ = time_cpu(cfg, x, max(1, args.steps), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": cfg.metric, "value": value, "unit": "params/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64" if cfg.dtype == "float64" else "f32",
        "data": "synthetic AR(1) phi=0.5 (host numpy)",
        "config": {"workload": cfg.workload, "sample": f"each step = {sample} of the {cfg.params} parameters (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": "params/s", "cores": used, "kind": "port",
                         "sample": f"{sample} params/step; {desc} — the Julia package cannot run here (no Julia in the image)"},
        "e2e": {"value": value, "unit": "params/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def gpu_arm(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import mcmcdiag_b200 as mcd
    ctx = mcd.get_context(local)

    total = args.params or cfg.params
    lo, hi = mcd.sharding.shard_range(total, rank, world)
    shard = hi - lo
    sigma = (1 - PHI * PHI) ** 0.5
    x = mcd.generate_ar1(PHI, sigma, cfg.draws, cfg.chains, shard, seed=1, param_offset=lo, device=local, dtype=cfg.dtype)
    torch.cuda.synchronize()

    def step():
        res = cfg.run(mcd, x)
        if world > 1:
            return mcd.sharding.gather_params(torch.stack(res), total, dst=0)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # device timing of the call(s) alone (CUDA events on the launching stream) next to the whole step
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.stat("kernel_launches")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        kev[i][0].record()
        res = cfg.run(mcd, x)
        kev[i][1].record()
        if world > 1:
            mcd.sharding.gather_params(torch.stack(res), total, dst=0)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = e0.elapsed_time(e1)
    launches = ctx.stat("kernel_launches") - l0
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    t = torch.tensor([elapsed_ms, kernel_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        elapsed_ms, kernel_ms, launches = float(tmax[0]), float(tmax[1]), int(tsum[2])
    ms_per_step = elapsed_ms / args.steps
    value = total / (ms_per_step * 1e-3)
    last_path = ctx.stat("last_path")

    # ---- e2e: public API, host (pinned) input, H2D + D2H inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        e2e_params = min(shard, args.e2e_params // world if args.e2e_params else shard)
        tdt = torch.float64 if cfg.dtype == "float64" else torch.float32
        # pinned staging of the whole shard; if the host cannot pin that much, every rank halves its sample
        # together (the choice is agreed with an all-reduce) and the line says so
        xh_t = None
        while True:
            try:
                xh_t = torch.empty((e2e_params, cfg.chains, cfg.draws), dtype=tdt, pin_memory=True)
                ok = 1
            except (RuntimeError, MemoryError):
                xh_t, ok = None, 0
            okt = torch.tensor([ok], dtype=torch.int32, device=dev)
            if world > 1:
                dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            if int(okt[0]) == 1:
                break
            xh_t = None
            e2e_params //= 2
            if e2e_params < 1:
                raise SystemExit("cannot pin host memory for the e2e measurement")
        xh_t.copy_(x.permute(2, 1, 0)[:e2e_params])
        torch.cuda.synchronize()
        xh = xh_t.numpy().transpose(2, 1, 0)          # (draws, chains, params) column-major view
        cfg.run(mcd, xh)                              # warm-up (allocates staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            hres = cfg.run(mcd, xh)
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": e2e_params * world / float(tt[0]), "unit": "params/s",
               "h2d_bytes_per_step": e2e_params * world * cfg.draws * cfg.chains * cfg.elem * cfg.reads,
               "d2h_bytes_per_step": e2e_params * world * cfg.outputs * cfg.elem,
               "params_per_step": e2e_params * world, "host_memory": "pinned", "ms_per_step": float(tt[0]) * 1e3,
               "sample": "whole array" if e2e_params == shard else f"first {e2e_params} parameters of each shard (host could not pin more)"}
        # parity spot check against the device-resident result
        assert np.array_equal(np.asarray(hres[0]), res[0][:e2e_params].cpu().numpy(), equal_nan=True), \
            "host-staged and device-resident results differ"
        del xh_t

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    achieved = shard * cfg.bytes_per_param / (kernel_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from its last `ncu --set full` capture (profiles/roofline_traffic.json, written
    # by scripts/ncu_traffic.py): used only while the kernel source it was taken on is unchanged (sha-256 stamp)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            import hashlib
            tj = json.load(open(tp)).get(cfg.name)
            if tj:
                with open(os.path.join(ROOT, tj["kernel_source"]), "rb") as f:
                    sha = hashlib.sha256(f.read()).hexdigest()[:16]
                if sha == tj["source_sha16"]:
                    traffic = tj["dram_bytes_per_param"] * shard
                    traffic_src = f"ncu --set full capture of this kernel source ({tj['kernel_source']} sha {sha}): {tj['capture']}"
                else:
                    traffic_src = "stale: the kernel source changed since the last ncu capture"
        except Exception:
            traffic, traffic_src = None, None
    line = {
        "metric": cfg.metric, "value": value, "unit": "params/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64" if cfg.dtype == "float64" else "f32",
        "data": "synthetic AR(1) phi=0.5, seed 1, generated on device (Philox4x32-10, Box-Muller)",
        "config": {"workload": cfg.workload, "name": cfg.name, "params_total": total, "params_per_gpu": shard,
                   "sharding": f"contiguous parameter ranges over {world} rank(s); results gathered to rank 0 (NCCL)",
                   "l2": "input per GPU (%.1f GB) is larger than the 126 MB L2; no flush needed"
                         % (shard * cfg.draws * cfg.chains * cfg.elem / 1e9)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "kernel": cfg.kernel, "path_code": last_path,
                     "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": shard * cfg.bytes_per_param,
                     "algorithmic_bytes_per_param": cfg.bytes_per_param},
        "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
    }
    if world == 1 and not args.no_cpu:
        from oracle import build_oracle
        build_oracle.build()
        cores = host_threads()
        sample = min(shard, cfg.cpu_sample(cores))
        xs = np.asfortranarray(x[:, :, :sample].cpu().numpy())
        v, used, sec, desc, cres = time_cpu(cfg, xs, 1, 0)
        ok, worst = True, 0.0
        for g, c in zip(res, cres):
            if c is None:
                continue
            g = g[:sample].double().cpu().numpy(); c = np.asarray(c, dtype=np.float64)
            rel = np.abs(g - c) / np.maximum(np.abs(c), 1e-300)
            rel = rel[np.isfinite(rel)]
            worst = max(worst, float(rel.max()) if rel.size else 0.0)
            ok = ok and bool(np.array_equal(np.isnan(g), np.isnan(c)))
        frac_ok = None
        try:
            bad = np.zeros(sample, dtype=bool)
            for g, c in zip(res, cres):
                if c is None:
                    continue
                g = g[:sample].double().cpu().numpy(); c = np.asarray(c, dtype=np.float64)
                with np.errstate(all="ignore"):
                    bad |= ~((np.abs(g - c) <= cfg.tol * np.abs(c)) | (np.isnan(g) & np.isnan(c)))
            frac_ok = float(1.0 - bad.mean())
        except Exception:
            pass
        line["cpu_baseline"] = {"value": v, "unit": "params/s", "cores": used, "kind": "port",
                                "sample": f"first {sample} of the GPU run's parameters (identical inputs), one pass; {desc}; "
                                          "Julia is not available in this image",
                                "max_rel_diff_vs_gpu": worst, "fraction_within_tolerance": frac_ok, "tolerance": cfg.tol,
                                "nan_pattern_matches": ok}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2rank", choices=sorted(CONFIGS),
                    help="workload (default: the headline, BASELINE.json's metric on configs[1]'s array)")
    ap.add_argument("--params", type=int, default=0, help="total parameters (default: the config's named size)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-params", type=int, default=0, help="cap on host-staged parameters (0 = all)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        return reference_arm(args, cfg)
    return gpu_arm(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
