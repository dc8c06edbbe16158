#!/usr/bin/env python
"""bench.py — headline benchmark: params/sec for ess_rhat(kind=:rank) on 1000 draws x 4 chains x
1e6 Float64 parameters (BASELINE.json configs[1] shape with the metric's call), on N B200s.

    python bench.py --gpus N --steps K --warmup W           # GPU arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference algorithm

One "step" = one complete `ess_rhat(x; kind=:rank)` over the whole (sharded) array, inputs
resident in HBM.  The parameter axis is sharded contiguously over ranks (parameters are
independent); the only collective is the gather of the per-parameter results
to rank 0 (NCCL), inside the timed region.  Total work is fixed (the named 1e6-parameter
array), so scaling is "strong".

`e2e` is the same metric through the public host API with HOST (pinned) input: every step
stages the shard over PCIe in overlapped chunks and reads the results back.

The CPU baseline / reference arm is the C++/OpenMP restatement of the reference algorithm
(oracle/ref_port.cpp): the real package is Julia, and no Julia exists in this image.
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

DRAWS, CHAINS, PARAMS = 1000, 4, 1_000_000
PHI = 0.5
BYTES_PER_PARAM = DRAWS * CHAINS * 8 + 2 * 8          # input once + two outputs
METRIC = "params/sec for ess_rhat(kind=:rank), 1000x4x1e6 f64"
WORKLOAD = "ess_rhat(kind=:rank, split_chains=2, maxlag=250, AutocovMethod) on 1000 draws x 4 chains x 1e6 params Float64, AR(1) phi=0.5"


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
# CPU arm: the reference algorithm on the host cores (oracle/ref_port.cpp)
# ---------------------------------------------------------------------------------------------
def host_threads():
    """Host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 for its workers; the CPU arm
    sets its own thread count (omp_set_num_threads), so the affinity mask is what counts."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample_params(cores):
    return max(2000, 1500 * cores)


def host_ar1(params, seed=1):
    import numpy as np
    rng = np.random.default_rng(seed)
    sigma = (1 - PHI * PHI) ** 0.5
    x = rng.standard_normal((params, CHAINS, DRAWS)) * sigma
    for t in range(1, DRAWS):
        x[:, :, t] += PHI * x[:, :, t - 1]
    return x.transpose(2, 1, 0)        # (draws, chains, params), column-major


def run_cpu(x, steps, warmup):
    """Times oracle/ref_port.cpp (all host threads) on the (draws, chains, sample) array x."""
    from oracle import ref_port as rp
    cores = host_threads()
    for _ in range(warmup):
        rp.ess_rhat(x, kind="rank", nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        rp.ess_rhat(x, kind="rank", nthreads=cores)
    dt = time.perf_counter() - t0
    return x.shape[2] * steps / dt, cores, dt / steps


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import build_oracle, ref_port as rp
    build_oracle.build()
    cores = host_threads()
    sample = cpu_sample_params(cores)
    x = host_ar1(sample)
    value, cores, sec = run_cpu(x, max(1, args.steps), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "params/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic AR(1) phi=0.5 (host numpy)",
        "config": {"workload": WORKLOAD, "sample": f"each step = {sample} of the 1e6 parameters (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": "params/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} params/step; C++/OpenMP restatement of the reference algorithm "
                                   "(oracle/ref_port.cpp) — the Julia package cannot run here (no Julia in the image)"},
        "e2e": {"value": value, "unit": "params/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def gpu_arm(args):
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

    total = args.params
    lo, hi = mcd.sharding.shard_range(total, rank, world)
    shard = hi - lo
    sigma = (1 - PHI * PHI) ** 0.5
    x = mcd.generate_ar1(PHI, sigma, DRAWS, CHAINS, shard, seed=1, param_offset=lo, device=local)
    torch.cuda.synchronize()

    def step():
        S, R = mcd.ess_rhat(x, kind="rank")
        if world > 1:
            return mcd.sharding.gather_params(torch.stack((S, R)), total, dst=0)
        return S, R

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # kernel-only timing of the dominant kernel (the shared-memory slab kernel = the whole device
    # step at N = 1), CUDA events on the launching stream
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
        S, R = mcd.ess_rhat(x, kind="rank")
        kev[i][1].record()
        if world > 1:
            mcd.sharding.gather_params(torch.stack((S, R)), total, dst=0)
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

    # ---- e2e: public API, host (pinned) input, H2D + D2H inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        e2e_params = min(shard, args.e2e_params // world if args.e2e_params else shard)
        # pinned staging of the whole shard (32 GB at N = 1); if the host cannot pin that much, every
        # rank halves its sample together (the choice is agreed with an all-reduce) and the line says so
        xh_t = None
        while True:
            try:
                xh_t = torch.empty((e2e_params, CHAINS, DRAWS), dtype=torch.float64, pin_memory=True)
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
            if e2e_params < 1000:
                raise SystemExit("cannot pin host memory for the e2e measurement")
        xh_t.copy_(x.permute(2, 1, 0)[:e2e_params])
        torch.cuda.synchronize()
        xh = xh_t.numpy().transpose(2, 1, 0)          # (draws, chains, params) column-major view
        mcd.ess_rhat(xh, kind="rank")                # warm-up (allocates staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            Sh, Rh = mcd.ess_rhat(xh, kind="rank")
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": e2e_params * world / float(tt[0]), "unit": "params/s",
               "h2d_bytes_per_step": e2e_params * world * DRAWS * CHAINS * 8,
               "d2h_bytes_per_step": e2e_params * world * 16,
               "params_per_step": e2e_params * world, "host_memory": "pinned", "ms_per_step": float(tt[0]) * 1e3,
               "sample": "whole array" if e2e_params == shard else f"first {e2e_params} parameters of each shard (host could not pin more)"}
        # parity spot check against the device-resident result
        assert np.array_equal(Sh, S[:e2e_params].cpu().numpy()), "host-staged and device-resident results differ"
        del xh_t

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    achieved = shard * BYTES_PER_PARAM / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_param")
            traffic = traffic * shard if traffic else None
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": "params/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64",
        "data": "synthetic AR(1) phi=0.5, seed 1, generated on device (Philox4x32-10, Box-Muller)",
        "config": {"workload": WORKLOAD, "params_total": total, "params_per_gpu": shard,
                   "sharding": f"contiguous parameter ranges over {world} rank(s); results gathered to rank 0 (NCCL)",
                   "l2": "input per GPU (%.1f GB) is larger than the 126 MB L2; no flush needed" % (shard * 32000 / 1e9)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": {1: "mcd::slab_kernel<double,256>", 2: "large-slab pipeline", 3: "mcd::fast_kernel<double>"}.get(ctx.stat("last_path"), "?"),
                     "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": shard * BYTES_PER_PARAM},
        "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
    }
    if world == 1 and not args.no_cpu:
        from oracle import build_oracle, ref_port as rp
        build_oracle.build()
        cores = host_threads()
        sample = min(shard, cpu_sample_params(cores))
        xs = np.asfortranarray(x[:, :, :sample].cpu().numpy())
        v, cores, sec = run_cpu(xs, 1, 0)
        Sc, Rc = rp.ess_rhat(xs[:, :, :256], kind="rank")
        ok = bool(np.allclose(Sc, S[:256].cpu().numpy(), rtol=1e-8) and np.allclose(Rc, R[:256].cpu().numpy(), rtol=1e-8))
        line["cpu_baseline"] = {"value": v, "unit": "params/s", "cores": cores, "kind": "port",
                                "sample": f"first {sample} of the GPU run's parameters (identical inputs), one pass; "
                                          "C++/OpenMP restatement of the reference algorithm (oracle/ref_port.cpp); "
                                          "Julia is not available in this image",
                                "matches_gpu_1e-8": ok}
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
    ap.add_argument("--params", type=int, default=PARAMS, help="total parameters (default: the named 1e6)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-params", type=int, default=0, help="cap on host-staged parameters (0 = all)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
