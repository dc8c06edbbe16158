#!/bin/bash
# Refresh of the round-2 bench lines after the last library change (no profiler): GPU tests, smoke, every config.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -2 | tee $O/r2_gpu_tests.log
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r2_bench_gpu.json 2> $O/r2_bench_gpu.err
for c in c2summary c3fft c4nested c5bda; do
  timeout 400 python bench.py --config $c --steps 3 --warmup 3 > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err
done
python - <<'PY'
import json
for c in ("gpu", "c2summary", "c3fft", "c4nested", "c5bda"):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_{c}.json"))
        r = d.get("roofline") or {}
        print(c, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "frac %.4f" % r.get("frac"), "e2e %.4g" % (d.get("e2e") or {}).get("value"),
              "cpu %.4g" % (d.get("cpu_baseline") or {}).get("value"), "within_tol %s" % (d.get("cpu_baseline") or {}).get("fraction_within_tolerance"))
    except Exception as e:
        print(c, "FAILED", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_ll_c3fft.csv python scripts/launch_list.py c3fft 8 1 > /dev/null 2>&1
python scripts/sum_launches.py $O/r2_ll_c3fft.csv 2>/dev/null | head -14
