#!/bin/bash
# Round-2 evidence run (one B200): GPU tests, the bench line of every BASELINE config, the CPU reference arm, the
# ncu launch list of the headline bench command and one ncu --set full capture of the headline kernel.
# gpurun -- 'bash scripts/r2_final.sh'      (text / csv exports only: gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $O/r2_gpu_tests.log
timeout 400 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r2_bench_gpu.json 2> $O/r2_bench_gpu.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err
for c in c2summary c3fft c4nested c5bda; do
  timeout 400 python bench.py --config $c --steps 3 --warmup 3 > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err
done
python - <<'PY'
import json
for c in ("gpu", "reference_arm", "c2summary", "c3fft", "c4nested", "c5bda"):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_{c}.json"))
        r = d.get("roofline") or {}
        print(c, "value %.4g" % d["value"], "ms %.2f" % d["ms_per_step"], "frac %s" % r.get("frac"), "e2e %s" % (d.get("e2e") or {}).get("value"),
              "cpu %s" % (d.get("cpu_baseline") or {}).get("value"), "within_tol %s" % (d.get("cpu_baseline") or {}).get("fraction_within_tolerance"))
    except Exception as e:
        print(c, "FAILED", e)
PY
# launch list of the headline bench command (cold-cache, serialised: shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/r2_bench_under_ncu.log 2>&1
python scripts/sum_launches.py $O/r2_launches_bench.csv 2>/dev/null | head -8
for c in c2summary:20000 c5bda:4000; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_ll_${c%%:*}.csv python scripts/launch_list.py ${c%%:*} ${c##*:} 1 > /dev/null 2>&1
done
# one full capture of the headline kernel (40 000 parameters)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rk2_kernel -s 2 -c 1 -f -o /tmp/r2_rk2 python scripts/prof_rank.py 40000 > $O/r2_rk2_ncu.log 2>&1
ncu -i /tmp/r2_rk2.ncu-rep --page details > $O/r2_rk2_details.txt 2>/dev/null
ncu -i /tmp/r2_rk2.ncu-rep --page raw --csv > $O/r2_rk2_raw.csv 2>/dev/null
ncu -i /tmp/r2_rk2.ncu-rep --page source --csv > $O/r2_rk2_source.csv 2>/dev/null
gzip -f $O/r2_rk2_source.csv
rm -f /tmp/r2_rk2.ncu-rep
du -sh $O
