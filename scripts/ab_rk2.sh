#!/bin/bash
# Same-box A/B of headline-kernel variants: for every ab_libs/lib<name>.so given, parity of the rk2 kernel against
# the round-1 kernel on assorted inputs and CUDA-event timings of both (the round-1 kernel is the same code in
# every variant: its time is the box-consistency control).   gpurun -- 'bash scripts/ab_rk2.sh base v1 v2'
mkdir -p gpurun_out
for name in "$@"; do
  echo "=== $name" | tee -a gpurun_out/ab_rk2.log
  MCMCDIAG_B200_LIB=$PWD/ab_libs/lib$name.so timeout 300 python scripts/rk2_probe.py 200000 0.5 > gpurun_out/ab_rk2_$name.log 2>&1
  cat gpurun_out/ab_rk2_$name.log >> gpurun_out/ab_rk2.log
  grep -v "max rel diff old vs rk2 = 0.000e+00" gpurun_out/ab_rk2_$name.log | tail -25
  grep -c "max rel diff old vs rk2 = 0.000e+00" gpurun_out/ab_rk2_$name.log
done
