#!/bin/bash
# ncu --set full captures of the kernels off the headline path (round-2 evidence: FFT path FP64 pipe, C4 sort
# pipeline, fused summary).  The reports are exported to text / csv pages on the box (gpurun_out/ is capped at
# 64 MiB) and deleted.   gpurun -- 'bash scripts/r2_evidence.sh'
mkdir -p gpurun_out
# counting-rank path of the large-slab pipeline: bit-identity against the sort path, timings
timeout 400 python scripts/crank_probe.py 400 2>&1 | tee gpurun_out/crank_probe.log | tail -45
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name regex count command...
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 300 $NCU -k regex:"$rx" -c $cnt -f -o /tmp/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  gzip -f gpurun_out/${name}_source.csv
  rm -f /tmp/$name.ncu-rep
}
cap r2_fft "fft4_|fft_chain|fft_block" 4 python scripts/launch_list.py c3fft 8 1
cap r2_c4 "crank_|chain_stats|nested_kernel|tile_sort|merge_pass|rank_kernel" 12 python scripts/launch_list.py c4nested 200 1
cap r2_summary "fastgen" 1 python scripts/launch_list.py c2summary 20000 1
ls -la gpurun_out | tail -20
du -sh gpurun_out
