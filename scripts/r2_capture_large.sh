mkdir -p gpurun_out
NCU="ncu --set full --clock-control none"
timeout 200 $NCU -k regex:"crank_|chain_stats|fft4p_" -c 12 -f -o /tmp/c4 python scripts/launch_list.py c4nested 200 1 > gpurun_out/r2_c4_final_ncu.log 2>&1
ncu -i /tmp/c4.ncu-rep --page details > gpurun_out/r2_c4_final_details.txt 2>/dev/null
ncu -i /tmp/c4.ncu-rep --page raw --csv > gpurun_out/r2_c4_final_raw.csv 2>/dev/null
timeout 200 $NCU -k regex:"fft4p_|fft4_cols_inv" -c 3 -f -o /tmp/c3 python scripts/launch_list.py c3fft 8 1 > gpurun_out/r2_c3_final_ncu.log 2>&1
ncu -i /tmp/c3.ncu-rep --page details > gpurun_out/r2_fft_final_details.txt 2>/dev/null
ncu -i /tmp/c3.ncu-rep --page raw --csv > gpurun_out/r2_fft_final_raw.csv 2>/dev/null
ls -la gpurun_out/r2_c4_final* gpurun_out/r2_fft_final*
