mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_tc_launches.csv python tools/tc_ncu_target.py 7 1922 8 > gpurun_out/r2b_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 1 -c 1 -o gpurun_out/r2b_prof_tc_gemm python tools/tc_ncu_target.py 7 1922 8 > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2b_prof_tc_gemm.ncu-rep --page raw --csv > gpurun_out/r2b_prof_tc_gemm_raw.csv 2>/dev/null
grep -E "tc_|Duration" gpurun_out/r2b_tc_launches.csv | head -30
