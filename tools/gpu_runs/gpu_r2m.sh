mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 1 -c 1 -o gpurun_out/r2m_prof_tc_gemm python tools/tc_ncu_target.py 7 1922 8 > gpurun_out/r2m_ncu_full.log 2>&1; echo "ncu full rc=$?"
