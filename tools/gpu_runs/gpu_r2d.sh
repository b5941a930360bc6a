mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gemm.py -m gpu -x -q > gpurun_out/r2e_pytest_tc.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2e_pytest_tc.log
timeout 900 python tools/tc_probe.py --quick --out gpurun_out/r2e_tc_probe.json > gpurun_out/r2e_tc_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2e_tc_probe.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/r2e_tc_launches.csv python tools/tc_ncu_target.py 7 1922 8 > gpurun_out/r2e_ncu_launch.log 2>&1
grep -E "tc_gemm" gpurun_out/r2e_tc_launches.csv | tail -2 | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-250
