mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gemm.py -m gpu -x -q > gpurun_out/r2c_pytest_tc.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest_tc.log
timeout 900 python tools/tc_probe.py --quick --out gpurun_out/r2c_tc_probe.json > gpurun_out/r2c_tc_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2c_tc_probe.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_tc_launches.csv python tools/tc_ncu_target.py 7 1922 8 > gpurun_out/r2c_ncu_launch.log 2>&1
grep -E "tc_" gpurun_out/r2c_tc_launches.csv | tail -6 | awk -F'","' '{print $5, $NF}' | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 1 -c 1 -o gpurun_out/r2c_prof_tc_gemm python tools/tc_ncu_target.py 7 1922 8 > gpurun_out/r2c_ncu_full.log 2>&1; echo "ncu full rc=$?"
