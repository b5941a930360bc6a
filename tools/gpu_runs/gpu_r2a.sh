mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 600 python tools/tc_debug.py > gpurun_out/r2a_tc_debug.log 2>&1; echo "debug rc=$?"; tail -40 gpurun_out/r2a_tc_debug.log
timeout 900 python -m pytest tests/test_tc_gemm.py -m gpu -x -q > gpurun_out/r2a_pytest_tc.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_pytest_tc.log
timeout 900 python tools/tc_probe.py --out gpurun_out/r2a_tc_probe.json > gpurun_out/r2a_tc_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2a_tc_probe.log | cut -c1-400
