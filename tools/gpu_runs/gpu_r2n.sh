mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -6 gpurun_out/r2n_pytest_gpu.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2n_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "parity_c128 or c64_api" 2>&1 | grep -E "new-c64|S-parameter max err" | tail -30
