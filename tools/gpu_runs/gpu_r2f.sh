mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_tc_gemm.py -m gpu -x -q -k "tc or wrappers" > gpurun_out/r2f_pytest_tc_stage.log 2>&1; echo "pytest stage rc=$?"; tail -5 gpurun_out/r2f_pytest_tc_stage.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "c64_api or digits" > gpurun_out/r2f_pytest_parity.log 2>&1; echo "pytest parity rc=$?"; grep -E "digits|ex1_o15|passed|failed|Error" gpurun_out/r2f_pytest_parity.log | tail -14
for dg in 0 5 4; do timeout 900 python tools/stage_timing.py --nb 64 --digits $dg --check > gpurun_out/r2f_stage_d$dg.log 2>&1; echo "stage digits=$dg rc=$?"; grep -E "layer_smatrix|redheffer|eig\(total\)|total|parity" gpurun_out/r2f_stage_d$dg.log; done
