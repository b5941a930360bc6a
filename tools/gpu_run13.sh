#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eig.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r1o_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/r1o_pytest_gpu.log
for v in "split:" "nosplit:11=0"; do
  name=${v%%:*}; tune=${v#*:}
  RCWA_B200_TUNE="$tune" timeout 300 python tools/stage_timing.py --nb 128 --check > gpurun_out/r1o_stage_$name.log 2>&1; echo "stage $name rc=$?"
  grep -h "parity\|eig(total)\|layers/s\|hessenberg(alone)" gpurun_out/r1o_stage_$name.log
done
